// Persistent version of the warp-specialised fused attention forward (attn_fwd_ws.cuh): same roles, streams and TMEM
// plan, but a CTA is launched once per SM slot and walks a strided list of work items (Q tile group, head, batch).
//
// Why: a non-persistent CTA costs ~5.9 us beyond its steady-state steps (launch, TMEM allocation, barrier init, cold Q / K
// loads, pipeline fill, epilogue) -- 23 % of the run time at T = 2048, 7 % at T = 8192 (profiles/r01_ws_tiles_per_cta_l.txt).
// Here the TMA producer runs ahead into the next item (Q is double buffered, the K / V rings simply continue), the MMA
// issuer queues the next item's first QK^T behind the last PV, and the softmax streams go from their epilogue straight
// into scores that are already waiting; only the O read-out of the epilogue is not overlapped.
//
// Bookkeeping: all ring / parity decisions use GLOBAL step counters (gk = steps so far, gv = PV steps so far), so the
// s_full / p_full / k_* / v_* barriers keep alternating across items; per-item barriers (q_full / q_empty by Q buffer,
// o_final, epi_done) flip once per item.  epi_done[t] (8 arrivals: every softmax warp of the tile, after its
// tcgen05.ld of O) gates the next item's first PV, which overwrites the accumulators.
#pragma once

#include "attn_fwd_ws.cuh"

namespace oeh {
namespace attn_wsp {

using attn_ws::BNW;
using attn_ws::kSubBytes;
using attn_ws::kLazyRescale;
using attn_ws::reg_alloc;
using attn_ws::reg_dealloc;

using namespace attn;

// kTiles = Q tiles (of 128 rows) per CTA.  2: one CTA per SM (640 threads, 512 TMEM columns), K / V steps shared by the two
// tiles.  1: TWO independent CTAs per SM (320 threads, 256 TMEM columns each): the same warps / TMEM per SM, K / V loaded
// once per tile (twice the L2 -> smem traffic), but a CTA's prologue (TMEM allocation, first Q / K loads, pipeline
// fill) and epilogue -- ~5.9 us per CTA, a quarter of the run time at T = 2048 -- overlap its neighbour's steady state.
template <int kTiles>
struct Cfg {
  static constexpr int kSoftmaxWarps = 8 * kTiles;
  static constexpr int kWarpMma0 = kSoftmaxWarps, kWarpTma = kSoftmaxWarps + 1, kWarpMma1 = kSoftmaxWarps + 2;
  // four auxiliary warps in both shapes (MMA issuer(s), TMA producer, idle): setmaxnreg moves registers between warps of
  // the SAME sub-partition, so each of the four schedulers needs one warp that gives its registers away
  static constexpr int kThreads = (kSoftmaxWarps + 4) * 32;          // 640 | 384
  // launch allocation 96 | 80 registers per thread; 16*112 + 4*24 <= 20*96, 8*104 + 4*24 <= 12*80
  static constexpr int kSoftmaxRegs = kTiles == 2 ? 112 : 104, kAuxRegs = 24;
  static constexpr int kCtasPerSm = kTiles == 2 ? 1 : 2;
  static constexpr int kQRows = 128 * kTiles;                        // query rows per CTA
  // Seven (two tiles) / three (one tile, two CTAs per SM) K / V stages and a 2 KB staging slot per softmax warp for the context store (the item's epilogue wrote
  // 16-byte pieces into 32 different lines per instruction; those 2048 partial-line requests per item also held up the MMA
  // issuer's and the producer's barrier polls behind them in the load / store pipeline: profiles/r02_timeline_wsp.txt)
  static constexpr int kStages = kTiles == 2 ? 7 : 3;
  static constexpr int kStageOut = kSoftmaxWarps * 2048;
  static constexpr int kSmemBytes = 2 * kTiles * kTileBytes + 2 * kStages * kSubBytes + kStageOut + 1024;   // 2 x Q tiles | K ring | V ring | store staging
  static constexpr uint32_t kTmemCols = 256 * kTiles;
  static constexpr uint32_t kColO0 = 128 * kTiles;                   // S_t[b] at (2t+b)*64, O_{t,b} at kColO0 + (2t+b)*64
};
template <int kStages>
struct Barriers {
  uint64_t q_full[2], q_empty[2];
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2][2], p_full[2][2];
  uint64_t o_final[2], epi_done[2];
};

#ifdef OEH_TIMELINE
// experiment builds only (tools/timeline_wsp.py): CTA 0, first 8 items, 4 recorders (MMA issuer of tile 0, TMA producer, softmax
// streams (tile 0, parity 0 / 1)), 8 stamps each
__device__ unsigned long long* g_wsp_timeline = nullptr;
__device__ __forceinline__ void wsp_stamp(int it, int who, int k) {
  if (blockIdx.x == 0 && it < 8 && (threadIdx.x & 31) == 0 && g_wsp_timeline) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_wsp_timeline[(it * 4 + who) * 8 + k] = t;
  }
}
#define WSP_STAMP(it, who, k) wsp_stamp(it, who, k)
#else
#define WSP_STAMP(it, who, k)
#endif

template <bool kBF16, bool kGeneric, bool kNorm, int kTiles>
__global__ void __launch_bounds__(Cfg<kTiles>::kThreads, Cfg<kTiles>::kCtasPerSm)
attn_fwd_wsp_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                    const __grid_constant__ CUtensorMap tmv, const __grid_constant__ AttnParams p) {
  using C = Cfg<kTiles>;
  constexpr int kStages = C::kStages, kQRows = C::kQRows;
  constexpr uint32_t kTmemCols = C::kTmemCols, kColO0 = C::kColO0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Barriers<kStages> bars;
  __shared__ uint32_t tmem_base_slot;
  // (item parity, tile, stream parity, row): statistics merged across streams.  Double-buffered by item: the barrier of item i+1's
  // merge orders item i's reads before item i+2's writes (one buffer relied on the pipeline handshakes in between for that)
  __shared__ float xch_m[2][kTiles][2][128], xch_l[2][kTiles][2][128];

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it
  // (role, TMEM lane base, barrier / descriptor addresses) on the uniform datapath instead of R2UR-ing it at every use
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem0;                                    // 2 buffers of kTiles tiles of 128 rows
  const uint32_t sK = smem0 + 2 * kTiles * kTileBytes;          // kStages steps of 64 keys
  const uint32_t sV = sK + kStages * kSubBytes;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars.q_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.q_empty[i]), kTiles);  // one tcgen05.commit from each MMA issuer
    }
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars.k_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.k_empty[i]), kTiles);  // one tcgen05.commit from each MMA issuer
      ptx::mbar_init(ptx::smem_u32(&bars.v_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.v_empty[i]), kTiles);
    }
#pragma unroll
    for (int t = 0; t < kTiles; ++t) {
#pragma unroll
      for (int sb = 0; sb < 2; ++sb) {
        ptx::mbar_init(ptx::smem_u32(&bars.s_full[t][sb]), 1);
        ptx::mbar_init(ptx::smem_u32(&bars.p_full[t][sb]), 4);  // one elected lane of each of the stream's 4 warps
      }
      ptx::mbar_init(ptx::smem_u32(&bars.o_final[t]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.epi_done[t]), 8);      // one elected lane of each softmax warp of the tile
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (warp == C::kWarpMma0) ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  // ---- work items and their step plans (every role walks the same list) ---------------------------------------------
  const int n_items = p.n_qt * p.H * p.B;
  const bool softmax1 = (p.kind == OEH_SOFTMAX_1 || p.kind == OEH_SOFTMAX_CLIPPED1);
  const bool clipped = p.kind >= OEH_SOFTMAX_CLIPPED;
  constexpr bool normalized = kNorm;                              // host: clipped || qdq_probs
  struct Plan {
    int q0, h, b, n_kv, nA, n_iter;
  };
  auto make_plan = [&](int item) {
    Plan pl;
    // causal: heaviest Q tiles first over chunks of heads (attn_common.cuh, causal_item)
    int qt = item % p.n_qt, hb = item / p.n_qt;
    if (p.causal) causal_item(item, p.n_qt, p.H * p.B, p.hb_chunk, qt, hb);
    pl.q0 = qt * kQRows;
    pl.h = hb % p.H;
    pl.b = hb / p.H;
    int n_kv = (p.Tk + BNW - 1) / BNW;
    if (p.causal) {
      const int last_row = min(pl.q0 + kQRows, p.Tq) - 1;        // last real query row of this item
      const int lim = (last_row + (p.Tk - p.Tq)) / BNW + 1;
      n_kv = lim < n_kv ? lim : n_kv;
    }
    pl.n_kv = n_kv;
    pl.nA = normalized ? n_kv : 0;                               // stats-pass steps (QK^T only)
    pl.n_iter = pl.nA + n_kv;
    return pl;
  };

  // Item `it` of this CTA.  Non-causal: stride gridDim.x.  Causal: the items are sorted by length (heaviest first), so a
  // plain stride would hand CTA 0 the longest item of every round; walking odd rounds backwards ("snake") evens the
  // sums out to within one item.  A CTA without an item in the last, partial round is done (no round follows it).
  auto item_of = [&](int it) {
    const int base = it * (int)gridDim.x;
    if (base >= n_items) return n_items;
    const int c = (p.causal && (it & 1)) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    return base + c;
  };

  if (warp >= C::kSoftmaxWarps) reg_dealloc<C::kAuxRegs>();
  if (warp == C::kWarpTma) {
    // ================================ TMA producer (whole warp waits, one elected lane issues) ============
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmq);
      ptx::prefetch_tmap(&tmk);
      ptx::prefetch_tmap(&tmv);
    }
    __syncwarp();
    int gk = 0, gv = 0;                                         // K loads / V loads issued by earlier items
    for (int it = 0, item; (item = item_of(it)) < n_items; ++it) {
      const Plan pl = make_plan(item);
      const int qb = it & 1;
      if (it >= 2) ptx::mbar_wait(ptx::smem_u32(&bars.q_empty[qb]), (uint32_t)((it >> 1) - 1) & 1u);
      if (ptx::elect_one()) {
        const uint32_t bq = ptx::smem_u32(&bars.q_full[qb]);
        const uint32_t dst = sQ + qb * (kTiles * kTileBytes);
        ptx::mbar_expect_tx(bq, kTiles * kTileBytes);
        tma_tile(dst, &tmq, bq, p.cq, pl.q0, pl.h, pl.b);
        if (kTiles == 2) tma_tile(dst + kTileBytes, &tmq, bq, p.cq, pl.q0 + 128, pl.h, pl.b);
      }
      __syncwarp();
      auto load_k = [&](int i) {                                // K of step i of this item -> ring slot (gk + i) % kStages
        const int g = gk + i, s_ = g % kStages, n = g / kStages;
        if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.k_empty[s_]), (uint32_t)(n - 1) & 1u);
        if (ptx::elect_one()) {
          const uint32_t bf = ptx::smem_u32(&bars.k_full[s_]);
          ptx::mbar_expect_tx(bf, kSubBytes);
          tma_tile(sK + s_ * kSubBytes, &tmk, bf, p.ck, (i < pl.nA ? i : i - pl.nA) * BNW, pl.h, pl.b);
        }
        __syncwarp();
      };
      load_k(0);
      if (pl.n_iter > 1) load_k(1);
      for (int i = 0; i < pl.n_iter; ++i) {                     // K runs two steps ahead of V, like the MMA issuers
        if (i + 2 < pl.n_iter) load_k(i + 2);
        if (i >= pl.nA) {
          const int v = i - pl.nA, g = gv + v, s_ = g % kStages, n = g / kStages;
          if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.v_empty[s_]), (uint32_t)(n - 1) & 1u);
          if (ptx::elect_one()) {
            const uint32_t bf = ptx::smem_u32(&bars.v_full[s_]);
            ptx::mbar_expect_tx(bf, kSubBytes);
            tma_tile(sV + s_ * kSubBytes, &tmv, bf, p.cv, v * BNW, pl.h, pl.b);
          }
          __syncwarp();
        }
      }
      gk += pl.n_iter;
      gv += pl.n_kv;
    }
  } else if (warp == C::kWarpMma0 || (kTiles == 2 && warp == C::kWarpMma1)) {
    // ================================ MMA issuers: one warp per Q tile =====================================
    const int t = (warp - C::kWarpMma0) >> 1;
    constexpr uint32_t kIdescQK = ptx::make_idesc_f16(kBF16, 128, BNW, false, false);
    constexpr uint32_t kIdescPV = ptx::make_idesc_f16(kBF16, 128, HD, false, true);
    constexpr uint64_t kTileUnits = kTileBytes >> 4;             // descriptor address units (16 B)
    constexpr uint64_t kSubUnits = kSubBytes >> 4;
    const uint64_t dq = ptx::make_smem_desc_sw128(sQ, 16, 1024) + t * kTileUnits;
    const uint64_t dk = ptx::make_smem_desc_sw128(sK, 16, 1024);
    const uint64_t dv = ptx::make_smem_desc_sw128(sV, 1024, 1024);
    const uint32_t tS0 = tmem + (uint32_t)(2 * t) * 64u, tO0 = tmem + kColO0 + (uint32_t)(2 * t) * 64u;
    auto issue_pv = [&](int sb, int stage, bool first) {         // O_{t,sb} (+)= P_t[sb] . V(stage)  (K = 64 keys)
#pragma unroll
      for (int k = 0; k < BNW / 16; ++k)
        ptx::mma_ts(tO0 + sb * 64, tS0 + sb * 64 + k * 8, dv + stage * kSubUnits + k * 128, kIdescPV,
                    (first && k == 0) ? 0u : 1u);
    };
    int gk = 0, gv = 0;                                          // steps / PV steps of earlier items
    bool pre_issued = false;                                     // this item's first two QK^T were issued by the previous item
    for (int it = 0, item; (item = item_of(it)) < n_items; ++it) {
      const Plan pl = make_plan(item);
      const int qb = it & 1;
      const uint64_t dqi = dq + (uint64_t)qb * (kTiles * kTileUnits);
      const uint64_t dqn = dq + (uint64_t)(qb ^ 1) * (kTiles * kTileUnits);   // the next item's Q buffer
      auto issue_qk = [&](uint64_t dqx, int sb, int stage) {     // S_t[sb] = Q_t . K(stage)^T with the Q buffer dqx
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::mma_ss(tS0 + sb * 64, dqx + k * 2, dk + stage * kSubUnits + k * 2, kIdescQK, k > 0);
      };
      // The score tiles run two steps ahead of the PV steps, and they keep doing so ACROSS items: the last two steps of an item
      // issue the first two QK^T of the next one (its Q and K are long there: the producer runs a ring ahead), so that the softmax
      // streams come out of their epilogue into scores that are waiting.  (Issued by the next item's own prologue they were ~0.5 us
      // late, and held up behind the epilogue's shared-memory / store burst: profiles/r02_timeline_wsp.txt.)
      const int next_item = item_of(it + 1);
      const bool cross = next_item < n_items && pl.n_iter >= 2 && make_plan(next_item).n_iter >= 2;
      if (t == 0) WSP_STAMP(it, 0, 0);
      if (!pre_issued) {
        ptx::mbar_wait(ptx::smem_u32(&bars.q_full[qb]), (uint32_t)(it >> 1) & 1u);
        if (t == 0) WSP_STAMP(it, 0, 1);
        for (int i = 0; i < 2 && i < pl.n_iter; ++i) {           // prologue: the item's first two score tiles
          const int g = gk + i, ks = g % kStages;
          ptx::mbar_wait(ptx::smem_u32(&bars.k_full[ks]), (uint32_t)(g / kStages) & 1u);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_qk(dqi, g & 1, ks);
            ptx::mma_commit(ptx::smem_u32(&bars.s_full[t][g & 1]));
            ptx::mma_commit(ptx::smem_u32(&bars.k_empty[ks]));
          }
          __syncwarp();
        }
      }
      if (t == 0) WSP_STAMP(it, 0, 2);
      for (int i = 0; i < pl.n_iter; ++i) {
        const int g = gk + i, sb = g & 1;
        const bool pv = i >= pl.nA;
        const int v = i - pl.nA, gvv = gv + v, vs = pv ? gvv % kStages : 0;
        const bool more = i + 2 < pl.n_iter;                     // this item's step i + 2 ...
        const bool ahead = !more && cross;                       // ... or the next item's step i + 2 - n_iter (0 or 1)
        const bool last = i + 1 == pl.n_iter;
        const int g2 = g + 2, ks = g2 % kStages;
        if (pv) ptx::mbar_wait(ptx::smem_u32(&bars.v_full[vs]), (uint32_t)(gvv / kStages) & 1u);
        if (ahead && i + 2 == pl.n_iter) ptx::mbar_wait(ptx::smem_u32(&bars.q_full[qb ^ 1]), (uint32_t)((it + 1) >> 1) & 1u);
        if (more || ahead) ptx::mbar_wait(ptx::smem_u32(&bars.k_full[ks]), (uint32_t)(g2 / kStages) & 1u);
        ptx::mbar_wait(ptx::smem_u32(&bars.p_full[t][sb]), (uint32_t)(g >> 1) & 1u);
        // the item's first PV overwrites an accumulator: the previous item's epilogue must have read both of them
        if (pv && v == 0 && it > 0) ptx::mbar_wait(ptx::smem_u32(&bars.epi_done[t]), (uint32_t)(it - 1) & 1u);
        ptx::tc_fence_after();
        if (t == 0 && i == 0) WSP_STAMP(it, 0, 3);
        if (t == 0 && i == 1) WSP_STAMP(it, 0, 4);
        if (t == 0 && i + 2 == pl.n_iter) WSP_STAMP(it, 0, 5);
        if (t == 0 && last) WSP_STAMP(it, 0, 6);
        if (ptx::elect_one()) {
          if (pv) {
            issue_pv(sb, vs, v < 2);                             // the first PV step of each parity starts its accumulator
            if (last) ptx::mma_commit(ptx::smem_u32(&bars.o_final[t]));
            ptx::mma_commit(ptx::smem_u32(&bars.v_empty[vs]));
          }
          if (last) ptx::mma_commit(ptx::smem_u32(&bars.q_empty[qb]));   // every QK^T of THIS item has been issued
          if (more || ahead) {
            issue_qk(more ? dqi : dqn, sb, ks);
            ptx::mma_commit(ptx::smem_u32(&bars.s_full[t][sb]));
            ptx::mma_commit(ptx::smem_u32(&bars.k_empty[ks]));
          }
        }
        __syncwarp();
      }
      pre_issued = cross;
      gk += pl.n_iter;
      gv += pl.n_kv;
    }
  } else if (warp < C::kSoftmaxWarps) {
    // ================================ softmax streams ======================================================
    reg_alloc<C::kSoftmaxRegs>();
    const int t = warp >> 3, par = (warp >> 2) & 1;               // Q tile, step parity
    const int r = (warp & 3) * 32 + (tid & 31);                   // row within the tile == TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tS = t_lane + (uint32_t)(2 * t + par) * 64u;              // this stream's S / P columns
    const uint32_t tO = t_lane + kColO0 + (uint32_t)(2 * t + par) * 64u;     // this stream's accumulator
    const uint32_t bar_s = ptx::smem_u32(&bars.s_full[t][par]), bar_p = ptx::smem_u32(&bars.p_full[t][par]);

    const bool has_qs = kGeneric && p.qdq_scores != nullptr;
    const bool has_qp = kGeneric && p.qdq_probs != nullptr;
    const bool has_mask = kGeneric && p.mask != nullptr;
    Qdq qd_s{1.f, 1.f, 0.f, p.int_max}, qd_p{1.f, 1.f, 0.f, p.int_max};
    if (has_qs) qd_s = load_qdq(p.qdq_scores, p.int_max);
    if (has_qp) qd_p = load_qdq(p.qdq_probs, p.int_max);
    int gk = 0;                                                   // steps of earlier items
    // the plan of an item (integer divisions, ~0.4 us) is made while the PREVIOUS item's epilogue waits for its statistics merge
    Plan pl_next = make_plan(item_of(0) < n_items ? item_of(0) : 0);
    for (int it = 0, item; (item = item_of(it)) < n_items; ++it) {
    const Plan pl = pl_next;
    const int q0 = pl.q0, h = pl.h, b = pl.b, n_kv = pl.n_kv, nA = pl.nA, n_iter = pl.n_iter;
    const int row = q0 + t * 128 + r;                             // this thread's query row
    const int64_t mask_row =
        p.mask ? ((int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)(row < p.Tq ? row : 0) * p.ms2) : 0;
    const int causal_lim = p.causal ? row + (p.Tk - p.Tq) : 0x7fffffff;

    float m_run = -INFINITY, l_run = 0.f, inv_den = 1.f, skip_thr = -INFINITY;
    bool merged = false, started = false;                         // statistics merged / accumulator holds a PV result

    // merge (m_b, l_b) of the two streams of a row: afterwards both hold the row's m in the common frame
    auto merge_streams = [&](float& a_own, float& a_other, float& l_other) {
      xch_m[it & 1][t][par][r] = m_run;
      xch_l[it & 1][t][par][r] = l_run;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      const float m_o = xch_m[it & 1][t][par ^ 1][r];
      l_other = xch_l[it & 1][t][par ^ 1][r];
      const float m = fmaxf(m_run, m_o);
      const float m_safe = m == -INFINITY ? 0.f : m;
      a_own = ptx::ex2(m_run - m_safe);                           // -inf -> 0
      a_other = ptx::ex2(m_o - m_safe);
      m_run = m;
    };

    for (int i = (par - gk) & 1; i < n_iter; i += 2) {            // the steps whose GLOBAL index has this stream's parity
      const int g = gk + i;
      const bool pv_iter = i >= nA;
      const int k0 = (pv_iter ? i - nA : i) * BNW;
      if (normalized && pv_iter && !merged) {                     // stats pass done: the row statistics become common
        float a_own, a_other, l_other;
        merge_streams(a_own, a_other, l_other);
        l_run = l_run * a_own + l_other * a_other;
        merged = true;
      }
      ptx::mbar_wait(bar_s, (uint32_t)(g >> 1) & 1u);
      ptx::tc_fence_after();
      if (t == 0 && (warp & 3) == 0 && i < 2) WSP_STAMP(it, 2 + par, 0);
      if (t == 0 && (warp & 3) == 0 && i >= 2 && i < 4) WSP_STAMP(it, 2 + par, 1);
      if (t == 0 && (warp & 3) == 0 && i + 2 >= n_iter) WSP_STAMP(it, 2 + par, 2);
      // (two-tile CTAs only: the one-tile shapes -- T <= 512 -- have two hidden warp-steps per item and the test costs them more)
      if (kTiles == 2 && p.causal && !(kGeneric && p.dbg_s)) {
        // Causal diagonal: a 64-key step none of this warp's 32 rows can see (12 of the 32 warp-steps of a 256-row item's
        // diagonal region) adds nothing to (m, l) and its P is 0 -- no tcgen05.ld, no exponentials; the PV MMA of the step
        // (issued regardless) accumulates zeros.  Once hidden, every later step of the item is hidden for this warp too.
        const int vis0 = min(min(p.Tk - k0, BNW) - 1, causal_lim - k0);
        if (__all_sync(0xffffffffu, vis0 < 0)) {
          if (pv_iter) {
            uint32_t z[BNW / 2];
#pragma unroll
            for (int j = 0; j < BNW / 2; ++j) z[j] = 0u;
            ptx::tmem_st32(tS, z);
            ptx::tmem_wait_st();
            if (!normalized) started = true;                       // the accumulator holds a (zero) PV result
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (ptx::elect_one()) ptx::mbar_arrive(bar_p);
          continue;
        }
      }
      float s[BNW];
      ptx::tmem_ld32(tS, reinterpret_cast<uint32_t*>(s));
      ptx::tmem_ld32(tS + 32, reinterpret_cast<uint32_t*>(s) + 32);
      ptx::tmem_wait_ld();
      if (!pv_iter) {                                             // stats pass: S is consumed, QK^T(i+2) may start
        ptx::tc_fence_before();
        __syncwarp();
        if (ptx::elect_one()) ptx::mbar_arrive(bar_p);
      }
      if (kGeneric && p.dbg_s && i < 2 && row < p.Tq) {           // raw S of keys [0,128)
        float* d = p.dbg_s + (((int64_t)b * p.H + h) * p.Tq + row) * BN + k0;
#pragma unroll
        for (int j = 0; j < BNW; ++j) d[j] = s[j];
      }

      const int valid = min(p.Tk - k0, BNW);
      const int vis = min(valid - 1, causal_lim - k0);
      const float mul = transform_scores<BNW, kGeneric>(s, p, qd_s, has_qs, has_mask, mask_row, k0, valid, vis);
      const float tile_max = row_max<BNW>(s) * mul;

      if (!pv_iter) {
        // stats pass: exact running (m, l) of this stream
        const float m_new = fmaxf(m_run, tile_max);
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float alpha = ptx::ex2(m_run - m_safe);
        l_run = fmaf(l_run, alpha, exp2_and_sum<BNW, !kGeneric>(s, mul, m_safe));
        m_run = m_new;
        continue;
      }
      if (normalized) {
        // PV pass of the two-pass mode: (m, l) are final and common to both streams
        const float m_safe = m_run == -INFINITY ? 0.f : m_run;
        if (!started) {
          const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
          inv_den = __frcp_rn(den);
          skip_thr = zero_prob_threshold(p, clipped, has_qp, qd_p, den);
        }
#ifndef OEH_NO_SPARSE_SKIP
        // Clipped softmax with gamma < 0 (and / or the probs tap) maps every probability below a fixed bound to an exact
        // 0: at most 1/bound keys of a row survive, whatever the sequence length.  If none of this warp's 32 rows x
        // 64 keys can survive, P is 0 without a single exponential (the stats pass already paid for them).
        // (Own store path: merging the two branches through s[] makes ptxas spill the packed row.)
        if (!__any_sync(0xffffffffu, tile_max - m_safe > skip_thr)) {
          uint32_t z[BNW / 2];
#pragma unroll
          for (int j = 0; j < BNW / 2; ++j) z[j] = 0u;
          ptx::tmem_st32(tS, z);
          ptx::tmem_wait_st();
          ptx::tc_fence_before();
          __syncwarp();
          if (ptx::elect_one()) ptx::mbar_arrive(bar_p);
          started = true;
          continue;
        }
#endif
        exp2_only<BNW, !kGeneric>(s, mul, m_safe);
        normalize_probs<BNW, kGeneric, kBF16>(s, p, inv_den, clipped, has_qp, qd_p);
      } else {
        // online softmax with lazy rescale
        const bool move = tile_max > m_run + kLazyRescale || m_run == -INFINITY;
        const float m_new = move ? tile_max : m_run;              // tile_max may be -inf only if m_run is too
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float alpha = move ? ptx::ex2(m_run - m_safe) : 1.f;
        l_run = fmaf(l_run, alpha, exp2_and_sum<BNW, !kGeneric>(s, mul, m_safe));
        m_run = m_new;
        if (started && !__all_sync(0xffffffffu, alpha == 1.f)) {
          // O_{t,par} *= alpha.  No barrier: S_t(i) was committed behind PV_t(i-2) -- the last MMA that wrote this
          // accumulator -- in the issuer's in-order stream, and PV_t(i) needs the P this thread has not written yet.
#pragma unroll
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            ptx::tmem_ld16(tO + c * 16, o);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            ptx::tmem_st16(tO + c * 16, o);
          }
        }
      }
      started = true;

      // P -> TMEM (packed 16-bit pairs) over the first 32 columns of the S buffer this thread has consumed
      {
        uint32_t pk[BNW / 2];
#pragma unroll
        for (int j = 0; j < BNW / 2; ++j) pk[j] = pack2<kBF16>(s[2 * j], s[2 * j + 1]);
        ptx::tmem_st32(tS, pk);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (ptx::elect_one()) ptx::mbar_arrive(bar_p);
    }

    // ---- epilogue: merge the streams, O / denominator -> gate / taps -> global -----------------------------------
    if (t == 0 && (warp & 3) == 0) WSP_STAMP(it, 2 + par, 3);
    {
      const int nx = item_of(it + 1);
      if (nx < n_items) pl_next = make_plan(nx);
    }
    float a_own = 1.f, a_other = 1.f, omul = 1.f;                 // scales of the two accumulators; 1/denominator
    if (!normalized) {
      float l_other;
      merge_streams(a_own, a_other, l_other);
      const float m_safe = m_run == -INFINITY ? 0.f : m_run;
      float den = l_run * a_own + l_other * a_other;
      if (softmax1) den += ptx::ex2(-m_safe);
      omul = __frcp_rn(den);
    } else if (!merged) {                                         // a stream without PV steps still owes the barrier
      float l_other;
      merge_streams(a_own, a_other, l_other);
      a_own = a_other = 1.f;
    }
    // an accumulator exists iff some PV step had its parity (only n_kv == 1 leaves one of them untouched)
    const bool has0 = n_kv > 1 || ((gk + nA) & 1) == 0, has1 = n_kv > 1 || ((gk + nA) & 1) == 1;
    const float a0 = par ? a_other : a_own, a1 = par ? a_own : a_other;
    if (t == 0 && (warp & 3) == 0) WSP_STAMP(it, 2 + par, 4);
    ptx::mbar_wait(ptx::smem_u32(&bars.o_final[t]), (uint32_t)it & 1u);
    ptx::tc_fence_after();
    if (t == 0 && (warp & 3) == 0) WSP_STAMP(it, 2 + par, 5);
    // this thread finishes head-dim columns [32par, 32par+32) of its row from both accumulators
    float o[32];
    {
      uint32_t u0[32], u1[32];
      const uint32_t tOt = t_lane + kColO0 + (uint32_t)(2 * t) * 64u + par * 32u;
      ptx::tmem_ld32(tOt, u0);
      ptx::tmem_ld32(tOt + 64, u1);
      ptx::tmem_wait_ld();
      if (t == 0 && warp == 0) WSP_STAMP(it, 1, 0);
      ptx::tc_fence_before();                                     // the accumulators are in registers: the next item's
      __syncwarp();                                               // first PV may overwrite them
      if (ptx::elect_one()) ptx::mbar_arrive(ptx::smem_u32(&bars.epi_done[t]));
      if (has0 && has1) {                                         // (always, unless the item has a single key step)
        // plain launches: 1 / denominator folded into the two scales (finish_row then only stores)
        const float f0 = kGeneric || p.gate_logits ? a0 : a0 * omul, f1 = kGeneric || p.gate_logits ? a1 : a1 * omul;
        const uint64_t f02 = ptx::pack2f(f0, f0), f12 = ptx::pack2f(f1, f1);
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          ptx::unpack2f(ptx::ffma2(ptx::pack2f(__uint_as_float(u1[j]), __uint_as_float(u1[j + 1])), f12,
                                   ptx::fmul2(ptx::pack2f(__uint_as_float(u0[j]), __uint_as_float(u0[j + 1])), f02)),
                        o[j], o[j + 1]);
        if (!(kGeneric || p.gate_logits)) omul = 1.f;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          o[j] = (has0 ? __uint_as_float(u0[j]) * a0 : 0.f) + (has1 ? __uint_as_float(u1[j]) * a1 : 0.f);
      }
    }
    if (t == 0 && warp == 0) WSP_STAMP(it, 1, 1);
    if ((kGeneric || p.gate_logits) && row < p.Tq) finish_row<kBF16, kGeneric, 32>(o, p, omul, b, h, row, par * 32);
    if (t == 0 && warp == 0) WSP_STAMP(it, 1, 2);
    store_half_rows_coalesced<kBF16>(o, p, sV + kStages * kSubBytes + (uint32_t)warp * 2048u, b, h, q0 + t * 128 + (warp & 3) * 32,
                                     par * 32);
    if (t == 0 && (warp & 3) == 0) WSP_STAMP(it, 2 + par, 6);
    gk += n_iter;
    }   // items
  }

  // ---- teardown ------------------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == C::kWarpMma0) {
    __syncwarp();
    ptx::tmem_dealloc<kTmemCols>(tmem);
  }
}

}  // namespace attn_wsp
}  // namespace oeh
