// Fused attention forward for sm_100a when all keys of a query tile fit TMEM at once: 128 < Tk <= 512 (the OPT eval
// shape, block 512: BASELINE configs 3 / 4).  ONE pass for every softmax kind, including the normalised ones (clipped
// softmax / probs fake-quant tap) that the long-sequence kernel runs as stats pass + recomputed-QK^T PV pass:
//
//   the whole score row block S [128 x Tk] (fp32) stays in TMEM (up to 512 columns = all of it), so QK^T is computed
//   once, every exponential is evaluated once, and the row is normalised in place -- the trick attn_fwd_one.cuh plays at
//   Tk <= 128, stretched over four K/V tiles.  At B=4, H=12, T=512 causal with the three INT8 taps the two-pass kernel
//   needs 30.7 us per launch (1.6 GFLOP: prologue / epilogue bound, profiles/r01_launches_int8_opt_step_aa.summary.txt).
//
// One work item = one 128-row Q tile of one (batch, head); Q, all K tiles and all V tiles resident in shared memory (144 KB).
// At most one CTA per SM (all 512 TMEM columns); with more items than SMs a CTA walks several (see "work items" below).
//   warp 8      producer / issuer: TMA loads (Q + K tile 0 | K tiles 1..3 | V, separate mbarriers), QK^T of tile t as soon
//               as K(t) lands (one commit per tile), P.V after the softmax warps have written all of P
//   warps 0-7   softmax, TWO threads per query row: warp w and warp w+4 share TMEM lane quarter w & 3 and split the
//               row's 64-key chunks by parity (half h = w >> 2 takes chunks 2i + h), so both halves start as soon as tile 0
//               is complete and a row's serial work is halved.
//     pass A  chunk by chunk: scale / taps / mask, chunk max, e = exp2(s - m_i) in the running-max frame m_i of that moment,
//             written back over S as fp32; running (m, l).  One exchange of (m, l) through shared memory between the two
//             halves of a row -> common max and denominator (softmax_1: + 2^-m, vutils/softmax_1.py:18-20).
//     pass B  e * 2^(m_i - m) / den -> clip((eta-gamma) p + gamma, 0, 1) -> [probs tap] -> 16-bit P, packed into the TMEM
//             columns the PV MMA reads its A operand from.  P(2i), P(2i+1) land exactly on the columns of e-chunk i, which
//             its owner read in iteration i / 2: a 64-thread named barrier per iteration orders the two halves of a row.
//   O [128 x 64] takes columns 448..511 (dead once the last e-chunk has been read); epilogue as in the other kernels.
#pragma once

#include "attn_common.cuh"

namespace oeh {
namespace attn_mid {

using namespace attn;

constexpr int BM = 128;                          // query rows per CTA
constexpr int kMaxTiles = 4;                     // K / V tiles of 128 keys resident at once (Tk <= 512)
// kParts threads per query row (2 or 4): warps w, w+4, ... share TMEM lane quarter w & 3 and split a 128-key tile into kParts
// chunks of CW = 128 / kParts keys
template <int kParts>
struct Shape {
  static constexpr int CW = 128 / kParts;
  static constexpr int kSoftmaxWarps = 4 * kParts, kThreads = (kSoftmaxWarps + 1) * 32;
};
constexpr int kSmemBytes = (1 + 2 * kMaxTiles) * kTileBytes + 1024;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColO = 448;

#ifdef OEH_TIMELINE
// experiment builds only (tools/timeline_mid.py): 8 stamps per work item by thread 0 (softmax warp 0): 0 item start, 1 (first item of
// a CTA) TMEM allocated, 2 first S tile ready, 3 pass A done, 4 pass B done (P written), 5 O ready, 6 stored; slot 7 = (n_kv << 16) | SM id
__device__ unsigned long long* g_mid_timeline = nullptr;
__device__ __forceinline__ void mid_stamp(int k, int item, int n_kv = 0) {
  if (threadIdx.x == 0 && g_mid_timeline) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_mid_timeline[item * 8 + k] = t;
    if (k == 0) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_mid_timeline[item * 8 + 7] = ((unsigned long long)n_kv << 16) | sm; }
  }
}
#define MID_STAMP(...) mid_stamp(__VA_ARGS__)
#else
#define MID_STAMP(...)
#endif

struct Barriers {
  uint64_t k_full[kMaxTiles];                    // K tile t landed (k_full[0] also covers Q)
  uint64_t v_full;                               // all V tiles landed
  uint64_t s_full[kMaxTiles];                    // S tile t complete
  uint64_t p_full;                               // P complete (one arrival per softmax warp)
  uint64_t o_full;                               // O complete
  uint64_t epi_done;                             // every softmax warp has read O (the next item's QK^T may overwrite the columns)
};

template <bool kBF16, bool kGeneric, int kParts>
__global__ void __launch_bounds__(Shape<kParts>::kThreads, 1)
attn_fwd_mid_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                    const __grid_constant__ CUtensorMap tmv, const __grid_constant__ AttnParams p) {
  constexpr int CW = Shape<kParts>::CW, kSoftmaxWarps = Shape<kParts>::kSoftmaxWarps;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Barriers bars;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float xch_m[kParts][BM], xch_l[kParts][BM];

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  // ---- work items ------------------------------------------------------------------------------------------------------------------
  // The launch has at most one CTA per SM (512 TMEM columns, 148 KB of shared memory per CTA); with more (Q tile, head, batch) items
  // than SMs a CTA takes several, one after the other.  (As one CTA per item, the 44 surplus CTAs of the config-4 shape -- 192
  // items on 148 SMs -- entered 1.1-1.6 us after an SM had drained and ran on their own for 4.6 us behind the 11 us of the
  // heaviest CTAs: 2.7 of the launch's 13.8 us, profiles/r02_timeline_mid.txt.)  Items are numbered heaviest first (causal:
  // causal_item) and dealt out in a snake -- round r forwards, round r + 1 backwards -- so the CTA with the lightest item of one
  // round gets the heaviest of the next.  The next item's Q / K are loaded as soon as the current item's QK^T MMAs have read
  // theirs, its V once the PV MMAs have.
  const int n_items = p.n_qt * p.H * p.B, G = (int)gridDim.x;
  struct Item { int item, q0, h, b, n_kv; };
  auto item_of = [&](int it) {                    // it-th item of this CTA, or n_items
    const int base = it * G;
    if (base >= n_items) return n_items;
    const int idx = base + ((it & 1) ? G - 1 - (int)blockIdx.x : (int)blockIdx.x);
    return idx < n_items ? idx : n_items;
  };
  auto decode = [&](int item) {
    Item w;
    w.item = item;
    int qt = item % p.n_qt, hb = item / p.n_qt;
    if (p.causal) causal_item(item, p.n_qt, p.H * p.B, p.hb_chunk, qt, hb);   // heaviest Q tiles first (attn_common.cuh)
    w.q0 = qt * BM;
    w.h = hb % p.H;
    w.b = hb / p.H;
    int n_kv = (p.Tk + BN - 1) / BN;              // K / V tiles this Q tile needs
    if (p.causal) {
      const int last_row = min(w.q0 + BM, p.Tq) - 1;
      const int lim = (last_row + (p.Tk - p.Tq)) / BN + 1;
      n_kv = lim < n_kv ? lim : n_kv;
    }
    w.n_kv = n_kv < 1 ? 1 : n_kv;                 // (Tk < Tq, causal: a tile whose rows see no key still walks tile 0, all hidden)
    return w;
  };

  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem0, sK = smem0 + kTileBytes, sV = sK + kMaxTiles * kTileBytes;

  // TMA loads of an item (one elected lane of the producer warp): Q + K tile 0 | K tiles 1.. | V, on their own mbarriers
  auto load_qk = [&](const Item& w) {
    const uint32_t b0 = ptx::smem_u32(&bars.k_full[0]);
    ptx::mbar_expect_tx(b0, 2 * kTileBytes);
    tma_tile(sQ, &tmq, b0, p.cq, w.q0, w.h, w.b);
    tma_tile(sK, &tmk, b0, p.ck, 0, w.h, w.b);
    for (int t = 1; t < w.n_kv; ++t) {
      const uint32_t bt = ptx::smem_u32(&bars.k_full[t]);
      ptx::mbar_expect_tx(bt, kTileBytes);
      tma_tile(sK + t * kTileBytes, &tmk, bt, p.ck, t * BN, w.h, w.b);
    }
  };
  auto load_v = [&](const Item& w) {
    const uint32_t bv = ptx::smem_u32(&bars.v_full);
    ptx::mbar_expect_tx(bv, w.n_kv * kTileBytes);
    for (int t = 0; t < w.n_kv; ++t) tma_tile(sV + t * kTileBytes, &tmv, bv, p.cv, t * BN, w.h, w.b);
  };

  const Item first = decode(item_of(0));          // (the grid never exceeds the number of items)
  MID_STAMP(0, first.item, first.n_kv);
  if (warp == kSoftmaxWarps && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
#pragma unroll
    for (int t = 0; t < kMaxTiles; ++t) {
      ptx::mbar_init(ptx::smem_u32(&bars.k_full[t]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.s_full[t]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&bars.v_full), 1);
    ptx::mbar_init(ptx::smem_u32(&bars.p_full), kSoftmaxWarps);
    ptx::mbar_init(ptx::smem_u32(&bars.o_full), 1);
    ptx::mbar_init(ptx::smem_u32(&bars.epi_done), kSoftmaxWarps);
    ptx::fence_mbar_init();
    // loads first: their latency covers the TMEM allocation and the block barrier below
    load_qk(first);
    load_v(first);
  }
  __syncwarp();
  if (warp == 0) ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  MID_STAMP(1, first.item);

  // Barrier phases across items: v_full, p_full, o_full and epi_done complete once per item (parity = item count & 1); k_full[t] /
  // s_full[t] only in items with more than t key tiles -- bit t of kph is their parity, flipped after every item that used them.
  if (warp == kSoftmaxWarps) {
    // ================================ producer / MMA issuer ===========================================================
    constexpr uint32_t kIdescQK = ptx::make_idesc_f16(kBF16, BM, BN, false, false);
    constexpr uint32_t kIdescPV = ptx::make_idesc_f16(kBF16, BM, HD, false, true);
    uint32_t kph = 0;
    Item w = first;
    for (int it = 0; w.item < n_items; ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const int n_kv = w.n_kv;
      // the previous item's epilogue has read O (columns 448..511, inside S tile 3), and its PV MMAs -- which read P from the S
      // columns -- completed before that (o_full, waited for at the end of the previous iteration)
      if (it > 0) ptx::mbar_wait(ptx::smem_u32(&bars.epi_done), ph ^ 1u);
      for (int t = 0; t < n_kv; ++t) {                         // S tile t = Q . K(t)^T  -> columns [128 t, 128 t + 128)
        ptx::mbar_wait(ptx::smem_u32(&bars.k_full[t]), kph >> t & 1u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ptx::mma_ss(tmem + t * BN, ptx::make_smem_desc_sw128(sQ + k * 32, 16, 1024),
                        ptx::make_smem_desc_sw128(sK + t * kTileBytes + k * 32, 16, 1024), kIdescQK, k > 0);
          ptx::mma_commit(ptx::smem_u32(&bars.s_full[t]));
        }
        __syncwarp();
      }
      const Item nx = decode(item_of(it + 1));
      const uint32_t kph_next = kph ^ ((1u << n_kv) - 1u);
      if (nx.item < n_items) {
        // Q and the K tiles are free once the last QK^T of this item has completed (commits complete in issue order): the next
        // item's Q / K loads run under this item's softmax passes
        ptx::mbar_wait(ptx::smem_u32(&bars.s_full[n_kv - 1]), kph >> (n_kv - 1) & 1u);
        if (ptx::elect_one()) load_qk(nx);
        __syncwarp();
      }
      ptx::mbar_wait(ptx::smem_u32(&bars.p_full), ph);         // every softmax warp has written its share of P
      ptx::mbar_wait(ptx::smem_u32(&bars.v_full), ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        for (int k = 0; k < n_kv * (BN / 16); ++k)             // O += P[:, 16k..16k+16) . V[16k..16k+16, :]
          ptx::mma_ts(tmem + kColO, tmem + k * 8, ptx::make_smem_desc_sw128(sV + k * 2048, 1024, 1024), kIdescPV, k > 0);
        ptx::mma_commit(ptx::smem_u32(&bars.o_full));
      }
      __syncwarp();
      if (nx.item < n_items) {                                 // the V tiles are free once the PV MMAs have completed
        ptx::mbar_wait(ptx::smem_u32(&bars.o_full), ph);
        if (ptx::elect_one()) load_v(nx);
        __syncwarp();
      }
      kph = kph_next;
      w = nx;
    }
  } else {
    // ================================ softmax: two threads per row ====================================================
    const int half = warp >> 2;                                 // which of the kParts chunks of a tile this thread takes
    const int r = (warp & 3) * 32 + (tid & 31);                 // row within the tile == TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int pair_bar = 2 + (warp & 3);                        // named barrier of the kParts warps of a lane quarter

    const bool softmax1 = (p.kind == OEH_SOFTMAX_1 || p.kind == OEH_SOFTMAX_CLIPPED1);
    const bool clipped = p.kind >= OEH_SOFTMAX_CLIPPED;
    const bool has_qs = kGeneric && p.qdq_scores != nullptr;
    const bool has_qp = kGeneric && p.qdq_probs != nullptr;
    const bool has_mask = kGeneric && p.mask != nullptr;
    const bool normalized = clipped || has_qp;
    Qdq qd_s{1.f, 1.f, 0.f, p.int_max}, qd_p{1.f, 1.f, 0.f, p.int_max};
    if (has_qs) qd_s = load_qdq(p.qdq_scores, p.int_max);
    if (has_qp) qd_p = load_qdq(p.qdq_probs, p.int_max);
    uint32_t kph = 0;
    Item w = first;
    for (int it = 0; w.item < n_items; ++it) {
    const uint32_t ph = (uint32_t)it & 1u;
    const int q0 = w.q0, h = w.h, b = w.b, n_kv = w.n_kv;
    const int row = q0 + r;
    if (it > 0) MID_STAMP(0, w.item, n_kv);
    const int64_t mask_row =
        p.mask ? ((int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)(row < p.Tq ? row : 0) * p.ms2) : 0;
    const int causal_lim = p.causal ? row + (p.Tk - p.Tq) : 0x7fffffff;

    // ---- pass A: e = exp2(s - m_i) over S, running (m, l) ------------------------------------------------------------
    float m_run = -INFINITY, l_run = 0.f;
    // frame (running max) of the e values of chunk kParts * i + half.  Four scalars behind selects, not an indexed array: the tile
    // loops below are NOT unrolled (OEH_MID_UNROLL_TILES restores the old form for A/B runs) -- fully unrolled the tapped
    // instantiation was 7544 SASS instructions (118 KB) and a fifth of its stall samples under ncu were instruction fetch
    // (stall_no_inst 19.8 %, profiles/r02_ncu_hot_kernels.txt); rolled up it is 3048 and times the same back to back
    float mc0 = 0.f, mc1 = 0.f, mc2 = 0.f, mc3 = 0.f;
    uint32_t hidden = 0;                                        // bit i: chunk i of this warp is invisible to all of its rows
    static_assert(kMaxTiles == 4, "mc0..mc3");
#ifdef OEH_MID_UNROLL_TILES
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int i = 0; i < kMaxTiles; ++i) {
      if (i < n_kv) {
        const int j = kParts * i + half, k0 = j * CW;
        ptx::mbar_wait(ptx::smem_u32(&bars.s_full[i]), kph >> i & 1u);
        ptx::tc_fence_after();
        if (i == 0) MID_STAMP(2, w.item);
        {
          // A chunk none of this warp's 32 rows can see (above the causal diagonal: 6 of the 16 (lane quarter, chunk) pairs of the
          // diagonal tile with four parts; or past Tk in the ragged last tile) contributes nothing to (m, l) and its P is 0:
          // no load, no exponentials -- pass B writes the zeros.
          const int valid_ = min(p.Tk - k0, CW), vis_ = min(valid_ - 1, causal_lim - k0);
          if (!(kGeneric && p.dbg_s) && __all_sync(0xffffffffu, vis_ < 0)) {
            hidden |= 1u << i;
            continue;
          }
        }
        float s[CW];
#pragma unroll
        for (int c = 0; c < CW / 32; ++c) ptx::tmem_ld32(t_lane + j * CW + 32 * c, reinterpret_cast<uint32_t*>(s) + 32 * c);
        ptx::tmem_wait_ld();
        if (kGeneric && p.dbg_s && j < kParts && row < p.Tq) {  // raw S of keys [0, 128)
          float* d = p.dbg_s + (((int64_t)b * p.H + h) * p.Tq + row) * BN + k0;
#pragma unroll
          for (int c = 0; c < CW; ++c) d[c] = s[c];
        }
        const int valid = min(p.Tk - k0, CW);                   // may be <= 0 in the ragged last tile: vis < 0 hides all
        const int vis = min(valid - 1, causal_lim - k0);
        const float mul = transform_scores<CW, kGeneric>(s, p, qd_s, has_qs, has_mask, mask_row, k0, max(valid, 0), vis);
        const float m_new = fmaxf(m_run, row_max<CW>(s) * mul);
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float alpha = ptx::ex2(m_run - m_safe);           // m_run = -inf -> 0
        l_run = fmaf(l_run, alpha, exp2_and_sum<CW, !kGeneric>(s, mul, m_safe));
        m_run = m_new;
        mc0 = i == 0 ? m_safe : mc0;
        mc1 = i == 1 ? m_safe : mc1;
        mc2 = i == 2 ? m_safe : mc2;
        mc3 = i == 3 ? m_safe : mc3;
#pragma unroll
        for (int c = 0; c < CW / 32; ++c) ptx::tmem_st32(t_lane + j * CW + 32 * c, reinterpret_cast<const uint32_t*>(s) + 32 * c);
      }
    }
    ptx::tmem_wait_st();
    MID_STAMP(3, w.item);

    // ---- the kParts threads of a row agree on (m, denominator) -----------------------------------------------------------
    xch_m[half][r] = m_run;
    xch_l[half][r] = l_run;
    // (only the kParts warps of this lane quarter exchange rows with each other: a 32 * kParts-thread barrier, not the CTA's)
    asm volatile("bar.sync %0, %1;" ::"r"(pair_bar), "n"(32 * kParts) : "memory");
    float m_all = m_run;
#pragma unroll
    for (int q2 = 0; q2 < kParts; ++q2) m_all = fmaxf(m_all, xch_m[q2][r]);
    const float m_fin = m_all == -INFINITY ? 0.f : m_all;
    float den = 0.f;
#pragma unroll
    for (int q2 = 0; q2 < kParts; ++q2) den = fmaf(xch_l[q2][r], ptx::ex2(xch_m[q2][r] - m_fin), den);
    if (softmax1) den += ptx::ex2(-m_fin);
    const float inv_den = __frcp_rn(den);
    const float omul = normalized ? 1.f : inv_den;              // one-pass kinds: P stays in the row-max frame (max P = 1)

    // ---- pass B: P ----------------------------------------------------------------------------------------------------------
#ifdef OEH_MID_UNROLL_TILES
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int i = 0; i < kMaxTiles; ++i) {
      if (i < n_kv) {
        const int j = kParts * i + half;
        const bool hid = (hidden >> i & 1u) != 0;               // (warp-uniform) P = 0, nothing to read
        float s[CW];
        if (!hid) {
#pragma unroll
          for (int c = 0; c < CW / 32; ++c) ptx::tmem_ld32(t_lane + j * CW + 32 * c, reinterpret_cast<uint32_t*>(s) + 32 * c);
          ptx::tmem_wait_ld();
        }
        // The P chunks of iteration i (kParts * CW / 2 = 64 packed columns) overwrite e-chunks read in iteration <= i (in this very
        // iteration when i == 0): every thread of the row must have COMPLETED its reads of this iteration before any writes
        // (ONE barrier instruction for both kinds of warp: compute-sanitizer's synccheck wants a barrier's arrivals at one PC)
        asm volatile("bar.sync %0, %1;" ::"r"(pair_bar), "n"(32 * kParts) : "memory");
        if (hid) {
          uint32_t z[CW / 2];
#pragma unroll
          for (int c = 0; c < CW / 2; ++c) z[c] = 0u;
          if constexpr (CW == 64) ptx::tmem_st32(t_lane + j * (CW / 2), z);
          else ptx::tmem_st16(t_lane + j * (CW / 2), z);
          continue;
        }
        const float mci = i == 0 ? mc0 : i == 1 ? mc1 : i == 2 ? mc2 : mc3;
        const float f = ptx::ex2(mci - m_fin);                  // chunk frame -> common frame (mc <= m_fin)
        if (normalized) {
          normalize_probs<CW, kGeneric, kBF16>(s, p, inv_den * f, clipped, has_qp, qd_p);
        } else {
#pragma unroll
          for (int c = 0; c < CW; ++c) s[c] *= f;
        }
        uint32_t pk[CW / 2];
#pragma unroll
        for (int c = 0; c < CW / 2; ++c) pk[c] = pack2<kBF16>(s[2 * c], s[2 * c + 1]);
        if constexpr (CW == 64) ptx::tmem_st32(t_lane + j * (CW / 2), pk);
        else ptx::tmem_st16(t_lane + j * (CW / 2), pk);
      }
    }
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncwarp();
    if (ptx::elect_one()) ptx::mbar_arrive(ptx::smem_u32(&bars.p_full));
    MID_STAMP(4, w.item);

    // ---- epilogue: each of the row's threads finishes 64 / kParts head-dim columns ---------------------------------------------
    const EpiPre pre = epi_prefetch<kBF16, kGeneric>(p, b, h, row);   // global reads of the epilogue, under the PV MMAs
    const Item nx = decode(item_of(it + 1));                    // (integer divisions: also under the PV MMAs)
    ptx::mbar_wait(ptx::smem_u32(&bars.o_full), ph);
    ptx::tc_fence_after();
    MID_STAMP(5, w.item);
    constexpr int ND = HD / kParts;
    float o[ND];
    if constexpr (ND == 32) ptx::tmem_ld32(t_lane + kColO + half * ND, reinterpret_cast<uint32_t*>(o));
    else ptx::tmem_ld16(t_lane + kColO + half * ND, reinterpret_cast<uint32_t*>(o));
    ptx::tmem_wait_ld();
    if (nx.item < n_items) {                                    // O is in registers: the next item's QK^T may write these columns
      ptx::tc_fence_before();
      __syncwarp();
      if (ptx::elect_one()) ptx::mbar_arrive(ptx::smem_u32(&bars.epi_done));
    }
    if (row < p.Tq) store_row<kBF16, kGeneric, ND>(o, p, omul, b, h, row, half * ND, pre);
    MID_STAMP(6, w.item);
    kph ^= (1u << n_kv) - 1u;
    w = nx;
    }   // items
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace attn_mid
}  // namespace oeh
