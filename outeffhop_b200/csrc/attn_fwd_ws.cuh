// Warp-specialised fused attention forward for sm_100a, "parity streams" (the long-sequence kernel, Tq > 128).
//
// One CTA per SM works on TWO 128-row Q tiles of one (batch, head) and walks the keys in 64-wide steps.  Each Q tile
// has two S buffers in TMEM (step parity) and each (tile, parity) pair is owned by ONE softmax stream of four warps:
//
//   warps 0-15  softmax: stream = warp >> 2 = 2*tile + parity; thread == query row == TMEM lane; the stream takes
//               every other step of its tile (64 keys) and keeps its OWN running max m_b, row sum l_b and accumulator
//               O_{t,b} += P(i) . V(i)  (i = b mod 2) -- a split of the key range by step parity, merged once in the
//               epilogue:  ctx = (O_0 2^(m_0-m) + O_1 2^(m_1-m)) / (l_0 2^(m_0-m) + l_1 2^(m_1-m) [+ 2^-m]).
//   warp  16    MMA issuer of Q tile 0 (one elected lane issues its tcgen05.mma / tcgen05.commit)
//   warp  17    TMA producer (one elected lane; Q once, K / V steps through kStages-deep rings)
//   warp  18    MMA issuer of Q tile 1
//   warp  19    idle (warps 16-19 give their registers to the softmax warps with setmaxnreg)
//
//   TMEM (512 columns):  S_t[b] at (2t+b)*64 (its packed P over the first 32 of them)     O_{t,b} at 256 + (2t+b)*64
//
// Why parity streams (an earlier version split every step of a tile in two column halves, one stream each): at
// d = 64 the loop is bound by the exp2 (MUFU/XU) pipe, and what keeps the XU below its rate is not issue slots but
// the fixed cost around every step (barrier wait, tcgen05.ld latency, wait::st, fence, arrive, issuer wake-up).
// A parity stream pays that once per 64 keys instead of once per 32, its two streams are half a period out of
// phase by construction (column halves run in lockstep: same barrier, same phase; 784 vs 898 TFLOP/s at T = 8192,
// profiles/r01_ws2_variants_f.txt), and because
// S_t(i) is committed behind PV_t(i-2) in the issuer's in-order MMA stream, owning S_t(i) implies O_{t,b} is
// quiescent: the O rescale needs no barrier of its own.
//
//   mbarriers:  q_full | k_full/k_empty[kStages] | v_full/v_empty[kStages]   (TMA <-> MMA; *_empty take one
//                                                                              tcgen05.commit from EACH issuer)
//               s_full[t][b]  (MMA -> softmax: S_t of a step with parity b is complete, and so is PV_t(i-2))
//               p_full[t][b]  (softmax -> MMA, 4 arrivals = one elected lane per warp of the stream: P is in TMEM /
//                              S_t has been consumed)
//               o_final[t]    (MMA -> softmax: the last PV_t has completed -> epilogue)
//   MMA issue order of issuer t per step i:  [p_full_t(i)] PV_t(i) QK_t(i+2)
//
// The running maximum is only moved when a step's maximum exceeds it by more than 2^8 ("lazy rescale"): P stays
// <= 256 (exact in fp16/bf16 range, fp32 row sums), the result is identical after the final normalisation, and the O
// round trip through TMEM almost never happens after the first steps.  softmax_1 adds exp2(-m) to the final
// denominator in the same frame (vutils/softmax_1.py:18-20).  Clipped / probs-quantised modes run a stats pass
// (QK^T only, exact running (m_b, l_b)), merge the streams' statistics through shared memory, and then a PV pass
// that recomputes QK^T (two-pass; see attn_fwd_sm100.cu).  In that PV pass a warp whose 32 rows x 64 keys all fall
// below the clip / quantisation threshold writes P = 0 without evaluating a single exponential ("sparse skip").
// Tried and dropped: staging P through shared memory so that S can be refilled before P is ready (SS-mode PV costs
// 48 instead of 32 clocks per MMA and the extra smem traffic made it 22 % slower, profiles/r01_ws3_variants_g.txt);
// one S buffer per tile released right after the tcgen05.ld, with P in its own TMEM region (S 128 + P 128 + O 256
// columns): the s_full wait halves but the step is no shorter, 951 vs 959 TFLOP/s (profiles/r01_ws4_variants_i.txt) --
// without any MUFU work both plans stop at ~1000-1030 TFLOP/s, i.e. the FMA / ALU / XU pipes of the four softmax warps
// per sub-partition overlap imperfectly, and more warps would need TMEM and registers this kernel does not have.
#pragma once

#include "attn_common.cuh"

#ifdef OEH_TRACE   // per-phase cycle trace of one lane per role in CTA (0,0,0) -> g_trace[warp][0..9], read back with oeh_trace_read()
__device__ long long g_trace[20][12];
#define OEH_T(k) do { if (trace) { long long c_ = clock64(); tacc[k] += c_ - tprev; tprev = c_; } } while (0)
#define OEH_T_DECL const bool trace = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid & 31) == 0; \
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64(), tstart = tprev
#define OEH_T_PRINT(role, n) do { if (trace) { g_trace[warp][0] = (n); g_trace[warp][1] = clock64() - tstart; \
  for (int k_ = 0; k_ < 8; ++k_) g_trace[warp][2 + k_] = tacc[k_]; } } while (0)
extern "C" int oeh_trace_read(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * 20 * 12);
}
#else
#define OEH_T(k) do { } while (0)
#define OEH_T_DECL do { } while (0)
#define OEH_T_PRINT(role, n) do { } while (0)
#endif


namespace oeh {
namespace attn_ws {

using namespace attn;

#ifdef OEH_TIMELINE
// experiment builds only (tools/timeline_ws.py): 8 stamps per CTA by thread 0 (softmax warp 0 = stream (tile 0, parity 0)):
// 0 entry, 1 TMEM allocated + block barrier, 2 first S of the stream ready, 3 its last P written, 4 statistics merged,
// 5 last PV complete, 6 stored; slot 7 = (n_kv << 16) | SM id
__device__ unsigned long long* g_ws_timeline = nullptr;
__device__ __forceinline__ void ws_stamp(int k, int n_kv = 0) {
  if (threadIdx.x == 0 && g_ws_timeline) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    g_ws_timeline[cta * 8 + k] = t;
    if (k == 0) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_ws_timeline[cta * 8 + 7] = ((unsigned long long)n_kv << 16) | sm; }
  }
}
#define WS_STAMP(...) ws_stamp(__VA_ARGS__)
#else
#define WS_STAMP(...)
#endif

constexpr int BNW = 64;                                            // keys per step
constexpr int kSubBytes = BNW * HD * 2;                            // 8 KB
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
  static constexpr int kStages = kTiles == 2 ? 8 : 5;
  static constexpr int kSmemBytes = kTiles * kTileBytes + 2 * kStages * kSubBytes + 1024;   // Q tiles | K ring | V ring
  static constexpr uint32_t kTmemCols = 256 * kTiles;
  static constexpr uint32_t kColO0 = 128 * kTiles;                   // S_t[b] at (2t+b)*64, O_{t,b} at kColO0 + (2t+b)*64
};
constexpr float kLazyRescale = 8.f;                                // log2 units

template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

template <int kStages>
struct Barriers {
  uint64_t q_full;
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2][2], p_full[2][2];
  uint64_t o_final[2];
};

template <bool kBF16, bool kGeneric, bool kNorm, int kTiles>
__global__ void __launch_bounds__(Cfg<kTiles>::kThreads, Cfg<kTiles>::kCtasPerSm)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                    const __grid_constant__ CUtensorMap tmv, const __grid_constant__ AttnParams p) {
  using C = Cfg<kTiles>;
  constexpr int kStages = C::kStages, kQRows = C::kQRows;
  constexpr uint32_t kTmemCols = C::kTmemCols, kColO0 = C::kColO0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Barriers<kStages> bars;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float xch_m[kTiles][2][128], xch_l[kTiles][2][128];  // (tile, parity, row): statistics merged across streams

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it
  // (role, TMEM lane base, barrier / descriptor addresses) on the uniform datapath instead of R2UR-ing it at every use
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // causal: heaviest Q tiles first over chunks of heads (attn_common.cuh, causal_item), so the last CTAs of the grid are
  // the short ones; the grid is walked in launch order (x fastest)
  int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (p.causal) {
    int hb;
    causal_item((int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)), (int)gridDim.x,
                (int)(gridDim.y * gridDim.z), p.hb_chunk, qt, hb);
    h = hb % (int)gridDim.y;
    b = hb / (int)gridDim.y;
  }
  const int q0 = qt * kQRows;

  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem0;                                    // kTiles tiles of 128 rows
  const uint32_t sK = smem0 + kTiles * kTileBytes;              // kStages steps of 64 keys
  const uint32_t sV = sK + kStages * kSubBytes;

  // ---- step plan (shared by all roles) ---------------------------------------------------------------------------
  int n_kv = (p.Tk + BNW - 1) / BNW;
  if (p.causal) {
    const int last_row = min(q0 + kQRows, p.Tq) - 1;            // last real query row of this CTA
    const int lim = (last_row + (p.Tk - p.Tq)) / BNW + 1;
    n_kv = lim < n_kv ? lim : n_kv;
  }
  const bool softmax1 = (p.kind == OEH_SOFTMAX_1 || p.kind == OEH_SOFTMAX_CLIPPED1);
  const bool clipped = p.kind >= OEH_SOFTMAX_CLIPPED;
  constexpr bool normalized = kNorm;                              // host: clipped || qdq_probs
  const int nA = normalized ? n_kv : 0;                          // stats-pass steps (QK^T only)
  const int n_iter = nA + n_kv;

  WS_STAMP(0, n_kv);
  if (tid == 0) {
    ptx::mbar_init(ptx::smem_u32(&bars.q_full), 1);
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
    }
    ptx::fence_mbar_init();
    // The first loads (Q, K of steps 0 and 1) go out HERE, before the TMEM allocation and the block barrier: their latency
    // (~1.2 us cold) covers both instead of following them (the allocations of two co-resident CTAs serialise: ~0.5 us each,
    // profiles/r02_micro_tmem_alloc.txt).  The producer warp takes over from step 2.
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
    const uint32_t bq = ptx::smem_u32(&bars.q_full);
    ptx::mbar_expect_tx(bq, kTiles * kTileBytes);
    tma_tile(sQ, &tmq, bq, p.cq, q0, h, b);
    if (kTiles == 2) tma_tile(sQ + kTileBytes, &tmq, bq, p.cq, q0 + 128, h, b);
    for (int i = 0; i < 2 && i < n_iter; ++i) {
      const uint32_t bf = ptx::smem_u32(&bars.k_full[i]);
      ptx::mbar_expect_tx(bf, kSubBytes);
      tma_tile(sK + i * kSubBytes, &tmk, bf, p.ck, (i < nA ? i : i - nA) * BNW, h, b);
    }
  }
  __syncwarp();
  if (warp == C::kWarpMma0) ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  WS_STAMP(1);

  if (warp >= C::kSoftmaxWarps) reg_dealloc<C::kAuxRegs>();
  if (warp == C::kWarpTma) {
    // ================================ TMA producer (whole warp waits, one elected lane issues) ============
    auto load_k = [&](int i) {                                  // K of step i -> ring slot i % kStages
      const int s = i % kStages, n = i / kStages;
      if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.k_empty[s]), (uint32_t)(n - 1) & 1u);
      if (ptx::elect_one()) {
        const uint32_t bf = ptx::smem_u32(&bars.k_full[s]);
        ptx::mbar_expect_tx(bf, kSubBytes);
        tma_tile(sK + s * kSubBytes, &tmk, bf, p.ck, (i < nA ? i : i - nA) * BNW, h, b);
      }
      __syncwarp();
    };
    for (int i = 0; i < n_iter; ++i) {                          // K runs two steps ahead of V, like the MMA issuers (steps 0, 1: above)
      if (i + 2 < n_iter) load_k(i + 2);
      if (i >= nA) {
        const int v = i - nA, s = v % kStages, n = v / kStages;
        if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.v_empty[s]), (uint32_t)(n - 1) & 1u);
        if (ptx::elect_one()) {
          const uint32_t bf = ptx::smem_u32(&bars.v_full[s]);
          ptx::mbar_expect_tx(bf, kSubBytes);
          tma_tile(sV + s * kSubBytes, &tmv, bf, p.cv, v * BNW, h, b);
        }
        __syncwarp();
      }
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
    auto issue_qk = [&](int sb, int stage) {                     // S_t[sb] = Q_t . K(stage)^T    (128 x 64, K = 64)
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        ptx::mma_ss(tS0 + sb * 64, dq + k * 2, dk + stage * kSubUnits + k * 2, kIdescQK, k > 0);
    };
    auto issue_pv = [&](int sb, int stage, bool first) {         // O_{t,sb} (+)= P_t[sb] . V(stage)  (K = 64 keys)
#pragma unroll
      for (int k = 0; k < BNW / 16; ++k)
        ptx::mma_ts(tO0 + sb * 64, tS0 + sb * 64 + k * 8, dv + stage * kSubUnits + k * 128, kIdescPV,
                    (first && k == 0) ? 0u : 1u);
    };
    ptx::mbar_wait(ptx::smem_u32(&bars.q_full), 0);
    OEH_T_DECL;
    for (int i = 0; i < 2 && i < n_iter; ++i) {                  // prologue: S_t(0), S_t(1)
      ptx::mbar_wait(ptx::smem_u32(&bars.k_full[i % kStages]), 0);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        issue_qk(i, i % kStages);
        ptx::mma_commit(ptx::smem_u32(&bars.s_full[t][i]));
        ptx::mma_commit(ptx::smem_u32(&bars.k_empty[i % kStages]));
      }
      __syncwarp();
    }
    for (int i = 0; i < n_iter; ++i) {
      const int sb = i & 1;
      const bool pv = i >= nA;
      const int v = i - nA, vs = pv ? v % kStages : 0;
      const bool more = i + 2 < n_iter;
      const bool last = i + 1 == n_iter;
      const int ks = (i + 2) % kStages;
      OEH_T(0);
      if (pv) ptx::mbar_wait(ptx::smem_u32(&bars.v_full[vs]), (uint32_t)(v / kStages) & 1u);
      if (more) ptx::mbar_wait(ptx::smem_u32(&bars.k_full[ks]), (uint32_t)((i + 2) / kStages) & 1u);
      OEH_T(1);
      ptx::mbar_wait(ptx::smem_u32(&bars.p_full[t][sb]), (uint32_t)(i >> 1) & 1u);
      ptx::tc_fence_after();
      OEH_T(2);
      if (ptx::elect_one()) {
        if (pv) {
          issue_pv(sb, vs, v < 2);                               // the first PV step of each parity starts its accumulator
          if (last) ptx::mma_commit(ptx::smem_u32(&bars.o_final[t]));
          ptx::mma_commit(ptx::smem_u32(&bars.v_empty[vs]));
        }
        if (more) {
          issue_qk(sb, ks);
          ptx::mma_commit(ptx::smem_u32(&bars.s_full[t][sb]));
          ptx::mma_commit(ptx::smem_u32(&bars.k_empty[ks]));
        }
      }
      __syncwarp();
      OEH_T(3);
    }
    OEH_T_PRINT("issuer: loop kv_full p_full issue", n_iter);
  } else if (warp < C::kSoftmaxWarps) {
    // ================================ softmax streams ======================================================
    reg_alloc<C::kSoftmaxRegs>();
    const int t = warp >> 3, par = (warp >> 2) & 1;               // Q tile, step parity
    const int r = (warp & 3) * 32 + (tid & 31);                   // row within the tile == TMEM lane
    const int row = q0 + t * 128 + r;                             // this thread's query row
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
    const int64_t mask_row =
        p.mask ? ((int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)(row < p.Tq ? row : 0) * p.ms2) : 0;
    const int causal_lim = p.causal ? row + (p.Tk - p.Tq) : 0x7fffffff;

    float m_run = -INFINITY, l_run = 0.f, inv_den = 1.f, skip_thr = -INFINITY;
    bool merged = false, started = false;                         // statistics merged / accumulator holds a PV result
    OEH_T_DECL;

    // merge (m_b, l_b) of the two streams of a row: afterwards both hold the row's m in the common frame
    auto merge_streams = [&](float& a_own, float& a_other, float& l_other) {
      xch_m[t][par][r] = m_run;
      xch_l[t][par][r] = l_run;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      const float m_o = xch_m[t][par ^ 1][r];
      l_other = xch_l[t][par ^ 1][r];
      const float m = fmaxf(m_run, m_o);
      const float m_safe = m == -INFINITY ? 0.f : m;
      a_own = ptx::ex2(m_run - m_safe);                           // -inf -> 0
      a_other = ptx::ex2(m_o - m_safe);
      m_run = m;
    };

    for (int i = par; i < n_iter; i += 2) {
      const bool pv_iter = i >= nA;
      const int k0 = (pv_iter ? i - nA : i) * BNW;
      if (normalized && pv_iter && !merged) {                     // stats pass done: the row statistics become common
        float a_own, a_other, l_other;
        merge_streams(a_own, a_other, l_other);
        l_run = l_run * a_own + l_other * a_other;
        merged = true;
      }
      OEH_T(0);
      ptx::mbar_wait(bar_s, (uint32_t)(i >> 1) & 1u);
      ptx::tc_fence_after();
      if (i == par) WS_STAMP(2);
      OEH_T(1);
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
      OEH_T(2);
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
      OEH_T(3);

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
      OEH_T(4);

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
      OEH_T(5);
    }
    OEH_T_PRINT("softmax: top s_full ld max exp pack+arrive", n_iter);

    // ---- epilogue: merge the streams, O / denominator -> gate / taps -> global -----------------------------------
    WS_STAMP(3);
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
    const bool has0 = n_kv > 1 || (nA & 1) == 0, has1 = n_kv > 1 || (nA & 1) == 1;
    const float a0 = par ? a_other : a_own, a1 = par ? a_own : a_other;
    const EpiPre pre = epi_prefetch<kBF16, kGeneric>(p, b, h, row);   // global reads of the epilogue, under the last PV
    WS_STAMP(4);
    ptx::mbar_wait(ptx::smem_u32(&bars.o_final[t]), 0);
    ptx::tc_fence_after();
    WS_STAMP(5);
    // this thread finishes head-dim columns [32par, 32par+32) of its row from both accumulators
    float o[32];
    {
      uint32_t u0[32], u1[32];
      const uint32_t tOt = t_lane + kColO0 + (uint32_t)(2 * t) * 64u + par * 32u;
      ptx::tmem_ld32(tOt, u0);
      ptx::tmem_ld32(tOt + 64, u1);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        o[j] = (has0 ? __uint_as_float(u0[j]) * a0 : 0.f) + (has1 ? __uint_as_float(u1[j]) * a1 : 0.f);
    }
    // Context rows leave as full 64-byte segments staged per warp (store_half_rows_coalesced) in the K ring: every K / V stage
    // has been consumed once the LAST PV of every tile of the CTA is complete (the two tiles share the rings).
    if (row < p.Tq) finish_row<kBF16, kGeneric, 32>(o, p, omul, b, h, row, par * 32, pre);
    if constexpr (kTiles == 2) {
      ptx::mbar_wait(ptx::smem_u32(&bars.o_final[t ^ 1]), 0);
    }
    static_assert(kStages * kSubBytes >= C::kSoftmaxWarps * 2048, "the K ring holds one staging slot per softmax warp");
    store_half_rows_coalesced<kBF16>(o, p, sK + (uint32_t)warp * 2048u, b, h, q0 + t * 128 + (warp & 3) * 32, par * 32);
    WS_STAMP(6);
  }

  // ---- teardown ------------------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == C::kWarpMma0) {
    __syncwarp();
    ptx::tmem_dealloc<kTmemCols>(tmem);
  }
}

}  // namespace attn_ws
}  // namespace oeh
