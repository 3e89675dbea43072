// Warp-specialised fused attention forward for sm_100a (the long-sequence kernel, Tq > 128).
//
// One CTA per SM works on TWO 128-row Q tiles of one (batch, head) and walks the keys in 64-wide sub-tiles.  Each
// S sub-tile [128 x 64] is consumed by TWO independent softmax streams, one per 32-column half, so that FOUR softmax
// warps share every SM sub-partition:
//
//   warps 0-15  softmax: stream s = warp >> 2 = 2*tile + half;  thread == query row == TMEM lane, 32 keys per step
//   warp  16    MMA issuer of Q tile 0 (one elected lane issues its tcgen05.mma / tcgen05.commit)
//   warp  17    TMA producer (one elected lane; Q once, K / V sub-tiles through kStages-deep rings)
//   warp  18    MMA issuer of Q tile 1
//   warp  19    idle (warps 16-19 give their registers to the softmax warps with setmaxnreg)
//
//   TMEM (512 columns):  S_t[b] at (2t+b)*64 (b = step parity; half h owns columns 32h..32h+31, its packed P the first
//                        16 of them)          O_{t,h} at 256 + (2t+h)*64
//
// Why four streams: at d = 64 this kernel is bound by the exp2 (MUFU/XU) pipe -- 16 results/clk/SM, against 512
// tensor cycles per 128x128 tile -- and ONE warp can drive its sub-partition's XU to at most 78 % (microbenchmark
// tools/micro/mufu_bench.cu), only while it is in its exp phase.  With two softmax warps per sub-partition (the
// FlashAttention-4 layout, versions 1-3 of this kernel) the measured XU utilisation stayed at 50 % (exactly 8
// exp/clk/SM) however S was buffered: the two warps run in lockstep through exp / non-exp phases (cycle trace in
// profiles/README.md).  Splitting each row in two column halves doubles the warps without extra TMEM: every half
// keeps its OWN running maximum m_h, row sum l_h and accumulator O_{t,h} += P_{t,h} . V[32h..32h+31] (a split-K of
// the PV product: same MMA work, K = 32 per half), so the halves never synchronise inside the key loop; they are
// merged once in the epilogue:  ctx = (O_0 2^(m_0-m) + O_1 2^(m_1-m)) / (l_0 2^(m_0-m) + l_1 2^(m_1-m) [+ 2^-m]).
//
//   mbarriers:  q_full | k_full/k_empty[kStages] | v_full/v_empty[kStages]   (TMA <-> MMA; *_empty take one
//                                                                              tcgen05.commit from EACH issuer)
//               s_full[t][b]  (MMA -> softmax: S_t of a step with parity b is complete)
//               p_full[t][b]  (softmax -> MMA, 8 arrivals = one elected lane per warp of tile t: P_{t,0}, P_{t,1} are
//                              in TMEM / S_t has been consumed)
//               o_done[t][b]  (MMA -> softmax: PV_t(v), v % 2 == b, has completed; waited only before an O rescale;
//                              one barrier per parity keeps the waiter at most one phase behind)
//               o_final[t]    (MMA -> softmax: the last PV_t has completed -> epilogue)
//   MMA issue order of issuer t per step i:  [p_full_t(i)] PV_{t,0}(i) PV_{t,1}(i) QK_t(i+2)
//   (tcgen05.mma executes in issue order, so QK_t(i+2) may overwrite the columns PV_t(i) just read P from).
//
// The running maximum is only moved when a sub-tile's maximum exceeds it by more than 2^8 ("lazy rescale"): P stays
// <= 256 (exact in fp16/bf16 range, fp32 row sums), the result is identical after the final normalisation, and the O
// round trip through TMEM almost never happens after the first sub-tiles.  softmax_1 adds exp2(-m) to the final
// denominator in the same frame (vutils/softmax_1.py:18-20).  Clipped / probs-quantised modes run a stats pass
// (QK^T only, exact running (m_h, l_h)), merge the halves' statistics through shared memory, and then a PV pass
// that recomputes QK^T (two-pass; see attn_fwd_sm100.cu).
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

constexpr int BNW = 64;                                            // keys per sub-tile (one MMA step)
constexpr int BNH = 32;                                            // keys per softmax stream and step
constexpr int kSubBytes = BNW * HD * 2;                            // 8 KB
constexpr int kStages = 8;
constexpr int kThreads = 640;                                      // 20 warps
constexpr int kQRows = 256;                                        // query rows per CTA (two tiles)
constexpr int kSmemBytes = 2 * kTileBytes + 2 * kStages * kSubBytes + 1024;   // Q0 Q1 | K ring | V ring
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColO0 = 256;                                   // S_t[b] at (2t+b)*64, O_{t,h} at 256 + (2t+h)*64
constexpr float kLazyRescale = 8.f;                                // log2 units
constexpr int kSoftmaxRegs = 112, kAuxRegs = 24;                   // 16*32*112 + 4*32*24 <= 640 * 96

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2][2], p_full[2][2], o_done[2][2];
  uint64_t o_final[2];
};

template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

// kNorm: P must be normalised before PV (clipped variants / probs fake-quant) -> two-pass; compiled separately so the
// one-pass hot loop carries none of the two-pass branches.
template <bool kBF16, bool kGeneric, bool kNorm>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Barriers bars;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float xch_m[2][2][128], xch_l[2][2][128];            // (tile, half, row): statistics merged across halves

  const int tid = threadIdx.x, warp = tid >> 5;
  // causal: heaviest Q tiles of a head first, so the last CTAs of the grid are the short ones
  const int qt = p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int q0 = qt * kQRows, h = blockIdx.y, b = blockIdx.z;

  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem0;                                    // 2 tiles of 128 rows
  const uint32_t sK = smem0 + 2 * kTileBytes;                   // kStages sub-tiles of 64 keys
  const uint32_t sV = sK + kStages * kSubBytes;                 // kStages sub-tiles

  if (tid == 0) {
    ptx::mbar_init(ptx::smem_u32(&bars.q_full), 1);
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars.k_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.k_empty[i]), 2);       // one tcgen05.commit from each MMA issuer
      ptx::mbar_init(ptx::smem_u32(&bars.v_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.v_empty[i]), 2);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int sb = 0; sb < 2; ++sb) {
        ptx::mbar_init(ptx::smem_u32(&bars.s_full[t][sb]), 1);
        ptx::mbar_init(ptx::smem_u32(&bars.p_full[t][sb]), 8);  // one elected lane of each of the tile's 8 softmax warps
        ptx::mbar_init(ptx::smem_u32(&bars.o_done[t][sb]), 1);
      }
      ptx::mbar_init(ptx::smem_u32(&bars.o_final[t]), 1);
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 16) ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

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

  if (warp >= 16) reg_dealloc<kAuxRegs>();
  if (warp == 17) {
    // ================================ TMA producer (whole warp waits, one elected lane issues) ============
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmq);
      ptx::prefetch_tmap(&tmk);
      ptx::prefetch_tmap(&tmv);
      const uint32_t bq = ptx::smem_u32(&bars.q_full);
      ptx::mbar_expect_tx(bq, 2 * kTileBytes);
      tma_tile(sQ, &tmq, bq, p.cq, q0, h, b);
      tma_tile(sQ + kTileBytes, &tmq, bq, p.cq, q0 + 128, h, b);
    }
    __syncwarp();
    auto load_k = [&](int i) {                                  // K sub-tile of step i -> ring slot i % kStages
      const int s = i % kStages, n = i / kStages;
      if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.k_empty[s]), (uint32_t)(n - 1) & 1u);
      if (ptx::elect_one()) {
        const uint32_t bf = ptx::smem_u32(&bars.k_full[s]);
        ptx::mbar_expect_tx(bf, kSubBytes);
        tma_tile(sK + s * kSubBytes, &tmk, bf, p.ck, (i < nA ? i : i - nA) * BNW, h, b);
      }
      __syncwarp();
    };
    load_k(0);
    if (n_iter > 1) load_k(1);
    for (int i = 0; i < n_iter; ++i) {                          // K runs two steps ahead of V, like the MMA issuers
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
  } else if (warp == 16 || warp == 18) {
    // ================================ MMA issuers: warp 16 serves Q tile 0, warp 18 Q tile 1 ===============
    // (whole warp waits, one elected lane issues)
    const int t = (warp - 16) >> 1;
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
    auto issue_pv = [&](int sb, int stage, bool first) {         // O_{t,hf} (+)= P_{t,hf}[sb] . V(stage)[32hf..]  (K = 32)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int k = 0; k < BNH / 16; ++k)
          ptx::mma_ts(tO0 + hf * 64, tS0 + sb * 64 + hf * 32 + k * 8, dv + stage * kSubUnits + (hf * 2 + k) * 128, kIdescPV,
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
          issue_pv(sb, vs, v == 0);
          ptx::mma_commit(ptx::smem_u32(last ? &bars.o_final[t] : &bars.o_done[t][v & 1]));
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
  } else if (warp < 16) {
    // ================================ softmax streams ======================================================
    reg_alloc<kSoftmaxRegs>();
    const int t = warp >> 3, hf = (warp >> 2) & 1;                // Q tile, column half
    const int r = (warp & 3) * 32 + (tid & 31);                   // row within the tile == TMEM lane
    const int row = q0 + t * 128 + r;                             // this thread's query row
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tS0 = t_lane + (uint32_t)(2 * t) * 64u + hf * 32u;       // + sb*64: this stream's S / P columns
    const uint32_t tO = t_lane + kColO0 + (uint32_t)(2 * t + hf) * 64u;    // this stream's accumulator

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
    OEH_T_DECL;

    // merge (m_h, l_h) of the two halves of a row: afterwards both hold the row's m and l in the common frame
    auto merge_halves = [&](float& a_own, float& a_other, float& l_other) {
      xch_m[t][hf][r] = m_run;
      xch_l[t][hf][r] = l_run;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      const float m_o = xch_m[t][hf ^ 1][r];
      l_other = xch_l[t][hf ^ 1][r];
      const float m = fmaxf(m_run, m_o);
      const float m_safe = m == -INFINITY ? 0.f : m;
      a_own = ptx::ex2(m_run - m_safe);                           // -inf -> 0
      a_other = ptx::ex2(m_o - m_safe);
      m_run = m;
    };

#ifdef OEH_STAGGER   // experiment: start Q tile 1 out of phase with Q tile 0
    if (t == 1) {
      const long long c0 = clock64();
      while (clock64() - c0 < OEH_STAGGER) { }
    }
#endif
    for (int i = 0; i < n_iter; ++i) {
      const int sb = i & 1;
      const uint32_t tS = tS0 + sb * 64u;
      const uint32_t bar_p = ptx::smem_u32(&bars.p_full[t][sb]);
      const bool pv_iter = i >= nA;
      const int k0 = (pv_iter ? i - nA : i) * BNW + hf * BNH;
      if (normalized && i == nA) {                                // stats pass done: the row statistics become common
        float a_own, a_other, l_other;
        merge_halves(a_own, a_other, l_other);
        l_run = l_run * a_own + l_other * a_other;
      }
      OEH_T(0);
      ptx::mbar_wait(ptx::smem_u32(&bars.s_full[t][sb]), (uint32_t)(i >> 1) & 1u);
      ptx::tc_fence_after();
      OEH_T(1);
      float s[BNH];
      ptx::tmem_ld32(tS, reinterpret_cast<uint32_t*>(s));
      ptx::tmem_wait_ld();
      OEH_T(2);
      if (!pv_iter) {                                             // stats pass: S is consumed, QK^T(i+2) may start
        ptx::tc_fence_before();
        __syncwarp();
        if (ptx::elect_one()) ptx::mbar_arrive(bar_p);
      }
      if (kGeneric && p.dbg_s && i < 2 && k0 == i * BNW + hf * BNH && row < p.Tq) {   // raw S of keys [0,128)
        float* d = p.dbg_s + (((int64_t)b * p.H + h) * p.Tq + row) * BN + k0;
#pragma unroll
        for (int j = 0; j < BNH; ++j) d[j] = s[j];
      }

      const int valid = min(p.Tk - k0, BNH);
      const int vis = min(valid - 1, causal_lim - k0);
      const float mul = transform_scores<BNH, kGeneric>(s, p, qd_s, has_qs, has_mask, mask_row, k0, valid, vis);
      const float tile_max = row_max<BNH>(s) * mul;
      OEH_T(3);

      if (!pv_iter) {
        // stats pass: exact running (m, l) of this half
        const float m_new = fmaxf(m_run, tile_max);
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float alpha = ptx::ex2(m_run - m_safe);
        l_run = fmaf(l_run, alpha, exp2_and_sum<BNH>(s, mul, m_safe));
        m_run = m_new;
        continue;
      }
      if (normalized) {
        // PV pass of the two-pass mode: (m, l) are final and common to both halves
        const float m_safe = m_run == -INFINITY ? 0.f : m_run;
        if (i == nA) {
          const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
          inv_den = __frcp_rn(den);
          skip_thr = zero_prob_threshold(p, clipped, has_qp, qd_p, den);
        }
#ifndef OEH_NO_SPARSE_SKIP
        // Clipped softmax with gamma < 0 (and / or the probs tap) maps every probability below a fixed bound to an
        // exact 0: at most 1/bound keys of a row survive, whatever the sequence length.  If none of this warp's
        // 32 rows x 32 keys can survive, P is 0 without a single exponential (the stats pass already paid for them).
        if (!__any_sync(0xffffffffu, tile_max - m_safe > skip_thr)) {
#pragma unroll
          for (int j = 0; j < BNH; ++j) s[j] = 0.f;
        } else
#endif
        {
          exp2_only<BNH>(s, mul, m_safe);
          normalize_probs<BNH, kGeneric>(s, p, inv_den, clipped, has_qp, qd_p);
        }
      } else {
        // online softmax with lazy rescale
        const bool move = tile_max > m_run + kLazyRescale || m_run == -INFINITY;
        const float m_new = move ? tile_max : m_run;              // tile_max may be -inf only if m_run is too
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float alpha = move ? ptx::ex2(m_run - m_safe) : 1.f;
        l_run = fmaf(l_run, alpha, exp2_and_sum<BNH>(s, mul, m_safe));
        m_run = m_new;
        if (i > nA && !__all_sync(0xffffffffu, alpha == 1.f)) {
          // O_{t,hf} *= alpha; needs PV_t(v-1) complete.  Its o_done barrier is at most one phase behind: PV_t(v-2)
          // is complete (S_t(i) was committed behind it) and PV_t(v) needs the P this thread has not written yet.
          const int v1 = i - nA - 1;
          ptx::mbar_wait(ptx::smem_u32(&bars.o_done[t][v1 & 1]), (uint32_t)(v1 >> 1) & 1u);
          ptx::tc_fence_after();
          uint32_t o[HD];
          ptx::tmem_ld32(tO, o);
          ptx::tmem_ld32(tO + 32, o + 32);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < HD; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
          ptx::tmem_st32(tO, o);
          ptx::tmem_st32(tO + 32, o + 32);
        }
      }
      OEH_T(4);

      // P_{t,hf} -> TMEM (packed 16-bit pairs) over the first 16 columns of the S half this thread has consumed
      {
        uint32_t pk[BNH / 2];
#pragma unroll
        for (int j = 0; j < BNH / 2; ++j) pk[j] = pack2<kBF16>(s[2 * j], s[2 * j + 1]);
        ptx::tmem_st16(tS, pk);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (ptx::elect_one()) ptx::mbar_arrive(bar_p);
      OEH_T(5);
    }
    OEH_T_PRINT("softmax: top s_full ld max exp pack+arrive", n_iter);

    // ---- epilogue: merge the halves, O / denominator -> gate / taps -> global ----------------------------------
    float a0 = 1.f, a1 = 1.f, omul = 1.f;                         // scales of O_{t,0}, O_{t,1}; 1/denominator
    if (!normalized) {
      float a_own, a_other, l_other;
      merge_halves(a_own, a_other, l_other);
      const float m_safe = m_run == -INFINITY ? 0.f : m_run;
      float den = l_run * a_own + l_other * a_other;
      if (softmax1) den += ptx::ex2(-m_safe);
      omul = __frcp_rn(den);
      a0 = hf ? a_other : a_own;
      a1 = hf ? a_own : a_other;
    }
    ptx::mbar_wait(ptx::smem_u32(&bars.o_final[t]), 0);
    ptx::tc_fence_after();
    // this thread finishes head-dim columns [32hf, 32hf+32) of its row from both accumulators
    float o[BNH];
    {
      uint32_t u0[BNH], u1[BNH];
      const uint32_t tOt = t_lane + kColO0 + (uint32_t)(2 * t) * 64u + hf * 32u;
      ptx::tmem_ld32(tOt, u0);
      ptx::tmem_ld32(tOt + 64, u1);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < BNH; ++j) o[j] = fmaf(__uint_as_float(u0[j]), a0, __uint_as_float(u1[j]) * a1);
    }
    if (row < p.Tq) store_row<kBF16, kGeneric, BNH>(o, p, omul, b, h, row, hf * BNH);
  }

  // ---- teardown ------------------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    __syncwarp();
    ptx::tmem_dealloc<kTmemCols>(tmem);
  }
}

}  // namespace attn_ws
}  // namespace oeh
