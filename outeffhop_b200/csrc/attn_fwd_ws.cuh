// Warp-specialised fused attention forward for sm_100a (the long-sequence kernel, Tq > 128).
//
// One CTA per SM works on TWO 128-row Q tiles of one (batch, head) and walks the keys in 64-wide sub-tiles.
//
//   warps 0-3   softmax warpgroup 0  (Q tile 0; thread == query row == TMEM lane)
//   warps 4-7   softmax warpgroup 1  (Q tile 1)
//   warp  8     MMA issuer of warpgroup 0 (one elected lane issues its tcgen05.mma / tcgen05.commit)
//   warp  9     TMA producer (one elected lane; Q once, K / V sub-tiles through kStages-deep rings)
//   warp  10    MMA issuer of warpgroup 1        (warp 11 idle; warps 8-11 give their registers to the softmax warps)
//
//   TMEM (512 columns):  S_w[b] / P_w[b] at (3w+b)*64, b = step % 3      O_w at 384 + 64w
//
// At d = 64 this kernel is bound by the exp2 (MUFU/XU) pipe, 16 results/clk/SM, not by the tensor pipe: the whole
// design keeps the XU busy.  History (profiles/README.md): v1 (one 128-wide S per warpgroup) and v2 (two 64-wide S
// buffers) both ran at exactly 8 exp/clk/SM -- an in-kernel cycle trace showed each warp spending 1300 cycles in a
// 512-cycle exp phase because the two warps that share an SM sub-partition (warp k of warpgroup 0 and warp k of
// warpgroup 1) ran in lockstep: both in the exp phase (sharing the XU), then both in the XU-idle phases (TMEM load,
// row max, P pack/store, barriers; ~740 cycles).  This version
//   * TRIPLE-buffers S per warpgroup: QK^T of step i+3 is issued into S_w[i%3] right behind PV_w(i), which consumed
//     P_w(i) from the same columns (tcgen05.mma executes in issue order).  S(i+1) is already in REGISTERS and S(i+2)
//     complete in TMEM while step i runs, so the row max of S(i+1) and the P pack of step i are plain ALU work the
//     compiler interleaves with the exp2 stream of step i (one basic block), and the TMEM load of S(i+2) is issued
//     as soon as the registers of step i are dead;
//   * passes an "XU token" between the two warps of a sub-partition with named barriers (bar.sync / bar.arrive,
//     FlashAttention-3 style ping-pong), so their exp phases alternate instead of colliding;
//   * gives each warpgroup its own MMA issuer warp, so the warpgroups are not coupled through a shared issue order.
//
//   mbarriers:  q_full | k_full/k_empty[kStages] | v_full/v_empty[kStages]   (TMA <-> MMA; *_empty take one
//                                                                              tcgen05.commit from EACH issuer)
//               s_full[w][b]  (MMA -> softmax w: S_w of a step with step % 3 == b is complete)
//               p_full[w][b]  (softmax w -> MMA, 128 arrivals: P_w is in TMEM / S_w has been consumed)
//               o_done[w][b]  (MMA -> softmax w: PV_w(v), v % 3 == b, has completed; waited only before an O rescale.
//                              One barrier per buffer index keeps the parity wait unambiguous: the waiter is never
//                              more than one phase behind.)
//               o_final[w]    (MMA -> softmax w: the last PV_w has completed -> epilogue)
//   MMA issue order of issuer w per step i:  [p_full_w(i)] PV_w(i) QK_w(i+3)
//
// The running maximum is only moved when a sub-tile's maximum exceeds it by more than 2^8 ("lazy rescale"): P stays
// <= 256 (exact in fp16/bf16 range, fp32 row sums), the result is identical after the final normalisation, and the O
// round trip through TMEM almost never happens after the first sub-tiles.  softmax_1 adds exp2(-m) to the final
// denominator in the same frame (vutils/softmax_1.py:18-20).  Clipped / probs-quantised modes run a stats pass
// (QK^T only, exact running (m,l)) followed by a PV pass that recomputes QK^T (two-pass; see attn_fwd_sm100.cu).
#pragma once

#include "attn_common.cuh"

#ifdef OEH_TRACE   // per-phase cycle trace of one lane per role in CTA (0,0,0); printed at the end of the kernel
#define OEH_T(k) do { if (trace) { long long c_ = clock64(); tacc[k] += c_ - tprev; tprev = c_; } } while (0)
#define OEH_T_DECL const bool trace = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid & 31) == 0 && (warp == 0 || warp == 4 || warp >= 8); \
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64(), tstart = tprev
#define OEH_T_PRINT(role, n) do { if (trace) printf("%s w%d steps %d total/step %lld | %lld %lld %lld %lld %lld %lld %lld %lld\n", role, warp, n, \
  (clock64() - tstart) / (n), tacc[0] / (n), tacc[1] / (n), tacc[2] / (n), tacc[3] / (n), tacc[4] / (n), tacc[5] / (n), tacc[6] / (n), tacc[7] / (n)); } while (0)
#else
#define OEH_T(k) do { } while (0)
#define OEH_T_DECL do { } while (0)
#define OEH_T_PRINT(role, n) do { } while (0)
#endif

namespace oeh {
namespace attn_ws {

using namespace attn;

constexpr int BNW = 64;                                            // keys per sub-tile (one softmax step)
constexpr int kSubBytes = BNW * HD * 2;                            // 8 KB
constexpr int kStages = 8;
constexpr int kSBufs = 3;                                          // S buffers per warpgroup
constexpr int kThreads = 384;                                      // 12 warps (warp 11 idle: setmaxnreg works on whole warpgroups)
constexpr int kQRows = 256;                                        // query rows per CTA (two tiles)
constexpr int kSmemBytes = 2 * kTileBytes + 2 * kStages * kSubBytes + 1024;   // Q0 Q1 | K ring | V ring
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColO0 = 2 * kSBufs * 64;                       // S_w[b] at (3w+b)*64, O_w at 384 + w*64
constexpr float kLazyRescale = 8.f;                                // log2 units

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2][kSBufs], p_full[2][kSBufs], o_done[2][kSBufs];
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

template <bool kBF16, bool kGeneric>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Barriers bars;
  __shared__ uint32_t tmem_base_slot;

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
    for (int w = 0; w < 2; ++w) {
#pragma unroll
      for (int sb = 0; sb < kSBufs; ++sb) {
        ptx::mbar_init(ptx::smem_u32(&bars.s_full[w][sb]), 1);
        ptx::mbar_init(ptx::smem_u32(&bars.p_full[w][sb]), 128);
        ptx::mbar_init(ptx::smem_u32(&bars.o_done[w][sb]), 1);
      }
      ptx::mbar_init(ptx::smem_u32(&bars.o_final[w]), 1);
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 8) ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(&tmem_base_slot));
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
  const bool normalized = clipped || (kGeneric && p.qdq_probs != nullptr);
  const int nA = normalized ? n_kv : 0;                          // stats-pass steps (QK^T only)
  const int n_iter = nA + n_kv;

  if (warp >= 8) reg_dealloc<40>();                              // 8 softmax warps x 232 + 4 x 40 registers <= 64 K
  if (warp == 9) {
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
    for (int i = 0; i < kSBufs && i < n_iter; ++i) load_k(i);
    for (int i = 0; i < n_iter; ++i) {                          // K runs kSBufs steps ahead of V, like the MMA issuers
      if (i + kSBufs < n_iter) load_k(i + kSBufs);
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
  } else if (warp == 8 || warp == 10) {
    // ================================ MMA issuers: warp 8 serves warpgroup 0, warp 10 warpgroup 1 ==========
    // (whole warp waits, one elected lane issues)
    const int w = (warp - 8) >> 1;
    constexpr uint32_t kIdescQK = ptx::make_idesc_f16(kBF16, 128, BNW, false, false);
    constexpr uint32_t kIdescPV = ptx::make_idesc_f16(kBF16, 128, HD, false, true);
    constexpr uint64_t kTileUnits = kTileBytes >> 4;             // descriptor address units (16 B)
    constexpr uint64_t kSubUnits = kSubBytes >> 4;
    const uint64_t dq = ptx::make_smem_desc_sw128(sQ, 16, 1024) + w * kTileUnits;
    const uint64_t dk = ptx::make_smem_desc_sw128(sK, 16, 1024);
    const uint64_t dv = ptx::make_smem_desc_sw128(sV, 1024, 1024);
    const uint32_t tS0 = tmem + (uint32_t)(kSBufs * w) * 64u, tO = tmem + kColO0 + w * 64;
    auto issue_qk = [&](int sb, int stage) {                     // S_w[sb] = Q_w . K(stage)^T    (128 x 64, K = 64)
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        ptx::mma_ss(tS0 + sb * 64, dq + k * 2, dk + stage * kSubUnits + k * 2, kIdescQK, k > 0);
    };
    auto issue_pv = [&](int sb, int stage, bool first) {         // O_w (+)= P_w[sb] . V(stage)  (128 x 64, K = 64)
#pragma unroll
      for (int k = 0; k < BNW / 16; ++k)
        ptx::mma_ts(tO, tS0 + sb * 64 + k * 8, dv + stage * kSubUnits + k * 128, kIdescPV, (first && k == 0) ? 0u : 1u);
    };
    ptx::mbar_wait(ptx::smem_u32(&bars.q_full), 0);
    OEH_T_DECL;
    for (int i = 0; i < kSBufs && i < n_iter; ++i) {             // prologue: S_w(0..2)
      ptx::mbar_wait(ptx::smem_u32(&bars.k_full[i % kStages]), 0);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        issue_qk(i, i % kStages);
        ptx::mma_commit(ptx::smem_u32(&bars.s_full[w][i]));
        ptx::mma_commit(ptx::smem_u32(&bars.k_empty[i % kStages]));
      }
      __syncwarp();
    }
    for (int i = 0; i < n_iter; ++i) {
      const int sb = i % kSBufs;
      const bool pv = i >= nA;
      const int v = i - nA, vs = pv ? v % kStages : 0;
      const bool more = i + kSBufs < n_iter;
      const bool last = i + 1 == n_iter;
      const int ks = (i + kSBufs) % kStages;
      OEH_T(0);
      if (pv) ptx::mbar_wait(ptx::smem_u32(&bars.v_full[vs]), (uint32_t)(v / kStages) & 1u);
      OEH_T(1);
      if (more) ptx::mbar_wait(ptx::smem_u32(&bars.k_full[ks]), (uint32_t)((i + kSBufs) / kStages) & 1u);
      OEH_T(2);
      ptx::mbar_wait(ptx::smem_u32(&bars.p_full[w][sb]), (uint32_t)(i / kSBufs) & 1u);
      ptx::tc_fence_after();
      OEH_T(3);
      if (ptx::elect_one()) {
        if (pv) {
          issue_pv(sb, vs, v == 0);
          ptx::mma_commit(ptx::smem_u32(last ? &bars.o_final[w] : &bars.o_done[w][v % kSBufs]));
          ptx::mma_commit(ptx::smem_u32(&bars.v_empty[vs]));
        }
        if (more) {
          issue_qk(sb, ks);
          ptx::mma_commit(ptx::smem_u32(&bars.s_full[w][sb]));
          ptx::mma_commit(ptx::smem_u32(&bars.k_empty[ks]));
        }
      }
      __syncwarp();
      OEH_T(4);
    }
    OEH_T_PRINT("issuer: loop v_full k_full p_full issue", n_iter);
  } else if (warp < 8) {
    // ================================ softmax warpgroups ==================================================
    reg_alloc<232>();
    const int w = warp >> 2;                                      // warpgroup == Q tile
    const int row = q0 + w * 128 + (tid & 127);                   // this thread's query row
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tS0 = t_lane + (uint32_t)(kSBufs * w) * 64u, tO = t_lane + kColO0 + w * 64;
    // XU token of this sub-partition: barrier id 1 + 2*(warp & 3) + w lets warp (w, warp & 3) start its exp phase
    const int bar_mine = 1 + 2 * (warp & 3) + w, bar_other = 1 + 2 * (warp & 3) + (1 - w);

    const bool has_qs = kGeneric && p.qdq_scores != nullptr;
    const bool has_qp = kGeneric && p.qdq_probs != nullptr;
    const bool has_mask = kGeneric && p.mask != nullptr;
    Qdq qd_s{1.f, 1.f, 0.f, p.int_max}, qd_p{1.f, 1.f, 0.f, p.int_max};
    if (has_qs) qd_s = load_qdq(p.qdq_scores, p.int_max);
    if (has_qp) qd_p = load_qdq(p.qdq_probs, p.int_max);
    const int64_t mask_row =
        p.mask ? ((int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)(row < p.Tq ? row : 0) * p.ms2) : 0;
    const int causal_lim = p.causal ? row + (p.Tk - p.Tq) : 0x7fffffff;

    float m_run = -INFINITY, l_run = 0.f, inv_den = 1.f, mul = 1.f;
    OEH_T_DECL;

    // wait for S_w(i) and start its TMEM -> register load (completed by the next tmem_wait_ld)
    auto load_s = [&](int i, float (&s)[BNW]) {
      ptx::mbar_wait(ptx::smem_u32(&bars.s_full[w][i % kSBufs]), (uint32_t)(i / kSBufs) & 1u);
      ptx::tc_fence_after();
      uint32_t* sr = reinterpret_cast<uint32_t*>(s);
      const uint32_t t = tS0 + (uint32_t)(i % kSBufs) * 64u;
#if defined(OEH_EXP) && (OEH_EXP == 2 || OEH_EXP == 5 || OEH_EXP == 6)
      (void)t; (void)sr;
#else
      ptx::tmem_ld32(t, sr);
      ptx::tmem_ld32(t + 32, sr + 32);
#endif
    };
    // score transform + row maximum of step i (log2 domain); pure ALU work on registers
    auto pre = [&](int i, float (&s)[BNW]) -> float {
      const int k0 = (i >= nA ? i - nA : i) * BNW;
      const int valid = min(p.Tk - k0, BNW);
      const int vis = min(valid - 1, causal_lim - k0);
      if (kGeneric && p.dbg_s && i < 2 && k0 == i * BNW && row < p.Tq) {   // raw S of keys [0,128)
        float* d = p.dbg_s + (((int64_t)b * p.H + h) * p.Tq + row) * BN + i * BNW;
#pragma unroll
        for (int j = 0; j < BNW; ++j) d[j] = s[j];
      }
      mul = transform_scores<BNW, kGeneric>(s, p, qd_s, has_qs, has_mask, mask_row, k0, valid, vis);
      return row_max<BNW>(s) * mul;
    };
    // in the stats pass S_w(i) is consumed as soon as it is in registers: QK^T(i+3) may overwrite its buffer
    auto release_s = [&](int i) {
      ptx::tc_fence_before();
      ptx::mbar_arrive(ptx::smem_u32(&bars.p_full[w][i % kSBufs]));
    };

    // one step: cur = S(i) (transformed, maximum tm_cur known), nxt = S(i+1) landing in registers
    auto step = [&](int i, float (&cur)[BNW], float tm_cur, float (&nxt)[BNW], float& tm_nxt) {
      const bool pv_iter = i >= nA;
      const bool has_next = i + 1 < n_iter;
      float m_safe, alpha = 1.f;
      if (!pv_iter) {                                             // stats pass: exact running maximum
        const float m_new = fmaxf(m_run, tm_cur);
        m_safe = m_new == -INFINITY ? 0.f : m_new;
        alpha = ptx::ex2(m_run - m_safe);
        m_run = m_new;
      } else if (normalized) {                                    // PV pass of the two-pass mode: (m, l) are final
        m_safe = m_run == -INFINITY ? 0.f : m_run;
        if (i == nA) {
          const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
          inv_den = __frcp_rn(den);
        }
      } else {                                                    // online softmax with lazy rescale
        const bool move = tm_cur > m_run + kLazyRescale || m_run == -INFINITY;
        const float m_new = move ? tm_cur : m_run;                // tm_cur may be -inf only if m_run is too
        m_safe = m_new == -INFINITY ? 0.f : m_new;
        alpha = move ? ptx::ex2(m_run - m_safe) : 1.f;
        m_run = m_new;
        if (i > nA && !__all_sync(0xffffffffu, alpha == 1.f)) {
          // O_w *= alpha; needs PV_w(v-1) complete (its o_done barrier is at most one phase behind: PV_w(v-4) is
          // complete, S_w(i) was committed behind PV_w(v-3))
          const int v1 = i - nA - 1;
          ptx::mbar_wait(ptx::smem_u32(&bars.o_done[w][v1 % kSBufs]), (uint32_t)(v1 / kSBufs) & 1u);
          ptx::tc_fence_after();
          uint32_t o[HD];
          ptx::tmem_ld32(tO, o);
          ptx::tmem_ld32(tO + 32, o + 32);
          ptx::tmem_wait_ld();                                    // (also completes the pending load of nxt)
#pragma unroll
          for (int j = 0; j < HD; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
          ptx::tmem_st32(tO, o);
          ptx::tmem_st32(tO + 32, o + 32);
        }
      }
      OEH_T(0);
      if (has_next) {
        ptx::tmem_wait_ld();                                      // S_w(i+1) is in nxt[]
        if (i + 1 < nA) release_s(i + 1);
      }
      OEH_T(1);

#if !(defined(OEH_EXP) && (OEH_EXP == 6 || OEH_EXP == 7))
      ptx::nbar_sync(bar_mine, 64, m_safe);                       // ---- XU token acquired ----
#endif
      OEH_T(2);
      float sum = 0.f;
#if defined(OEH_EXP) && (OEH_EXP == 5 || OEH_EXP == 6)
      sum = m_safe; tm_nxt = tm_cur;
#else
      if (pv_iter && normalized) {
#pragma unroll
        for (int j = 0; j < BNW; ++j) cur[j] = ptx::ex2(fmaf(cur[j], mul, -m_safe));
        sum = cur[BNW - 1];
      } else {
        sum = exp2_and_sum<BNW>(cur, mul, m_safe);
      }
#if !(defined(OEH_EXP) && (OEH_EXP == 6 || OEH_EXP == 7))
      ptx::nbar_arrive(bar_other, 64, sum);                       // ---- XU token passed on ----
#endif
      if (has_next) tm_nxt = pre(i + 1, nxt);                     // independent ALU work: interleaves with the exps
#endif
      if (!(pv_iter && normalized)) l_run = fmaf(l_run, alpha, sum);
      OEH_T(3);

      if (pv_iter) {
        if (normalized) normalize_probs<BNW, kGeneric>(cur, p, inv_den, clipped, has_qp, qd_p);
        // P_w -> TMEM (packed 16-bit pairs) over the first 32 columns of the S buffer this step consumed
        uint32_t pk[BNW / 2];
#pragma unroll
        for (int j = 0; j < BNW / 2; ++j) pk[j] = pack2<kBF16>(cur[2 * j], cur[2 * j + 1]);
#if !(defined(OEH_EXP) && (OEH_EXP == 3 || OEH_EXP == 5 || OEH_EXP == 6))
        ptx::tmem_st32(tS0 + (uint32_t)(i % kSBufs) * 64u, pk);
#else
        if (pk[0] == 0x12345u && pk[31] == 0x777u) ptx::tmem_st32(tS0 + (uint32_t)(i % kSBufs) * 64u, pk);
#endif
        OEH_T(4);
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        ptx::mbar_arrive(ptx::smem_u32(&bars.p_full[w][i % kSBufs]));
      }
      OEH_T(5);
      if (i + 2 < n_iter) load_s(i + 2, cur);                     // cur[] is dead: S_w(i+2) is complete in TMEM
      OEH_T(6);
    };

    float sa[BNW], sb[BNW];
    float tm_a = -INFINITY, tm_b = -INFINITY;
    load_s(0, sa);
    ptx::tmem_wait_ld();
    if (0 < nA) release_s(0);
    tm_a = pre(0, sa);
    if (n_iter > 1) load_s(1, sb);
#if !(defined(OEH_EXP) && (OEH_EXP == 6 || OEH_EXP == 7))
    if (w == 1) {                                                 // warpgroup 0 holds the XU token first
      float dummy = 0.f;
      ptx::nbar_arrive(bar_other, 64, dummy);
    }
#endif
    for (int i = 0; i < n_iter; i += 2) {
      step(i, sa, tm_a, sb, tm_b);
      if (i + 1 < n_iter) step(i + 1, sb, tm_b, sa, tm_a);
    }
#if !(defined(OEH_EXP) && (OEH_EXP == 6 || OEH_EXP == 7))
    if (w == 0) {                                                 // consume the last token so no arrival is left pending
      float dummy = 0.f;
      ptx::nbar_sync(bar_mine, 64, dummy);
    }
#endif

    OEH_T_PRINT("softmax: decide wait_ld token exp+pre pack+st wait_st+arrive load_s", n_iter);
    // ---- epilogue: O_w / denominator -> gate / taps -> global ----------------------------------------------
    ptx::mbar_wait(ptx::smem_u32(&bars.o_final[w]), 0);
    ptx::tc_fence_after();
    float o[HD];
    {
      uint32_t* orr = reinterpret_cast<uint32_t*>(o);
      ptx::tmem_ld32(tO, orr);
      ptx::tmem_ld32(tO + 32, orr + 32);
      ptx::tmem_wait_ld();
    }
    if (row < p.Tq) {
      float omul = 1.f;
      if (!normalized) {
        const float m_safe = m_run == -INFINITY ? 0.f : m_run;
        const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
        omul = __frcp_rn(den);
      }
      store_row<kBF16, kGeneric>(o, p, omul, b, h, row);
    }
  }

  // ---- teardown ------------------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    ptx::tmem_dealloc<kTmemCols>(tmem);
  }
}

}  // namespace attn_ws
}  // namespace oeh
