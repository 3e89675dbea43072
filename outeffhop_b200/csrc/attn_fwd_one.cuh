// Fused attention forward for sm_100a when all keys fit ONE K/V tile (Tk <= 128): the BERT eval shape
// (B=32, H=12, T=128: 384 (batch, head) items of 1.6 GFLOP / 25 MB in total -- HBM- and latency-bound, not tensor-bound).
//
// What matters here is how many independent items an SM has in flight, so this variant is built to be small:
//   * 128 threads (thread == query row == TMEM lane), 3 smem tiles (Q, K, V: 48 KB), 128 TMEM columns
//     -> FOUR CTAs per SM (the multi-tile kernel: two), i.e. the whole C-BERT grid (384 CTAs) is resident at once
//     on 148 x 4 slots and one CTA's TMA / MMA latency hides behind its neighbours' softmax.
//   * Softmax over the row, two shapes (template kRegS):
//     kRegS = false (generic instantiation: taps / masks): S is never held in registers as a whole row.  The row is
//     walked in 32-column chunks straight out of TMEM:
//       pass 1  max  (generic mode: the transformed scores are written back in place, so masks / taps are read once)
//       pass 2  e = exp2(s*mul - m), row sum; online kinds pack e to 16 bit and store it as P right away,
//               normalised kinds (clipped / probs tap) park e (fp32) in place
//       pass 3  (normalised kinds only) p = clip(stretch*e/den + gamma, 0, 1) [-> tap] -> P
//     kRegS = true (plain instantiation): the whole row (128 fp32) is read from TMEM once and stays in registers through
//     max -> exp / sum -> clip -> P (152 registers, THREE CTAs per SM: 444 slots still hold the 384-CTA grid); 5 % faster
//     for the clipped kinds, 2 % for softmax1 (profiles/r01_one_tile_regs_y.txt).  The share of polynomial exp2 makes no
//     difference here (same file): the launch is bound by its serial latency chain, not by the XU pipe.
//     P chunk c lands in columns [16c, 16c+16), which the same thread has already consumed (lane-private, in order).
//   * O [128 x 64] fp32 reuses columns 64..127: by the time the PV MMA is issued every thread has turned its S row
//     into P (columns 0..63), so S's upper half is dead.
// Same arithmetic as the other two kernels (attn_common.cuh); the two-pass "stats then PV" of the clipped modes
// collapses into one QK^T because the whole row is on chip.
#pragma once

#include "attn_common.cuh"

namespace oeh {
namespace attn_one {

using namespace attn;

constexpr int BM = 128;                          // query rows per CTA == threads per CTA
constexpr int CH = 32;                           // columns per softmax chunk
constexpr int kSmemBytes = 3 * kTileBytes + 1024;
constexpr uint32_t kTmemCols = 128;
constexpr uint32_t kColO = 64;

#ifdef OEH_TIMELINE
__device__ unsigned long long* g_timeline = nullptr;   // experiment builds only (tools/timeline_one.py): 8 stamps per CTA
__device__ __forceinline__ void stamp(int k) {
  if (threadIdx.x == 0 && g_timeline) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    g_timeline[cta * 8 + k] = t;
    if (k == 0) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_timeline[cta * 8 + 7] = sm; }
  }
}
#define OEH_STAMP(k) stamp(k)
#else
#define OEH_STAMP(k)
#endif

template <int N>
__device__ __forceinline__ void hide_invisible(float (&s)[N], int vis) {
  if (vis < N - 1) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j > vis) s[j] = -INFINITY;
  }
}

// One query row of a one-tile item: S (this thread's TMEM lane, columns 0..127) -> P (16 bit, columns 0..63 of the same lane).
// `wait_s` blocks until the QK^T MMA has landed; what is set up before it overlaps the loads.  Returns the factor the epilogue
// applies to the PV accumulator (1 / denominator for the online kinds, 1 for the normalised ones).
template <bool kBF16, bool kGeneric, bool kRegS, class WaitS>
__device__ __forceinline__ float softmax_row_to_p(uint32_t t_lane, const AttnParams& p, int b, int h, int row, WaitS&& wait_s) {
  // ---- per-row state (overlaps the loads / QK^T) -------------------------------------------------------------------
  const bool softmax1 = (p.kind == OEH_SOFTMAX_1 || p.kind == OEH_SOFTMAX_CLIPPED1);
  const bool clipped = p.kind >= OEH_SOFTMAX_CLIPPED;
  const bool has_qs = kGeneric && p.qdq_scores != nullptr;
  const bool has_qp = kGeneric && p.qdq_probs != nullptr;
  const bool has_mask = kGeneric && p.mask != nullptr;
  const bool normalized = clipped || has_qp;
  const bool xf = kGeneric && (has_qs || has_mask || p.clamp_min);   // transform_scores rewrites the scores
  Qdq qd_s{1.f, 1.f, 0.f, p.int_max}, qd_p{1.f, 1.f, 0.f, p.int_max};
  if (has_qs) qd_s = load_qdq(p.qdq_scores, p.int_max);
  if (has_qp) qd_p = load_qdq(p.qdq_probs, p.int_max);
  const int64_t mask_row = p.mask ? ((int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)(row < p.Tq ? row : 0) * p.ms2) : 0;
  const int causal_lim = p.causal ? row + (p.Tk - p.Tq) : 0x7fffffff;
  const int n_ch = (p.Tk + CH - 1) / CH;          // chunks that hold real keys (1..4)

  wait_s();

  float omul = 1.f;
  if constexpr (kRegS) {
    // ---- register-resident row: load once, max -> exp / sum -> normalise -> P ---------------------------------------
    float s[BN];
#pragma unroll
    for (int c = 0; c < BN / CH; ++c) ptx::tmem_ld32(t_lane + c * CH, reinterpret_cast<uint32_t*>(s) + c * CH);
    ptx::tmem_wait_ld();
    if (kGeneric && p.dbg_s && row < p.Tq) {
      float* d = p.dbg_s + (((int64_t)b * p.H + h) * p.Tq + row) * BN;
#pragma unroll
      for (int j = 0; j < BN; ++j) d[j] = s[j];
    }
    float mul = p.scale * kLog2e, m_run = -INFINITY;
#pragma unroll
    for (int c = 0; c < BN / CH; ++c) {
      if (c < n_ch) {
        float(&sc)[CH] = *reinterpret_cast<float(*)[CH]>(s + c * CH);
        const int k0 = c * CH, valid = min(p.Tk - k0, CH), vis = min(valid - 1, causal_lim - k0);
        mul = transform_scores<CH, kGeneric>(sc, p, qd_s, has_qs, has_mask, mask_row, k0, valid, vis);
        m_run = fmaxf(m_run, row_max<CH>(sc) * mul);
      }
    }
    const float m_safe = m_run == -INFINITY ? 0.f : m_run;
    float l_run = 0.f;
#pragma unroll
    for (int c = 0; c < BN / CH; ++c)
      if (c < n_ch) l_run += exp2_and_sum<CH, !kGeneric>(*reinterpret_cast<float(*)[CH]>(s + c * CH), mul, m_safe);
    const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
    const float inv_den = __frcp_rn(den);
    if (!normalized) omul = inv_den;
#pragma unroll
    for (int c = 0; c < BN / CH; ++c) {
      uint32_t pk[CH / 2];
      if (c < n_ch) {
        float(&sc)[CH] = *reinterpret_cast<float(*)[CH]>(s + c * CH);
        if (normalized) normalize_probs<CH, kGeneric, kBF16>(sc, p, inv_den, clipped, has_qp, qd_p);
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) pk[j] = pack2<kBF16>(sc[2 * j], sc[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) pk[j] = 0u;     // keys beyond Tk: P = 0 (the PV MMA always spans 128 keys)
      }
      ptx::tmem_st16(t_lane + c * (CH / 2), pk);
    }
  } else {
  // ---- pass 1: row maximum --------------------------------------------------------------------------------------
  float mul = p.scale * kLog2e, m_run = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < n_ch; ++c) {
    float s[CH];
    ptx::tmem_ld32(t_lane + c * CH, reinterpret_cast<uint32_t*>(s));
    ptx::tmem_wait_ld();
    if (kGeneric && p.dbg_s && row < p.Tq) {
      float* d = p.dbg_s + (((int64_t)b * p.H + h) * p.Tq + row) * BN + c * CH;
#pragma unroll
      for (int j = 0; j < CH; ++j) d[j] = s[j];
    }
    const int k0 = c * CH, valid = min(p.Tk - k0, CH), vis = min(valid - 1, causal_lim - k0);
    mul = transform_scores<CH, kGeneric>(s, p, qd_s, has_qs, has_mask, mask_row, k0, valid, vis);
    m_run = fmaxf(m_run, row_max<CH>(s) * mul);
    if (xf) ptx::tmem_st32(t_lane + c * CH, reinterpret_cast<const uint32_t*>(s));
  }
  if (xf) ptx::tmem_wait_st();
  const float m_safe = m_run == -INFINITY ? 0.f : m_run;

  // ---- pass 2: exponentials and the row sum ------------------------------------------------------------------------
  float l_run = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t pk[CH / 2];
    if (c < n_ch) {
      float s[CH];
      ptx::tmem_ld32(t_lane + c * CH, reinterpret_cast<uint32_t*>(s));
      ptx::tmem_wait_ld();
      if (!xf) {
        const int k0 = c * CH, valid = min(p.Tk - k0, CH);
        hide_invisible<CH>(s, min(valid - 1, causal_lim - k0));
      }
      l_run += exp2_and_sum<CH, !kGeneric>(s, mul, m_safe);
      if (normalized) {
        ptx::tmem_st32(t_lane + c * CH, reinterpret_cast<const uint32_t*>(s));
        continue;
      }
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) pk[j] = pack2<kBF16>(s[2 * j], s[2 * j + 1]);
    } else {
      if (normalized) continue;
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) pk[j] = 0u;   // keys beyond Tk: P = 0 (the PV MMA always spans 128 keys)
    }
    ptx::tmem_st16(t_lane + c * (CH / 2), pk);
  }

  // ---- pass 3 (normalised kinds): clip / tap, then P ------------------------------------------------------------------
  {
    const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
    const float inv_den = __frcp_rn(den);
    if (normalized) {
      ptx::tmem_wait_st();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[CH / 2];
        if (c < n_ch) {
          float s[CH];
          ptx::tmem_ld32(t_lane + c * CH, reinterpret_cast<uint32_t*>(s));
          ptx::tmem_wait_ld();
          normalize_probs<CH, kGeneric, kBF16>(s, p, inv_den, clipped, has_qp, qd_p);
          // (columns past Tk inside this chunk may hold clip(gamma) != 0: their V rows are TMA zero fill)
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) pk[j] = pack2<kBF16>(s[2 * j], s[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) pk[j] = 0u;
        }
        ptx::tmem_st16(t_lane + c * (CH / 2), pk);
      }
    } else {
      omul = inv_den;
    }
  }
  }   // chunked passes
  return omul;
}

// kRegS: the whole S row (128 fp32) lives in registers -- ONE TMEM read instead of the 2-3 chunked passes, at 168
// registers per thread, i.e. THREE CTAs per SM (444 slots still hold the whole C-BERT grid of 384).
template <bool kBF16, bool kGeneric, bool kRegS = false>
__global__ void __launch_bounds__(BM, kRegS ? 3 : 4)
attn_fwd_one_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                    const __grid_constant__ CUtensorMap tmv, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3];      // q+k, v, mma
  __shared__ uint32_t tmem_base_slot;

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it
  // (role, TMEM lane base, barrier / descriptor addresses) on the uniform datapath instead of R2UR-ing it at every use
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int row = q0 + tid;
  OEH_STAMP(0);

  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem0, sK = smem0 + kTileBytes, sV = smem0 + 2 * kTileBytes;
  const uint32_t bar_qk = ptx::smem_u32(&bars[0]), bar_v = ptx::smem_u32(&bars[1]), bar_mma = ptx::smem_u32(&bars[2]);

  constexpr uint32_t kIdescQK = ptx::make_idesc_f16(kBF16, BM, BN, false, false);
  constexpr uint32_t kIdescPV = ptx::make_idesc_f16(kBF16, BM, HD, false, true);

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmq);                    // descriptor fetches overlap the barrier initialisation
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
    // loads first: their latency covers the TMEM allocation and the barrier below
    ptx::mbar_expect_tx(bar_qk, 2 * kTileBytes);
    tma_tile(sQ, &tmq, bar_qk, p.cq, q0, h, b);
    tma_tile(sK, &tmk, bar_qk, p.ck, 0, h, b);
    ptx::mbar_expect_tx(bar_v, kTileBytes);
    tma_tile(sV, &tmv, bar_v, p.cv, 0, h, b);
  }
  __syncwarp();
  if (warp == 0) ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  OEH_STAMP(1);

  if (warp == 0 && ptx::elect_one()) {
    ptx::mbar_wait(bar_qk, 0);
    ptx::tc_fence_after();
#pragma unroll
    for (int k = 0; k < HD / 16; ++k)
      ptx::mma_ss(tmem, ptx::make_smem_desc_sw128(sQ + k * 32, 16, 1024), ptx::make_smem_desc_sw128(sK + k * 32, 16, 1024),
                  kIdescQK, k > 0);
    ptx::mma_commit(bar_mma);
  }
  __syncwarp();

  const float omul = softmax_row_to_p<kBF16, kGeneric, kRegS>(t_lane, p, b, h, row, [&] {
    ptx::mbar_wait(bar_mma, 0);
    ptx::tc_fence_after();
    OEH_STAMP(2);
  });
  ptx::tmem_wait_st();
  ptx::tc_fence_before();
  __syncthreads();
  OEH_STAMP(3);

  if (warp == 0 && ptx::elect_one()) {
    ptx::tc_fence_after();
    ptx::mbar_wait(bar_v, 0);
    ptx::tc_fence_after();
#pragma unroll
    for (int k = 0; k < BN / 16; ++k)
      ptx::mma_ts(tmem + kColO, tmem + k * 8, ptx::make_smem_desc_sw128(sV + k * 2048, 1024, 1024), kIdescPV, k > 0);
    ptx::mma_commit(bar_mma);
  }
  __syncwarp();

  // ---- epilogue -----------------------------------------------------------------------------------------------------
  const EpiPre pre = epi_prefetch<kBF16, kGeneric>(p, b, h, row);     // global reads of the epilogue, under the PV MMA
  ptx::mbar_wait(bar_mma, 1);
  ptx::tc_fence_after();
  OEH_STAMP(4);
  float o[HD];
  {
    uint32_t* orr = reinterpret_cast<uint32_t*>(o);
    ptx::tmem_ld32(t_lane + kColO, orr);
    ptx::tmem_ld32(t_lane + kColO + 32, orr + 32);
    ptx::tmem_wait_ld();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<kTmemCols>(tmem);
  OEH_STAMP(5);
  if (row < p.Tq) finish_row<kBF16, kGeneric>(o, p, omul, b, h, row, 0, pre);
  store_rows_coalesced<kBF16>(o, p, sQ + (uint32_t)warp * 4096u, b, h, q0 + warp * 32);   // Q is dead since the QK^T MMA
  OEH_STAMP(6);
}

}  // namespace attn_one
}  // namespace oeh
