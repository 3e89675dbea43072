// One-tile fused attention (Tk <= 128) with TWO threads per query row: 256 threads per CTA, three CTAs per SM.
//
// Why (profiles/r02_timeline_one.txt): with one thread per row (attn_fwd_one.cuh) a CTA is a serial chain -- loads, QK^T, ~1.6 us
// of softmax on ONE warp per scheduler (128 exponentials, 2-3 TMEM passes per thread; ~3 us with the INT8 taps), PV, store --
// and the launch ends when the last CTA's chain does, about 3 us after the last byte of Q / K / V has landed.  The SM is far from
// busy during that tail (issue slots ~25 % used), so the row is split: warps w and w+4 share a TMEM lane quarter, warp w takes
// keys [0, 64), warp w+4 keys [64, 128), in 32-column chunks straight out of TMEM like the chunked one-thread kernel:
//   pass 1  max of the own half (generic mode: transformed scores written back in place)        -> (m) exchanged through smem
//   pass 2  e = exp2(s*mul - m), partial sum; online kinds pack e as P right away,
//           normalised kinds (clipped / probs tap) park e in place                              -> (l) exchanged through smem
//   pass 3  (normalised kinds) p = clip(stretch*e/den + gamma, 0, 1) [-> tap] -> P
// P chunk c lands in columns [16c, 16c+16), as in the one-thread kernel.  For the upper half those columns belong to the LOWER
// half's second chunk, so in the pass that writes P the lower warp announces "chunk 1 is in my registers" on a named barrier
// (bar.arrive, it does not wait) and the upper warp passes the same barrier (bar.sync) before its first P store.
// Epilogue: each thread finishes 32 of the row's 64 context columns; rows leave as full 64-byte segments staged per warp.
// Same arithmetic per element as attn_fwd_one.cuh; the row sum is added in a different order (two partial sums).
#pragma once

#include "attn_fwd_one.cuh"

namespace oeh {
namespace attn_one2 {

using namespace attn;
using attn_one::hide_invisible;

constexpr int BM = 128;                          // query rows per CTA
constexpr int CH = 32;                           // columns per softmax chunk
constexpr int kThreads = 2 * BM;
constexpr int kSmemBytes = 3 * kTileBytes + 1024;
constexpr uint32_t kTmemCols = 128;
constexpr uint32_t kColO = 64;

template <bool kBF16, bool kGeneric>
__global__ void __launch_bounds__(kThreads, 3)
attn_fwd_one2_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                     const __grid_constant__ CUtensorMap tmv, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3];      // q+k, v, mma
  __shared__ uint32_t tmem_base_slot;
  __shared__ float xch_m[2][BM], xch_l[2][BM];   // (half, row): the halves' row statistics

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int half = warp >> 2, wq = warp & 3;     // key half; TMEM lane quarter
  const int r = wq * 32 + (tid & 31);            // row within the tile == TMEM lane
  const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int row = q0 + r;

  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem0, sK = smem0 + kTileBytes, sV = smem0 + 2 * kTileBytes;
  const uint32_t bar_qk = ptx::smem_u32(&bars[0]), bar_v = ptx::smem_u32(&bars[1]), bar_mma = ptx::smem_u32(&bars[2]);

  constexpr uint32_t kIdescQK = ptx::make_idesc_f16(kBF16, BM, BN, false, false);
  constexpr uint32_t kIdescPV = ptx::make_idesc_f16(kBF16, BM, HD, false, true);

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
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
  const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);

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
  const int c_lo = 2 * half;                      // this thread's chunks: c_lo, c_lo + 1
  const int bar_x = 1 + wq, bar_hz = 5 + wq;      // named barriers of the row's two warps: statistics exchange, P hazard

  // Writes P chunk c (16 packed columns at [16c, 16c+16)).  The upper half's targets are the lower half's chunk 1: the lower
  // warp announces that chunk 1 is in registers (lower_consumed), the upper warp waits for that once.
  bool hz_passed = false;
  auto store_p = [&](int c, const uint32_t (&pk)[CH / 2]) {
    if (half == 1 && !hz_passed) {
      asm volatile("bar.sync %0, 64;" ::"r"(bar_hz) : "memory");
      ptx::tc_fence_after();
      hz_passed = true;
    }
    ptx::tmem_st16(t_lane + c * (CH / 2), pk);
  };
  auto lower_consumed = [&](int c) {               // after tcgen05.wait::ld of chunk c in the pass that writes P
    if (half == 0 && c == 1) {
      ptx::tc_fence_before();
      asm volatile("bar.arrive %0, 64;" ::"r"(bar_hz) : "memory");
    }
  };
  auto exchange = [&](float (&x)[2][BM], float own) {  // returns the other half's value
    x[half][r] = own;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_x) : "memory");
    return x[half ^ 1][r];
  };

  ptx::mbar_wait(bar_mma, 0);
  ptx::tc_fence_after();

  // ---- pass 1: maximum of the own half -----------------------------------------------------------------------------------
  float mul = p.scale * kLog2e, m_run = -INFINITY;
#pragma unroll 1
  for (int c = c_lo; c < c_lo + 2; ++c) {
    if (c >= n_ch) break;
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
  m_run = fmaxf(m_run, exchange(xch_m, m_run));
  const float m_safe = m_run == -INFINITY ? 0.f : m_run;

  // ---- pass 2: exponentials and the partial row sum ----------------------------------------------------------------------
  float l_run = 0.f;
#pragma unroll 1
  for (int c = c_lo; c < c_lo + 2; ++c) {
    uint32_t pk[CH / 2];
    if (c < n_ch) {
      float s[CH];
      ptx::tmem_ld32(t_lane + c * CH, reinterpret_cast<uint32_t*>(s));
      ptx::tmem_wait_ld();
      if (!normalized) lower_consumed(c);
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
      if (!normalized) lower_consumed(c);
      if (normalized) continue;
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) pk[j] = 0u;   // keys beyond Tk: P = 0 (the PV MMA always spans 128 keys)
    }
    store_p(c, pk);
  }
  const float l_all = l_run + exchange(xch_l, l_run);
  const float den = softmax1 ? l_all + ptx::ex2(-m_safe) : l_all;
  const float inv_den = __frcp_rn(den);
  float omul = 1.f;

  // ---- pass 3 (normalised kinds): clip / tap, then P ------------------------------------------------------------------
  if (normalized) {
    ptx::tmem_wait_st();
#pragma unroll 1
    for (int c = c_lo; c < c_lo + 2; ++c) {
      uint32_t pk[CH / 2];
      if (c < n_ch) {
        float s[CH];
        ptx::tmem_ld32(t_lane + c * CH, reinterpret_cast<uint32_t*>(s));
        ptx::tmem_wait_ld();
        lower_consumed(c);
        normalize_probs<CH, kGeneric, kBF16>(s, p, inv_den, clipped, has_qp, qd_p);
        // (columns past Tk inside this chunk may hold clip(gamma) != 0: their V rows are TMA zero fill)
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) pk[j] = pack2<kBF16>(s[2 * j], s[2 * j + 1]);
      } else {
        lower_consumed(c);
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) pk[j] = 0u;
      }
      store_p(c, pk);
    }
  } else {
    omul = inv_den;
  }
  ptx::tmem_wait_st();
  ptx::tc_fence_before();
  __syncthreads();

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

  // ---- epilogue: each of the row's threads finishes 32 context columns --------------------------------------------------
  const EpiPre pre = epi_prefetch<kBF16, kGeneric>(p, b, h, row);     // global reads of the epilogue, under the PV MMA
  ptx::mbar_wait(bar_mma, 1);
  ptx::tc_fence_after();
  float o[32];
  ptx::tmem_ld32(t_lane + kColO + half * 32, reinterpret_cast<uint32_t*>(o));
  ptx::tmem_wait_ld();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<kTmemCols>(tmem);
  if (row < p.Tq) finish_row<kBF16, kGeneric, 32>(o, p, omul, b, h, row, half * 32, pre);
  store_half_rows_coalesced<kBF16>(o, p, sQ + (uint32_t)warp * 2048u, b, h, q0 + wq * 32, half * 32);   // Q is dead since QK^T
}

}  // namespace attn_one2
}  // namespace oeh
