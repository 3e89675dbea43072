// Warp-specialised fused attention forward for sm_100a (the long-sequence kernel).
//
// One CTA per SM works on TWO 128-row Q tiles of one (batch, head) at a time, FlashAttention-4 style, so the
// tensor core always has the other tile's MMAs to run while one tile is in its softmax:
//
//   warps 0-3   softmax warpgroup 0  (Q tile 0; thread == query row == TMEM lane)
//   warps 4-7   softmax warpgroup 1  (Q tile 1)
//   warp  8     MMA issuer   (one elected thread issues every tcgen05.mma and tcgen05.commit)
//   warp  9     TMA producer (one elected thread; Q once, K / V through 3-stage rings)
//   warps 10-11 idle (keep the warpgroup complete for setmaxnreg)
//
//   TMEM (512 columns):  S0/P0 [0,128)   S1/P1 [128,256)   O0 [256,320)   O1 [320,384)
//   mbarriers:  q_full | k_full/k_empty[3] | v_full/v_empty[3] (TMA <-> MMA)
//               s_full[w]  (MMA -> softmax w: S_w(it) is complete and so is every earlier MMA, incl. PV_w(it-1))
//               p_full[w]  (softmax w -> MMA, 128 arrivals: P_w(it) is in TMEM and O_w has been rescaled)
//   MMA issue order per K/V tile j:  [p_full0] PV0(j) QK0(j+1) commit->s_full0   [p_full1] PV1(j) QK1(j+1) commit->s_full1
//   so while warpgroup 0 runs exp() on S0(j+1), the tensor pipe executes PV1(j)/QK1(j+1), and vice versa.
//
// The softmax warps get 224 registers (setmaxnreg), the MMA/TMA warpgroup drops to 56.  The running maximum is only
// moved when a tile's maximum exceeds it by more than 2^8 ("lazy rescale"): P stays <= 256 (exact in fp16/bf16
// range, fp32 row sums), the result is identical after the final normalisation, and the O round trip through TMEM
// almost never happens after the first tiles.  softmax_1 adds exp2(-m) to the final denominator in the same frame.
#pragma once

#include "attn_common.cuh"

namespace oeh {
namespace attn_ws {

using namespace attn;

constexpr int kStages = 3;
constexpr int kThreads = 384;
constexpr int kQRows = 256;                                        // query rows per CTA (two tiles)
constexpr int kSmemBytes = (2 + 2 * kStages) * kTileBytes + 1024;  // Q0 Q1 | K ring | V ring
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS0 = 0, kColO0 = 256;                       // S_w at w*128, O_w at 256 + w*64
constexpr float kLazyRescale = 8.f;                                // log2 units

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2], p_full[2];
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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kQRows, h = blockIdx.y, b = blockIdx.z;

  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem0;                                    // 2 tiles
  const uint32_t sK = smem0 + 2 * kTileBytes;                   // kStages tiles
  const uint32_t sV = sK + kStages * kTileBytes;                // kStages tiles

  if (tid == 0) {
    ptx::mbar_init(ptx::smem_u32(&bars.q_full), 1);
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars.k_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.k_empty[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.v_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.v_empty[i]), 1);
    }
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      ptx::mbar_init(ptx::smem_u32(&bars.s_full[w]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.p_full[w]), 128);
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 8) ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  // ---- iteration plan (shared by all roles) ---------------------------------------------------------------------
  int n_kv = (p.Tk + BN - 1) / BN;
  if (p.causal) {
    const int last_row = min(q0 + kQRows, p.Tq) - 1;            // last real query row of this CTA
    const int lim = (last_row + (p.Tk - p.Tq)) / BN + 1;
    n_kv = lim < n_kv ? lim : n_kv;
  }
  const bool softmax1 = (p.kind == OEH_SOFTMAX_1 || p.kind == OEH_SOFTMAX_CLIPPED1);
  const bool clipped = p.kind >= OEH_SOFTMAX_CLIPPED;
  const bool normalized = clipped || (kGeneric && p.qdq_probs != nullptr);
  const bool one_tile = normalized && n_kv == 1;
  const int nA = (normalized && !one_tile) ? n_kv : 0;           // stats-pass iterations (QK^T only)
  const int n_iter = nA + n_kv;

  if (warp >= 8) {
    reg_dealloc<56>();
    if (warp == 9 && lane == 0) {
      // ================================ TMA producer ================================
      ptx::prefetch_tmap(&tmq);
      ptx::prefetch_tmap(&tmk);
      ptx::prefetch_tmap(&tmv);
      const uint32_t bq = ptx::smem_u32(&bars.q_full);
      ptx::mbar_expect_tx(bq, 2 * kTileBytes);
      tma_tile(sQ, &tmq, bq, p.cq, q0, h, b);
      tma_tile(sQ + kTileBytes, &tmq, bq, p.cq, q0 + BN, h, b);
      int kc = 0, vc = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int tile = it < nA ? it : it - nA;
        {
          const int s = kc % kStages, n = kc / kStages;
          if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.k_empty[s]), (uint32_t)(n - 1) & 1u);
          const uint32_t bf = ptx::smem_u32(&bars.k_full[s]);
          ptx::mbar_expect_tx(bf, kTileBytes);
          tma_tile(sK + s * kTileBytes, &tmk, bf, p.ck, tile * BN, h, b);
          ++kc;
        }
        if (it >= nA) {
          const int s = vc % kStages, n = vc / kStages;
          if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.v_empty[s]), (uint32_t)(n - 1) & 1u);
          const uint32_t bf = ptx::smem_u32(&bars.v_full[s]);
          ptx::mbar_expect_tx(bf, kTileBytes);
          tma_tile(sV + s * kTileBytes, &tmv, bf, p.cv, tile * BN, h, b);
          ++vc;
        }
      }
    } else if (warp == 8 && lane == 0) {
      // ================================ MMA issuer ==================================
      constexpr uint32_t kIdescQK = ptx::make_idesc_f16(kBF16, 128, BN, false, false);
      constexpr uint32_t kIdescPV = ptx::make_idesc_f16(kBF16, 128, HD, false, true);
      constexpr uint64_t kTileUnits = kTileBytes >> 4;           // descriptor address units (16 B)
      const uint64_t dq = ptx::make_smem_desc_sw128(sQ, 16, 1024);
      const uint64_t dk = ptx::make_smem_desc_sw128(sK, 16, 1024);
      const uint64_t dv = ptx::make_smem_desc_sw128(sV, 1024, 1024);
      auto issue_qk = [&](int w, int stage) {                    // S_w = Q_w . K(stage)^T
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::mma_ss(tmem + kColS0 + w * 128, dq + w * kTileUnits + k * 2, dk + stage * kTileUnits + k * 2, kIdescQK,
                      k > 0);
      };
      auto issue_pv = [&](int w, int stage, bool first) {        // O_w (+)= P_w . V(stage)
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)
          ptx::mma_ts(tmem + kColO0 + w * 64, tmem + kColS0 + w * 128 + k * 8, dv + stage * kTileUnits + k * 128,
                      kIdescPV, (first && k == 0) ? 0u : 1u);
      };
      ptx::mbar_wait(ptx::smem_u32(&bars.q_full), 0);
      int kc = 0, vc = 0;
      ptx::mbar_wait(ptx::smem_u32(&bars.k_full[0]), 0);
      ptx::tc_fence_after();
      issue_qk(0, 0);
      ptx::mma_commit(ptx::smem_u32(&bars.s_full[0]));
      issue_qk(1, 0);
      ptx::mma_commit(ptx::smem_u32(&bars.s_full[1]));
      ptx::mma_commit(ptx::smem_u32(&bars.k_empty[0]));
      kc = 1;
      for (int it = 0; it < n_iter; ++it) {
        const bool pv = it >= nA;
        const bool more = it + 1 < n_iter;
        const int vs = vc % kStages, ks = kc % kStages;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          ptx::mbar_wait(ptx::smem_u32(&bars.p_full[w]), (uint32_t)it & 1u);
          ptx::tc_fence_after();
          if (pv) {
            if (w == 0) {
              ptx::mbar_wait(ptx::smem_u32(&bars.v_full[vs]), (uint32_t)(vc / kStages) & 1u);
              ptx::tc_fence_after();
            }
            issue_pv(w, vs, it == nA);
            if (w == 1) ptx::mma_commit(ptx::smem_u32(&bars.v_empty[vs]));
          }
          if (more) {
            if (w == 0) {
              ptx::mbar_wait(ptx::smem_u32(&bars.k_full[ks]), (uint32_t)(kc / kStages) & 1u);
              ptx::tc_fence_after();
            }
            issue_qk(w, ks);
          }
          ptx::mma_commit(ptx::smem_u32(&bars.s_full[w]));       // S_w(it+1) ready / O_w final
          if (more && w == 1) ptx::mma_commit(ptx::smem_u32(&bars.k_empty[ks]));
        }
        if (pv) ++vc;
        if (more) ++kc;
      }
    }
  } else {
    // ================================ softmax warpgroups ==============================
    reg_alloc<224>();
    const int w = warp >> 2;                                      // warpgroup == Q tile
    const int row = q0 + w * BN + (tid & 127);                    // this thread's query row
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tS = t_lane + kColS0 + w * 128, tO = t_lane + kColO0 + w * 64;
    const uint32_t bar_s = ptx::smem_u32(&bars.s_full[w]), bar_p = ptx::smem_u32(&bars.p_full[w]);

    const bool has_qs = kGeneric && p.qdq_scores != nullptr;
    const bool has_qp = kGeneric && p.qdq_probs != nullptr;
    const bool has_mask = kGeneric && p.mask != nullptr;
    Qdq qd_s{1.f, 1.f, 0.f, p.int_max}, qd_p{1.f, 1.f, 0.f, p.int_max};
    if (has_qs) qd_s = load_qdq(p.qdq_scores, p.int_max);
    if (has_qp) qd_p = load_qdq(p.qdq_probs, p.int_max);
    const int64_t mask_row =
        p.mask ? ((int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)(row < p.Tq ? row : 0) * p.ms2) : 0;
    const int causal_lim = p.causal ? row + (p.Tk - p.Tq) : 0x7fffffff;

    float m_run = -INFINITY, l_run = 0.f, inv_den = 1.f;

    for (int it = 0; it < n_iter; ++it) {
      ptx::mbar_wait(bar_s, (uint32_t)it & 1u);
      ptx::tc_fence_after();
      const bool pv_iter = it >= nA;
      const int tile = pv_iter ? it - nA : it;
      const int k0 = tile * BN;

      float s[BN];
      {
        uint32_t* sr = reinterpret_cast<uint32_t*>(s);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) ptx::tmem_ld32(tS + c * 32, sr + c * 32);
        ptx::tmem_wait_ld();
      }
      if (!pv_iter) {                                             // stats pass: S is consumed, next QK^T may start
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_p);
      }
      if (kGeneric && p.dbg_s && it == 0 && row < p.Tq) {
        float* d = p.dbg_s + (((int64_t)b * p.H + h) * p.Tq + row) * BN;
#pragma unroll
        for (int j = 0; j < BN; ++j) d[j] = s[j];
      }

      const int valid = min(p.Tk - k0, BN);
      const int vis = min(valid - 1, causal_lim - k0);
      const float mul = transform_scores<kGeneric>(s, p, qd_s, has_qs, has_mask, mask_row, k0, valid, vis);
      const float tile_max = row_max(s) * mul;

      if (!pv_iter) {
        // stats pass: exact running (m, l)
        const float m_new = fmaxf(m_run, tile_max);
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float alpha = ptx::ex2(m_run - m_safe);
        l_run = fmaf(l_run, alpha, exp2_and_sum(s, mul, m_safe));
        m_run = m_new;
        continue;
      }
      if (normalized && !one_tile) {
        // PV pass of the two-pass mode: (m, l) are final
        const float m_safe = m_run == -INFINITY ? 0.f : m_run;
        if (it == nA) {
          const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
          inv_den = __frcp_rn(den);
        }
#pragma unroll
        for (int j = 0; j < BN; ++j) s[j] = ptx::ex2(fmaf(s[j], mul, -m_safe));
      } else {
        // online softmax with lazy rescale (or the single-tile normalised case, where m is simply the tile max)
        const bool move = tile_max > m_run + kLazyRescale || m_run == -INFINITY;
        const float m_new = move ? tile_max : m_run;              // tile_max may be -inf only if m_run is too
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float alpha = move ? ptx::ex2(m_run - m_safe) : 1.f;
        l_run = fmaf(l_run, alpha, exp2_and_sum(s, mul, m_safe));
        m_run = m_new;
        if (one_tile) {
          const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
          inv_den = __frcp_rn(den);
        } else if (it > nA && !__all_sync(0xffffffffu, alpha == 1.f)) {
          // O_w *= alpha.  PV_w(it-1) has completed: s_full(it) was committed after it.
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
      if (normalized) normalize_probs<kGeneric>(s, p, inv_den, clipped, has_qp, qd_p);

      // P_w -> TMEM (packed 16-bit pairs) over the S_w columns this thread has consumed
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack2<kBF16>(s[c * 32 + 2 * j], s[c * 32 + 2 * j + 1]);
        ptx::tmem_st16(tS + c * 16, pk);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_p);
    }

    // ---- epilogue: O_w / denominator -> gate / taps -> global ----------------------------------------------
    ptx::mbar_wait(bar_s, (uint32_t)n_iter & 1u);
    ptx::tc_fence_after();
    float o[HD];
    {
      uint32_t* orr = reinterpret_cast<uint32_t*>(o);
      ptx::tmem_ld32(tO, orr);
      ptx::tmem_ld32(tO + 32, orr + 32);
      ptx::tmem_wait_ld();
    }
    if (row < p.Tq) {
      float mul = 1.f;
      if (!normalized) {
        const float m_safe = m_run == -INFINITY ? 0.f : m_run;
        const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
        mul = __frcp_rn(den);
      }
      store_row<kBF16, kGeneric>(o, p, mul, b, h, row);
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
