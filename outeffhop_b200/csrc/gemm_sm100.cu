// Linear layer with fused epilogue for sm_100a:  Y = tap( act( X . W^T + bias ) )      (X: M x K, W: N x K, both K-major)
//
// The reference's QuantLinear quantizes the Linear output inside the same module call (quantization/hijacker.py:78-134,
// autoquant_utils.py:18-91: F.linear -> [activation] -> activation_quantizer); HF's BertIntermediate applies the erf GELU
// right behind its dense layer.  As separate kernels each of those is a full extra pass over the (M, N) activation
// (profiles/r01_launches_int8_opt_step_aa.summary.txt: 24 % of the INT8 OPT step, 12 % GELU in the fp16 BERT step).  Here
// they run in the TMEM -> register epilogue of a hand-written tcgen05 GEMM:
//
//   warp 0       TMA producer: A (128 x 64) and B (BN x 64) k-blocks, SWIZZLE_128B, through a kStages-deep mbarrier ring
//   warp 1       MMA issuer: tcgen05.mma kind::f16 M128 x N(BN) x K16, fp32 accumulators in TMEM, DOUBLE-BUFFERED
//                (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2..17  epilogue, FOUR warps per TMEM lane quarter (each takes a quarter of the tile's columns; with two the erf GELU
//                was latency-bound: 26.8 -> 23.9 us at 4096 x 3072 x 768, profiles/r02_gemm_epilogue_*.txt): tcgen05.ld 32 columns
//                -> + bias -> round to the tensor dtype (what the unfused GEMM would have stored) -> ReLU / erf GELU
//                -> fixed-range fake-quant tap (per-column-segment qparams: the fused q|k|v projection keeps three ranges)
//                -> 16-bit pack -> the warp's own swizzled staging tile (32 rows x 32 columns) -> TMA store (clipped at the edges);
//                no barrier between epilogue warps
//   persistent CTAs (one per SM) walk the output tiles with stride gridDim.x.
//
// kCta2 = true: the CTA-PAIR form (cta_group::2, clusters of two CTAs on one TPC).  The pair computes a 256 x BN tile: each
// CTA loads its own 128 rows of A and HALF of the B tile (BN/2 rows), the leader CTA issues tcgen05.mma.cta_group::2
// (M = 256) which reads both halves of B, and every CTA finishes its own 128 accumulator rows.  Per CTA and k-block that is
// 32 KB through L2 and shared memory instead of 48 KB: the single-CTA form is bound by exactly that traffic (operands
// written by TMA and read back by the tensor core: 96 + 96 B/clk against a 128 B/clk shared-memory port).
//
// Roofline: tensor pipe, 2 M N K flop per call.  Like the fused attention taps, the epilogue tap is tolerance-parity with
// `GEMM -> oeh_fake_quant_asym` (accumulation order moves values across bin edges); the standalone quantizer keeps the
// bit-exact contract.
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

#include "gelu_math.cuh"
#include "oeh_common.cuh"
#include "sm100_ptx.cuh"

namespace oeh {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int kStageA = BM * BK * 2;                     // 16 KB
// Epilogue warps: kParts column parts x 4 TMEM lane quarters.  128 / 256-column tiles: 4 parts (16 warps); 192-column tiles: SIX
// parts of one 32-column chunk each (24 warps, <= 72 registers) -- the erf-GELU epilogue of a 128 x 256 tile on 16 warps takes
// longer (4.3 us) than the tile's main loop (4.0 us), profiles/r02_timeline_gemm.txt.
template <int BN>
struct Epi {
  static constexpr int kParts = BN == 192 ? 6 : 4;
  static constexpr int kEpiWarps = 4 * kParts;
  static constexpr int kThreads = (2 + kEpiWarps) * 32;
};
constexpr int kChunk = 32;                               // columns per epilogue step / TMA store box
constexpr int kOutBuf = BM * kChunk * 2;                 // staging per column part: 4 warps x 2 KB (32 rows x 64 B, SWIZZLE_64B), x 2 buffers
constexpr int kMaxSeg = 4;

template <int BN, bool kCta2>
struct Cfg {
  static constexpr int kParts = Epi<BN>::kParts;
  static constexpr int kStageB = (kCta2 ? BN / 2 : BN) * BK * 2;        // rows of B this CTA loads per k-block
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kRing = 226 * 1024 - 2 * kParts * kOutBuf;       // what the staging tiles leave of 227 KB
  static constexpr int kStages = kRing / kStage > 8 ? 8 : kRing / kStage;
  static constexpr uint32_t kTmemCols = BN <= 128 ? 256 : 512;          // two accumulators
  static constexpr int kSmemBytes = kStages * kStage + 2 * kParts * kOutBuf + 1024;
};

struct GemmParams {
  int M, N, K;
  int n_mblk, n_nblk, n_kblk;
  const void* bias;
  int act;                                               // OEH_ACT_*
  const float* qp[kMaxSeg];                              // device [scale, zero_point] per column segment (NULL = no tap)
  int seg_cols;                                          // columns per segment (multiple of the tile width)
  float int_max;
  int dbg;                                               // timing experiments (tile_n / 10000): 1 = no output stores
};

template <int kStages>
struct Barriers {
  uint64_t full[kStages], empty[kStages];
  uint64_t tmem_full[2], tmem_empty[2];
};

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) forms -----------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;            // shared::cluster address of the same offset in the EVEN CTA of a pair
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_result) {   // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// all MMAs issued so far by this thread arrive (when complete) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// the bytes land in THIS CTA's shared memory, the transaction count goes to the leader's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {        // arrive on the leader CTA's barrier (own, if leader)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

template <bool kBF16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if constexpr (kBF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
template <bool kBF16>
__device__ __forceinline__ void unpack16(uint32_t u, float& a, float& b) {
  if constexpr (kBF16) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
    a = f.x; b = f.y;
  } else {
    const float2 f = __half22float2(*reinterpret_cast<__half2*>(&u));
    a = f.x; b = f.y;
  }
}

#ifdef OEH_TIMELINE
__device__ unsigned long long* g_gemm_timeline = nullptr;   // experiment builds only (tools/timeline_gemm.py): 16 stamps per CTA
__device__ __forceinline__ void gemm_stamp(int k) {
  if (g_gemm_timeline && k < 16) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_gemm_timeline[blockIdx.x * 16 + k] = t;
  }
}
#define GEMM_STAMP(k) gemm_stamp(k)
#else
#define GEMM_STAMP(k)
#endif

template <bool kBF16, int BN, bool kCta2>
__global__ void __launch_bounds__(Epi<BN>::kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb,
            const __grid_constant__ CUtensorMap tmy, const __grid_constant__ GemmParams p) {
  using C = Cfg<BN, kCta2>;
  constexpr int kStages = C::kStages, kParts = Epi<BN>::kParts, kEpiWarps = Epi<BN>::kEpiWarps;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Barriers<kStages> bars;
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint4 s_bias[2][Epi<BN>::kParts][BN / Epi<BN>::kParts / 8];   // (tile parity, column part): the part's bias slice

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sOut = smem0 + kStages * C::kStage;             // 4 staging tiles of 8 KB (2 per column half)
  // work items: output tiles (single CTA) or 256-row tile pairs (CTA pair: rank r takes rows [256 i + 128 r, +128))
  const int rank = kCta2 ? (int)cluster_ctarank() : 0;
  const int n_mitems = kCta2 ? (p.n_mblk + 1) / 2 : p.n_mblk;
  const int n_items = n_mitems * p.n_nblk;
  const int item0 = kCta2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_stride = kCta2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_m0 = [&](int item) { return ((item % n_mitems) * (kCta2 ? 2 : 1) + rank) * BM; };
  auto tile_n0 = [&](int item) { return (item / n_mitems) * BN; };

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars.full[i]), kCta2 ? 2 : 1);      // pair: one arrival from each CTA's producer
      ptx::mbar_init(ptx::smem_u32(&bars.empty[i]), 1);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars.tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.tmem_empty[i]), kCta2 ? 2 * kEpiWarps : kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) {
    if constexpr (kCta2) tmem_alloc_pair<C::kTmemCols>(ptx::smem_u32(&tmem_base_slot));
    else ptx::tmem_alloc<C::kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCta2) cluster_sync_all();        // the peer's barriers are initialised before anything is sent to them
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp == 0) {
    // ================================ TMA producer ======================================================================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tma);
      ptx::prefetch_tmap(&tmb);
    }
    int it = 0;                                                  // k-block counter across tiles -> ring slot / parity
    for (int item = item0; item < n_items; item += item_stride) {
      const int m0 = tile_m0(item), n0 = tile_n0(item);
      for (int kb = 0; kb < p.n_kblk; ++kb, ++it) {
        const int s = it % kStages, n = it / kStages;
        if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.empty[s]), (uint32_t)(n - 1) & 1u);
        if (ptx::elect_one()) {
          const uint32_t bf = ptx::smem_u32(&bars.full[s]);
          if constexpr (kCta2) {
            // both CTAs' bytes are counted on the LEADER's barrier; the leader announces them, the other CTA just arrives
            if (rank == 0) ptx::mbar_expect_tx(bf, 2 * C::kStage);
            else mbar_arrive_leader(bf);
            tma_load_2d_pair(smem0 + s * C::kStage, &tma, bf, kb * BK, m0);
            tma_load_2d_pair(smem0 + s * C::kStage + kStageA, &tmb, bf, kb * BK, n0 + rank * (BN / 2));
          } else {
            ptx::mbar_expect_tx(bf, C::kStage);
            tma_load_2d(smem0 + s * C::kStage, &tma, bf, kb * BK, m0);
            tma_load_2d(smem0 + s * C::kStage + kStageA, &tmb, bf, kb * BK, n0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ========================================================================
    constexpr uint32_t kIdesc = ptx::make_idesc_f16(kBF16, kCta2 ? 2 * BM : BM, BN, false, false);
    int it = 0, ti = 0;
    if (!kCta2 || rank == 0) {                                   // pair: the leader issues for both CTAs
      for (int item = item0; item < n_items; item += item_stride, ++ti) {
        const int acc = ti & 1;
        if (ti >= 2) ptx::mbar_wait(ptx::smem_u32(&bars.tmem_empty[acc]), (uint32_t)((ti >> 1) - 1) & 1u);
        ptx::tc_fence_after();
        if ((tid & 31) == 0) GEMM_STAMP(4 * ti + 0);
        for (int kb = 0; kb < p.n_kblk; ++kb, ++it) {
          const int s = it % kStages;
          ptx::mbar_wait(ptx::smem_u32(&bars.full[s]), (uint32_t)(it / kStages) & 1u);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t a = smem0 + s * C::kStage, b = a + kStageA;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = ptx::make_smem_desc_sw128(a + k * 32, 16, 1024), db = ptx::make_smem_desc_sw128(b + k * 32, 16, 1024);
              if constexpr (kCta2) mma_ss_pair(tmem + acc * BN, da, db, kIdesc, (kb | k) != 0);
              else ptx::mma_ss(tmem + acc * BN, da, db, kIdesc, (kb | k) != 0);
            }
            if constexpr (kCta2) {
              mma_commit_pair(ptx::smem_u32(&bars.empty[s]));    // frees the slot in BOTH CTAs
              if (kb + 1 == p.n_kblk) mma_commit_pair(ptx::smem_u32(&bars.tmem_full[acc]));
            } else {
              ptx::mma_commit(ptx::smem_u32(&bars.empty[s]));    // slot free once these MMAs have read it
              if (kb + 1 == p.n_kblk) ptx::mma_commit(ptx::smem_u32(&bars.tmem_full[acc]));
            }
          }
          __syncwarp();
        }
        if ((tid & 31) == 0) GEMM_STAMP(4 * ti + 1);
      }
    }
  } else {
    // ================================ epilogue: 2 warps per TMEM lane quarter ============================================
    const int q = warp & 3, part = (warp - 2) >> 2;                // lane quarter; column part
    const int r = q * 32 + (tid & 31);                             // row within the tile == TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    // Everything below is WARP-PRIVATE -- no barrier between epilogue warps: a warp stages its own 32 rows x 32 columns (2 KB,
    // double-buffered), its lane 0 issues and tracks the TMA stores of those boxes, and the bias slice of its column part is
    // read into its own 128 bytes.  (With one staging tile and one named barrier per column part, bar.sync / fence stalls
    // were ~20 % of the epilogue warps' samples, and the erf-GELU epilogue -- 4.3 us per tile against a 4.0 us main loop --
    // is what bounds that GEMM: profiles/r02_timeline_gemm.txt.)
    const int ew = warp - 2;                                       // 0 .. kEpiWarps-1
    const bool issuer = (tid & 31) == 0;
    const uint32_t sBuf = sOut + (uint32_t)ew * (2u * 2048u);
    const uint32_t lane = (uint32_t)tid & 31u;
    const uint32_t row_off = lane * 64u, sw = (lane >> 1) & 3u;    // SWIZZLE_64B: chunk ^= (row / 2) % 4
    if (warp == 2 && ptx::elect_one()) ptx::prefetch_tmap(&tmy);
    int ti = 0, chunk_it = 0;
    for (int item = item0; item < n_items; item += item_stride, ++ti) {
      const int m0 = tile_m0(item), n0 = tile_n0(item);
      const int acc = ti & 1;
      // per-tile tap parameters (a tile never straddles two column segments: host checks seg_cols % BN == 0)
      const float* qp = p.seg_cols > 0 ? p.qp[n0 / p.seg_cols] : p.qp[0];
      float q_scale = 1.f, q_inv = 1.f, q_lo = 0.f, q_hi = 0.f;
      if (qp) {
        q_scale = __ldg(qp);
        const float zp = __ldg(qp + 1);
        q_inv = __frcp_rn(q_scale);
        q_lo = -zp;
        q_hi = p.int_max - zp;
      }
      // The bias of this column part -> shared memory before the accumulator is waited for (a global load per chunk on the
      // critical path of the epilogue cost ~0.3 us each); columns beyond N read as 0.  Written by the part's first warp into
      // the slot of this tile's parity; ONE named barrier of the part's four warps per tile (off the critical path: the main
      // loop is still running) publishes it, and the same barrier one tile later orders these reads before the slot's reuse.
      uint4* my_bias = &s_bias[acc][part][0];
      {
        constexpr int kV = BN / kParts / 8;
        if (q == 0 && (int)lane < kV) {
          uint4 bv = make_uint4(0u, 0u, 0u, 0u);
          const int gc = n0 + part * (BN / kParts) + 8 * (int)lane;
          if (p.bias) {
            if (gc + 8 <= p.N) {
              bv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias) + gc));
            } else {
              uint16_t tmp[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) tmp[j] = gc + j < p.N ? __ldg(reinterpret_cast<const uint16_t*>(p.bias) + gc + j) : (uint16_t)0;
              bv = *reinterpret_cast<const uint4*>(tmp);
            }
          }
          my_bias[lane] = bv;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
      }
      ptx::mbar_wait(ptx::smem_u32(&bars.tmem_full[acc]), (uint32_t)(ti >> 1) & 1u);
      ptx::tc_fence_after();
      if (tid == 64) GEMM_STAMP(4 * ti + 2);
#pragma unroll 1
      for (int c = 0; c < BN / kParts / kChunk; ++c, ++chunk_it) {
        const int col = part * (BN / kParts) + c * kChunk;              // column within the tile
        float v[kChunk];
        ptx::tmem_ld32(t_lane + acc * BN + col, reinterpret_cast<uint32_t*>(v));
        uint4 bvec[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) bvec[g] = my_bias[c * (kChunk / 8) + g];
        ptx::tmem_wait_ld();
        if (c + 1 == BN / kParts / kChunk) {                            // last read of this accumulator (and of the bias) by this warp
          ptx::tc_fence_before();
          __syncwarp();
          if (ptx::elect_one()) {
            if constexpr (kCta2) mbar_arrive_leader(ptx::smem_u32(&bars.tmem_empty[acc]));   // the leader's issuer waits for both CTAs
            else ptx::mbar_arrive(ptx::smem_u32(&bars.tmem_empty[acc]));
          }
        }
        const int gcol = n0 + col;
        if (p.bias) {                                              // 32 bias values = 64 B, the same for every lane (broadcast)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 bv = bvec[g];
            const uint32_t w[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float b0, b1;
              unpack16<kBF16>(w[e], b0, b1);
              v[g * 8 + 2 * e] += b0;
              v[g * 8 + 2 * e + 1] += b1;
            }
          }
        }
        uint32_t pk[kChunk / 2];
        if (p.act == OEH_ACT_NONE && !qp) {
#pragma unroll
          for (int j = 0; j < kChunk / 2; ++j) pk[j] = pack16<kBF16>(v[2 * j], v[2 * j + 1]);
        } else {
          // what the unfused path would have stored after the GEMM, then the activation and the tap on that value
#pragma unroll
          for (int j = 0; j < kChunk / 2; ++j) unpack16<kBF16>(pack16<kBF16>(v[2 * j], v[2 * j + 1]), v[2 * j], v[2 * j + 1]);
          if (p.act == OEH_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < kChunk; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (p.act == OEH_ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < kChunk; j += 2) {
              gelu_math::gelu2<false>(v[j], v[j + 1], v[j], v[j + 1]);
              if (qp) unpack16<kBF16>(pack16<kBF16>(v[j], v[j + 1]), v[j], v[j + 1]);   // the activation's own 16-bit store, which the tap reads
            }
          }
          if (qp) {
            const uint64_t inv2 = ptx::pack2f(q_inv, q_inv), sc2 = ptx::pack2f(q_scale, q_scale);
            const uint64_t mg = ptx::pack2f(12582912.f, 12582912.f), nmg = ptx::pack2f(-12582912.f, -12582912.f);
#pragma unroll
            for (int j = 0; j < kChunk; j += 2) {
              float a, b;
              ptx::unpack2f(ptx::fmul2(ptx::pack2f(v[j], v[j + 1]), inv2), a, b);
              a = fminf(fmaxf(a, q_lo), q_hi);
              b = fminf(fmaxf(b, q_lo), q_hi);
              const uint64_t c2 = ptx::fadd2(ptx::fadd2(ptx::pack2f(a, b), mg), nmg);   // rint, exact for |.| < 2^22
              ptx::unpack2f(ptx::fmul2(c2, sc2), v[j], v[j + 1]);
            }
          }
#pragma unroll
          for (int j = 0; j < kChunk / 2; ++j) pk[j] = pack16<kBF16>(v[2 * j], v[2 * j + 1]);
        }
        if (p.dbg & 1) continue;
        // warp-private staging tile (double-buffered) -> TMA store of a 32 x 32 box
        const uint32_t buf = sBuf + (uint32_t)(chunk_it & 1) * 2048u;
        if (issuer) tma_store_wait_read<1>();                      // the store that last used this buffer has read it
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 4; ++g)
          ptx::st_shared_v4(buf + row_off + (((uint32_t)g ^ sw) << 4), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (issuer) {
          tma_store_2d(&tmy, buf, gcol, m0 + q * 32);
          tma_store_commit();
        }
      }
      if (tid == 64) GEMM_STAMP(4 * ti + 3);
    }
    if (issuer) tma_store_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCta2) cluster_sync_all();        // nobody leaves while the peer may still signal its barriers / read its TMEM
  if (warp == 1) {
    if constexpr (kCta2) tmem_dealloc_pair<C::kTmemCols>(tmem);
    else ptx::tmem_dealloc<C::kTmemCols>(tmem);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// 2-D map over a row-major (rows, cols) 16-bit matrix with row stride ld (elements): dim0 = cols (contiguous)
int make_map2d(CUtensorMap* tm, const void* base, int dtype, int64_t rows, int64_t cols, int64_t ld, int box_cols,
               int box_rows, CUtensorMapSwizzle swz, const char* what) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return set_error(OEH_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 != 0 || ld < cols)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_fwd: %s must be 16-byte aligned with a row stride that is a multiple of 8 "
                     "elements and >= the row length", what);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult res = enc(tm, dtype == OEH_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_fwd: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)res);
  return OEH_OK;
}

int ensure_smem(const void* fn, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  OEH_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, fn})) return OEH_OK;
  OEH_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({dev, fn});
  return OEH_OK;
}

template <bool kBF16, int BN, bool kCta2>
int launch(const OehLinearArgs* a, GemmParams& p, int grid_cap) {
  using C = Cfg<BN, kCta2>;
  CUtensorMap tma, tmb, tmy;
  int rc;
  if ((rc = make_map2d(&tma, a->x, a->dtype, a->M, a->K, a->ldx, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B, "x"))) return rc;
  if ((rc = make_map2d(&tmb, a->w, a->dtype, a->N, a->K, a->ldw, BK, kCta2 ? BN / 2 : BN, CU_TENSOR_MAP_SWIZZLE_128B, "w")))
    return rc;
  if ((rc = make_map2d(&tmy, a->y, a->dtype, a->M, a->N, a->ldy, kChunk, 32, CU_TENSOR_MAP_SWIZZLE_64B, "y"))) return rc;
  p.n_mblk = (a->M + BM - 1) / BM;
  p.n_nblk = (a->N + BN - 1) / BN;
  p.n_kblk = (a->K + BK - 1) / BK;
  auto fn = gemm_kernel<kBF16, BN, kCta2>;
  if ((rc = ensure_smem(reinterpret_cast<const void*>(fn), C::kSmemBytes))) return rc;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (kCta2) {
    const long long items = (long long)((p.n_mblk + 1) / 2) * p.n_nblk;
    const long long pairs = items < grid_cap / 2 ? items : grid_cap / 2;
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    const long long tiles = (long long)p.n_mblk * p.n_nblk;
    cfg.gridDim = dim3((unsigned)(tiles < grid_cap ? tiles : grid_cap));
  }
  cfg.blockDim = dim3(Epi<BN>::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = (cudaStream_t)a->stream;
  OEH_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, tma, tmb, tmy, p));
  return check_launch("gemm_kernel");
}

}  // namespace
}  // namespace oeh

using namespace oeh;

#ifdef OEH_TIMELINE
extern "C" int oeh_set_gemm_timeline(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_gemm_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int oeh_linear_fwd(const OehLinearArgs* a) {
  OEH_CHECK_ARG(a, "oeh_linear_fwd: null args");
  OEH_CHECK_ARG(a->M >= 0 && a->N >= 0 && a->K >= 0, "oeh_linear_fwd: negative dimension");
  if (a->M == 0 || a->N == 0) return OEH_OK;
  OEH_CHECK_ARG(a->K > 0, "oeh_linear_fwd: K == 0");
  OEH_CHECK_ARG(a->x && a->w && a->y, "oeh_linear_fwd: null tensor pointer");
  if (a->dtype != OEH_F16 && a->dtype != OEH_BF16)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_fwd: dtype %d (fp16 / bf16 operands, fp32 accumulate)", a->dtype);
  OEH_CHECK_ARG(a->act >= OEH_ACT_NONE && a->act <= OEH_ACT_GELU_ERF, "oeh_linear_fwd: unknown activation %d", a->act);
  OEH_CHECK_ARG(a->n_seg >= 0 && a->n_seg <= kMaxSeg, "oeh_linear_fwd: n_seg=%d outside [0,%d]", a->n_seg, kMaxSeg);
  if (a->n_seg > 0 && (a->n_bits < 1 || a->n_bits > 16))
    return set_error(OEH_ERR_BAD_ARG, "oeh_linear_fwd: n_bits=%d outside [1,16]", a->n_bits);
  if (a->bias && (reinterpret_cast<uintptr_t>(a->bias) & 15))
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_fwd: bias must be 16-byte aligned");
  if (a->N % 8 != 0) return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_fwd: N must be a multiple of 8");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias;
  p.act = a->act;
  p.int_max = (float)((1 << (a->n_bits > 0 ? a->n_bits : 8)) - 1);
  for (int i = 0; i < a->n_seg; ++i) {
    OEH_CHECK_ARG(a->qparams[i], "oeh_linear_fwd: null qparams[%d]", i);
    p.qp[i] = a->qparams[i];
  }
  int seg_cols = 0;
  if (a->n_seg > 1) {
    if (a->N % a->n_seg) return set_error(OEH_ERR_BAD_ARG, "oeh_linear_fwd: N=%d is not divisible into %d segments", a->N, a->n_seg);
    seg_cols = a->N / a->n_seg;
  }
  p.seg_cols = seg_cols;

  // Tile shape: fewest rounds x width over the SMs; a tile must not straddle two tap segments.  The CTA-pair form
  // (256 x BN per pair) moves a third less data per flop and is preferred whenever it fills the machine as well.
  // tile_n: 0 = choose; 128 / 256 = single-CTA tiles of that width; 2256 / 2128 = CTA pairs, 256 x 256 / 256 x 128.
  const int sms = num_sms();
  struct Shape { int bn; bool pair; };
  constexpr int kShapes = 6;
  const Shape shapes[kShapes] = {{256, true}, {128, true}, {256, false}, {128, false}, {192, true}, {192, false}};
  int best = -1;
  double best_cost = 1e30;
  for (int i = 0; i < kShapes; ++i) {
    const int bn = shapes[i].bn;
    if (bn == 192) continue;                                     // (chosen below, for the activation epilogues only)
    if (seg_cols && seg_cols % bn) continue;
    const long long mi = shapes[i].pair ? ((a->M + 2 * BM - 1) / (2 * BM)) : ((a->M + BM - 1) / BM);
    const long long items = mi * ((a->N + bn - 1) / bn);
    const long long slots = shapes[i].pair ? sms / 2 : sms;
    // time ~ rounds x (work per CTA per item) x (bytes-per-flop penalty of the single-CTA form)
    const double cost = (double)((items + slots - 1) / slots) * bn * (shapes[i].pair ? 1.0 : 1.3);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = i; }
  }
  p.dbg = a->tile_n / 10000;
  if (a->tile_n % 10000 != 0) {
    best = -1;
    for (int i = 0; i < kShapes; ++i)
      if (a->tile_n % 10000 == shapes[i].bn + (shapes[i].pair ? 2000 : 0)) best = i;
    if (best < 0) return set_error(OEH_ERR_BAD_ARG, "oeh_linear_fwd: tile_n=%d (0, 128, 192, 256, 2128, 2192, 2256)", a->tile_n);
    if (seg_cols && seg_cols % shapes[best].bn)
      return set_error(OEH_ERR_BAD_ARG, "oeh_linear_fwd: tile_n=%d does not divide the segment width %d", a->tile_n, seg_cols);
  }
  if (best < 0) return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_fwd: segment width %d is not a multiple of 128", seg_cols);
  const bool bf = a->dtype == OEH_BF16;
  switch (best) {
    case 0: return bf ? launch<true, 256, true>(a, p, sms) : launch<false, 256, true>(a, p, sms);
    case 1: return bf ? launch<true, 128, true>(a, p, sms) : launch<false, 128, true>(a, p, sms);
    case 2: return bf ? launch<true, 256, false>(a, p, sms) : launch<false, 256, false>(a, p, sms);
    case 3: return bf ? launch<true, 128, false>(a, p, sms) : launch<false, 128, false>(a, p, sms);
    case 4: return bf ? launch<true, 192, true>(a, p, sms) : launch<false, 192, true>(a, p, sms);
    default: return bf ? launch<true, 192, false>(a, p, sms) : launch<false, 192, false>(a, p, sms);
  }
}
