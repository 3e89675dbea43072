// Linear -> (+ residual) -> LayerNorm in ONE launch for sm_100a, with the INT8 path's fixed-range taps fused in:
//
//   v   = round_T(x . w^T + bias)                       the Linear's output as the unfused GEMM would have stored it
//   s   = [qdq_res]( [qdq_a](v) + residual )            sum rounded once to the tensor dtype; optionally stored (sum_out)
//   y   = [qdq_out]( LayerNorm(s) * gamma + beta )
//
// i.e. HF BertSelfOutput / BertOutput.forward (dense -> dropout -> LayerNorm(hidden + input)), their quantized wrappers
// (quantized_bert.py:541-606) and the two residual joints of the OPT decoder layer (quantized_opt.py:338-370), which the repo ran
// as `GEMM -> oeh_add_layer_norm`: two launches, the (M, N) Linear output written and read back, and a ~6-8 us latency-bound
// LayerNorm launch per joint (24 per model step, ~12 % of the fp16 BERT step, ~17 % of the INT8 OPT step).
//
// LayerNorm needs whole rows (N = 768 ... 2048) while one CTA's TMEM holds 512 fp32 columns, so a row block of 128 rows is computed
// by a CLUSTER of N / BN CTAs (BN = 192 or 256: 4 CTAs for N = 768 / 1024, 8 for N = 2048), each owning 128 rows x BN columns:
//   warp 0       TMA producer: A (128 x 64) / B (BN x 64) k-blocks through a kStages ring, plus the CTA's residual tile
//   warp 1       MMA issuer: tcgen05.mma kind::f16 M128 x N(BN) x K16 into one TMEM accumulator
//   warps 2..17  epilogue: thread == (row, quarter of the CTA's columns).  accumulator -> + bias -> round -> taps / residual add in
//                the tensor dtype (the arithmetic of layernorm.cu, vector for vector) -> fp32 row values in registers;
//                row statistics in two passes like layernorm.cu: partial sums are reduced inside the CTA through shared memory,
//                every CTA writes its row partials into ALL CTAs of the cluster (st.shared::cluster, distributed shared memory),
//                barrier.cluster, each thread adds the cluster's partials in a fixed order -> exact mean; the same again for
//                sum (x - mean)^2 -> rstd; y -> swizzled staging tile -> TMA store.
// One wave: (M / 128) x (N / BN) CTAs (BERT 4096 x 768: 128 CTAs), no persistence, no double-buffered accumulator.
//
// Parity: same per-element arithmetic as `library GEMM -> oeh_add_layer_norm`; the fp32 accumulation order of the GEMM and of
// the row statistics differs, so results agree to the tensor dtype's rounding (tests/test_gpu_gemm_ln.py).
#include <cstring>
#include <mutex>
#include <set>
#include <type_traits>
#include <utility>

#include "oeh_common.cuh"
#include "sm100_ptx.cuh"

namespace oeh {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int kStageA = BM * BK * 2;                       // 16 KB
constexpr int kEpiWarps = 16, kParts = 4, kThreads = (2 + kEpiWarps) * 32;
constexpr int kBoxBytes = BM * 128;                        // one [128 rows x 64 columns] 16-bit box, 128-byte swizzle
constexpr int kMaxCluster = 8;

template <int BN>
struct LCfg {
  static constexpr int kStageB = BN * BK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kBoxes = BN / 64;                    // residual / output boxes per CTA tile
  static constexpr int kTileBytes = kBoxes * kBoxBytes;     // 48 | 64 KB
  static constexpr int kStages = BN == 192 ? 4 : 3;         // ring >= one output tile (reused as the y staging tile)
  static constexpr int kSmemBytes = kStages * kStage + kTileBytes + 1024;
  static constexpr uint32_t kTmemCols = 256;
  static constexpr int kCols = BN / kParts;                 // columns per epilogue thread: 48 | 64
  static_assert(kStages * kStage >= kTileBytes, "the ring doubles as the output staging tile");
};

struct LnParams {
  int M, N, K, n_kblk;
  const void* bias;
  const void* gamma;
  const void* beta;
  const float* q_a;
  const float* q_res;
  const float* q_out;
  float int_max, eps;
  int has_res, has_sum;
};

struct Barriers {
  uint64_t full[4], empty[4];
  uint64_t res_full, tmem_full;
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
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

#ifdef OEH_TIMELINE
__device__ unsigned long long* g_ln_timeline = nullptr;   // experiment builds only (tools/timeline_gemm_ln.py): 8 stamps per CTA
__device__ __forceinline__ void ln_stamp(int k) {
  if (threadIdx.x == 64 && g_ln_timeline) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_ln_timeline[blockIdx.x * 8 + k] = t;
  }
}
#define LN_STAMP(k) ln_stamp(k)
#else
#define LN_STAMP(k)
#endif

template <typename T, int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb,
               const __grid_constant__ CUtensorMap tmr, const __grid_constant__ CUtensorMap tmy,
               const __grid_constant__ CUtensorMap tms, const __grid_constant__ LnParams p) {
  using C = LCfg<BN>;
  using V = Vec16<T>;
  constexpr bool kBF16 = std::is_same<T, __nv_bfloat16>::value;
  constexpr int kStages = C::kStages, kCols = C::kCols, kVecs = kCols / 8;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Barriers bars;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float2 part_red[kParts][BM];                  // (column part, row): (mean, M2) partials inside the CTA
  __shared__ float2 clus_red[kMaxCluster][BM];             // (source CTA, row): written by every CTA of the cluster into every one
  __shared__ uint4 s_par[3][BN / 8];                       // this CTA's slice of bias / gamma / beta

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sRes = smem0 + kStages * C::kStage;       // residual tile; the tapped sum is written back in place
  const uint32_t sY = smem0;                               // y staging tile: the ring, once the main loop is done
  const uint32_t n_cta = cluster_nctarank(), rank = cluster_ctarank();
  const int m0 = (int)(blockIdx.x / n_cta) * BM, n0 = (int)rank * BN;
  LN_STAMP(0);

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bars.full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars.empty[i]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&bars.res_full), 1);
    ptx::mbar_init(ptx::smem_u32(&bars.tmem_full), 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) ptx::tmem_alloc<C::kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // every CTA of the cluster is running before anything is sent to it
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  LN_STAMP(1);

  if (warp == 0) {
    // ================================ TMA producer ======================================================================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tma);
      ptx::prefetch_tmap(&tmb);
      if (p.has_res) {
        ptx::prefetch_tmap(&tmr);
        const uint32_t br = ptx::smem_u32(&bars.res_full);
        ptx::mbar_expect_tx(br, C::kTileBytes);
#pragma unroll
        for (int j = 0; j < C::kBoxes; ++j) tma_load_2d(sRes + j * kBoxBytes, &tmr, br, n0 + j * 64, m0);
      }
    }
    __syncwarp();
    for (int kb = 0; kb < p.n_kblk; ++kb) {
      const int s = kb % kStages, n = kb / kStages;
      if (n > 0) ptx::mbar_wait(ptx::smem_u32(&bars.empty[s]), (uint32_t)(n - 1) & 1u);
      if (ptx::elect_one()) {
        const uint32_t bf = ptx::smem_u32(&bars.full[s]);
        ptx::mbar_expect_tx(bf, C::kStage);
        tma_load_2d(smem0 + s * C::kStage, &tma, bf, kb * BK, m0);
        tma_load_2d(smem0 + s * C::kStage + kStageA, &tmb, bf, kb * BK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ========================================================================
    constexpr uint32_t kIdesc = ptx::make_idesc_f16(kBF16, BM, BN, false, false);
    for (int kb = 0; kb < p.n_kblk; ++kb) {
      const int s = kb % kStages;
      ptx::mbar_wait(ptx::smem_u32(&bars.full[s]), (uint32_t)(kb / kStages) & 1u);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t a = smem0 + s * C::kStage, b = a + kStageA;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          ptx::mma_ss(tmem, ptx::make_smem_desc_sw128(a + k * 32, 16, 1024), ptx::make_smem_desc_sw128(b + k * 32, 16, 1024), kIdesc,
                      (kb | k) != 0);
        ptx::mma_commit(ptx::smem_u32(&bars.empty[s]));
        if (kb + 1 == p.n_kblk) ptx::mma_commit(ptx::smem_u32(&bars.tmem_full));
      }
      __syncwarp();
    }
  }

  // ================================ epilogue (warps 2..17); warps 0 / 1 only keep the barriers company ========================
  const bool epi = warp >= 2;
  const int q = warp & 3, part = epi ? (warp - 2) >> 2 : 0;     // TMEM lane quarter; column part
  const int r = q * 32 + (tid & 31);                            // row within the tile == TMEM lane
  const uint32_t sw = (uint32_t)r & 7u, row_off = (uint32_t)r * 128u;
  // 16-byte chunk i (0 .. kVecs-1) of this thread: columns [part*kCols + 8i, +8) of the CTA tile
  auto chunk_addr = [&](uint32_t base, int i) {
    const int g = part * kVecs + i;
    return base + (uint32_t)(g >> 3) * kBoxBytes + row_off + ((((uint32_t)g & 7u) ^ sw) << 4);
  };

  if (epi) {
    // While the main loop runs: this CTA's slice of bias / gamma / beta -> shared memory.
    {
      const int e = tid - 64;                                   // 0 .. 511
      constexpr int kV = BN / 8;                                // 16-byte vectors per parameter
      if (e < 3 * kV) {
        const int which = e / kV, j = e % kV;
        const void* src = which == 0 ? p.bias : (which == 1 ? p.gamma : p.beta);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (src) v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(src) + n0) + j);
        s_par[which][j] = v;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    float as = 1.f, az = 0.f, rs = 1.f, rz = 0.f;
    if (p.q_a) { as = DT<T>::rnd(__ldg(p.q_a)); az = DT<T>::rnd(__ldg(p.q_a + 1)); }
    if (p.q_res) { rs = DT<T>::rnd(__ldg(p.q_res)); rz = DT<T>::rnd(__ldg(p.q_res + 1)); }
    const QdqVec<T> qa(as, az, p.int_max), qr(rs, rz, p.int_max);
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * kCols);

    ptx::mbar_wait(ptx::smem_u32(&bars.tmem_full), 0);
    ptx::tc_fence_after();
    LN_STAMP(2);
    if (p.has_res) ptx::mbar_wait(ptx::smem_u32(&bars.res_full), 0);

    // The three passes below are LOOPS over the thread's 16-byte chunks, not unrolled: the tap code is large (several exact
    // 16-bit formulations per tap), a CTA executes its epilogue exactly once, and unrolled over six chunks the kernel was 160 KB
    // of straight-line code whose instruction fetch dominated the epilogue (profiles/r02_timeline_gemm_ln.txt).  The row value
    // s is a 16-bit number, so it is parked in the residual tile in place instead of in 48 registers.
    float sum = 0.f;
#pragma unroll 1
    for (int i = 0; i < kVecs; ++i) {
      float acc[8];
      ptx::tmem_ld8(t_lane + i * 8, reinterpret_cast<uint32_t*>(acc));
      ptx::tmem_wait_ld();
      V va;
      const uint4 braw = s_par[0][part * kVecs + i];            // zeros without a bias: x + 0 rounds like x
      const V& bvec = *reinterpret_cast<const V*>(&braw);
#pragma unroll
      for (int e = 0; e < 8; ++e) va.v[e] = DT<T>::from_f(acc[e] + DT<T>::to_f(bvec.v[e]));
      if (p.q_a) qa.apply(va, false);
      const uint32_t addr = chunk_addr(sRes, i);
      if (p.has_res) {
        const uint4 raw = ld_shared_v4(addr);
        add_vec<T>(va, *reinterpret_cast<const V*>(&raw));
      }
      if (p.q_res) qr.apply(va, false);
      const uint4 o = *reinterpret_cast<const uint4*>(&va);
      ptx::st_shared_v4(addr, o.x, o.y, o.z, o.w);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += DT<T>::to_f(va.v[e]);
    }
    // Row statistics in ONE exchange: every thread takes the exact two-pass (mean, M2) of its kCols values; the partials of a row
    // are merged with Chan's formula (equal counts), in a fixed order, by every thread of the row alike.
    const float lmean = sum * (1.f / (float)kCols);
    float m2 = 0.f;
#pragma unroll 1
    for (int i = 0; i < kVecs; ++i) {
      const uint4 raw = ld_shared_v4(chunk_addr(sRes, i));
      const V& v = *reinterpret_cast<const V*>(&raw);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = DT<T>::to_f(v.v[e]) - lmean;
        m2 = fmaf(d, d, m2);
      }
    }
    // merge the four column parts inside the CTA (equal counts: mean = average of means, M2 = sum M2_i + n sum (mean_i - mean)^2),
    // then the CTA's (mean, M2) goes to every CTA of the cluster (distributed shared memory)
    part_red[part][r] = make_float2(lmean, m2);
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (part == 0) {
      float ms = 0.f;
#pragma unroll
      for (int i = 0; i < kParts; ++i) ms += part_red[i][r].x;
      const float cmean = ms * (1.f / (float)kParts);
      float cm2 = 0.f, dev = 0.f;
#pragma unroll
      for (int i = 0; i < kParts; ++i) {
        const float2 pr = part_red[i][r];
        const float d = pr.x - cmean;
        cm2 += pr.y;
        dev = fmaf(d, d, dev);
      }
      cm2 = fmaf((float)kCols, dev, cm2);
      const uint32_t local = ptx::smem_u32(&clus_red[rank][r]);
      for (uint32_t c = 0; c < n_cta; ++c) {
        const uint32_t remote = map_to_cta(local, c);
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(cmean), "f"(cm2) : "memory");
      }
    }
    LN_STAMP(3);
  }
  cluster_sync_all();
  if (epi) {
    // merge the cluster's partials (equal counts BN), in a fixed order, by every thread of the row alike
    float msum = 0.f;
    for (uint32_t i = 0; i < n_cta; ++i) msum += clus_red[i][r].x;
    const float mean = msum / (float)n_cta;
    float m2 = 0.f, dev = 0.f;
    for (uint32_t i = 0; i < n_cta; ++i) {
      const float2 pr = clus_red[i][r];
      const float d = pr.x - mean;
      m2 += pr.y;
      dev = fmaf(d, d, dev);
    }
    const float rstd = rsqrtf((m2 + (float)BN * dev) / (float)p.N + p.eps);
    LN_STAMP(4);

    float os = 1.f, oz = 0.f;
    if (p.q_out) { os = DT<T>::rnd(__ldg(p.q_out)); oz = DT<T>::rnd(__ldg(p.q_out + 1)); }
    const QdqVec<T> qo(os, oz, p.int_max);
#pragma unroll 1
    for (int i = 0; i < kVecs; ++i) {
      const uint4 graw = s_par[1][part * kVecs + i], beraw = s_par[2][part * kVecs + i];
      const V& g = *reinterpret_cast<const V*>(&graw);
      const V& be = *reinterpret_cast<const V*>(&beraw);
      const uint4 raw = ld_shared_v4(chunk_addr(sRes, i));
      const V& xv = *reinterpret_cast<const V*>(&raw);
      V out;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = DT<T>::to_f(g.v[e]) * (rstd * (DT<T>::to_f(xv.v[e]) - mean));
        if (p.beta) v += DT<T>::to_f(be.v[e]);
        out.v[e] = DT<T>::from_f(v);                            // one rounding to the tensor dtype
      }
      if (p.q_out) qo.apply(out, false);
      const uint4 o = *reinterpret_cast<const uint4*>(&out);
      ptx::st_shared_v4(chunk_addr(sY, i), o.x, o.y, o.z, o.w);
    }
    ptx::fence_proxy_async_smem();
    LN_STAMP(5);
  }
  __syncthreads();
  if (warp == 2 && ptx::elect_one()) {
#pragma unroll
    for (int j = 0; j < C::kBoxes; ++j) tma_store_2d(&tmy, sY + j * kBoxBytes, n0 + j * 64, m0);
    if (p.has_sum) {
#pragma unroll
      for (int j = 0; j < C::kBoxes; ++j) tma_store_2d(&tms, sRes + j * kBoxBytes, n0 + j * 64, m0);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory has been read; the writes drain on their own
  }
  LN_STAMP(6);
  // No cluster barrier at the end: the only remote accesses are the statistics written BEFORE the barrier above.
  ptx::tc_fence_before();
  __syncthreads();
  LN_STAMP(7);
  if (warp == 1) ptx::tmem_dealloc<C::kTmemCols>(tmem);
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

// 2-D map over a row-major (rows, cols) 16-bit matrix with row stride ld (elements), boxes of 64 columns, 128-byte swizzle
int make_map(CUtensorMap* tm, const void* base, int dtype, int64_t rows, int64_t cols, int64_t ld, int box_rows, const char* what) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return set_error(OEH_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 != 0 || ld < cols)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_add_layer_norm: %s must be 16-byte aligned with a row stride that is a "
                     "multiple of 8 elements and >= the row length", what);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult res = enc(tm, dtype == OEH_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_add_layer_norm: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)res);
  return OEH_OK;
}

int ensure_attrs(const void* fn, int bytes, bool big_cluster) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  OEH_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, fn})) return OEH_OK;
  OEH_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  (void)big_cluster;
  done.insert({dev, fn});
  return OEH_OK;
}

template <typename T, int BN>
int launch(const OehLinearLnArgs* a, LnParams& p, int n_cta) {
  using C = LCfg<BN>;
  CUtensorMap tma, tmb, tmr, tmy, tms;
  int rc;
  if ((rc = make_map(&tma, a->x, a->dtype, a->M, a->K, a->ldx, BM, "x"))) return rc;
  if ((rc = make_map(&tmb, a->w, a->dtype, a->N, a->K, a->ldw, BN, "w"))) return rc;
  if ((rc = make_map(&tmy, a->y, a->dtype, a->M, a->N, a->N, BM, "y"))) return rc;
  tmr = tmy;
  tms = tmy;
  if (a->residual && (rc = make_map(&tmr, a->residual, a->dtype, a->M, a->N, a->N, BM, "residual"))) return rc;
  if (a->sum_out && (rc = make_map(&tms, a->sum_out, a->dtype, a->M, a->N, a->N, BM, "sum_out"))) return rc;
  auto fn = gemm_ln_kernel<T, BN>;
  if ((rc = ensure_attrs(reinterpret_cast<const void*>(fn), C::kSmemBytes, n_cta > 8))) return rc;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)n_cta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)(n_cta * ((a->M + BM - 1) / BM)));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = (cudaStream_t)a->stream;
  OEH_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, tma, tmb, tmr, tmy, tms, p));
  return check_launch("gemm_ln_kernel");
}

}  // namespace
}  // namespace oeh

using namespace oeh;

// N this entry point can normalise in one cluster: 192 c or 256 c columns with c <= 8 CTAs
static int pick_bn(int N, int* n_cta) {
  if (N % 192 == 0 && N / 192 <= kMaxCluster) { *n_cta = N / 192; return 192; }
  if (N % 256 == 0 && N / 256 <= kMaxCluster) { *n_cta = N / 256; return 256; }
  return 0;
}

#ifdef OEH_TIMELINE
extern "C" int oeh_set_ln_timeline(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_ln_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int oeh_linear_add_layer_norm_supported(int32_t N) {
  int c;
  return pick_bn(N, &c) != 0;
}

extern "C" int oeh_linear_add_layer_norm(const OehLinearLnArgs* a) {
  OEH_CHECK_ARG(a, "oeh_linear_add_layer_norm: null args");
  OEH_CHECK_ARG(a->M >= 0 && a->N > 0 && a->K > 0, "oeh_linear_add_layer_norm: bad dimension");
  if (a->M == 0) return OEH_OK;
  OEH_CHECK_ARG(a->x && a->w && a->y && a->ln_weight, "oeh_linear_add_layer_norm: null tensor pointer");
  if (a->dtype != OEH_F16 && a->dtype != OEH_BF16)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_add_layer_norm: dtype %d (fp16 / bf16 operands, fp32 accumulate)", a->dtype);
  if ((a->q_a || a->q_res || a->q_out) && (a->n_bits < 1 || a->n_bits > 16))
    return set_error(OEH_ERR_BAD_ARG, "oeh_linear_add_layer_norm: n_bits=%d outside [1,16]", a->n_bits);
  for (const void* ptr : {a->bias, a->ln_weight, a->ln_bias})
    if (ptr && (reinterpret_cast<uintptr_t>(ptr) & 15))
      return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_add_layer_norm: bias / LayerNorm parameters must be 16-byte aligned");
  OEH_CHECK_ARG(!a->sum_out || a->residual, "oeh_linear_add_layer_norm: sum_out without a residual");
  int n_cta = 0;
  const int bn = pick_bn(a->N, &n_cta);
  if (!bn)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_linear_add_layer_norm: N=%d (192 c or 256 c columns, c <= %d)", a->N, kMaxCluster);
  LnParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.n_kblk = (a->K + BK - 1) / BK;
  p.bias = a->bias;
  p.gamma = a->ln_weight;
  p.beta = a->ln_bias;
  p.q_a = a->q_a; p.q_res = a->q_res; p.q_out = a->q_out;
  p.int_max = (float)((1 << (a->n_bits > 0 ? a->n_bits : 8)) - 1);
  p.eps = (float)a->eps;
  p.has_res = a->residual != nullptr;
  p.has_sum = a->sum_out != nullptr;
  const bool bf = a->dtype == OEH_BF16;
  if (bn == 192) return bf ? launch<__nv_bfloat16, 192>(a, p, n_cta) : launch<__half, 192>(a, p, n_cta);
  return bf ? launch<__nv_bfloat16, 256>(a, p, n_cta) : launch<__half, 256>(a, p, n_cta);
}
