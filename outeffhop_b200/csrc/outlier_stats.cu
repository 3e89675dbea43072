// Outlier metrics of an activation tensor, one pass, on the device (sm_100a) -- SURVEY section 8f rank 3.
//   inf_norm[r] = max_j |x[r,j]|                      x.norm(dim=1, p=inf)      validate_mlm_config.py:513-518
//   kurt[r]     = mean((x-mu)^4) / (std^4 + eps)      kurtosis()                transformers_language/utils.py:9-20
// per sample r of x viewed as (rows, cols) = (batch, T*hidden).  The reference reads every per-sample value back
// with .item() (32 x 36 host syncs per BERT batch); here one launch pair writes all rows and can fold them into a
// device-resident AverageMeter (sum, sum, count), so an evaluation needs a single read-back at the end.
//
// HBM-bound: bytes = rows*cols*size_in, read once with 16-byte streaming loads.  Every row is cut into chunks so that
// the grid fills the machine whatever the batch size; a thread keeps fp32 power sums of at most 4 vectors, folds
// them into float64, and the (row, chunk) partials are combined in a fixed order by stage 2 (deterministic).
// Central moments come from the float64 raw power sums:  m2 = S2/n - mu^2, m4 = S4/n - 4 mu S3/n + 6 mu^2 S2/n - 3 mu^4.
#include "oeh_common.cuh"

namespace oeh {

namespace {

constexpr int kThreads = 256;

struct Part {
  double s1, s2, s3, s4;
  float amax;
  int nan;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) outlier_stage1_kernel(const T* __restrict__ x, int64_t cols, int chunks,
                                                                  int64_t chunk_elems, int aligned,
                                                                  Part* __restrict__ parts) {
  using V = Vec16<T>;
  constexpr int N = V::N;
  const int row = blockIdx.y, chunk = blockIdx.x;
  const T* xr = x + (int64_t)row * cols;
  const int64_t c0 = (int64_t)chunk * chunk_elems;
  const int64_t c1 = c0 + chunk_elems < cols ? c0 + chunk_elems : cols;
  double d1 = 0, d2 = 0, d3 = 0, d4 = 0;
  float amax = 0.f;
  int nan = 0;
  auto fold = [&](float f1, float f2, float f3, float f4) {
    d1 += (double)f1;
    d2 += (double)f2;
    d3 += (double)f3;
    d4 += (double)f4;
  };
  auto acc = [&](float v, float& f1, float& f2, float& f3, float& f4) {
    nan |= (v != v);
    amax = fmaxf(amax, fabsf(v));
    const float v2 = v * v;
    f1 += v;
    f2 += v2;
    f3 = fmaf(v2, v, f3);
    f4 = fmaf(v2, v2, f4);
  };
  int64_t lo = c0, hi = c1;                              // [lo, hi): the part read with 16-byte loads
  if (aligned) {                                         // chunk_elems and cols are multiples of N when aligned
    const V* xv = reinterpret_cast<const V*>(xr + c0);
    const int64_t nvec = (c1 - c0) / N;
    int64_t i = threadIdx.x;
    for (; i + 3 * kThreads < nvec; i += 4 * kThreads) {
      V in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int4 raw = __ldcs(reinterpret_cast<const int4*>(xv + i + u * kThreads));
        in[u] = *reinterpret_cast<V*>(&raw);
      }
      float f1 = 0, f2 = 0, f3 = 0, f4 = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < N; ++e) acc(DT<T>::to_f(in[u].v[e]), f1, f2, f3, f4);
      fold(f1, f2, f3, f4);
    }
    for (; i < nvec; i += kThreads) {
      V in = xv[i];
      float f1 = 0, f2 = 0, f3 = 0, f4 = 0;
#pragma unroll
      for (int e = 0; e < N; ++e) acc(DT<T>::to_f(in.v[e]), f1, f2, f3, f4);
      fold(f1, f2, f3, f4);
    }
    lo = c0 + nvec * N;
  }
  for (int64_t t = lo + threadIdx.x; t < hi; t += kThreads) {
    float f1 = 0, f2 = 0, f3 = 0, f4 = 0;
    acc(DT<T>::to_f(xr[t]), f1, f2, f3, f4);
    fold(f1, f2, f3, f4);
  }
  // block reduction (fixed tree: deterministic)
  __shared__ double sh[4][kThreads / 32];
  __shared__ float sh_max[kThreads / 32];
  __shared__ int sh_nan[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    d3 += __shfl_xor_sync(0xffffffffu, d3, o);
    d4 += __shfl_xor_sync(0xffffffffu, d4, o);
  }
  amax = warp_max(amax);
  nan = __any_sync(0xffffffffu, nan);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh[0][w] = d1;
    sh[1][w] = d2;
    sh[2][w] = d3;
    sh[3][w] = d4;
    sh_max[w] = amax;
    sh_nan[w] = nan;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Part p{0, 0, 0, 0, 0.f, 0};
    for (int i = 0; i < kThreads / 32; ++i) {
      p.s1 += sh[0][i];
      p.s2 += sh[1][i];
      p.s3 += sh[2][i];
      p.s4 += sh[3][i];
      p.amax = fmaxf(p.amax, sh_max[i]);
      p.nan |= sh_nan[i];
    }
    parts[(int64_t)row * chunks + chunk] = p;
  }
}

// stage 2: one thread per row folds its chunks in order; thread 0 then updates the meter over rows in order
__global__ void outlier_stage2_kernel(const Part* __restrict__ parts, int rows, int chunks, int64_t cols, double eps,
                                      float* __restrict__ out_inf, float* __restrict__ out_kurt,
                                      double* __restrict__ meter) {
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    Part p{0, 0, 0, 0, 0.f, 0};
    for (int c = 0; c < chunks; ++c) {
      const Part q = parts[(int64_t)r * chunks + c];
      p.s1 += q.s1;
      p.s2 += q.s2;
      p.s3 += q.s3;
      p.s4 += q.s4;
      p.amax = fmaxf(p.amax, q.amax);
      p.nan |= q.nan;
    }
    const double n = (double)cols;
    const double mu = p.s1 / n;
    const double e2 = p.s2 / n, e3 = p.s3 / n, e4 = p.s4 / n;
    const double m2 = e2 - mu * mu;                                   // biased variance
    const double m4 = e4 - 4.0 * mu * e3 + 6.0 * mu * mu * e2 - 3.0 * mu * mu * mu * mu;
    const double var_unbiased = cols > 1 ? m2 * n / (n - 1.0) : __longlong_as_double(0x7ff8000000000000LL);   // x.std(): N-1
    const double kurt = m4 / (var_unbiased * var_unbiased + eps);     // s**4 + eps
    const float qnan = __int_as_float(0x7fc00000);
    out_inf[r] = p.nan ? qnan : p.amax;
    out_kurt[r] = p.nan ? qnan : (float)kurt;
  }
  __syncthreads();
  if (meter && threadIdx.x == 0) {
    double si = 0, sk = 0;
    for (int r = 0; r < rows; ++r) {
      si += (double)out_inf[r];
      sk += (double)out_kurt[r];
    }
    meter[0] += si;
    meter[1] += sk;
    meter[2] += (double)rows;
  }
}

// chunks per row: enough (row, chunk) blocks for >= 2 waves of 8 blocks per SM, each chunk >= 4096 elements and a
// multiple of 16 elements (so that aligned rows keep aligned chunks)
void plan(int64_t rows, int64_t cols, int* chunks, int64_t* chunk_elems) {
  const int64_t target_blocks = (int64_t)num_sms() * 16;
  int64_t c = (target_blocks + rows - 1) / rows;
  const int64_t max_c = (cols + 4095) / 4096;
  if (c > max_c) c = max_c;
  if (c < 1) c = 1;
  int64_t ce = (cols + c - 1) / c;
  ce = (ce + 15) / 16 * 16;
  c = (cols + ce - 1) / ce;
  *chunks = (int)c;
  *chunk_elems = ce;
}

template <typename T>
int launch(const void* x, int64_t rows, int64_t cols, double eps, float* out_inf, float* out_kurt, double* meter,
           Part* parts, cudaStream_t st) {
  int chunks;
  int64_t ce;
  plan(rows, cols, &chunks, &ce);
  const int aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (cols * (int64_t)sizeof(T)) % 16 == 0;
  outlier_stage1_kernel<T><<<dim3(chunks, (unsigned)rows), kThreads, 0, st>>>((const T*)x, cols, chunks, ce, aligned, parts);
  int rc = check_launch("outlier_stage1_kernel");
  if (rc) return rc;
  outlier_stage2_kernel<<<1, 256, 0, st>>>(parts, (int)rows, chunks, cols, eps, out_inf, out_kurt, meter);
  return check_launch("outlier_stage2_kernel");
}

}  // namespace
}  // namespace oeh

using namespace oeh;

extern "C" size_t oeh_outlier_stats_workspace_bytes(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  int chunks;
  int64_t ce;
  plan(rows, cols, &chunks, &ce);
  return (size_t)rows * chunks * sizeof(Part);
}

extern "C" int oeh_outlier_stats(const void* x, int64_t rows, int64_t cols, int32_t dtype, double eps, float* out_inf,
                                 float* out_kurt, double* meter, void* workspace, size_t workspace_bytes,
                                 oeh_stream_t stream) {
  OEH_CHECK_ARG(rows > 0 && cols > 0, "oeh_outlier_stats: empty tensor");
  OEH_CHECK_ARG(rows <= 65535, "oeh_outlier_stats: at most 65535 rows (samples) per call");
  OEH_CHECK_ARG(x && out_inf && out_kurt, "oeh_outlier_stats: null pointer");
  const size_t need = oeh_outlier_stats_workspace_bytes(rows, cols);
  if (!workspace || workspace_bytes < need)
    return set_error(OEH_ERR_WORKSPACE, "oeh_outlier_stats: workspace of %zu bytes needed", need);
  OEH_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "oeh_outlier_stats: workspace must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  Part* parts = (Part*)workspace;
  switch (dtype) {
    case OEH_F32: return launch<float>(x, rows, cols, eps, out_inf, out_kurt, meter, parts, st);
    case OEH_F16: return launch<__half>(x, rows, cols, eps, out_inf, out_kurt, meter, parts, st);
    case OEH_BF16: return launch<__nv_bfloat16>(x, rows, cols, eps, out_inf, out_kurt, meter, parts, st);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_outlier_stats: unknown dtype %d", dtype);
}
