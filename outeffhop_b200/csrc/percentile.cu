// Exact percentile range estimation on the device (sm_100a), SURVEY section 8f rank 1:
//   np.percentile(x, (q_lo, q_hi)) with the default 'linear' method on the WHOLE activation tensor, which the
//   reference evaluates on a host copy (quantization/range_estimators.py:89-94 RunningMinMaxEstimator,
//   :62-69 CurrentMinMaxEstimator; every OPT INT8 command passes --percentile 99.999).
//
// numpy's algorithm (numpy/lib/_function_base_impl.py `_quantile`, method 'linear'; numpy is an un-vendored
// dependency of the reference, pinned numpy==1.24.1 in requirement.txt:23, 2.3 in this image -- same arithmetic):
//   virtual = (n-1) * (q/100)            [float64]
//   prev = floor(virtual), next = prev+1 [clamped to the last element when virtual >= n-1]
//   a = sorted[prev], b = sorted[next]   [order statistics, in the tensor's dtype]
//   t = virtual - prev                   [float64]
//   d = b - a                            [ROUNDED TO THE TENSOR'S DTYPE]
//   r = t >= 0.5 ? b - d*(1-t) : a + d*t [float64]           any NaN in x -> NaN
// oeh_percentile_plan does the index arithmetic on the host (it depends on n and q only); the device finds the four
// order statistics EXACTLY by most-significant-digit radix select on order-preserving integer keys:
//   pass p: histogram of digit p of every element whose higher digits equal the selected prefix (one shared-memory
//           histogram per DISTINCT prefix among the four ranks, flushed with integer atomics -> deterministic),
//           then a one-block scan picks each rank's bin and the residual rank inside it.
//   fp32: digits of 11 + 11 + 10 bits (3 reads of x); fp16: 11 + 5 bits (2 reads).  HBM-bound: bytes = passes*n*size.
// Between a histogram and its scan the caller may all-reduce (SUM) the histogram over ranks: the select then runs on
// the union of all ranks' tensors, bit-identical to a single process holding the concatenated batch
// (outeffhop_b200/ops.py::percentile).
#include "oeh_common.cuh"

namespace oeh {

namespace {

constexpr int kBins = 2048;
constexpr int kRanks = 4;
constexpr int kHistThreads = 512;
constexpr int kScanThreads = 1024;

struct Ctl {
  unsigned long long krem[kRanks];   // rank still to descend inside the selected prefix
  uint32_t prefix[kRanks];           // selected higher digits
  uint32_t group[kRanks];            // first rank with the same prefix (histograms are shared per group)
  uint32_t has_nan;
  uint32_t pad;
};

struct Ws {
  uint32_t hist[kRanks][kBins];      // first: the caller may all-reduce it as a (4*2048,) int32 tensor
  Ctl ctl;
};

__host__ __device__ inline int n_passes(int dtype) { return dtype == OEH_F32 ? 3 : 2; }
__host__ __device__ inline int pass_shift(int dtype, int pass) {
  return dtype == OEH_F32 ? (pass == 0 ? 21 : (pass == 1 ? 10 : 0)) : (pass == 0 ? 5 : 0);
}
__host__ __device__ inline int pass_bits(int dtype, int pass) {
  return dtype == OEH_F32 ? (pass == 2 ? 10 : 11) : (pass == 0 ? 11 : 5);
}

// order-preserving keys; every NaN sorts last (numpy sorts NaN to the end), -0.0 == +0.0
__device__ __forceinline__ uint32_t key_of(float v, uint32_t& nan) {
  uint32_t u = __float_as_uint(v);
  if ((u & 0x7fffffffu) > 0x7f800000u) {
    nan = 1u;
    return 0xffffffffu;
  }
  if (u == 0x80000000u) u = 0u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint32_t key_of(__half v, uint32_t& nan) {
  uint32_t u = __half_as_ushort(v);
  if ((u & 0x7fffu) > 0x7c00u) {
    nan = 1u;
    return 0xffffu;
  }
  if (u == 0x8000u) u = 0u;
  return (u & 0x8000u) ? (~u & 0xffffu) : (u | 0x8000u);
}
__device__ __forceinline__ float fp32_of_key(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ __half fp16_of_key(uint32_t k) {
  return __ushort_as_half((unsigned short)((k & 0x8000u) ? (k ^ 0x8000u) : (~k & 0xffffu)));
}

__global__ void begin_kernel(Ws* ws, long long k0, long long k1, long long k2, long long k3) {
  for (int i = threadIdx.x; i < kRanks * kBins; i += blockDim.x) (&ws->hist[0][0])[i] = 0u;
  if (threadIdx.x == 0) {
    ws->ctl.krem[0] = (unsigned long long)k0;
    ws->ctl.krem[1] = (unsigned long long)k1;
    ws->ctl.krem[2] = (unsigned long long)k2;
    ws->ctl.krem[3] = (unsigned long long)k3;
    for (int r = 0; r < kRanks; ++r) {
      ws->ctl.prefix[r] = 0u;
      ws->ctl.group[r] = 0u;           // pass 0: no prefix yet, every rank reads histogram 0
    }
    ws->ctl.has_nan = 0u;
  }
}

template <typename T, bool kFirst>
__global__ void __launch_bounds__(kHistThreads) hist_kernel(const T* __restrict__ x, int64_t n, int shift, int bits,
                                                            Ws* __restrict__ ws) {
  // kFirst: all elements count, the four copies are sub-histograms by warp (less same-address contention)
  __shared__ uint32_t sh[kRanks][kBins];
  for (int i = threadIdx.x; i < kRanks * kBins; i += kHistThreads) (&sh[0][0])[i] = 0u;
  uint32_t prefix[kRanks], lead[kRanks];
#pragma unroll
  for (int r = 0; r < kRanks; ++r) {
    prefix[r] = kFirst ? 0u : ws->ctl.prefix[r];
    lead[r] = kFirst ? (r == 0) : (ws->ctl.group[r] == (uint32_t)r);
  }
  __syncthreads();
  const uint32_t mask = (1u << bits) - 1u;
  const int hi_shift = shift + bits;
  const int sub = (threadIdx.x >> 5) & (kRanks - 1);
  uint32_t nan = 0u;
  auto count = [&](T v) {
    const uint32_t key = key_of(v, nan);
    const uint32_t digit = (key >> shift) & mask;
    if (kFirst) {
      atomicAdd(&sh[sub][digit], 1u);
    } else {
      const uint32_t hi = key >> hi_shift;
#pragma unroll
      for (int r = 0; r < kRanks; ++r)
        if (lead[r] && hi == prefix[r]) atomicAdd(&sh[r][digit], 1u);
    }
  };
  using V = Vec16<T>;
  constexpr int N = V::N;
  // 16-byte loads from the first aligned element on; the ragged head / tail is read element-wise by block 0
  const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
  int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
  if (head > n) head = n;
  const int64_t nvec = (n - head) / N;
  const V* xv = reinterpret_cast<const V*>(x + head);
  const int64_t stride = (int64_t)gridDim.x * kHistThreads;
  int64_t i = (int64_t)blockIdx.x * kHistThreads + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    V in[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int4 raw = __ldcs(reinterpret_cast<const int4*>(xv + i + u * stride));
      in[u] = *reinterpret_cast<V*>(&raw);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int e = 0; e < N; ++e) count(in[u].v[e]);
  }
  for (; i < nvec; i += stride) {
    V in = xv[i];
#pragma unroll
    for (int e = 0; e < N; ++e) count(in.v[e]);
  }
  if (blockIdx.x == 0) {
    for (int64_t t = threadIdx.x; t < head; t += kHistThreads) count(x[t]);
    for (int64_t t = head + nvec * N + threadIdx.x; t < n; t += kHistThreads) count(x[t]);
  }
  __syncthreads();
  if (kFirst) {
    for (int b = threadIdx.x; b < kBins; b += kHistThreads) {
      const uint32_t c = sh[0][b] + sh[1][b] + sh[2][b] + sh[3][b];
      if (c) atomicAdd(&ws->hist[0][b], c);
    }
    // NaN blocks are counted in hist[1][0] (unused in pass 0), so a cross-rank SUM of the histogram carries the flag too
    if (__syncthreads_or((int)nan) && threadIdx.x == 0) atomicAdd(&ws->hist[1][0], 1u);
  } else {
    for (int i2 = threadIdx.x; i2 < kRanks * kBins; i2 += kHistThreads) {
      const uint32_t c = (&sh[0][0])[i2];
      if (c) atomicAdd(&(&ws->hist[0][0])[i2], c);
    }
  }
}

// One block: for every rank find the bin holding it, descend, then prepare the next pass.
__global__ void __launch_bounds__(kScanThreads) scan_kernel(Ws* __restrict__ ws, int bits, int first) {
  __shared__ unsigned long long warp_tot[kScanThreads / 32];
  __shared__ uint32_t sel_bin[kRanks];
  __shared__ unsigned long long sel_below[kRanks];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = 1 << bits;
  const uint32_t nan_blocks = first ? ws->hist[1][0] : 0u;
  for (int r = 0; r < kRanks; ++r) {
    const uint32_t g = ws->ctl.group[r];
    if (g != (uint32_t)r) continue;                      // uniform branch: resolved below from its group leader
    // each thread owns bins 2*tid, 2*tid+1
    const uint32_t c0 = 2 * tid < bins ? ws->hist[r][2 * tid] : 0u;
    const uint32_t c1 = 2 * tid + 1 < bins ? ws->hist[r][2 * tid + 1] : 0u;
    unsigned long long v = (unsigned long long)c0 + c1, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    const unsigned long long excl = base + incl - v;     // elements in bins below 2*tid
    // every rank of this group (same prefix, hence the same histogram) picks its bin
    for (int r2 = r; r2 < kRanks; ++r2) {
      if (ws->ctl.group[r2] != (uint32_t)r) continue;
      const unsigned long long k = ws->ctl.krem[r2];
      if (k >= excl && k < excl + c0) {
        sel_bin[r2] = 2 * tid;
        sel_below[r2] = excl;
      } else if (k >= excl + c0 && k < excl + c0 + c1) {
        sel_bin[r2] = 2 * tid + 1;
        sel_below[r2] = excl + c0;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) {
    if (nan_blocks) ws->ctl.has_nan = 1u;
    for (int r = 0; r < kRanks; ++r) {
      ws->ctl.prefix[r] = (ws->ctl.prefix[r] << bits) | sel_bin[r];
      ws->ctl.krem[r] -= sel_below[r];
    }
    for (int r = 0; r < kRanks; ++r) {
      uint32_t g = r;
      for (int q = r - 1; q >= 0; --q)
        if (ws->ctl.prefix[q] == ws->ctl.prefix[r]) g = q;
      ws->ctl.group[r] = g;
    }
  }
  __syncthreads();
  for (int i = tid; i < kRanks * kBins; i += kScanThreads) (&ws->hist[0][0])[i] = 0u;
}

// after the last pass ctl.prefix[r] is the full key of order statistic r: (a_lo, b_lo, a_hi, b_hi)
template <typename T>
__global__ void finish_kernel(const Ws* __restrict__ ws, double t_lo, double t_hi, double* __restrict__ out2) {
  if (threadIdx.x >= 2) return;
  const int j = threadIdx.x;
  const double t = j == 0 ? t_lo : t_hi;
  const uint32_t ka = ws->ctl.prefix[2 * j], kb = ws->ctl.prefix[2 * j + 1];
  double a, b, d;
  if constexpr (sizeof(T) == 4) {
    const float fa = fp32_of_key(ka), fb = fp32_of_key(kb);
    a = (double)fa;
    b = (double)fb;
    d = (double)__fsub_rn(fb, fa);                       // numpy: subtract(b, a) in the array dtype
  } else {
    const __half ha = fp16_of_key(ka), hb = fp16_of_key(kb);
    a = (double)__half2float(ha);
    b = (double)__half2float(hb);
    d = (double)__half2float(__float2half_rn(__fsub_rn(__half2float(hb), __half2float(ha))));   // npy_half: via float
  }
  double r = t >= 0.5 ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(a, __dmul_rn(d, t));   // _lerp
  if (ws->ctl.has_nan) r = __longlong_as_double(0x7ff8000000000000LL);
  out2[j] = r;
}

// EMA of the float64 percentile range + set_quant_range in float64, mirrored into the fp32 kernel state.
__global__ void range_update_f64_kernel(double* __restrict__ st64, float* __restrict__ st32, const double* __restrict__ new2,
                                        double one_minus_m, double m, double int_max, double eps) {
  if (threadIdx.x != 0) return;
  double lo = new2[0], hi = new2[1];
  if (st64[4] != 0.0) {                                  // range_estimators.py:103-104, python-float coefficients
    lo = __dadd_rn(__dmul_rn(one_minus_m, lo), __dmul_rn(m, st64[0]));
    hi = __dadd_rn(__dmul_rn(one_minus_m, hi), __dmul_rn(m, st64[1]));
  }
  st64[0] = lo;
  st64[1] = hi;
  st64[4] = 1.0;
  // uniform_quantizers.py:199-200, 218-219 on float64 tensors (torch.min/max propagate NaN)
  const double x_min = lo < 0.0 ? lo : (lo != lo ? lo : 0.0);
  const double x_max = hi > eps ? hi : (hi != hi ? hi : eps);
  const double delta = __ddiv_rn(__dsub_rn(x_max, x_min), int_max);
  const double zf = __ddiv_rn(-x_min, delta);
  st64[2] = delta;
  st64[3] = zf;
  double scale = delta > eps ? delta : (delta != delta ? delta : eps);   // :72-76
  double zp = rint(zf);                                                   // :79-82
  zp = zp < 0.0 ? 0.0 : (zp > int_max ? int_max : zp);
  // the quantize kernels use fp32 scalars: ATen casts the 0-dim float64 scale / zero_point to the tensor dtype at use
  st32[OEH_QS_CUR_MIN] = (float)lo;
  st32[OEH_QS_CUR_MAX] = (float)hi;
  st32[OEH_QS_DELTA] = (float)delta;
  st32[OEH_QS_ZERO_FLOAT] = (float)zf;
  st32[OEH_QS_SCALE] = (float)scale;
  st32[OEH_QS_ZERO_POINT] = (float)zp;
  st32[OEH_QS_INITIALIZED] = 1.f;
}

int hist_grid(int64_t n, size_t esize) {
  const int64_t nvec = n * (int64_t)esize / 16 + 1;
  int64_t blocks = (nvec + kHistThreads * 4 - 1) / (kHistThreads * 4);
  const int64_t cap = (int64_t)num_sms() * 4;            // 4 x 512 threads per SM: a whole number of waves
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace oeh

using namespace oeh;

extern "C" size_t oeh_percentile_workspace_bytes(void) { return sizeof(Ws); }

extern "C" int oeh_percentile_passes(int32_t dtype) { return (dtype == OEH_F32 || dtype == OEH_F16) ? n_passes(dtype) : 0; }

extern "C" int oeh_percentile_plan(int64_t n_total, double q_lo, double q_hi, int64_t ranks_out[4], double gamma_out[2]) {
  OEH_CHECK_ARG(n_total > 0, "oeh_percentile_plan: empty tensor");
  OEH_CHECK_ARG(ranks_out && gamma_out, "oeh_percentile_plan: null output");
  OEH_CHECK_ARG(q_lo >= 0.0 && q_lo <= 100.0 && q_hi >= 0.0 && q_hi <= 100.0,
                "Percentiles must be in the range [0, 100]");
  const double qs[2] = {q_lo, q_hi};
  for (int j = 0; j < 2; ++j) {
    const double quant = qs[j] / 100.0;                          // np.true_divide(q, 100)
    const double virt = (double)(n_total - 1) * quant;           // (n - 1) * quantiles
    int64_t prev = (int64_t)__builtin_floor(virt), next = prev + 1;
    double prev_for_gamma = (double)prev;
    if (virt >= (double)(n_total - 1)) {                          // _get_indexes: above bounds -> index -1
      prev = next = n_total - 1;
      prev_for_gamma = -1.0;
    }
    if (virt < 0) prev = next = 0;
    ranks_out[2 * j] = prev;
    ranks_out[2 * j + 1] = next;
    gamma_out[j] = virt - prev_for_gamma;                         // _get_gamma
  }
  return OEH_OK;
}

extern "C" int oeh_percentile_begin(void* workspace, size_t workspace_bytes, const int64_t ranks[4], oeh_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OEH_CHECK_ARG(workspace && ranks, "oeh_percentile_begin: null pointer");
  if (workspace_bytes < sizeof(Ws)) return set_error(OEH_ERR_WORKSPACE, "oeh_percentile: workspace too small (%zu < %zu)",
                                                     workspace_bytes, sizeof(Ws));
  OEH_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "oeh_percentile: workspace must be 8-byte aligned");
  begin_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<Ws*>(workspace), ranks[0], ranks[1], ranks[2], ranks[3]);
  return check_launch("percentile begin_kernel");
}

extern "C" int oeh_percentile_hist(const void* x, int64_t n, int32_t dtype, int32_t pass, void* workspace, oeh_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OEH_CHECK_ARG(x && workspace, "oeh_percentile_hist: null pointer");
  OEH_CHECK_ARG(n > 0 && n < ((int64_t)1 << 32), "oeh_percentile_hist: n must be in [1, 2^32)");
  if (dtype != OEH_F32 && dtype != OEH_F16)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_percentile: fp32 / fp16 only (numpy has no bfloat16; the reference raises)");
  OEH_CHECK_ARG(pass >= 0 && pass < n_passes(dtype), "oeh_percentile_hist: bad pass %d", pass);
  Ws* ws = reinterpret_cast<Ws*>(workspace);
  const int shift = pass_shift(dtype, pass), bits = pass_bits(dtype, pass);
  const int grid = hist_grid(n, dtype_size(dtype));
  if (dtype == OEH_F32) {
    if (pass == 0) hist_kernel<float, true><<<grid, kHistThreads, 0, stream>>>((const float*)x, n, shift, bits, ws);
    else hist_kernel<float, false><<<grid, kHistThreads, 0, stream>>>((const float*)x, n, shift, bits, ws);
  } else {
    if (pass == 0) hist_kernel<__half, true><<<grid, kHistThreads, 0, stream>>>((const __half*)x, n, shift, bits, ws);
    else hist_kernel<__half, false><<<grid, kHistThreads, 0, stream>>>((const __half*)x, n, shift, bits, ws);
  }
  return check_launch("percentile hist_kernel");
}

extern "C" int oeh_percentile_scan(int32_t dtype, int32_t pass, void* workspace, oeh_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OEH_CHECK_ARG(workspace, "oeh_percentile_scan: null pointer");
  OEH_CHECK_ARG((dtype == OEH_F32 || dtype == OEH_F16) && pass >= 0 && pass < n_passes(dtype), "oeh_percentile_scan: bad pass");
  scan_kernel<<<1, kScanThreads, 0, stream>>>(reinterpret_cast<Ws*>(workspace), pass_bits(dtype, pass), pass == 0);
  return check_launch("percentile scan_kernel");
}

extern "C" int oeh_percentile_finish(int32_t dtype, const double gamma[2], double* out2, void* workspace, oeh_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OEH_CHECK_ARG(workspace && out2 && gamma, "oeh_percentile_finish: null pointer");
  const Ws* ws = reinterpret_cast<const Ws*>(workspace);
  if (dtype == OEH_F32) finish_kernel<float><<<1, 32, 0, stream>>>(ws, gamma[0], gamma[1], out2);
  else if (dtype == OEH_F16) finish_kernel<__half><<<1, 32, 0, stream>>>(ws, gamma[0], gamma[1], out2);
  else return set_error(OEH_ERR_UNSUPPORTED, "oeh_percentile: fp32 / fp16 only");
  return check_launch("percentile finish_kernel");
}

extern "C" int oeh_percentile(const void* x, int64_t n, int32_t dtype, double q_lo, double q_hi, double* out2,
                              void* workspace, size_t workspace_bytes, oeh_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int64_t ranks[4];
  double gamma[2];
  int rc = oeh_percentile_plan(n, q_lo, q_hi, ranks, gamma);
  if (rc != OEH_OK) return rc;
  if (dtype != OEH_F32 && dtype != OEH_F16)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_percentile: fp32 / fp16 only (numpy has no bfloat16; the reference raises)");
  if ((rc = oeh_percentile_begin(workspace, workspace_bytes, ranks, stream_)) != OEH_OK) return rc;
  for (int32_t pass = 0; pass < n_passes(dtype); ++pass) {
    if ((rc = oeh_percentile_hist(x, n, dtype, pass, workspace, stream_)) != OEH_OK) return rc;
    if ((rc = oeh_percentile_scan(dtype, pass, workspace, stream_)) != OEH_OK) return rc;
  }
  return oeh_percentile_finish(dtype, gamma, out2, workspace, stream_);
}

extern "C" int oeh_range_update_f64(double* state64, float* qstate, const double* new_minmax, double momentum, int32_t n_bits,
                                    double eps, oeh_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OEH_CHECK_ARG(state64 && qstate && new_minmax, "oeh_range_update_f64: null pointer");
  OEH_CHECK_ARG(n_bits >= 1 && n_bits <= 16, "oeh_range_update_f64: n_bits %d out of range", n_bits);
  range_update_f64_kernel<<<1, 32, 0, stream>>>(state64, qstate, new_minmax, 1.0 - momentum, momentum,
                                                (double)((1 << n_bits) - 1), eps);
  return check_launch("range_update_f64_kernel");
}
