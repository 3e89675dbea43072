// HBM-bandwidth-bound kernels of the activation fake-quant path (sm_100a):
//   fake_quant_asym  -- AsymmetricUniformQuantizer.forward/to_integer_forward
//                       (quantization/quantizers/uniform_quantizers.py:92-148)
//   minmax           -- torch.min/torch.max of RunningMinMaxEstimator (range_estimators.py:96-97)
//   range_update     -- EMA (range_estimators.py:99-104) + set_quant_range (uniform_quantizers.py:204-224)
// Roofline: bytes = n*(size_in+size_out) for fake-quant, n*size_in for min/max, against the measured
// HBM copy bandwidth.  All traffic is 16-byte vectorised and coalesced; grids are sized as a
// multiple of the SM count and grid-stride over the tensor with 4 independent 16-byte loads in
// flight per thread.  No fast-math: the fp32 path must be bit-exact with ATen (IEEE div, rintf =
// round-half-even, no FMA contraction -- every op below is a single rounding by construction).
#include <type_traits>

#include "oeh_common.cuh"

namespace oeh {

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kBlocksPerSm = 8;

// streaming (evict-first) 16-byte accesses: every byte is touched exactly once
template <typename V>
__device__ __forceinline__ V ld_stream(const V* p) {
  int4 raw = __ldcs(reinterpret_cast<const int4*>(p));
  return *reinterpret_cast<V*>(&raw);
}
template <typename V>
__device__ __forceinline__ void st_stream(V* p, const V& v) {
  __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<const int4*>(&v));
}

template <typename T>
__global__ void __launch_bounds__(kThreads) fake_quant_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n,
                                                              const float* __restrict__ qparams, float int_max,
                                                              int codes) {
  using V = Vec16<T>;
  constexpr int N = V::N;
  // 0-dim fp32 scale / zero_point are converted to the tensor dtype first (ATen-CUDA dynamic cast)
  const float scale = DT<T>::rnd(__ldg(qparams));
  const float zp = DT<T>::rnd(__ldg(qparams + 1));
  const QdqVec<T> qv(scale, zp, int_max);
  const int64_t nvec = n / N;
  const V* xv = reinterpret_cast<const V*>(x);
  V* yv = reinterpret_cast<V*>(y);
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  // The main loop exists twice: with only the quantizer's common chain compiled in (QdqVec::fast(): straight-line code), and
  // with every rare-case chain.  Interleaved in one loop body the hot path was a string of jumps over code it never runs --
  // instruction-cache misses at every landing (measured on the LayerNorm taps: profiles/r02_ln_code_size.txt)
  auto main_loop = [&](auto fast_tag) {
    constexpr bool kFast = decltype(fast_tag)::value;
    for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
      V in[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) in[u] = ld_stream(xv + i + u * stride);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        qv.template apply_t<kFast>(in[u], codes);
        st_stream(yv + i + u * stride, in[u]);
      }
    }
  };
  if (qv.fast()) main_loop(std::true_type{});
  else main_loop(std::false_type{});
  for (; i < nvec; i += stride) {
    V in = xv[i];
    qv.apply(in, codes);
    yv[i] = in;
  }
  // scalar tail
  for (int64_t t = nvec * N + (int64_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride)
    y[t] = DT<T>::from_f(qdq_one<T>(DT<T>::to_f(x[t]), scale, zp, int_max, codes));
}

// unaligned fallback (pointer not 16-byte aligned): scalar, still coalesced
template <typename T>
__global__ void __launch_bounds__(kThreads) fake_quant_scalar_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                     int64_t n, const float* __restrict__ qparams,
                                                                     float int_max, int codes) {
  const float scale = DT<T>::rnd(__ldg(qparams));
  const float zp = DT<T>::rnd(__ldg(qparams + 1));
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride)
    y[t] = DT<T>::from_f(qdq_one<T>(DT<T>::to_f(x[t]), scale, zp, int_max, codes));
}


// ---- residual add + fake-quant (quantized_bert.py:566-567, 602-603; quantized_opt.py:342-345, 367-369) ---------
// sum = round_T(a + b) exactly as ATen's add kernel (one rounding to the tensor dtype), then the qdq chain above.
template <typename T>
__global__ void __launch_bounds__(kThreads) add_fake_quant_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                                  T* __restrict__ y, int64_t n,
                                                                  const float* __restrict__ qparams, float int_max,
                                                                  int aligned, const float* __restrict__ qparams_b) {
  using V = Vec16<T>;
  constexpr int N = V::N;
  const float scale = DT<T>::rnd(__ldg(qparams));
  const float zp = DT<T>::rnd(__ldg(qparams + 1));
  // optional tap on b before the add (the output tap of the QuantLinear that produced b: quantized_opt.py:362-369)
  const float bscale = qparams_b ? DT<T>::rnd(__ldg(qparams_b)) : 1.f;
  const float bzp = qparams_b ? DT<T>::rnd(__ldg(qparams_b + 1)) : 0.f;
  const QdqVec<T> qv(scale, zp, int_max), qb(bscale, bzp, int_max);
  auto one = [&](T av, T bv) {
    float bf = DT<T>::to_f(bv);
    if (qparams_b) bf = qdq_one<T>(bf, bscale, bzp, int_max, false);
    const float s = DT<T>::rnd(__fadd_rn(DT<T>::to_f(av), bf));
    return DT<T>::from_f(qdq_one<T>(s, scale, zp, int_max, false));
  };
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int64_t done = 0;
  if (aligned) {
    const int64_t nvec = n / N;
    const V* av = reinterpret_cast<const V*>(a);
    const V* bv = reinterpret_cast<const V*>(b);
    V* yv = reinterpret_cast<V*>(y);
    int64_t i = tid;
    for (; i + stride < nvec; i += 2 * stride) {          // 2 x (two 16-byte loads) in flight per thread
      V ia[2], ib[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        ia[u] = ld_stream(av + i + u * stride);
        ib[u] = ld_stream(bv + i + u * stride);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (qparams_b) qb.apply(ib[u], false);
        add_vec<T>(ia[u], ib[u]);
        qv.apply(ia[u], false);
        st_stream(yv + i + u * stride, ia[u]);
      }
    }
    for (; i < nvec; i += stride) {
      V ia = av[i], ib = bv[i], out;
#pragma unroll
      for (int e = 0; e < N; ++e) out.v[e] = one(ia.v[e], ib.v[e]);
      yv[i] = out;
    }
    done = nvec * N;
  }
  for (int64_t t = done + tid; t < n; t += stride) y[t] = one(a[t], b[t]);
}

// ---- min/max ---------------------------------------------------------------------------------
struct MinMax {
  float lo, hi;
  int nan;
};

__device__ __forceinline__ void mm_acc(MinMax& m, float v) {
  m.nan |= (v != v);
  m.lo = fminf(m.lo, v);
  m.hi = fmaxf(m.hi, v);
}

__device__ __forceinline__ MinMax block_reduce(MinMax m) {
  __shared__ float s_lo[kThreads / 32], s_hi[kThreads / 32];
  __shared__ int s_nan[kThreads / 32];
  m.lo = warp_min(m.lo);
  m.hi = warp_max(m.hi);
  m.nan = __any_sync(0xffffffffu, m.nan);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    s_lo[w] = m.lo;
    s_hi[w] = m.hi;
    s_nan[w] = m.nan;
  }
  __syncthreads();
  if (w == 0) {
    constexpr int NW = kThreads / 32;
    m.lo = l < NW ? s_lo[l] : INFINITY;
    m.hi = l < NW ? s_hi[l] : -INFINITY;
    m.nan = l < NW ? s_nan[l] : 0;
    m.lo = warp_min(m.lo);
    m.hi = warp_max(m.hi);
    m.nan = __any_sync(0xffffffffu, m.nan);
  }
  return m;   // valid in warp 0
}

// stage 1: every block reduces a grid-strided slice to partial[3*block]
template <typename T>
__global__ void __launch_bounds__(kThreads) minmax_stage1_kernel(const T* __restrict__ x, int64_t n,
                                                                 float* __restrict__ partial, int aligned) {
  using V = Vec16<T>;
  constexpr int N = V::N;
  MinMax m{INFINITY, -INFINITY, 0};
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int64_t done = 0;
  if (aligned) {
    const int64_t nvec = n / N;
    const V* xv = reinterpret_cast<const V*>(x);
    int64_t i = tid;
    for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
      V in[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) in[u] = ld_stream(xv + i + u * stride);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int e = 0; e < N; ++e) mm_acc(m, DT<T>::to_f(in[u].v[e]));
    }
    for (; i < nvec; i += stride) {
      V in = xv[i];
#pragma unroll
      for (int e = 0; e < N; ++e) mm_acc(m, DT<T>::to_f(in.v[e]));
    }
    done = nvec * N;
  }
  for (int64_t t = done + tid; t < n; t += stride) mm_acc(m, DT<T>::to_f(x[t]));
  m = block_reduce(m);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x + 0] = m.lo;
    partial[3 * blockIdx.x + 1] = m.hi;
    partial[3 * blockIdx.x + 2] = m.nan ? 1.f : 0.f;
  }
}

// stage 2: one block folds the partials (min/max are exact and associative: deterministic)
__global__ void __launch_bounds__(kThreads) minmax_stage2_kernel(const float* __restrict__ partial, int nblocks,
                                                                 float* __restrict__ out2) {
  MinMax m{INFINITY, -INFINITY, 0};
  for (int i = threadIdx.x; i < nblocks; i += kThreads) {
    m.lo = fminf(m.lo, partial[3 * i]);
    m.hi = fmaxf(m.hi, partial[3 * i + 1]);
    m.nan |= partial[3 * i + 2] != 0.f;
  }
  m = block_reduce(m);
  if (threadIdx.x == 0) {
    const float qnan = __int_as_float(0x7fc00000);
    out2[0] = m.nan ? qnan : m.lo;
    out2[1] = m.nan ? qnan : m.hi;
  }
}

// ---- EMA + set_quant_range, one thread, no host round trip -----------------------------------
__global__ void range_update_kernel(float* __restrict__ state, const float* __restrict__ new_minmax, float one_minus,
                                    float mom, int n_bits, float eps) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float lo = new_minmax[0], hi = new_minmax[1];
  if (state[OEH_QS_INITIALIZED] != 0.f) {
    // (1 - momentum) * x_min + momentum * current  (range_estimators.py:103-104): mul, mul, add
    lo = __fadd_rn(__fmul_rn(one_minus, lo), __fmul_rn(mom, state[OEH_QS_CUR_MIN]));
    hi = __fadd_rn(__fmul_rn(one_minus, hi), __fmul_rn(mom, state[OEH_QS_CUR_MAX]));
  }
  state[OEH_QS_CUR_MIN] = lo;
  state[OEH_QS_CUR_MAX] = hi;
  set_quant_range_f32(lo, hi, n_bits, eps, &state[OEH_QS_DELTA], &state[OEH_QS_ZERO_FLOAT], &state[OEH_QS_SCALE],
                      &state[OEH_QS_ZERO_POINT]);
  state[OEH_QS_INITIALIZED] = 1.f;
}

int grid_for(int64_t work_items) {
  const int64_t cap = (int64_t)num_sms() * kBlocksPerSm;
  int64_t g = (work_items + kThreads - 1) / kThreads;
  if (g > cap) g = cap;                       // persistent-style: a multiple of the SM count
  if (g < 1) g = 1;
  return (int)g;
}

template <typename T>
int launch_fake_quant(const void* x, void* y, int64_t n, const float* qparams, int n_bits, int codes,
                      cudaStream_t st) {
  const float int_max = (float)((1 << n_bits) - 1);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (aligned) {
    const int grid = grid_for((n / Vec16<T>::N + kUnroll - 1) / kUnroll + 1);
    fake_quant_kernel<T><<<grid, kThreads, 0, st>>>((const T*)x, (T*)y, n, qparams, int_max, codes);
  } else {
    fake_quant_scalar_kernel<T><<<grid_for(n), kThreads, 0, st>>>((const T*)x, (T*)y, n, qparams, int_max, codes);
  }
  return check_launch("fake_quant_kernel");
}

template <typename T>
int launch_minmax(const void* x, int64_t n, float* out2, float* partial, cudaStream_t st) {
  const int aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int grid = grid_for((n / Vec16<T>::N + kUnroll - 1) / kUnroll + 1);
  minmax_stage1_kernel<T><<<grid, kThreads, 0, st>>>((const T*)x, n, partial, aligned);
  int rc = check_launch("minmax_stage1_kernel");
  if (rc) return rc;
  minmax_stage2_kernel<<<1, kThreads, 0, st>>>(partial, grid, out2);
  return check_launch("minmax_stage2_kernel");
}

template <typename T>
int launch_add_fake_quant(const void* a, const void* b, void* y, int64_t n, const float* qparams, int n_bits,
                          cudaStream_t st, const float* qparams_b = nullptr) {
  const float int_max = (float)((1 << n_bits) - 1);
  const int aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const int grid = grid_for(aligned ? (n / Vec16<T>::N + 1) / 2 + 1 : n);
  add_fake_quant_kernel<T><<<grid, kThreads, 0, st>>>((const T*)a, (const T*)b, (T*)y, n, qparams, int_max, aligned, qparams_b);
  return check_launch("add_fake_quant_kernel");
}

// ---- fake-quant of column segments with their own ranges ---------------------------------------------------------------
// x, y: (rows, n_seg * seg_cols) with explicit row strides; segment s (columns [s*seg_cols, (s+1)*seg_cols)) uses
// qp[s].  One launch taps the output of a fused query | key | value projection (quantized_bert.py:317-330: three
// QuantLinear modules, three ranges).
struct SegQ {
  const float* qp[4];
};
template <typename T>
__global__ void __launch_bounds__(kThreads) fake_quant_segments_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                       int64_t rows, int seg_cols, int n_seg,
                                                                       int64_t ldx, int64_t ldy, SegQ q, float int_max) {
  using V = Vec16<T>;
  constexpr int N = V::N;
  const int vec_per_seg = seg_cols / N, vec_per_row = vec_per_seg * n_seg;
  const int64_t total = rows * vec_per_row;
  float scale[4], zp[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    scale[s] = s < n_seg ? DT<T>::rnd(__ldg(q.qp[s])) : 1.f;
    zp[s] = s < n_seg ? DT<T>::rnd(__ldg(q.qp[s] + 1)) : 0.f;
  }
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int64_t r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row);
    const int s = c / vec_per_seg;
    const float sc = s == 0 ? scale[0] : (s == 1 ? scale[1] : (s == 2 ? scale[2] : scale[3]));
    const float z = s == 0 ? zp[0] : (s == 1 ? zp[1] : (s == 2 ? zp[2] : zp[3]));
    V in = *reinterpret_cast<const V*>(x + r * ldx + (int64_t)c * N);
    QdqVec<T>(sc, z, int_max).apply(in, false);
    *reinterpret_cast<V*>(y + r * ldy + (int64_t)c * N) = in;
  }
}

template <typename T>
int launch_segments(const void* x, void* y, int64_t rows, int seg_cols, int n_seg, int64_t ldx, int64_t ldy, const SegQ& q,
                    int n_bits, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  if (seg_cols % N || ldx % N || ldy % N || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15))
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_fake_quant_asym_segments: segments, strides and pointers must be 16-byte aligned");
  const float int_max = (float)((1 << n_bits) - 1);
  const int grid = grid_for(rows * (seg_cols / N) * n_seg);
  fake_quant_segments_kernel<T><<<grid, kThreads, 0, st>>>((const T*)x, (T*)y, rows, seg_cols, n_seg, ldx, ldy, q, int_max);
  return check_launch("fake_quant_segments_kernel");
}

}  // namespace
}  // namespace oeh

using namespace oeh;

extern "C" int oeh_fake_quant_asym(const void* x, void* y, int64_t n, int32_t dtype, const float* qparams,
                                   int32_t n_bits, int32_t emit_codes, oeh_stream_t stream) {
  OEH_CHECK_ARG(n >= 0, "oeh_fake_quant_asym: n=%lld < 0", (long long)n);
  OEH_CHECK_ARG(n_bits >= 1 && n_bits <= 16, "oeh_fake_quant_asym: n_bits=%d outside [1,16]", n_bits);
  if (n == 0) return OEH_OK;
  OEH_CHECK_ARG(x && y && qparams, "oeh_fake_quant_asym: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case OEH_F32: return launch_fake_quant<float>(x, y, n, qparams, n_bits, emit_codes, st);
    case OEH_F16: return launch_fake_quant<__half>(x, y, n, qparams, n_bits, emit_codes, st);
    case OEH_BF16: return launch_fake_quant<__nv_bfloat16>(x, y, n, qparams, n_bits, emit_codes, st);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_fake_quant_asym: unknown dtype %d", dtype);
}

extern "C" int oeh_add_fake_quant(const void* a, const void* b, void* y, int64_t n, int32_t dtype, const float* qparams,
                                  int32_t n_bits, oeh_stream_t stream) {
  OEH_CHECK_ARG(n >= 0, "oeh_add_fake_quant: n=%lld < 0", (long long)n);
  OEH_CHECK_ARG(n_bits >= 1 && n_bits <= 16, "oeh_add_fake_quant: n_bits=%d outside [1,16]", n_bits);
  if (n == 0) return OEH_OK;
  OEH_CHECK_ARG(a && b && y && qparams, "oeh_add_fake_quant: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case OEH_F32: return launch_add_fake_quant<float>(a, b, y, n, qparams, n_bits, st);
    case OEH_F16: return launch_add_fake_quant<__half>(a, b, y, n, qparams, n_bits, st);
    case OEH_BF16: return launch_add_fake_quant<__nv_bfloat16>(a, b, y, n, qparams, n_bits, st);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_add_fake_quant: unknown dtype %d", dtype);
}

extern "C" int oeh_tap_add_fake_quant(const void* a, const void* b, void* y, int64_t n, int32_t dtype, const float* qparams_b,
                                      const float* qparams_sum, int32_t n_bits, oeh_stream_t stream) {
  OEH_CHECK_ARG(n >= 0, "oeh_tap_add_fake_quant: n=%lld < 0", (long long)n);
  OEH_CHECK_ARG(n_bits >= 1 && n_bits <= 16, "oeh_tap_add_fake_quant: n_bits=%d outside [1,16]", n_bits);
  if (n == 0) return OEH_OK;
  OEH_CHECK_ARG(a && b && y && qparams_sum, "oeh_tap_add_fake_quant: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case OEH_F32: return launch_add_fake_quant<float>(a, b, y, n, qparams_sum, n_bits, st, qparams_b);
    case OEH_F16: return launch_add_fake_quant<__half>(a, b, y, n, qparams_sum, n_bits, st, qparams_b);
    case OEH_BF16: return launch_add_fake_quant<__nv_bfloat16>(a, b, y, n, qparams_sum, n_bits, st, qparams_b);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_tap_add_fake_quant: unknown dtype %d", dtype);
}

extern "C" int oeh_fake_quant_asym_segments(const void* x, void* y, int64_t rows, int32_t seg_cols, int32_t n_seg,
                                            int64_t x_row_stride, int64_t y_row_stride, int32_t dtype,
                                            const float* const* qparams, int32_t n_bits, oeh_stream_t stream) {
  OEH_CHECK_ARG(rows >= 0 && seg_cols > 0 && n_seg >= 1 && n_seg <= 4, "oeh_fake_quant_asym_segments: bad shape");
  OEH_CHECK_ARG(n_bits >= 1 && n_bits <= 16, "oeh_fake_quant_asym_segments: n_bits=%d outside [1,16]", n_bits);
  if (rows == 0) return OEH_OK;
  OEH_CHECK_ARG(x && y && qparams, "oeh_fake_quant_asym_segments: null pointer");
  OEH_CHECK_ARG(x_row_stride >= (int64_t)seg_cols * n_seg && y_row_stride >= (int64_t)seg_cols * n_seg,
                "oeh_fake_quant_asym_segments: row stride smaller than the row");
  SegQ q;
  for (int s = 0; s < 4; ++s) {
    q.qp[s] = s < n_seg ? qparams[s] : nullptr;
    if (s < n_seg) OEH_CHECK_ARG(q.qp[s], "oeh_fake_quant_asym_segments: null qparams[%d]", s);
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case OEH_F32: return launch_segments<float>(x, y, rows, seg_cols, n_seg, x_row_stride, y_row_stride, q, n_bits, st);
    case OEH_F16: return launch_segments<__half>(x, y, rows, seg_cols, n_seg, x_row_stride, y_row_stride, q, n_bits, st);
    case OEH_BF16:
      return launch_segments<__nv_bfloat16>(x, y, rows, seg_cols, n_seg, x_row_stride, y_row_stride, q, n_bits, st);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_fake_quant_asym_segments: unknown dtype %d", dtype);
}

extern "C" size_t oeh_minmax_workspace_bytes(int64_t n) {
  (void)n;
  return (size_t)num_sms() * kBlocksPerSm * 3 * sizeof(float);
}

extern "C" int oeh_minmax(const void* x, int64_t n, int32_t dtype, float* out2, void* workspace,
                          size_t workspace_bytes, oeh_stream_t stream) {
  OEH_CHECK_ARG(n > 0, "oeh_minmax: empty tensor (torch.min raises on numel()==0 too)");
  OEH_CHECK_ARG(x && out2, "oeh_minmax: null pointer");
  if (!workspace || workspace_bytes < oeh_minmax_workspace_bytes(n))
    return set_error(OEH_ERR_WORKSPACE, "oeh_minmax: workspace of %zu bytes needed", oeh_minmax_workspace_bytes(n));
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  switch (dtype) {
    case OEH_F32: return launch_minmax<float>(x, n, out2, partial, st);
    case OEH_F16: return launch_minmax<__half>(x, n, out2, partial, st);
    case OEH_BF16: return launch_minmax<__nv_bfloat16>(x, n, out2, partial, st);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_minmax: unknown dtype %d", dtype);
}

extern "C" int oeh_range_update(float* state, const float* new_minmax, double momentum, int32_t n_bits, double eps,
                                oeh_stream_t stream) {
  OEH_CHECK_ARG(state && new_minmax, "oeh_range_update: null pointer");
  OEH_CHECK_ARG(n_bits >= 1 && n_bits <= 16, "oeh_range_update: n_bits=%d outside [1,16]", n_bits);
  // Python evaluates (1 - momentum) in double; the product with an fp32 tensor converts it to fp32.
  range_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, new_minmax, (float)(1.0 - momentum), (float)momentum,
                                                          n_bits, (float)eps);
  return check_launch("range_update_kernel");
}

extern "C" int oeh_set_quant_range_host(float x_min, float x_max, int32_t n_bits, double eps, float* out4) {
  OEH_CHECK_ARG(out4, "oeh_set_quant_range_host: null pointer");
  OEH_CHECK_ARG(n_bits >= 1 && n_bits <= 16, "oeh_set_quant_range_host: n_bits=%d outside [1,16]", n_bits);
  set_quant_range_f32(x_min, x_max, n_bits, (float)eps, &out4[0], &out4[1], &out4[2], &out4[3]);
  return OEH_OK;
}

extern "C" int oeh_estimate_and_quantize(const void* x, void* y, int64_t n, int32_t dtype, float* state,
                                         double momentum, int32_t n_bits, double eps, void* workspace,
                                         size_t workspace_bytes, oeh_stream_t stream) {
  OEH_CHECK_ARG(state, "oeh_estimate_and_quantize: null state");
  const size_t need = oeh_minmax_workspace_bytes(n) + 256;
  if (!workspace || workspace_bytes < need)
    return set_error(OEH_ERR_WORKSPACE, "oeh_estimate_and_quantize: workspace of %zu bytes needed", need);
  float* mm = (float*)workspace;                       // [min,max] scratch, then the partials
  void* ws = (char*)workspace + 256;
  int rc = oeh_minmax(x, n, dtype, mm, ws, workspace_bytes - 256, stream);
  if (rc) return rc;
  rc = oeh_range_update(state, mm, momentum, n_bits, eps, stream);
  if (rc) return rc;
  return oeh_fake_quant_asym(x, y, n, dtype, state + OEH_QS_SCALE, n_bits, 0, stream);
}
