// Shared helpers for liboeh.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/oeh.h"

namespace oeh {

// ---- error reporting (thread-local, never throws) --------------------------------------------
char* last_error_buf();
int set_error(int status, const char* fmt, ...);

#define OEH_CHECK_ARG(cond, ...)                              \
  do {                                                        \
    if (!(cond)) return ::oeh::set_error(OEH_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define OEH_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t err__ = (expr);                                                                   \
    if (err__ != cudaSuccess)                                                                     \
      return ::oeh::set_error(OEH_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                              __FILE__, __LINE__);                                                \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t err = cudaPeekAtLastError();
  if (err != cudaSuccess) {
    cudaGetLastError();
    return set_error(OEH_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(err));
  }
  return OEH_OK;
}

int num_sms();

inline size_t dtype_size(int dtype) { return dtype == OEH_F32 ? 4 : 2; }

// ---- dtype traits: load/store as fp32, with the ATen "round after every op" helper ------------
template <typename T>
struct DT;
template <>
struct DT<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
};
template <>
struct DT<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
};
template <>
struct DT<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

// 16-byte vector of T
template <typename T>
struct alignas(16) Vec16 {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// AsymmetricUniformQuantizer.forward / to_integer_forward for ONE element (uniform_quantizers.py:92-148), shared by the
// standalone quantizer (quant.cu) and the fused taps (quant.cu add+qdq, layernorm.cu).
template <typename T>
__device__ __forceinline__ float qdq_one(float x, float scale, float zp, float int_max, bool codes) {
  // For T=float DT::rnd is the identity and this is exactly the reference's fp32 chain.
  // For 16-bit T every intermediate is rounded to T, as ATen does per elementwise op.
  using D = DT<T>;
  // x_float / scale.  (A reciprocal + FMA-corrected quotient was tried instead of the IEEE division: it needs guards for
  // tiny / huge operands to stay bit-exact and was no faster -- 3.2 vs 3.8 TB/s on fp16 tensors, which are issue-bound
  // on the four per-op roundings to the tensor dtype, not on the division.)
  float q = D::rnd(__fdiv_rn(x, scale));
  float r = rintf(q);                                // round_ste_func (half-to-even)
  float xi = D::rnd(__fadd_rn(r, zp));               // + zero_point
  xi = xi < 0.f ? 0.f : (xi > int_max ? int_max : xi);   // clamp keeps NaN like torch.clamp
  if (codes) return xi;
  float diff = D::rnd(__fsub_rn(xi, zp));            // x_int - zero_point
  return D::rnd(__fmul_rn(scale, diff));             // scale * (...)
}

// ---- the same chain on a 16-byte vector --------------------------------------------------------------------------------
// 16-bit tensors: two elements per instruction with the hardware's own fp16 / bf16 arithmetic.  Every ATen op of the
// chain is "compute in fp32, round once to the tensor dtype"; for +, -, x, / on 16-bit operands that equals the correctly
// rounded 16-bit operation (24 >= 2*11 + 2 mantissa bits: the double rounding is innocuous), which is what HADD2 / HMUL2
// (.rn) deliver.  Two steps of the chain have no packed instruction (the compiler expands __h2div into a per-element fp32
// division with its slow-path call, h2rint into per-element FRND) and are restated exactly:
//   * x / scale  ->  RN16(fp32(x) * RN32(1 / scale)).  The quotient of two p-bit significands is never a rounding midpoint
//     of the p-bit grid and stays at least 2^-(2p) away from one, while the fp32 product is within 2^-22 of the quotient
//     (p = 11: 2^-22 vs an error < 2^-22(1 - 2^-11); p = 8: 2^-16), so both round to the same 16-bit value.  (Results in the
//     subnormal range may differ in the last bit; they are < 2^-14 and quantise to the same code.)  Used when 1 / scale
//     is a normal fp32 number; anything else keeps the division.
//   * fp16, n_bits <= 8:  clamp(rint(q) + zp, 0, M) - zp == rint(clamp(q, -zp, M - zp))  (zp integer), and inside
//     [-255, 255] round-half-even is (q + 1536) - 1536 in fp16 (unit spacing on [1024, 2048), 1536 even).
// Both identities were checked on the host over every fp16 input x 3000 scales x 4 zero points, and the kernels are
// compared with ATen over all 65536 bit patterns on the GPU (tests/test_gpu_quant.py).
// Used when 2^n_bits - 1 is exactly representable in the dtype (fp16: n <= 11, bf16: n <= 8).
template <typename T>
struct P2 {
  static constexpr bool kPacked = false;
};
template <>
struct P2<__half> {
  static constexpr bool kPacked = true;
  static constexpr float kMagic = 1536.f;          // 1.5 * 2^10; 0 = no magic-number rounding for this dtype
  using V2 = __half2;
  static __device__ __forceinline__ V2 bcast(float f) { return __float2half2_rn(f); }
  static __device__ __forceinline__ bool exact_int(float m) { return m <= 2047.f; }
  static __device__ __forceinline__ float2 to_f2(V2 v) { return __half22float2(v); }
  static __device__ __forceinline__ V2 from_f2(float a, float b) { return __floats2half2_rn(a, b); }
};
template <>
struct P2<__nv_bfloat16> {
  static constexpr bool kPacked = true;
  static constexpr float kMagic = 0.f;             // [128, 256) has unit spacing: too narrow for 8-bit codes
  using V2 = __nv_bfloat162;
  static __device__ __forceinline__ V2 bcast(float f) { return __float2bfloat162_rn(f); }
  static __device__ __forceinline__ bool exact_int(float m) { return m <= 255.f; }
  static __device__ __forceinline__ float2 to_f2(V2 v) { return __bfloat1622float2(v); }
  static __device__ __forceinline__ V2 from_f2(float a, float b) { return __floats2bfloat162_rn(a, b); }
};
template <typename T, bool kPacked = P2<T>::kPacked>
struct QdqVec;

template <typename T>
struct QdqVec<T, false> {                         // fp32 tensors: the scalar chain
  float scale, zp, int_max;
  __device__ __forceinline__ QdqVec(float scale_, float zp_, float int_max_) : scale(scale_), zp(zp_), int_max(int_max_) {}
  __device__ __forceinline__ void apply(Vec16<T>& v, bool codes) const {
#pragma unroll
    for (int e = 0; e < Vec16<T>::N; ++e) v.v[e] = DT<T>::from_f(qdq_one<T>(DT<T>::to_f(v.v[e]), scale, zp, int_max, codes));
  }
  __device__ __forceinline__ bool fast() const { return false; }
  template <bool kFast>
  __device__ __forceinline__ void apply_t(Vec16<T>& v, bool codes) const { apply(v, codes); }
};

template <typename T>
struct QdqVec<T, true> {
  using V2 = typename P2<T>::V2;
  // kept small on purpose (add_layer_norm holds three of these): rinv = RN32(1 / scale) or 0 (= divide);
  // a2 = -zp (magic rounding) or int_max; b2 = int_max - zp (magic rounding only)
  float rinv, int_max;
  V2 s2, z2, a2, b2;
  bool packed, magic;
  // scale_ / zp_: already rounded to T (the 0-dim fp32 scalars are cast to the tensor dtype first)
  __device__ __forceinline__ QdqVec(float scale_, float zp_, float int_max_) : int_max(int_max_) {
    s2 = P2<T>::bcast(scale_);
    z2 = P2<T>::bcast(zp_);
    packed = P2<T>::exact_int(int_max_);
    rinv = (scale_ >= 1e-30f && scale_ <= 1e30f) ? __frcp_rn(scale_) : 0.f;   // 1 / scale is a normal fp32 number
    magic = P2<T>::kMagic > 0.f && int_max_ <= 255.f && zp_ >= 0.f && zp_ <= int_max_;
    a2 = P2<T>::bcast(magic ? -zp_ : int_max_);
    b2 = P2<T>::bcast(int_max_ - zp_);
  }
  __device__ __forceinline__ V2 one2(V2 x, bool codes) const {
    V2 q;
    if (rinv != 0.f) {
      const float2 xf = P2<T>::to_f2(x);
      q = P2<T>::from_f2(__fmul_rn(xf.x, rinv), __fmul_rn(xf.y, rinv));   // == x_float / scale (see above)
    } else {
      q = __h2div(x, s2);
    }
    if (magic) {
      const V2 k2 = P2<T>::bcast(P2<T>::kMagic);
      q = __hmin2_nan(__hmax2_nan(q, a2), b2);    // clamp first (NaN propagates like torch.clamp) ...
      q = __hsub2_rn(__hadd2_rn(q, k2), k2);      // ... then round half to even: == x_int - zero_point
      if (codes) return __hadd2_rn(q, z2);
      return __hmul2_rn(s2, q);                   // scale * (...)
    }
    q = h2rint(q);                                // round half to even
    q = __hadd2_rn(q, z2);                        // + zero_point
    q = __hmin2_nan(__hmax2_nan(q, P2<T>::bcast(0.f)), a2);   // clamp (NaN propagates like torch.clamp)
    if (codes) return q;
    q = __hsub2_rn(q, z2);                        // x_int - zero_point
    return __hmul2_rn(s2, q);                     // scale * (...)
  }
  // the common case: packed 16-bit codes, reciprocal quotient, magic-number rounding -- exactly the first branches of one2()
  __device__ __forceinline__ bool fast() const { return packed && magic && rinv != 0.f; }
  // kFast: the caller has checked fast(); only that chain is compiled in (same instructions, same results as apply())
  template <bool kFast>
  __device__ __forceinline__ void apply_t(Vec16<T>& v, bool codes) const {
    if constexpr (kFast) {
      V2* p = reinterpret_cast<V2*>(v.v);
      const V2 k2 = P2<T>::bcast(P2<T>::kMagic);
#pragma unroll
      for (int k = 0; k < Vec16<T>::N / 2; ++k) {
        const float2 xf = P2<T>::to_f2(p[k]);
        V2 q = P2<T>::from_f2(__fmul_rn(xf.x, rinv), __fmul_rn(xf.y, rinv));
        q = __hmin2_nan(__hmax2_nan(q, a2), b2);
        q = __hsub2_rn(__hadd2_rn(q, k2), k2);
        p[k] = codes ? __hadd2_rn(q, z2) : __hmul2_rn(s2, q);
      }
    } else {
      apply(v, codes);
    }
  }
  __device__ __forceinline__ void apply(Vec16<T>& v, bool codes) const {
    if (packed) {
      V2* p = reinterpret_cast<V2*>(v.v);
#pragma unroll
      for (int k = 0; k < Vec16<T>::N / 2; ++k) p[k] = one2(p[k], codes);
    } else {
      const float scale = P2<T>::to_f2(s2).x, zp = P2<T>::to_f2(z2).x;      // exact: both were rounded to T
#pragma unroll
      for (int e = 0; e < Vec16<T>::N; ++e) v.v[e] = DT<T>::from_f(qdq_one<T>(DT<T>::to_f(v.v[e]), scale, zp, int_max, codes));
    }
  }
};
// a + b rounded once to the tensor dtype (ATen's add), on a 16-byte vector
template <typename T>
__device__ __forceinline__ void add_vec(Vec16<T>& a, const Vec16<T>& b) {
  if constexpr (P2<T>::kPacked) {
    using V2 = typename P2<T>::V2;
    V2* pa = reinterpret_cast<V2*>(a.v);
    const V2* pb = reinterpret_cast<const V2*>(b.v);
#pragma unroll
    for (int k = 0; k < Vec16<T>::N / 2; ++k) pa[k] = __hadd2_rn(pa[k], pb[k]);
  } else {
#pragma unroll
    for (int e = 0; e < Vec16<T>::N; ++e) a.v[e] = DT<T>::from_f(DT<T>::rnd(__fadd_rn(DT<T>::to_f(a.v[e]), DT<T>::to_f(b.v[e]))));
  }
}

// AsymmetricUniformQuantizer.set_quant_range + scale/zero_point (uniform_quantizers.py:72-82,
// 199-200, 218-219) with the reference's fp32 op order.  Usable on host and device.
__host__ __device__ inline void set_quant_range_f32(float x_min, float x_max, int n_bits, float eps, float* delta,
                                                    float* zero_float, float* scale, float* zp) {
  x_min = x_min < 0.f ? x_min : (x_min != x_min ? x_min : 0.f);   // torch.min(x_min, 0) (NaN propagates)
  x_max = x_max > eps ? x_max : (x_max != x_max ? x_max : eps);   // torch.max(x_max, eps)
  const float int_max = (float)((1 << n_bits) - 1);
#ifdef __CUDA_ARCH__
  const float d = __fdiv_rn(__fsub_rn(x_max, x_min), int_max);
  const float zf = __fdiv_rn(-x_min, d);
  float z = rintf(zf);
#else
  const float d = (x_max - x_min) / int_max;
  const float zf = -x_min / d;
  float z = __builtin_rintf(zf);
#endif
  const float s = d > eps ? d : (d != d ? d : eps);               // clamp(delta, min=eps)
  z = z < 0.f ? 0.f : (z > int_max ? int_max : z);
  *delta = d;
  *zero_float = zf;
  *scale = s;
  *zp = z;
}

}  // namespace oeh
