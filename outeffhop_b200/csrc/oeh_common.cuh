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
