// erf GELU of two fp32 values (shared by the standalone kernel, gelu.cu, and the GEMM epilogue, gemm_sm100.cu); see gelu.cu
// for the derivation:  gelu(x) = max(x, 0) - |x| * Phi(-|x|),  Phi(-|x|) = 0.5 * t * exp(-x^2/2 + P(t)),  t = 1 / (1 + |x| / (2 sqrt 2)).
#pragma once

#include "sm100_ptx.cuh"

namespace oeh {
namespace gelu_math {

constexpr float kLog2e = 1.4426950408889634f;
// log2(e) * P(t) - 1  (the -1 is the factor 0.5 of Phi)
constexpr float kC0 = -1.26551223f * kLog2e - 1.f, kC1 = 1.00002368f * kLog2e, kC2 = 0.37409196f * kLog2e,
                kC3 = 0.09678418f * kLog2e, kC4 = -0.18628806f * kLog2e, kC5 = 0.27886807f * kLog2e,
                kC6 = -1.13520398f * kLog2e, kC7 = 1.48851587f * kLog2e, kC8 = -0.82215223f * kLog2e,
                kC9 = 0.17087277f * kLog2e;

__device__ __forceinline__ uint64_t bc2(float c) { return ptx::pack2f(c, c); }
__device__ __forceinline__ float rcp_approx(float x) {       // MUFU.RCP (1 ulp); the operand is a normal number in [1, 8.1]
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 1 / d for two values in [1, 8.1] WITHOUT the MUFU pipe: seed from the bit pattern (relative error <= 12.5 %), three Newton steps
// y <- y + y (1 - d y) as packed fp32x2 FMAs (error squared per step: ~6e-8, the size of MUFU.RCP's own).  For the GEMM epilogue,
// where the exponentials and reciprocals of a 128 x 256 tile (2 MUFU ops per element, 16 per clock and SM) were HALF of the
// epilogue time (profiles/r02_timeline_gemm.txt).
__device__ __forceinline__ uint64_t rcp2_fma(uint64_t d2) {
  float d0, d1;
  ptx::unpack2f(d2, d0, d1);
  uint64_t y2 = ptx::pack2f(__int_as_float(0x7EF311C7 - __float_as_int(d0)), __int_as_float(0x7EF311C7 - __float_as_int(d1)));
  const uint64_t nd2 = d2 ^ 0x8000000080000000ull;                       // -d
#pragma unroll
  for (int it = 0; it < 3; ++it) y2 = ptx::ffma2(y2, ptx::ffma2(nd2, y2, bc2(1.f)), y2);
  return y2;
}

// gelu of two fp32 values.  kFmaRcp: the reciprocal on the FMA pipe (rcp2_fma) instead of MUFU.RCP.
template <bool kFmaRcp = false>
__device__ __forceinline__ void gelu2(float x0, float x1, float& y0, float& y1) {
  const float na0 = -fabsf(x0), na1 = -fabsf(x1);                        // -|x|
  // beyond |x| = 20 Phi(-|x|) < 1e-88 is 0 in any format: cap the argument there, so that |x| * Phi stays finite for inf (gelu(+inf) = inf; NaN still propagates through the x^2 term below)
  const float n0 = fmaxf(na0, -20.f), n1 = fmaxf(na1, -20.f);
  const uint64_t n2 = ptx::pack2f(n0, n1);
  const uint64_t r2 = ptx::pack2f(fmaxf(x0, 0.f), fmaxf(x1, 0.f));       // max(x, 0)
  const uint64_t d2 = ptx::ffma2(n2, bc2(-0.35355339059327373f), bc2(1.f));   // 1 + |x| / (2 sqrt 2)  in [1, 8.1]
  uint64_t t2;
  if constexpr (kFmaRcp) {
    t2 = rcp2_fma(d2);
  } else {
    float d0, d1;
    ptx::unpack2f(d2, d0, d1);
    t2 = ptx::pack2f(rcp_approx(d0), rcp_approx(d1));                    // every element on its own: a pure function of x
  }
  uint64_t p2 = ptx::ffma2(t2, bc2(kC9), bc2(kC8));
  p2 = ptx::ffma2(p2, t2, bc2(kC7));
  p2 = ptx::ffma2(p2, t2, bc2(kC6));
  p2 = ptx::ffma2(p2, t2, bc2(kC5));
  p2 = ptx::ffma2(p2, t2, bc2(kC4));
  p2 = ptx::ffma2(p2, t2, bc2(kC3));
  p2 = ptx::ffma2(p2, t2, bc2(kC2));
  p2 = ptx::ffma2(p2, t2, bc2(kC1));
  p2 = ptx::ffma2(p2, t2, bc2(kC0));
  const uint64_t sq2 = ptx::ffma2(ptx::pack2f(na0, na1), n2, bc2(0.f));   // x^2 (|x| * min(|x|, 20); NaN stays NaN)
  const uint64_t a2 = ptx::ffma2(sq2, bc2(-0.5f * kLog2e), p2);           // log2 of 0.5 * erfc / t
  float a0, a1;
  ptx::unpack2f(a2, a0, a1);
  const uint64_t e2 = ptx::pack2f(ptx::ex2(a0), ptx::ex2(a1));
  const uint64_t h2 = ptx::ffma2(t2, e2, bc2(0.f));                       // Phi(-|x|)
  ptx::unpack2f(ptx::ffma2(n2, h2, r2), y0, y1);                          // max(x,0) - |x| Phi(-|x|)
}

}  // namespace gelu_math
}  // namespace oeh
