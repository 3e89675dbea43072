// Definitions shared by the two fused attention kernels (attn_fwd_sm100.cu: one 128-row Q tile per CTA, two CTAs per
// SM; attn_fwd_ws.cuh: warp-specialised, two Q tiles per CTA): launch parameters, the in-kernel fake-quant tap, the
// per-row score transform (scale -> qdq -> +mask -> clamp -> visibility) and the output epilogue.
#pragma once

#include <cfloat>

#include "oeh_common.cuh"
#include "sm100_ptx.cuh"

namespace oeh {
namespace attn {

constexpr int BN = 128;                 // keys per K/V tile
constexpr int HD = 64;                  // head dim (BERT-base / OPT-125m..1.3b)
constexpr int kTileBytes = BN * HD * 2; // 16 KB
constexpr float kLog2e = 1.4426950408889634f;

struct TmapCoord {      // which logical index (0 = token row, 1 = head, 2 = batch) each tensor-map dim 1..3 takes
  int8_t perm[3];
};

struct AttnParams {
  int B, H, Tq, Tk;
  float scale;          // natural-domain score scale
  int kind;
  float stretch, gamma; // clip((eta-gamma)*p + gamma, 0, 1)
  int causal;
  const void* mask;
  int mask_dtype;
  int64_t ms0, ms1, ms2;
  int clamp_min;
  float fmin;
  const float* gate_logits;
  float gate_scale;
  const float* qdq_scores;
  const float* qdq_probs;
  const float* qdq_ctx;
  float int_max;
  int ctx_qdq_before_gate;
  void* out;
  int64_t os0, os1, os2;
  TmapCoord cq, ck, cv;
  int n_qt;             // Q tile groups per (batch, head); work items = n_qt * H * B
  int hb_chunk;         // causal work order: (batch, head) pairs per chunk (see causal_item); 0 = per-head order
  float* dbg_s;         // optional: raw S of K tile 0, (B,H,Tq,BN) fp32
  float* dbg_o;         // optional: raw O accumulator, (B,H,Tq,HD) fp32
};

// Fixed-range fake-quant tap inside the fused kernel.  x/scale is evaluated as x*(1/scale): the tap's input
// already differs from the unfused reference in the last bits (tensor-core accumulation order), so codes can
// flip at bin edges either way; the bit-exact contract belongs to the standalone quantizer (quant.cu).
//   clamp(rint(v) + zp, 0, M) - zp == rint(clamp(v, -zp, M - zp))   (zp is an integer, rint is monotone)
// and inside that range (|v| <= 65535) round-half-even is two FADDs with the 1.5 * 2^23 constant -- no FRND on the XU /
// conversion pipe, which the exponentials of the same loop already load.
struct Qdq {
  float scale, inv_scale, zp, int_max;
  __device__ __forceinline__ float operator()(float x) const {
#ifdef OEH_QDQ_FRND                                            // the literal chain (A/B switch, tools/build_variant.sh)
    float xi = rintf(x * inv_scale) + zp;
    xi = fminf(fmaxf(xi, 0.f), int_max);
    return scale * (xi - zp);
#else
    const float lo = -zp, hi = int_max - zp;                  // loop invariants
    float v = fminf(fmaxf(x * inv_scale, lo), hi);
    v = __fadd_rn(__fadd_rn(v, 12582912.f), -12582912.f);     // rint(v), exact for |v| < 2^22
    return scale * v;
#endif
  }
};
// rint(clamp(x2 * a2, lo, hi)) on two values: the quotient and both halves of the magic-number rounding run as packed
// fp32x2 instructions (FMUL2 / FADD2: one FMA-pipe slot per PAIR), the clamp as scalar FMNMX on the ALU pipe.
__device__ __forceinline__ uint64_t code2(uint64_t x2, uint64_t a2, float lo, float hi) {
  float a, b;
  ptx::unpack2f(ptx::fmul2(x2, a2), a, b);
  a = fminf(fmaxf(a, lo), hi);
  b = fminf(fmaxf(b, lo), hi);
  const uint64_t r2 = ptx::fadd2(ptx::pack2f(a, b), ptx::pack2f(12582912.f, 12582912.f));
  return ptx::fadd2(r2, ptx::pack2f(-12582912.f, -12582912.f));
}

// The probs tap can hand the PV MMA the integer (code - zero_point) itself -- exact in the 16-bit operand format up to
// 2^8 (bf16) / 2^11 (fp16) -- and leave the multiplication by the quantizer's scale to the epilogue:
//     sum_j scale_p (c_j - z) v_j  ==  scale_p * sum_j (c_j - z) v_j.
// One multiply per element less, and the P operand carries NO rounding error at all.
template <bool kBF16>
__device__ __forceinline__ bool probs_as_codes(const AttnParams& p) {
  return p.qdq_probs != nullptr && p.int_max <= (kBF16 ? 255.f : 2047.f);
}

__device__ __forceinline__ Qdq load_qdq(const float* qparams, float int_max) {
  Qdq q;
  q.scale = __ldg(qparams);
  q.zp = __ldg(qparams + 1);
  q.inv_scale = __frcp_rn(q.scale);
  q.int_max = int_max;
  return q;
}

// Causal work order.  Q tile groups of one head differ in length (tile qt attends (qt+1) tile widths of keys), so the
// order in which the hardware (or a persistent CTA) picks them up decides the tail of the launch: items are numbered
// HEAVIEST FIRST over a chunk of `hb_chunk` (batch, head) pairs -- all of the chunk's longest tiles, then the next
// longest, ... -- and chunk after chunk, so that the last CTAs of the grid are the short ones and an SM never gets
// two long tiles while its neighbour gets two short ones (T = 512, 48 heads, 4 tiles each on 148 SMs x 2 slots:
// longest SM 5 tile widths instead of 8).  The chunk bounds the K / V working set of the CTAs in flight (the host
// sizes it to ~32 MB, well inside L2), because tiles of one head share their keys.  Bijective on [0, n_qt*HB).
__device__ __forceinline__ void causal_item(int item, int n_qt, int HB, int hb_chunk, int& qt, int& hb) {
  const int G = hb_chunk > 0 ? hb_chunk : 1;
  const int per_chunk = G * n_qt;
  const int c = item / per_chunk, rem = item - c * per_chunk;
  const int g = min(G, HB - c * G);                    // the last chunk may be smaller
  const int lvl = rem / g;
  qt = n_qt - 1 - lvl;
  hb = c * G + (rem - lvl * g);
}

__device__ __forceinline__ float max4(float a, float b, float c, float d) { return fmaxf(fmaxf(a, b), fmaxf(c, d)); }

__device__ __forceinline__ float load_mask_elem(const void* mask, int dtype, int64_t off) {
  if (dtype == OEH_F32) return __ldg(reinterpret_cast<const float*>(mask) + off);
  if (dtype == OEH_F16) return __half2float(__ldg(reinterpret_cast<const __half*>(mask) + off));
  return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(mask) + off));
}

template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (kBF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

__device__ __forceinline__ void tma_tile(uint32_t dst, const CUtensorMap* m, uint32_t bar, const TmapCoord& c, int row,
                                         int h, int b) {
  auto pick = [&](int k) { return k == 0 ? row : (k == 1 ? h : b); };
  ptx::tma_load_4d(dst, m, bar, 0, pick(c.perm[0]), pick(c.perm[1]), pick(c.perm[2]));
}


// Per-row score transform of one 128-key tile, in the reference's op order.  On return s[j]*mul is the log2-domain
// score of key k0+j (-inf where the key is out of range / causally hidden) and the function returns mul.
//   plain   : s stays the raw MMA output, mul = scale*log2e (scale multiply and max subtraction fold into one FFMA)
//   generic : scale -> fake-quant tap -> + additive mask -> clamp(finfo.min) in natural units, then * log2e, mul = 1
//             (log2e is applied explicitly: with finfo.min-sized masked scores the rounding error of a folded multiply
//             -- exact product inside the FFMA vs the rounded tile maximum -- would be ~1e31).
template <int N, bool kGeneric>
__device__ __forceinline__ float transform_scores(float (&s)[N], const AttnParams& p, const Qdq& qd_s, bool has_qs,
                                                  bool has_mask, int64_t mask_row, int k0, int valid, int vis) {
  float mul = p.scale * kLog2e;
  if (kGeneric && has_qs && !has_mask && !p.clamp_min) {
    // Scores tap with nothing between it and the softmax (causal flag / no mask: the BERT eval data and the causal-hint OPT
    // path): keep the INTEGER c = code - zero_point, i.e. s_q = scale_q * c, and fold scale_q * log2e into the multiplier
    // the exponential applies anyway.  3.5 issue slots per element (packed) instead of 9.
    const float a = p.scale * qd_s.inv_scale, lo = -qd_s.zp, hi = qd_s.int_max - qd_s.zp;
    const uint64_t a2 = ptx::pack2f(a, a);
#pragma unroll
    for (int j = 0; j < N; j += 2) ptx::unpack2f(code2(ptx::pack2f(s[j], s[j + 1]), a2, lo, hi), s[j], s[j + 1]);
    mul = qd_s.scale * kLog2e;
  } else if (kGeneric && (has_qs || has_mask || p.clamp_min)) {
    mul = 1.f;
    const float scale = p.scale;
#pragma unroll
    for (int j = 0; j < N; ++j) s[j] *= scale;
    if (has_qs) {
#pragma unroll
      for (int j = 0; j < N; ++j) s[j] = qd_s(s[j]);
    }
    if (has_mask) {
      if (p.mask_dtype == OEH_F32) {
        const float* mp = reinterpret_cast<const float*>(p.mask) + mask_row + k0;
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (j < valid) s[j] += __ldg(mp + j);
      } else if (p.mask_dtype == OEH_F16) {
        const __half* mp = reinterpret_cast<const __half*>(p.mask) + mask_row + k0;
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (j < valid) s[j] += __half2float(__ldg(mp + j));
      } else {
        const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(p.mask) + mask_row + k0;
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (j < valid) s[j] += __bfloat162float(__ldg(mp + j));
      }
    }
    // clamp (opt_attention.py:221-223); the lower bound also keeps finfo.min-masked scores finite after the log2e
    // multiply, so a row masked entirely with finfo.min stays uniform like the reference's
    const float lo = p.clamp_min ? fmaxf(p.fmin, -FLT_MAX / 2.f) : -FLT_MAX / 2.f;
#pragma unroll
    for (int j = 0; j < N; ++j) s[j] = fmaxf(s[j], lo) * kLog2e;
  }
  if (vis < N - 1) {                               // ragged last tile / causal diagonal tile
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j > vis) s[j] = -INFINITY;
  }
  return mul;
}

template <int N>
__device__ __forceinline__ float row_max(const float (&s)[N]) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;   // 4 independent chains
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    m0 = fmaxf(m0, s[j]);
    m1 = fmaxf(m1, s[j + 1]);
    m2 = fmaxf(m2, s[j + 2]);
    m3 = fmaxf(m3, s[j + 3]);
  }
  return max4(m0, m1, m2, m3);
}

// exp2 of two log2-domain values on the FMA / ALU pipes instead of the MUFU (XU) pipe: at d = 64 the attention loop is
// bound by MUFU.EX2 (16 results/clk/SM against 512 tensor-pipe clocks per 128x128 tile), while the FMA pipe is
// mostly idle, so a fixed share of every row's exponentials is evaluated as
//     2^t = 2^n * p(f),  n = round(t),  f = t - n in [-0.5, 0.5],  p = minimax cubic (max rel. error 7.5e-5,
// below the 2^-12 rounding of the fp16 / bf16 P operand it feeds), 2^n applied by adding n to the exponent bits.
// Cost: 11 issue slots per PAIR (packed fp32x2 arithmetic) instead of 2 MUFU slots of 8 XU clocks each.
__device__ __forceinline__ void exp2_poly2(uint64_t t2, float& e0, float& e1) {
  float a, b;
  ptx::unpack2f(t2, a, b);
  a = fmaxf(a, -126.f);                                // keeps 2^n a normal number; also maps -inf (masked) to ~0
  b = fmaxf(b, -126.f);
  t2 = ptx::pack2f(a, b);
  const uint64_t r2 = ptx::fadd2(t2, ptx::pack2f(12582912.f, 12582912.f));        // 1.5 * 2^23: n in the low mantissa bits
  const uint64_t n2 = ptx::fadd2(r2, ptx::pack2f(-12582912.f, -12582912.f));
  const uint64_t f2 = ptx::ffma2(n2, ptx::pack2f(-1.f, -1.f), t2);
  uint64_t p2 = ptx::ffma2(f2, ptx::pack2f(0.0551716685f, 0.0551716685f), ptx::pack2f(0.242611125f, 0.242611125f));
  p2 = ptx::ffma2(p2, f2, ptx::pack2f(0.693260968f, 0.693260968f));
  p2 = ptx::ffma2(p2, f2, ptx::pack2f(0.999928057f, 0.999928057f));
  float ra, rb, pa, pb;
  ptx::unpack2f(r2, ra, rb);
  ptx::unpack2f(p2, pa, pb);
  e0 = __int_as_float(__float_as_int(pa) + (__float_as_int(ra) << 23));
  e1 = __int_as_float(__float_as_int(pb) + (__float_as_int(rb) << 23));
}

#ifndef OEH_POLY_PAIRS
#define OEH_POLY_PAIRS 3     // pairs out of every 16 (= 32 columns) that take the polynomial path
#endif
// pair pi of a row chunk goes to the FMA pipe if this is true (evenly spread over the chunk)
__host__ __device__ constexpr bool poly_pair(int pi) { return ((pi % 16) * OEH_POLY_PAIRS) % 16 < OEH_POLY_PAIRS; }

// kPoly = false: every exponential on the MUFU pipe.  The generic kernels (fake-quant taps / masks) use that: with the
// probs tap a 7.5e-5 relative error moves P across quantisation bin edges more often than the ex2.approx one does.
template <bool kPoly>
__device__ __forceinline__ void exp2_pair(uint64_t t2, int pi, float& e0, float& e1) {
  if (kPoly && poly_pair(pi)) {
    exp2_poly2(t2, e0, e1);
  } else {
    float a, b;
    ptx::unpack2f(t2, a, b);
    e0 = ptx::ex2(a);
    e1 = ptx::ex2(b);
  }
}

// s[j] <- exp2(s[j]*mul - m); returns the row sum.  The scale/subtract FMAs and the sum run as packed fp32x2
// instructions (FFMA2 / FADD2: half the issue slots, same per-lane IEEE results); 4 independent partial sums.
template <int N, bool kPoly = true>
__device__ __forceinline__ float exp2_and_sum(float (&s)[N], float mul, float m) {
  const uint64_t mul2 = ptx::pack2f(mul, mul), nm2 = ptx::pack2f(-m, -m);
  uint64_t acc0 = ptx::pack2f(0.f, 0.f), acc1 = acc0;
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    exp2_pair<kPoly>(ptx::ffma2(ptx::pack2f(s[j], s[j + 1]), mul2, nm2), j / 2, s[j], s[j + 1]);
    exp2_pair<kPoly>(ptx::ffma2(ptx::pack2f(s[j + 2], s[j + 3]), mul2, nm2), j / 2 + 1, s[j + 2], s[j + 3]);
    acc0 = ptx::fadd2(acc0, ptx::pack2f(s[j], s[j + 1]));
    acc1 = ptx::fadd2(acc1, ptx::pack2f(s[j + 2], s[j + 3]));
  }
  float x, y;
  ptx::unpack2f(ptx::fadd2(acc0, acc1), x, y);
  return x + y;
}

// s[j] <- exp2(s[j]*mul - m) without the sum (PV pass of the two-pass modes)
template <int N, bool kPoly = true>
__device__ __forceinline__ void exp2_only(float (&s)[N], float mul, float m) {
  const uint64_t mul2 = ptx::pack2f(mul, mul), nm2 = ptx::pack2f(-m, -m);
#pragma unroll
  for (int j = 0; j < N; j += 2)
    exp2_pair<kPoly>(ptx::ffma2(ptx::pack2f(s[j], s[j + 1]), mul2, nm2), j / 2, s[j], s[j + 1]);
}

// normalised modes: p = e * inv_den, then clip((eta-gamma)*p + gamma, 0, 1) and / or the probs fake-quant tap.
// With the tap and a code range that is exact in the operand format (probs_as_codes) the result is the integer
// (code - zero_point); the epilogue multiplies the context by the tap's scale (store_row).
template <int N, bool kGeneric, bool kBF16>
__device__ __forceinline__ void normalize_probs(float (&s)[N], const AttnParams& p, float inv_den, bool clipped,
                                                bool has_qp, const Qdq& qd_p) {
  const bool codes = kGeneric && has_qp && probs_as_codes<kBF16>(p);
  if (clipped) {
    const float st = p.stretch * inv_den, gm = p.gamma;
#pragma unroll
    for (int j = 0; j < N; ++j) s[j] = __saturatef(fmaf(s[j], st, gm));
  } else if (!codes) {
#pragma unroll
    for (int j = 0; j < N; ++j) s[j] *= inv_den;
  }
  if (codes) {
    const float a = clipped ? qd_p.inv_scale : inv_den * qd_p.inv_scale, lo = -qd_p.zp, hi = qd_p.int_max - qd_p.zp;
    const uint64_t a2 = ptx::pack2f(a, a);
#pragma unroll
    for (int j = 0; j < N; j += 2) ptx::unpack2f(code2(ptx::pack2f(s[j], s[j + 1]), a2, lo, hi), s[j], s[j + 1]);
  } else if (kGeneric && has_qp) {
#pragma unroll
    for (int j = 0; j < N; ++j) s[j] = qd_p(s[j]);
  }
}

// Normalised modes: log2-domain bound below which an unnormalised e = 2^t is GUARANTEED to end as P == 0 exactly:
//   clipped:  clip(stretch*p + gamma, 0, 1) == 0            <=>  p <= -gamma/stretch            (gamma < 0 < stretch)
//   probs tap: qdq(x) == 0  <=>  rint(x/scale) == 0          <=   x <  scale/2                   (zero_point is in range)
// with p = e/den.  1 % safety margin (ex2.approx and the reciprocal are good to ~2^-22).  -inf = nothing can be skipped.
__device__ __forceinline__ float zero_prob_threshold(const AttnParams& p, bool clipped, bool has_qp, const Qdq& qd_p,
                                                     float den) {
  const float x0 = has_qp ? 0.49f * qd_p.scale : 0.f;          // post-clip values below x0 quantise to 0
  float pmin;
  if (clipped) {
    if (!(p.stretch > 0.f)) return -INFINITY;
    pmin = (x0 - p.gamma) / p.stretch;
  } else {
    pmin = x0;
  }
  pmin *= 0.99f;
  if (!(pmin > 0.f) || !(den > 0.f) || !(den < INFINITY)) return -INFINITY;
  return log2f(pmin * den);
}

// Epilogue of one query row (ND = 64: the whole row; ND = 32: the half starting at head-dim column d0):
// o[] holds the PV accumulator; mul = 1/denominator (1 in the normalised modes).
// ctx -> [qdq before gate] -> * sigmoid(gate logit)*gate_scale -> [qdq after gate] -> ND/8 x 16-byte stores.
// What a row's epilogue reads from global memory -- the probs tap's scale, the gate logit, the context tap's parameters --
// fetched BEFORE the wait for the last PV (epi_prefetch), so that these L2 round trips (~0.4 us, serialised behind the
// accumulator in profiles/r02_ncu_hot_kernels.txt: 4 % of the mid kernel's samples sat on the gate's sigmoid) overlap it.
struct EpiPre {
  float pscale, gate;
  Qdq qd_c;
};
template <bool kBF16, bool kGeneric>
__device__ __forceinline__ EpiPre epi_prefetch(const AttnParams& p, int b, int h, int row) {
  EpiPre e{1.f, 1.f, Qdq{1.f, 1.f, 0.f, p.int_max}};
  if (kGeneric && probs_as_codes<kBF16>(p)) e.pscale = __ldg(p.qdq_probs);   // P was (code - zp): see normalize_probs
  if (p.gate_logits && row < p.Tq) {
    const float z = __ldg(p.gate_logits + ((int64_t)b * p.H + h) * p.Tq + row);
    e.gate = __frcp_rn(1.f + expf(-z)) * p.gate_scale;
  }
  if (kGeneric && p.qdq_ctx) e.qd_c = load_qdq(p.qdq_ctx, p.int_max);
  return e;
}

template <bool kBF16, bool kGeneric, int ND = HD>
__device__ __forceinline__ void finish_row(float (&o)[ND], const AttnParams& p, float mul, int b, int h, int row, int d0,
                                           const EpiPre& pre) {
  if (kGeneric && p.dbg_o) {
    float* d = p.dbg_o + (((int64_t)b * p.H + h) * p.Tq + row) * HD + d0;
#pragma unroll
    for (int j = 0; j < ND; ++j) d[j] = o[j];
  }
  mul *= pre.pscale;
  const float gate = pre.gate;
  if (kGeneric && p.qdq_ctx) {
    const Qdq qd_c = pre.qd_c;
    if (p.ctx_qdq_before_gate) {                       // OPT: tap on PV, then gate (quantized_opt.py:210)
#pragma unroll
      for (int j = 0; j < ND; ++j) o[j] = qd_c(o[j] * mul) * gate;
    } else {                                           // BERT: gate, then tap (quantized_bert.py:434)
#pragma unroll
      for (int j = 0; j < ND; ++j) o[j] = qd_c(o[j] * mul * gate);
    }
  } else {
    mul *= gate;
#pragma unroll
    for (int j = 0; j < ND; ++j) o[j] *= mul;
  }
}

template <bool kBF16, bool kGeneric, int ND = HD>
__device__ __forceinline__ void finish_row(float (&o)[ND], const AttnParams& p, float mul, int b, int h, int row,
                                           int d0 = 0) {
  finish_row<kBF16, kGeneric, ND>(o, p, mul, b, h, row, d0, epi_prefetch<kBF16, kGeneric>(p, b, h, row));
}

template <bool kBF16, bool kGeneric, int ND = HD>
__device__ __forceinline__ void store_row(float (&o)[ND], const AttnParams& p, float mul, int b, int h, int row, int d0,
                                          const EpiPre& pre) {
  finish_row<kBF16, kGeneric, ND>(o, p, mul, b, h, row, d0, pre);
  uint8_t* optr = reinterpret_cast<uint8_t*>(p.out) +
                  2 * ((int64_t)b * p.os0 + (int64_t)h * p.os1 + (int64_t)row * p.os2 + d0);
  if constexpr (ND == 16) {
    // a quarter row is one 32-byte sector: written whole by one 256-bit store when it is aligned (sm_100 STG.256), instead of
    // as two half-sector pieces
    if ((reinterpret_cast<uintptr_t>(optr) & 31) == 0) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = pack2<kBF16>(o[2 * j], o[2 * j + 1]);
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(optr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                   "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
      return;
    }
  }
#pragma unroll
  for (int c = 0; c < ND / 8; ++c) {
    uint4 v;
    v.x = pack2<kBF16>(o[c * 8 + 0], o[c * 8 + 1]);
    v.y = pack2<kBF16>(o[c * 8 + 2], o[c * 8 + 3]);
    v.z = pack2<kBF16>(o[c * 8 + 4], o[c * 8 + 5]);
    v.w = pack2<kBF16>(o[c * 8 + 6], o[c * 8 + 7]);
    reinterpret_cast<uint4*>(optr)[c] = v;
  }
}

template <bool kBF16, bool kGeneric, int ND = HD>
__device__ __forceinline__ void store_row(float (&o)[ND], const AttnParams& p, float mul, int b, int h, int row,
                                          int d0 = 0) {
  store_row<kBF16, kGeneric, ND>(o, p, mul, b, h, row, d0, epi_prefetch<kBF16, kGeneric>(p, b, h, row));
}

__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// one whole 32-byte sector per thread (sm_100 STG.256); ptr must be 32-byte aligned
__device__ __forceinline__ void st_global_256(void* ptr, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
               "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

// Warp-cooperative store of 32 finished context rows (lane == row, 64 values each; the caller has applied finish_row).
// A thread storing its own 128-byte row sends 16 bytes to 32 different lines per instruction -- 256 partial-line requests
// per warp, ~1 us at the end of every one-tile CTA (profiles/r02_timeline_one.txt).  Here the rows pass through the warp's
// own 4 KB of shared memory (`stage`; 128-byte swizzle, conflict free both ways) and leave as FULL lines: 8 lanes per row,
// 4 rows per instruction, 8 instructions.  Only __syncwarp() -- no CTA barrier.  row0 = global query row of lane 0.
template <bool kBF16>
__device__ __forceinline__ void store_rows_coalesced(const float (&o)[HD], const AttnParams& p, uint32_t stage, int b, int h,
                                                     int row0) {
  const uint32_t lane = threadIdx.x & 31u;
  {
    const uint32_t dst = stage + lane * 128u, sw = lane & 7u;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c)
      ptx::st_shared_v4(dst + (((uint32_t)c ^ sw) << 4), pack2<kBF16>(o[c * 8 + 0], o[c * 8 + 1]),
                        pack2<kBF16>(o[c * 8 + 2], o[c * 8 + 3]), pack2<kBF16>(o[c * 8 + 4], o[c * 8 + 5]),
                        pack2<kBF16>(o[c * 8 + 6], o[c * 8 + 7]));
  }
  __syncwarp();
  uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + 2 * ((int64_t)b * p.os0 + (int64_t)h * p.os1);
  if (((reinterpret_cast<uintptr_t>(obase) | (uintptr_t)(2 * p.os2)) & 31) == 0) {
    // 32-byte sectors: 4 lanes per row, 8 rows per instruction, 4 instructions (256-bit stores)
    const uint32_t cc = lane & 3u, rsub = lane >> 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t rr = (uint32_t)i * 8u + rsub, base = stage + rr * 128u, s7 = rr & 7u;
      const uint4 lo = ld_shared_u4(base + (((2u * cc) ^ s7) << 4)), hi = ld_shared_u4(base + (((2u * cc + 1u) ^ s7) << 4));
      const int row = row0 + (int)rr;
      if (row < p.Tq) st_global_256(obase + 2 * (int64_t)row * p.os2 + cc * 32u, lo, hi);
    }
    return;
  }
  const uint32_t cc = lane & 7u, rsub = lane >> 3;
  obase += cc * 16u;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t rr = (uint32_t)i * 4u + rsub;
    const uint4 v = ld_shared_u4(stage + rr * 128u + ((cc ^ (rr & 7u)) << 4));
    const int row = row0 + (int)rr;
    if (row < p.Tq) *reinterpret_cast<uint4*>(obase + 2 * (int64_t)row * p.os2) = v;
  }
}

// 32 rows x 32 context columns (64 bytes per row) of one warp -> global, as full 64-byte segments: 4 lanes per row, 8 rows per
// instruction (store_rows_coalesced of attn_common.cuh for half rows).  stage: the warp's own 2 KB of shared memory.
template <bool kBF16>
__device__ __forceinline__ void store_half_rows_coalesced(const float (&o)[32], const AttnParams& p, uint32_t stage, int b, int h,
                                                          int row0, int d0) {
  const uint32_t lane = threadIdx.x & 31u;
  {
    const uint32_t dst = stage + lane * 64u, sw = (lane >> 1) & 3u;     // 16-byte chunk c of row r at ((c ^ (r/2 & 3)) << 4)
#pragma unroll
    for (int c = 0; c < 4; ++c)
      ptx::st_shared_v4(dst + (((uint32_t)c ^ sw) << 4), pack2<kBF16>(o[c * 8 + 0], o[c * 8 + 1]),
                        pack2<kBF16>(o[c * 8 + 2], o[c * 8 + 3]), pack2<kBF16>(o[c * 8 + 4], o[c * 8 + 5]),
                        pack2<kBF16>(o[c * 8 + 6], o[c * 8 + 7]));
  }
  __syncwarp();
  uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + 2 * ((int64_t)b * p.os0 + (int64_t)h * p.os1 + d0);
  if (((reinterpret_cast<uintptr_t>(obase) | (uintptr_t)(2 * p.os2)) & 31) == 0) {
    // 32-byte sectors: 2 lanes per half row, 16 rows per instruction, 2 instructions (256-bit stores)
    const uint32_t cc = lane & 1u, rsub = lane >> 1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t rr = (uint32_t)i * 16u + rsub, base = stage + rr * 64u, s3 = (rr >> 1) & 3u;
      const uint4 lo = ld_shared_u4(base + (((2u * cc) ^ s3) << 4)), hi = ld_shared_u4(base + (((2u * cc + 1u) ^ s3) << 4));
      const int row = row0 + (int)rr;
      if (row < p.Tq) st_global_256(obase + 2 * (int64_t)row * p.os2 + cc * 32u, lo, hi);
    }
    return;
  }
  const uint32_t cc = lane & 3u, rsub = lane >> 2;
  obase += cc * 16u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t rr = (uint32_t)i * 8u + rsub;
    const uint4 v = ld_shared_u4(stage + rr * 64u + ((cc ^ ((rr >> 1) & 3u)) << 4));
    const int row = row0 + (int)rr;
    if (row < p.Tq) *reinterpret_cast<uint4*>(obase + 2 * (int64_t)row * p.os2) = v;
  }
}

}  // namespace attn
}  // namespace oeh
