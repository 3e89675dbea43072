/* CPU oracle (C restatement) for the OutEffHop attention + fake-quant hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product (outeffhop_b200/) never
 * touches it.
 *
 * Plain C99 + OpenMP, fp32, no fast-math, no FMA contraction (built with
 * -ffp-contract=off) so that the quantizer / range functions are bit-identical to the
 * reference's ATen fp32 ops.  Citations are relative to /root/reference/OutEffHop.
 * Pinned against the live reference through tests/golden/ (see oracle/oeh_oracle.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { KIND_VANILLA = 0, KIND_SOFTMAX1 = 1, KIND_CLIPPED = 2, KIND_CLIPPED1 = 3 };

int oeh_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oeh_ref_set_threads(int n) {   /* torchrun exports OMP_NUM_THREADS=1: the bench sets the count explicitly */
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* One row of softmax.py / vutils/softmax_1.py, in place.
 * vanilla  : e/sum(e)                               (softmax.py:23, F.softmax)
 * softmax1 : e/(sum(e) + exp(-m))                   (vutils/softmax_1.py:11-21)
 * clipped* : clip(p*(eta-gamma)+gamma, 0, 1)        (softmax.py:10-19)
 * stretch = (float)(eta-gamma) and gamma are converted from double by the caller, the way
 * Python scalars are converted at the tensor multiply/add. */
static void softmax_row(float* s, int n, int kind, float stretch, float gamma) {
  float m = -INFINITY;
  for (int j = 0; j < n; ++j) m = s[j] > m ? s[j] : m;
  float sum = 0.f;
  for (int j = 0; j < n; ++j) {
    s[j] = expf(s[j] - m);
    sum += s[j];
  }
  float den = sum;
  if (kind == KIND_SOFTMAX1 || kind == KIND_CLIPPED1) den = sum + expf(m * -1.0f) * 1.0f;
  for (int j = 0; j < n; ++j) {
    float p = s[j] / den;
    if (kind >= KIND_CLIPPED) {
      p = p * stretch + gamma;
      p = p < 0.f ? 0.f : (p > 1.f ? 1.f : p);
    }
    s[j] = p;
  }
}

void oeh_ref_softmax_rows(const float* x, float* y, int64_t rows, int cols, int kind, double gamma,
                          double eta) {
  const float stretch = (float)(eta - gamma), g = (float)gamma;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    float* yr = y + r * cols;
    memcpy(yr, x + r * cols, sizeof(float) * (size_t)cols);
    softmax_row(yr, cols, kind, stretch, g);
  }
}

/* Attention core: S = (Q K^T) * scale (+causal finfo.min) (+mask) (clamp) -> softmax_fn -> P V -> gate.
 * bert_attention.py:222,265,270-276,292,294-331 / opt_attention.py:204,215-232,263,276-313.
 * q,k,v: element strides (b,h,t) with d contiguous; out likewise (so (B,T,H*d) output is a stride choice).
 * mask: additive fp32, element strides (b,h,q) with k contiguous, any stride may be 0; NULL = none.
 * gate: sigmoid output (B,H,Tq) fp32 contiguous or NULL; ctx *= gate*gate_scale. */
void oeh_ref_attention(const float* q, const float* k, const float* v, float* out, int B, int H, int Tq,
                       int Tk, int d, const int64_t* qs, const int64_t* ks, const int64_t* vs,
                       const int64_t* os, float scale, int kind, double gamma, double eta, int causal,
                       const float* mask, const int64_t* ms, int clamp_min, const float* gate,
                       float gate_scale) {
  const float stretch = (float)(eta - gamma), g = (float)gamma;
#pragma omp parallel
  {
    float* s = (float*)malloc(sizeof(float) * (size_t)Tk);
#pragma omp for collapse(2) schedule(dynamic, 8)
    for (int bh = 0; bh < B * H; ++bh) {
      for (int i = 0; i < Tq; ++i) {
        const int b = bh / H, h = bh % H;
        const float* qi = q + b * qs[0] + h * qs[1] + i * qs[2];
        for (int j = 0; j < Tk; ++j) {
          const float* kj = k + b * ks[0] + h * ks[1] + j * ks[2];
          float acc = 0.f;
          for (int c = 0; c < d; ++c) acc += qi[c] * kj[c];
          acc = acc * scale;
          if (causal && j > i + (Tk - Tq)) acc = acc + -FLT_MAX; /* HF-4.31 causal mask value */
          if (mask) acc = acc + mask[b * ms[0] + h * ms[1] + i * ms[2] + j];
          if (clamp_min && acc < -FLT_MAX) acc = -FLT_MAX;        /* opt_attention.py:221-223 */
          s[j] = acc;
        }
        softmax_row(s, Tk, kind, stretch, g);
        float* o = out + b * os[0] + h * os[1] + i * os[2];
        for (int c = 0; c < d; ++c) o[c] = 0.f;
        for (int j = 0; j < Tk; ++j) {
          const float p = s[j];
          if (p == 0.f) continue;
          const float* vj = v + b * vs[0] + h * vs[1] + j * vs[2];
          for (int c = 0; c < d; ++c) o[c] += p * vj[c];
        }
        if (gate) {
          const float gg = gate[((int64_t)b * H + h) * Tq + i] * gate_scale;
          for (int c = 0; c < d; ++c) o[c] *= gg;
        }
      }
    }
    free(s);
  }
}

/* AsymmetricUniformQuantizer.set_quant_range (uniform_quantizers.py:173-224):
 * out[0]=_delta, out[1]=_zero_float, out[2]=scale (:72-76), out[3]=zero_point (:79-82). */
void oeh_ref_set_quant_range(float x_min, float x_max, int n_bits, float eps, float* out4) {
  x_min = x_min < 0.f ? x_min : 0.f;       /* :199 */
  x_max = x_max > eps ? x_max : eps;       /* :200 */
  const float int_max = (float)((1 << n_bits) - 1);
  const float delta = (x_max - x_min) / int_max;  /* :218 */
  const float zero_float = -x_min / delta;        /* :219 */
  const float scale = delta > eps ? delta : eps;
  float zp = rintf(zero_float); /* round-half-even under the default rounding mode */
  zp = zp < 0.f ? 0.f : (zp > int_max ? int_max : zp);
  out4[0] = delta; out4[1] = zero_float; out4[2] = scale; out4[3] = zp;
}

/* quantize-dequantize, fp32 (uniform_quantizers.py:114-115,146); codes optional. */
void oeh_ref_fake_quant(const float* x, float* y, float* codes, int64_t n, float scale, float zp, int n_bits) {
  const float int_max = (float)((1 << n_bits) - 1);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    float xi = rintf(x[i] / scale) + zp;
    xi = xi < 0.f ? 0.f : (xi > int_max ? int_max : xi);
    if (xi != xi) xi = x[i] / scale; /* NaN propagates like torch.clamp */
    if (codes) codes[i] = xi;
    y[i] = scale * (xi - zp);
  }
}

/* torch.min / torch.max over the whole tensor (range_estimators.py:96-97). */
void oeh_ref_minmax(const float* x, int64_t n, float* out2) {
  float lo = INFINITY, hi = -INFINITY;
  int has_nan = 0;
#pragma omp parallel for reduction(min : lo) reduction(max : hi) reduction(| : has_nan) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float t = x[i];
    has_nan |= (t != t);
    lo = t < lo ? t : lo;
    hi = t > hi ? t : hi;
  }
  if (has_nan) lo = hi = NAN;
  out2[0] = lo; out2[1] = hi;
}

/* RunningMinMaxEstimator EMA (range_estimators.py:99-104); cur2 is updated in place. */
void oeh_ref_running_minmax_update(float* cur2, const float* new2, double momentum, int initialized) {
  if (!initialized) {
    cur2[0] = new2[0]; cur2[1] = new2[1];
    return;
  }
  const float one_minus = (float)(1.0 - momentum), mom = (float)momentum;
  const float a0 = one_minus * new2[0], b0 = mom * cur2[0];
  const float a1 = one_minus * new2[1], b1 = mom * cur2[1];
  cur2[0] = a0 + b0;
  cur2[1] = a1 + b1;
}

/* Per-token gate logits, Linear(d,1) per head on the attention INPUT split by head
 * (bert_attention.py:311-320). x: (B,T,H*d) contiguous; w: (H,d); bias: (H); out: (B,H,T). */
void oeh_ref_gate_logits(const float* x, const float* w, const float* bias, float* out, int B, int T, int H, int d) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t)
      for (int h = 0; h < H; ++h) {
        const float* xr = x + ((int64_t)b * T + t) * H * d + h * d;
        float acc = 0.f;
        for (int c = 0; c < d; ++c) acc += xr[c] * w[h * d + c];
        out[((int64_t)b * H + h) * T + t] = acc + bias[h];
      }
}
