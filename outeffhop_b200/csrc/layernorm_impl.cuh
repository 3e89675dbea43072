// Residual add + LayerNorm in one pass, with the optional fake-quant taps of the INT8 path fused in:
//   FP  :  y = LayerNorm(a + b)                      BertSelfOutput / BertOutput.forward (HF shell of the eval step; the
//                                                    module the reference wraps at quantized_bert.py:540-605)
//   INT8:  y = qdq_out(LayerNorm(qdq_res(qdq_a(a) + b)))   QuantizedBertSelfOutput / QuantizedBertOutput.forward with fixed
//                                                    ranges (quantized_bert.py:566-568, :602-604): res_act_quantizer on
//                                                    the sum, QuantLayerNorm's output tap (autoquant_utils.py:62-70);
//                                                    qdq_a = the output tap of the QuantLinear `dense` that produced a
// ATen needs add (3 passes) + LayerNorm (2 passes) [+ 2 x ~11 passes of qdq elementwise ops]; here every element of a, b
// is read once and y written once: bytes = rows*cols*3*size (2*size without b).  HBM-bound.
// One warp per row, the row lives in registers (<= 8 x 16 B per lane): exact two-pass mean / variance in fp32.
// Arithmetic mirrors ATen: sum rounded once to the tensor dtype (add kernel), statistics and the affine map in fp32,
// y = gamma * (rstd * (x - mean)) + beta rounded once to the tensor dtype (vectorized_layer_norm_kernel); taps as in
// quant.cu (every op rounded to the tensor dtype).
#pragma once

#include "oeh_common.cuh"

namespace oeh {
namespace {

#ifndef OEH_LN_WARPS
#define OEH_LN_WARPS 8
#endif
constexpr int kWarpsPerBlock = OEH_LN_WARPS;

// Cache policy of the row traffic.  Inside a model step the inputs were just written by a GEMM and the output is read by the
// next one, all within the 126 MB L2: plain loads / stores keep it there.  Streaming (evict-first) accesses are ~10 % faster
// on big standalone tensors that do not fit L2 (5.6 vs 6.2 us cold at 4096 x 768) but push the result out of L2 under the
// next GEMM: config 2 runs 1.131 ms with plain accesses, 1.151 ms with streaming ones (tools/bench_ln.py, bench.py; round 2).
#ifdef OEH_LN_STREAM
__device__ __forceinline__ int4 ld_row(const int4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_row(int4* p, const int4& v) { __stcs(p, v); }
#else
__device__ __forceinline__ int4 ld_row(const int4* p) { return *p; }
__device__ __forceinline__ void st_row(int4* p, const int4& v) { *p = v; }
#endif

// kTaps = false: the launch has no fake-quant tap, and the three tap chains are compiled out (the floating-point model steps: with
// them the fp16 instantiation is 4920 SASS instructions, without ~1100 -- 25 launches per step, each starting from a cold
// instruction cache)
//
// kFast (16-bit tensors with taps): every active tap is on the quantizer's common path (8-bit codes exact in the tensor dtype, a
// normal 1 / scale, zero point in range: QdqVec::fast()), and only that path is compiled in -- straight-line code.  With the
// rare-case chains (IEEE division, rint, the scalar fp32 detour) inlined at each of the nine tap sites, the hot path was a
// sequence of far jumps over code it never runs, each landing on a cold instruction-cache line.
template <typename T, int kVecs, bool kGate, bool kTaps, bool kFast>
__device__ __forceinline__ void add_layer_norm_row(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ gamma,
                      const T* __restrict__ beta, T* __restrict__ y, int64_t rows, int cols, float eps,
                      const float* __restrict__ q_a_, const float* __restrict__ q_res_, const float* __restrict__ q_out_,
                      float int_max, T* __restrict__ sum_out, const float* __restrict__ gate_w,
                      const float* __restrict__ gate_b, float* __restrict__ gate_out, int T_len, int H,
                      float as, float az, float rs, float rz, float os, float oz, int64_t row,
                      Vec16<T> (&va)[kVecs], Vec16<T> (&vb)[kVecs]) {
  const float* q_a = kTaps ? q_a_ : nullptr;
  const float* q_res = kTaps ? q_res_ : nullptr;
  const float* q_out = kTaps ? q_out_ : nullptr;
  using V = Vec16<T>;
  constexpr int N = V::N;
  constexpr int LPH = 64 / N;                              // lanes that hold one 64-wide head of the row
  const int lane = threadIdx.x & 31;
  const int nvec = cols / N;
  const bool bv = b != nullptr;

  float x[kVecs][N];
  const QdqVec<T> qa(as, az, int_max), qr(rs, rz, int_max), qo(os, oz, int_max);
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < kVecs; ++v) {
    const int i = lane + 32 * v;
    if (i < nvec) {                                        // taps and the sum in the tensor dtype, 16 bytes at a time
      if (q_a) qa.template apply_t<kFast>(va[v], false);
      if (bv) add_vec<T>(va[v], vb[v]);
      if (q_res) qr.template apply_t<kFast>(va[v], false);
      if (sum_out)                                         // pre-LN blocks (OPT) keep the sum as the next residual
        st_row(reinterpret_cast<int4*>(reinterpret_cast<V*>(sum_out + row * cols) + i), *reinterpret_cast<const int4*>(&va[v]));
    }
#pragma unroll
    for (int e = 0; e < N; ++e) {
      const float s = i < nvec ? DT<T>::to_f(va[v].v[e]) : 0.f;
      x[v][e] = s;
      sum += s;
    }
  }
  const float mean = warp_sum(sum) / (float)cols;
  float ssq = 0.f;
#pragma unroll
  for (int v = 0; v < kVecs; ++v) {
    if (lane + 32 * v < nvec) {
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const float d = x[v][e] - mean;
        ssq = fmaf(d, d, ssq);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ssq) / (float)cols + eps);

  const V* gv = reinterpret_cast<const V*>(gamma);
  const V* bev = reinterpret_cast<const V*>(beta);
  V* yv = reinterpret_cast<V*>(y + row * cols);
  float gacc[kGate ? kVecs : 1];
  if (kGate) {
#pragma unroll
    for (int v = 0; v < kVecs; ++v) gacc[v] = 0.f;
  }
#pragma unroll
  for (int v = 0; v < kVecs; ++v) {
    const int i = lane + 32 * v;
    if (i < nvec) {
      const V g = gv[i];
      V be, out;
      if (beta) be = bev[i];
#pragma unroll
      for (int e = 0; e < N; ++e) {
        float r = DT<T>::to_f(g.v[e]) * (rstd * (x[v][e] - mean));
        if (beta) r += DT<T>::to_f(be.v[e]);
        out.v[e] = DT<T>::from_f(r);                       // one rounding to the tensor dtype
      }
      if (q_out) qo.template apply_t<kFast>(out, false);
      st_row(reinterpret_cast<int4*>(yv + i), *reinterpret_cast<const int4*>(&out));
      if (kGate) {                                         // per-token gate logit of head i / LPH from the row just written
        const float4* wp = reinterpret_cast<const float4*>(gate_w + (int64_t)i * N);   // (H, 64) contiguous == the row's layout
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < N; e += 4) {                   // N fp32 weights as N / 4 16-byte loads (N * 4 bytes per lane, aligned)
          const float4 w4 = __ldg(wp + e / 4);
          acc = fmaf(DT<T>::to_f(out.v[e]), w4.x, acc);
          acc = fmaf(DT<T>::to_f(out.v[e + 1]), w4.y, acc);
          acc = fmaf(DT<T>::to_f(out.v[e + 2]), w4.z, acc);
          acc = fmaf(DT<T>::to_f(out.v[e + 3]), w4.w, acc);
        }
        gacc[v] = acc;
      }
    }
  }
  if (kGate) {
    // out[b,h,t] = fc_h(y[b,t,h*64:(h+1)*64]) (bert_attention.py:303-320; opt_attention.py:285-302) for the attention that
    // consumes this LayerNorm output: a head is LPH consecutive lanes of one vector round
    const int64_t bidx = row / T_len, t = row % T_len;
#pragma unroll
    for (int v = 0; v < kVecs; ++v) {
      float acc = gacc[v];
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const int i = lane + 32 * v;
      if (i < nvec && (lane % LPH) == 0) {
        const int h = i / LPH;
        gate_out[(bidx * H + h) * T_len + t] = acc + __ldg(gate_b + h);
      }
    }
  }
}

template <typename T, int kVecs, bool kGate, bool kTaps = true>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
add_layer_norm_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ gamma,
                      const T* __restrict__ beta, T* __restrict__ y, int64_t rows, int cols, float eps,
                      const float* __restrict__ q_a, const float* __restrict__ q_res, const float* __restrict__ q_out,
                      float int_max, T* __restrict__ sum_out, const float* __restrict__ gate_w,
                      const float* __restrict__ gate_b, float* __restrict__ gate_out, int T_len, int H) {
  using V = Vec16<T>;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = cols / V::N;
  float as = 1.f, az = 0.f, rs = 1.f, rz = 0.f, os = 1.f, oz = 0.f;
  if constexpr (kTaps) {
    if (q_a) { as = DT<T>::rnd(__ldg(q_a)); az = DT<T>::rnd(__ldg(q_a + 1)); }
    if (q_res) { rs = DT<T>::rnd(__ldg(q_res)); rz = DT<T>::rnd(__ldg(q_res + 1)); }
    if (q_out) { os = DT<T>::rnd(__ldg(q_out)); oz = DT<T>::rnd(__ldg(q_out + 1)); }
  }
  const V* av = reinterpret_cast<const V*>(a + row * cols);
  const V* bv = b ? reinterpret_cast<const V*>(b + row * cols) : nullptr;
  V va[kVecs], vb[kVecs];
#pragma unroll
  for (int v = 0; v < kVecs; ++v) {                        // all loads of the row in flight before the first use
    const int i = lane + 32 * v;
    if (i < nvec) {
      int4 raw = ld_row(reinterpret_cast<const int4*>(av + i));
      va[v] = *reinterpret_cast<V*>(&raw);
      if (bv) {
        raw = ld_row(reinterpret_cast<const int4*>(bv + i));
        vb[v] = *reinterpret_cast<V*>(&raw);
      }
    }
  }
  if constexpr (kTaps && P2<T>::kPacked) {
    const bool fast = QdqVec<T>(as, az, int_max).fast() && QdqVec<T>(rs, rz, int_max).fast() && QdqVec<T>(os, oz, int_max).fast();
    if (fast) {                                              // (an inactive tap has scale 1, zero point 0: fast)
      add_layer_norm_row<T, kVecs, kGate, true, true>(a, b, gamma, beta, y, rows, cols, eps, q_a, q_res, q_out, int_max, sum_out,
                                                      gate_w, gate_b, gate_out, T_len, H, as, az, rs, rz, os, oz, row, va, vb);
      return;
    }
  }
  add_layer_norm_row<T, kVecs, kGate, kTaps, false>(a, b, gamma, beta, y, rows, cols, eps, q_a, q_res, q_out, int_max, sum_out,
                                                    gate_w, gate_b, gate_out, T_len, H, as, az, rs, rz, os, oz, row, va, vb);
}

// BertEmbeddings.forward (HF modeling_bert, the shell around the reference's encoder): y = LayerNorm((word[id] + type[tt])
// + pos[p]) -- three gathers, two adds and a LayerNorm in ATen (6 launches, the sums materialised); here one warp gathers
// the three rows, adds them in the tensor dtype in the same order (each sum rounded once, like the two ATen adds) and
// normalises in registers.  token_type_ids / position_ids may be NULL (type 0 / position = token index within its sequence).
template <typename T, int kVecs>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_layer_norm_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids,
                        const int64_t* __restrict__ pos_ids, const T* __restrict__ word, const T* __restrict__ type_tab,
                        const T* __restrict__ pos_tab, const T* __restrict__ gamma, const T* __restrict__ beta,
                        T* __restrict__ y, int64_t rows, int cols, int T_len, int64_t vocab, float eps) {
  using V = Vec16<T>;
  constexpr int N = V::N;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = cols / N;
  int64_t id = __ldg(ids + row);
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);        // out-of-range ids raise in torch; never fault here
  const int64_t tt = type_ids ? __ldg(type_ids + row) : 0;
  const int64_t pp = pos_ids ? __ldg(pos_ids + row) : row % T_len;
  const V* wv = reinterpret_cast<const V*>(word + id * cols);
  const V* tv = type_tab ? reinterpret_cast<const V*>(type_tab + tt * cols) : nullptr;
  const V* pv = pos_tab ? reinterpret_cast<const V*>(pos_tab + pp * cols) : nullptr;
  float x[kVecs][N];
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < kVecs; ++v) {
    const int i = lane + 32 * v;
    V a;
    if (i < nvec) {
      a = wv[i];
      if (tv) add_vec<T>(a, tv[i]);
      if (pv) add_vec<T>(a, pv[i]);
    }
#pragma unroll
    for (int e = 0; e < N; ++e) {
      const float sv = i < nvec ? DT<T>::to_f(a.v[e]) : 0.f;
      x[v][e] = sv;
      sum += sv;
    }
  }
  const float mean = warp_sum(sum) / (float)cols;
  float ssq = 0.f;
#pragma unroll
  for (int v = 0; v < kVecs; ++v)
    if (lane + 32 * v < nvec) {
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const float d = x[v][e] - mean;
        ssq = fmaf(d, d, ssq);
      }
    }
  const float rstd = rsqrtf(warp_sum(ssq) / (float)cols + eps);
  const V* gv = reinterpret_cast<const V*>(gamma);
  const V* bev = reinterpret_cast<const V*>(beta);
  V* yv = reinterpret_cast<V*>(y + row * cols);
#pragma unroll
  for (int v = 0; v < kVecs; ++v) {
    const int i = lane + 32 * v;
    if (i < nvec) {
      const V g = gv[i];
      V be, out;
      if (beta) be = bev[i];
#pragma unroll
      for (int e = 0; e < N; ++e) {
        float r = DT<T>::to_f(g.v[e]) * (rstd * (x[v][e] - mean));
        if (beta) r += DT<T>::to_f(be.v[e]);
        out.v[e] = DT<T>::from_f(r);
      }
      yv[i] = out;
    }
  }
}

template <typename T>
int launch_embed(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const void* word, const void* type_tab,
                 const void* pos_tab, const void* gamma, const void* beta, void* y, int64_t rows, int64_t cols, int T_len,
                 int64_t vocab, float eps, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  if (cols % N != 0 || cols / N > 32 * 8)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_embed_layer_norm: cols=%lld must be a multiple of %d and at most %d",
                     (long long)cols, N, 32 * 8 * N);
  const uintptr_t al = reinterpret_cast<uintptr_t>(word) | reinterpret_cast<uintptr_t>(type_tab) | reinterpret_cast<uintptr_t>(pos_tab) |
                       reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(y);
  if (al & 15) return set_error(OEH_ERR_UNSUPPORTED, "oeh_embed_layer_norm: pointers must be 16-byte aligned");
  const int kv = (int)((cols / N + 31) / 32);
  const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
#define OEH_EMB_CASE(K)                                                                                                       \
  case K:                                                                                                                     \
    embed_layer_norm_kernel<T, K><<<grid, kWarpsPerBlock * 32, 0, st>>>(ids, type_ids, pos_ids, (const T*)word,               \
                                                                        (const T*)type_tab, (const T*)pos_tab, (const T*)gamma, \
                                                                        (const T*)beta, (T*)y, rows, (int)cols, T_len, vocab, eps); \
    break;
  switch (kv) {
    OEH_EMB_CASE(1) OEH_EMB_CASE(2) OEH_EMB_CASE(3) OEH_EMB_CASE(4) OEH_EMB_CASE(5) OEH_EMB_CASE(6) OEH_EMB_CASE(7) OEH_EMB_CASE(8)
  }
#undef OEH_EMB_CASE
  return check_launch("embed_layer_norm_kernel");
}

// OPTDecoder's input (HF modeling_opt: inputs_embeds + OPTLearnedPositionalEmbedding(position_ids), whose forward adds its
// offset of 2 to the positions): out[r] = word[ids[r]] + pos[pos_ids[r] + pos_offset], one rounding to the tensor dtype like the
// ATen add -- two gathers, the int64 offset add and the sum in ATen (4 launches); one warp per row here.
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_sum_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ pos_ids, const T* __restrict__ word,
                 const T* __restrict__ pos_tab, T* __restrict__ y, int64_t rows, int cols, int T_len, int64_t vocab,
                 int64_t n_pos, int64_t pos_offset) {
  using V = Vec16<T>;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  int64_t id = __ldg(ids + row);
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);        // out-of-range ids raise in torch; never fault here
  int64_t pp = (pos_ids ? __ldg(pos_ids + row) : row % T_len) + pos_offset;
  pp = pp < 0 ? 0 : (pp >= n_pos ? n_pos - 1 : pp);
  const V* wv = reinterpret_cast<const V*>(word + id * cols);
  const V* pv = reinterpret_cast<const V*>(pos_tab + pp * cols);
  V* yv = reinterpret_cast<V*>(y + row * cols);
  for (int i = lane; i < cols / V::N; i += 32) {
    V a = wv[i];
    add_vec<T>(a, pv[i]);
    yv[i] = a;
  }
}

template <typename T>
int launch(const void* a, const void* b, const void* gamma, const void* beta, void* y, int64_t rows, int64_t cols, float eps,
           const float* q_a, const float* q_res, const float* q_out, int n_bits, void* sum_out, cudaStream_t st,
           const float* gate_w = nullptr, const float* gate_b = nullptr, float* gate_out = nullptr, int T_len = 1, int H = 0) {
  constexpr int N = Vec16<T>::N;
  if (cols % N != 0 || cols / N > 32 * 8)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_add_layer_norm: cols=%lld must be a multiple of %d and at most %d",
                     (long long)cols, N, 32 * 8 * N);
  const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(gamma) |
                       reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(sum_out);
  if (al & 15) return set_error(OEH_ERR_UNSUPPORTED, "oeh_add_layer_norm: pointers must be 16-byte aligned");
  const int kv = (int)((cols / N + 31) / 32);
  const float int_max = (float)((1 << n_bits) - 1);
  const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const bool taps = q_a || q_res || q_out;
#define OEH_LN_LAUNCH(K, G, Q)                                                                                       \
  add_layer_norm_kernel<T, K, G, Q><<<grid, kWarpsPerBlock * 32, 0, st>>>(                                           \
      (const T*)a, (const T*)b, (const T*)gamma, (const T*)beta, (T*)y, rows, (int)cols, eps, q_a, q_res, q_out,     \
      int_max, (T*)sum_out, gate_w, gate_b, gate_out, T_len, H)
#define OEH_LN_CASE(K)                                                                                               \
  case K:                                                                                                            \
    if (gate_out) {                                                                                                  \
      if (taps) OEH_LN_LAUNCH(K, true, true);                                                                        \
      else OEH_LN_LAUNCH(K, true, false);                                                                            \
    } else {                                                                                                         \
      if (taps) OEH_LN_LAUNCH(K, false, true);                                                                       \
      else OEH_LN_LAUNCH(K, false, false);                                                                           \
    }                                                                                                                \
    break;
  switch (kv) {
    OEH_LN_CASE(1) OEH_LN_CASE(2) OEH_LN_CASE(3) OEH_LN_CASE(4) OEH_LN_CASE(5) OEH_LN_CASE(6) OEH_LN_CASE(7) OEH_LN_CASE(8)
  }
#undef OEH_LN_CASE
#undef OEH_LN_LAUNCH
  return check_launch("add_layer_norm_kernel");
}

}  // namespace
}  // namespace oeh

