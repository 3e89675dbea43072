// Per-token attention-gate logits (sm_100a): out[b,h,t] = fc_h(x[b,t,h*d:(h+1)*d])
// (bert_attention.py:303-320; opt_attention.py:285-302).  The reference runs H tiny Linear(d,1)
// GEMVs (or d->m->ReLU->1 MLPs, bert_attention.py:126-143) in a Python loop plus torch.stack;
// here one launch reads the (B,T,H*d) hidden states exactly once, fully coalesced (HBM-bound:
// bytes = B*T*H*d*sizeof(T) + 4*B*H*T).  The sigmoid is applied in the attention epilogue.
#include "oeh_common.cuh"

namespace oeh {
namespace {

// Linear(d,1): one warp per token; each lane owns 16-byte chunks of the row, a head is a group of
// d/VN consecutive lanes, reduced with xor-shuffles inside the group.
template <typename T, int D>
__global__ void __launch_bounds__(256) gate_linear_kernel(const T* __restrict__ x, int64_t row_stride, int64_t tokens,
                                                          int T_len, int H, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out) {
  using V = Vec16<T>;
  constexpr int VN = V::N;             // elements per 16-byte chunk
  constexpr int LPH = D / VN;          // lanes per head (8 for fp16 d=64, 16 for fp32 d=64)
  static_assert(LPH >= 1 && LPH <= 32 && (LPH & (LPH - 1)) == 0, "head dim must map to a power-of-two lane group");
  const int lane = threadIdx.x & 31;
  const int64_t tok = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= tokens) return;
  const int b = (int)(tok / T_len), t = (int)(tok % T_len);
  const V* xr = reinterpret_cast<const V*>(x + tok * row_stride);
  const int chunks = H * LPH;          // 16-byte chunks per token row
  for (int c0 = 0; c0 < chunks; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    if (c < chunks) {
      const V v = xr[c];
      const float* wp = w + (int64_t)c * VN;   // w is (H,d) contiguous == the row's layout
#pragma unroll
      for (int e = 0; e < VN; ++e) acc = fmaf(DT<T>::to_f(v.v[e]), __ldg(wp + e), acc);
    }
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (c < chunks && (lane % LPH) == 0) {
      const int h = c / LPH;
      out[((int64_t)b * H + h) * T_len + t] = acc + __ldg(bias + h);
    }
  }
}

// MLP gate: Linear(d,m) -> ReLU -> Linear(m,1).  One block = 128 tokens of one head; the head's
// first-layer weights are staged in shared memory, each thread owns one token.
template <typename T, int D>
__global__ void __launch_bounds__(128) gate_mlp_kernel(const T* __restrict__ x, int64_t row_stride, int64_t tokens,
                                                       int T_len, int H, int M, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ out) {
  extern __shared__ float sw[];        // w1[h]: M*D, then b1[h]: M, then w2[h]: M
  const int h = blockIdx.y;
  for (int i = threadIdx.x; i < M * D; i += 128) sw[i] = w1[(int64_t)h * M * D + i];
  for (int i = threadIdx.x; i < M; i += 128) {
    sw[M * D + i] = b1[h * M + i];
    sw[M * D + M + i] = w2[h * M + i];
  }
  __syncthreads();
  const int64_t tok = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (tok >= tokens) return;
  using V = Vec16<T>;
  constexpr int VN = V::N;
  float xv[D];
  const V* xr = reinterpret_cast<const V*>(x + tok * row_stride + (int64_t)h * D);
#pragma unroll
  for (int c = 0; c < D / VN; ++c) {
    const V v = xr[c];
#pragma unroll
    for (int e = 0; e < VN; ++e) xv[c * VN + e] = DT<T>::to_f(v.v[e]);
  }
  float acc = __ldg(b2 + h);
  for (int j = 0; j < M; ++j) {
    float hid = sw[M * D + j];
#pragma unroll
    for (int c = 0; c < D; ++c) hid = fmaf(xv[c], sw[j * D + c], hid);
    acc = fmaf(fmaxf(hid, 0.f), sw[M * D + M + j], acc);
  }
  const int b = (int)(tok / T_len), t = (int)(tok % T_len);
  out[((int64_t)b * H + h) * T_len + t] = acc;
}

template <typename T>
int launch(const void* x, int B, int T_len, int H, int d, int64_t row_stride, const float* w1, const float* b1,
           const float* w2, const float* b2, int hidden, float* out, cudaStream_t st) {
  const int64_t tokens = (int64_t)B * T_len;
  if (d != 64) return set_error(OEH_ERR_UNSUPPORTED, "oeh_gate_logits: head dim %d (only 64 is built)", d);
  if (hidden == 0) {
    gate_linear_kernel<T, 64><<<(unsigned)((tokens + 7) / 8), 256, 0, st>>>((const T*)x, row_stride, tokens, T_len, H,
                                                                           w1, b1, out);
  } else {
    const size_t smem = (size_t)(hidden * 64 + 2 * hidden) * sizeof(float);
    if (smem > 48 * 1024) return set_error(OEH_ERR_UNSUPPORTED, "oeh_gate_logits: hidden=%d too large", hidden);
    dim3 grid((unsigned)((tokens + 127) / 128), (unsigned)H);
    gate_mlp_kernel<T, 64><<<grid, 128, smem, st>>>((const T*)x, row_stride, tokens, T_len, H, hidden, w1, b1, w2, b2,
                                                   out);
  }
  return check_launch("gate_logits_kernel");
}

}  // namespace
}  // namespace oeh

using namespace oeh;

extern "C" int oeh_gate_logits(const void* x, int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t d,
                               int64_t x_row_stride, const float* w1, const float* b1, const float* w2,
                               const float* b2, int32_t hidden, float* out, oeh_stream_t stream) {
  OEH_CHECK_ARG(B >= 0 && T >= 0 && H > 0 && d > 0 && hidden >= 0, "oeh_gate_logits: bad dimension");
  if ((int64_t)B * T == 0) return OEH_OK;
  OEH_CHECK_ARG(x && w1 && b1 && out, "oeh_gate_logits: null pointer");
  OEH_CHECK_ARG(hidden == 0 || (w2 && b2), "oeh_gate_logits: MLP gate needs w2/b2");
  OEH_CHECK_ARG(x_row_stride >= (int64_t)H * d, "oeh_gate_logits: row stride %lld < H*d", (long long)x_row_stride);
  OEH_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_row_stride * (int64_t)dtype_size(dtype)) % 16 == 0,
                "oeh_gate_logits: x must be 16-byte aligned with a 16-byte multiple row pitch");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case OEH_F32: return launch<float>(x, B, T, H, d, x_row_stride, w1, b1, w2, b2, hidden, out, st);
    case OEH_F16: return launch<__half>(x, B, T, H, d, x_row_stride, w1, b1, w2, b2, hidden, out, st);
    case OEH_BF16: return launch<__nv_bfloat16>(x, B, T, H, d, x_row_stride, w1, b1, w2, b2, hidden, out, st);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_gate_logits: unknown dtype %d", dtype);
}
