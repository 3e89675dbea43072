// Row softmax family on a materialised score tensor (unfused / range-estimation path, sm_100a).
//   y = softmax_fn(clamp(x*in_scale + causal + mask))  over the last dimension
// softmax_fn: F.softmax | softmax_1 (vutils/softmax_1.py:4-28) | clipped_softmax (models/softmax.py:10-13)
//             | clipped_softmax1 (models/softmax.py:16-19).
// HBM-bound: one read + one write of the (rows, Tk) tensor; the row lives in registers (warp per
// row, Tk <= 1024) or shared memory (block per row, Tk <= 12288).  Math in fp32 with accurate
// expf/division -- this is the parity path, not the fused tensor-core kernel.
#include <cfloat>

#include "oeh_common.cuh"

namespace oeh {
namespace {

struct RowParams {
  int B, H, Tq, Tk;
  float in_scale;
  int kind;
  float stretch, gamma;
  const void* mask;
  int mask_dtype;
  int64_t ms0, ms1, ms2;
  int causal;
  int clamp_min;
  float fmin;   // finfo(dtype).min
};

__device__ __forceinline__ float load_mask(const RowParams& p, int64_t off) {
  if (p.mask_dtype == OEH_F32) return reinterpret_cast<const float*>(p.mask)[off];
  if (p.mask_dtype == OEH_F16) return __half2float(reinterpret_cast<const __half*>(p.mask)[off]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.mask)[off]);
}

__device__ __forceinline__ float score(const RowParams& p, float x, int q, int j, int64_t mask_row) {
  float s = x * p.in_scale;
  if (p.causal && j > q + (p.Tk - p.Tq)) s = s + p.fmin;   // HF causal mask value (finfo.min), added
  if (p.mask) s = s + load_mask(p, mask_row + j);
  if (p.clamp_min) s = fmaxf(s, p.fmin);                   // opt_attention.py:221-223
  return s;
}

__device__ __forceinline__ float finish(const RowParams& p, float e, float den) {
  float v = e / den;
  if (p.kind >= OEH_SOFTMAX_CLIPPED) {
    v = __fadd_rn(__fmul_rn(v, p.stretch), p.gamma);       // sm_out * (eta - gamma) + gamma: two roundings
    v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
  }
  return v;
}

// one warp per row, VPL values per lane kept in registers
template <typename T, int VPL>
__global__ void __launch_bounds__(128) softmax_rows_warp_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                int64_t rows, RowParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int q = (int)(row % p.Tq);
  const int h = (int)((row / p.Tq) % p.H);
  const int b = (int)(row / ((int64_t)p.Tq * p.H));
  const int64_t mask_row = (int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)q * p.ms2;
  const T* xr = x + row * p.Tk;
  T* yr = y + row * p.Tk;
  float s[VPL];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int j = lane + 32 * i;
    s[i] = j < p.Tk ? score(p, DT<T>::to_f(xr[j]), q, j, mask_row) : -INFINITY;
    m = fmaxf(m, s[i]);
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    s[i] = (lane + 32 * i) < p.Tk ? expf(s[i] - m) : 0.f;
    sum += s[i];
  }
  sum = warp_sum(sum);
  float den = sum;
  if (p.kind == OEH_SOFTMAX_1 || p.kind == OEH_SOFTMAX_CLIPPED1) den = sum + expf(-m);   // the "+1" term
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int j = lane + 32 * i;
    if (j < p.Tk) yr[j] = DT<T>::from_f(finish(p, s[i], den));
  }
}

// one block per row, the row staged in shared memory
template <typename T>
__global__ void __launch_bounds__(256) softmax_rows_block_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                 int64_t rows, RowParams p) {
  extern __shared__ float srow[];
  __shared__ float red[8];
  const int64_t row = blockIdx.x;
  const int q = (int)(row % p.Tq);
  const int h = (int)((row / p.Tq) % p.H);
  const int b = (int)(row / ((int64_t)p.Tq * p.H));
  const int64_t mask_row = (int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)q * p.ms2;
  const T* xr = x + row * p.Tk;
  T* yr = y + row * p.Tk;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < p.Tk; j += 256) {
    const float s = score(p, DT<T>::to_f(xr[j]), q, j, mask_row);
    srow[j] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  if (l == 0) red[w] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < p.Tk; j += 256) {
    const float e = expf(srow[j] - m);
    srow[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (l == 0) red[w] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  float den = sum;
  if (p.kind == OEH_SOFTMAX_1 || p.kind == OEH_SOFTMAX_CLIPPED1) den = sum + expf(-m);
  for (int j = threadIdx.x; j < p.Tk; j += 256) yr[j] = DT<T>::from_f(finish(p, srow[j], den));
}

template <typename T>
int launch(const void* x, void* y, int64_t rows, const RowParams& p, cudaStream_t st) {
  const int Tk = p.Tk;
  if (Tk <= 1024) {
    const unsigned grid = (unsigned)((rows + 3) / 4);
#define OEH_WARP_CASE(V) softmax_rows_warp_kernel<T, V><<<grid, 128, 0, st>>>((const T*)x, (T*)y, rows, p)
    if (Tk <= 64) OEH_WARP_CASE(2);
    else if (Tk <= 128) OEH_WARP_CASE(4);
    else if (Tk <= 256) OEH_WARP_CASE(8);
    else if (Tk <= 512) OEH_WARP_CASE(16);
    else OEH_WARP_CASE(32);
#undef OEH_WARP_CASE
  } else {
    if (Tk > 12288) return set_error(OEH_ERR_UNSUPPORTED, "oeh_softmax_rows: Tk=%d > 12288", Tk);
    softmax_rows_block_kernel<T><<<(unsigned)rows, 256, (size_t)Tk * sizeof(float), st>>>((const T*)x, (T*)y, rows, p);
  }
  return check_launch("softmax_rows_kernel");
}

}  // namespace
}  // namespace oeh

using namespace oeh;

extern "C" int oeh_softmax_rows(const void* x, void* y, int32_t B, int32_t H, int32_t Tq, int32_t Tk, int32_t dtype,
                                float in_scale, int32_t kind, double gamma, double eta, const void* mask,
                                int32_t mask_dtype, const int64_t* mask_strides, int32_t causal, int32_t clamp_min,
                                oeh_stream_t stream) {
  OEH_CHECK_ARG(B >= 0 && H >= 0 && Tq >= 0 && Tk >= 0, "oeh_softmax_rows: negative dimension");
  OEH_CHECK_ARG(kind >= OEH_SOFTMAX_VANILLA && kind <= OEH_SOFTMAX_CLIPPED1, "oeh_softmax_rows: unknown kind %d", kind);
  const int64_t rows = (int64_t)B * H * Tq;
  if (rows == 0 || Tk == 0) return OEH_OK;
  OEH_CHECK_ARG(x && y, "oeh_softmax_rows: null pointer");
  OEH_CHECK_ARG(!mask || mask_strides, "oeh_softmax_rows: mask without strides");
  OEH_CHECK_ARG(rows <= 0x7fffffffLL, "oeh_softmax_rows: too many rows");
  RowParams p;
  p.B = B; p.H = H; p.Tq = Tq; p.Tk = Tk;
  p.in_scale = in_scale;
  p.kind = kind;
  p.stretch = (float)(eta - gamma);   // Python double arithmetic, converted at the tensor multiply
  p.gamma = (float)gamma;
  p.mask = mask;
  p.mask_dtype = mask_dtype;
  p.ms0 = mask ? mask_strides[0] : 0;
  p.ms1 = mask ? mask_strides[1] : 0;
  p.ms2 = mask ? mask_strides[2] : 0;
  p.causal = causal;
  p.clamp_min = clamp_min;
  p.fmin = dtype == OEH_F32 ? -FLT_MAX : (dtype == OEH_F16 ? -65504.f : -3.3895313892515355e38f);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case OEH_F32: return launch<float>(x, y, rows, p, st);
    case OEH_F16: return launch<__half>(x, y, rows, p, st);
    case OEH_BF16: return launch<__nv_bfloat16>(x, y, rows, p, st);
  }
  return set_error(OEH_ERR_BAD_ARG, "oeh_softmax_rows: unknown dtype %d", dtype);
}
