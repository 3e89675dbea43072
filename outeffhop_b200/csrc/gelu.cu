// erf GELU of the feed-forward block on 16-bit tensors (sm_100a):  y = x * 0.5 * (1 + erf(x / sqrt(2))).
//
// The reference's BERT is HF's (transformers 4.31 BertIntermediate: dense -> ACT2FN["gelu"], the exact erf form); in the
// fused eval step of this repo the 13 GELU launches were the largest non-GEMM item (13 % of the step,
// profiles/r01_launches_bert_step_n.summary.txt) because ATen's kernel is not HBM-bound on B200: erff() costs ~35 issue
// slots per element and the 25 MB fc1 output is still L2-resident when it runs.  This kernel evaluates
//     Phi(-|x|) = 0.5 * erfc(|x| / sqrt 2) = 0.5 * t * exp(-x^2/2 + P(t)),   t = 1 / (1 + |x| / (2 sqrt 2))
// (the classic erfc form with a degree-9 P: relative error < 1.2e-7 over the whole axis -- also in the negative tail, where
// an absolute-error erf would lose the fp16 result) and  gelu(x) = max(x, 0) - |x| * Phi(-|x|),  with packed fp32x2
// arithmetic (FFMA2), one MUFU.RCP and one MUFU.EX2 per element: ~13 issue slots and 2 XU operations per element.  The result is rounded once to the tensor dtype like ATen's (fp32 opmath): correctly rounded
// for all but 5 of the 63488 finite fp16 inputs (1 ulp) and for every bf16 input.  ATen's own chain x*0.5*(1+erff(x/sqrt 2))
// cancels in the negative tail, so the two differ by a few 16-bit ulps on ~1 % of the bit patterns (all below -2, where
// |gelu| < 0.05; max abs difference 3e-5 in fp16) -- tests/test_gpu_gelu.py.
// Roofline: HBM / L2, bytes = n * (size_in + size_out).
#include "oeh_common.cuh"
#include "gelu_math.cuh"
#include "sm100_ptx.cuh"

namespace oeh {
namespace {

#ifndef OEH_GELU_UNROLL
#define OEH_GELU_UNROLL 2
#endif
#ifndef OEH_GELU_BLOCKS
#define OEH_GELU_BLOCKS 8
#endif
constexpr int kThreads = 256;
constexpr int kUnroll = OEH_GELU_UNROLL;         // 16-byte vectors in flight per thread
constexpr int kBlocksPerSm = OEH_GELU_BLOCKS;

using gelu_math::gelu2;

template <typename T>
__device__ __forceinline__ void gelu_vec(Vec16<T>& v) {
#pragma unroll
  for (int e = 0; e < Vec16<T>::N; e += 2) {
    float y0, y1;
    gelu2<false>(DT<T>::to_f(v.v[e]), DT<T>::to_f(v.v[e + 1]), y0, y1);
    v.v[e] = DT<T>::from_f(y0);
    v.v[e + 1] = DT<T>::from_f(y1);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, kBlocksPerSm) gelu_erf_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n,
                                                            int aligned) {
  using V = Vec16<T>;
  constexpr int N = V::N;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int64_t done = 0;
  if (aligned) {
    const int64_t nvec = n / N;
    const V* xv = reinterpret_cast<const V*>(x);
    V* yv = reinterpret_cast<V*>(y);
    int64_t i = tid;
    // plain (not streaming) accesses: the input was just written by the fc1 GEMM and the output feeds the fc2 GEMM -- both
    // live in L2
    for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
      V in[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) in[u] = xv[i + u * stride];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        gelu_vec<T>(in[u]);
        yv[i + u * stride] = in[u];
      }
    }
    for (; i < nvec; i += stride) {
      V a = xv[i];
      gelu_vec<T>(a);
      yv[i] = a;
    }
    done = nvec * N;
  }
  for (int64_t t = done + tid; t < n; t += stride) {
    float y0, y1;
    gelu2<false>(DT<T>::to_f(x[t]), 0.f, y0, y1);
    y[t] = DT<T>::from_f(y0);
  }
}

template <typename T>
int launch_gelu(const void* x, void* y, int64_t n, cudaStream_t st) {
  const int aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const int64_t work = aligned ? (n / Vec16<T>::N + kUnroll - 1) / kUnroll + 1 : n;
  const int64_t want = (work + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)num_sms() * kBlocksPerSm;
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  gelu_erf_kernel<T><<<grid, kThreads, 0, st>>>((const T*)x, (T*)y, n, aligned);
  return check_launch("gelu_erf_kernel");
}

}  // namespace
}  // namespace oeh

using namespace oeh;

extern "C" int oeh_gelu_erf(const void* x, void* y, int64_t n, int32_t dtype, oeh_stream_t stream) {
  OEH_CHECK_ARG(n >= 0, "oeh_gelu_erf: negative size");
  if (n == 0) return OEH_OK;
  OEH_CHECK_ARG(x && y, "oeh_gelu_erf: null tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OEH_F16) return launch_gelu<__half>(x, y, n, st);
  if (dtype == OEH_BF16) return launch_gelu<__nv_bfloat16>(x, y, n, st);
  return set_error(OEH_ERR_UNSUPPORTED, "oeh_gelu_erf: dtype %d (fp16 / bf16 tensors only; fp32 keeps the library erff)", dtype);
}
