// Microbenchmark: what a kernel -> kernel edge costs inside a CUDA graph on this box, with and without programmatic dependent
// launch (griddepcontrol.launch_dependents / griddepcontrol.wait + cudaLaunchAttributeProgrammaticStreamSerialization).
// The kernel mimics the small attention launch: a prologue (barrier init, TMEM allocation), then `work_us` of memory-bound work on a
// buffer the previous launch wrote.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../outeffhop_b200/csrc -I../../include -o pdl_bench pdl_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

using namespace oeh;

template <bool kPdl, bool kTmem>
__global__ void __launch_bounds__(128) k(const float4* __restrict__ in, float4* __restrict__ out, int per_thread) {
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  if (kPdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (kTmem) {
    if (threadIdx.x < 32) ptx::tmem_alloc<128>(ptx::smem_u32(&slot));
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
  }
  if (kPdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int i = 0; i < per_thread; ++i) {
    float4 v = in[base + i * stride];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  out[base] = acc;
  if (kTmem) {
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<128>(slot);
  }
}

template <bool kPdl, bool kTmem>
static float run(int grid, int per_thread, int n_launch, float4* a, float4* b, cudaStream_t st) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n_launch; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = kPdl ? 1 : 0;
    const float4* in = (i & 1) ? b : a;
    float4* out = (i & 1) ? a : b;
    cudaLaunchKernelEx(&cfg, k<kPdl, kTmem>, in, out, per_thread);
  }
  cudaStreamEndCapture(st, &g);
  if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  const int reps = 20;
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("error: %s\n", cudaGetErrorString(err));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1000.f / reps / n_launch;
}

int main() {
  cudaStream_t st;
  cudaStreamCreate(&st);
  const size_t n = (size_t)64 << 20;   // float4 elements: 1 GiB per buffer is too much; 64 Mi float4 = 1 GiB.  use 16 Mi
  float4 *a, *b;
  const size_t elems = (size_t)16 << 20;
  (void)n;
  cudaMalloc(&a, elems * sizeof(float4));
  cudaMalloc(&b, elems * sizeof(float4));
  cudaMemset(a, 0, elems * sizeof(float4));
  cudaMemset(b, 0, elems * sizeof(float4));
  const int grids[] = {148, 384, 592, 1184};
  const int works[] = {0, 8, 32};
  for (int grid : grids)
    for (int w : works) {
      if ((size_t)grid * 128 * (w ? w : 1) > elems) continue;
      const float t00 = run<false, false>(grid, w, 200, a, b, st);
      const float t10 = run<true, false>(grid, w, 200, a, b, st);
      const float t01 = run<false, true>(grid, w, 200, a, b, st);
      const float t11 = run<true, true>(grid, w, 200, a, b, st);
      printf("grid %5d  loads/thread %3d  (%.1f MB/launch)   plain %.2f us   pdl %.2f us   | with tmem alloc: plain %.2f us   pdl %.2f us\n",
             grid, w, (double)grid * 128 * w * 16 / 1e6, t00, t10, t01, t11);
    }
  return 0;
}
