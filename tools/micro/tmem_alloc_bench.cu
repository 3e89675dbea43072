// Microbenchmark: what tcgen05.alloc / dealloc cost per CTA when several CTAs share an SM (the one-tile attention kernel runs three per
// SM).  Reports, per variant, the average in-graph time of an otherwise empty launch and the in-kernel clock64 span of alloc and dealloc.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../outeffhop_b200/csrc -I../../include -o tmem_alloc_bench tmem_alloc_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

using namespace oeh;

// MODE 0: barrier init + __syncthreads only   1: alloc<kCols> + relinquish + dealloc   2: alloc<kCols> without relinquish + dealloc
template <int MODE, uint32_t kCols>
__global__ void __launch_bounds__(128) k(long long* cyc) {
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  long long t0 = clock64(), t1 = t0, t2 = t0, t3 = t0;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (MODE >= 1 && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&slot)), "n"(kCols) : "memory");
    if (MODE == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    t1 = clock64();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (MODE >= 1 && threadIdx.x < 32) {
    t2 = clock64();
    ptx::tmem_dealloc<kCols>(slot);
    t3 = clock64();
  }
  if (threadIdx.x == 0 && cyc) { cyc[2 * blockIdx.x] = t1 - t0; cyc[2 * blockIdx.x + 1] = t3 - t2; }
}

template <int MODE, uint32_t kCols>
static void run(int grid, long long* cyc, cudaStream_t st) {
  const int n_launch = 200;
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n_launch; ++i) k<MODE, kCols><<<grid, 128, 0, st>>>(cyc);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < 20; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  static long long h[2 * 2048];
  cudaMemcpy(h, cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
  double sa = 0, sd = 0; long long ma = 0, md = 0;
  for (int i = 0; i < grid; ++i) { sa += h[2 * i]; sd += h[2 * i + 1]; if (h[2 * i] > ma) ma = h[2 * i]; if (h[2 * i + 1] > md) md = h[2 * i + 1]; }
  printf("grid %5d mode %d cols %3u: %.2f us/launch   alloc clk avg %.0f max %lld   dealloc clk avg %.0f max %lld   (%s)\n", grid, MODE, kCols,
         ms * 1000.f / 20 / n_launch, sa / grid, ma, sd / grid, md, cudaGetErrorString(cudaGetLastError()));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
}

int main() {
  cudaStream_t st;
  cudaStreamCreate(&st);
  long long* cyc;
  cudaMalloc(&cyc, sizeof(long long) * 2 * 2048);
  for (int grid : {148, 384, 592}) {
    run<0, 128>(grid, cyc, st);
    run<1, 32>(grid, cyc, st);
    run<1, 128>(grid, cyc, st);
    run<2, 128>(grid, cyc, st);
    run<1, 512>(grid, cyc, st);
  }
  return 0;
}
