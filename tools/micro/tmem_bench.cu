// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM vs warps per SMSP and vector width (x16 / x32 / x64 fp32 columns per
// instruction), and a softmax-like mix (ld x32 -> 32 FFMA -> st x16).  One CTA per SM, 512 TMEM columns.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../outeffhop_b200/csrc -I../../include -o tmem_bench tmem_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

using namespace oeh;

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
  ptx::tmem_ld32(taddr, r);
  ptx::tmem_ld32(taddr + 32, r + 32);
}

// MODE 0: ld x32 + wait   1: ld 2 x x32 + wait (64 columns in flight)   2: st x32 + wait   3: ld x32, 32 FFMA, st x16 (P pack)
// 4: ld x16 + wait
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) ptx::tmem_alloc<512>(ptx::smem_u32(&slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tl = slot + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t col0 = (uint32_t)((warp >> 2) * 64) & 511u;          // warps of one SMSP use different columns
  uint32_t r[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) r[j] = threadIdx.x + j;
  ptx::tmem_st32(tl + col0, r);
  ptx::tmem_st32(tl + ((col0 + 32) & 511u), r + 32);
  ptx::tmem_wait_st();
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      ptx::tmem_ld32(tl + col0, r);
      ptx::tmem_wait_ld();
      acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
    } else if (MODE == 1) {
      tmem_ld64(tl + col0, r);
      ptx::tmem_wait_ld();
      acc += __uint_as_float(r[0]) + __uint_as_float(r[63]);
    } else if (MODE == 2) {
      r[0] = i;
      ptx::tmem_st32(tl + col0, r);
      ptx::tmem_wait_st();
    } else if (MODE == 3) {
      ptx::tmem_ld32(tl + col0, r);
      ptx::tmem_wait_ld();
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float v = fmaf(__uint_as_float(r[j]), 1.0001f, 0.5f);
        s += v;
        r[j] = __float_as_uint(v);
      }
      acc += s;
      ptx::tmem_st16(tl + col0, r);
      ptx::tmem_wait_st();
    } else {
      ptx::tmem_ld16(tl + col0, r);
      ptx::tmem_wait_ld();
      acc += __uint_as_float(r[0]) + __uint_as_float(r[15]);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(slot);
}

template <int MODE>
void run(const char* name, int warps, int cols_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 4000;
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double elems = (double)iters * cols_per_iter * warps * 32;
  printf("%-22s warps/SM %2d (per SMSP %4.1f): %8.0f cycles, %7.2f fp32 elem/clk/SM = %6.1f B/clk/SM  err=%s\n", name, warps,
         warps / 4.0, c, elems / c, elems * 4 / c, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<4>("ld x16", w, 16);
  for (int w : {4, 8, 16, 32}) run<0>("ld x32", w, 32);
  for (int w : {4, 8, 16, 32}) run<1>("ld 2 x x32", w, 64);
  for (int w : {4, 8, 16, 32}) run<2>("st x32", w, 32);
  for (int w : {4, 8, 16}) run<3>("ld x32, 32 FFMA, st x16", w, 32);
  return 0;
}
