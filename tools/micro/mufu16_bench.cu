// Microbenchmark: ex2.approx.f16x2 vs f32, fma.rn.f32x2, cvt.rn.f16x2.f32 throughput per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// MODE 0: f16x2 MUFU only   1: f16 softmax mix per pair: 2 FFMA + F2FP + MUFU.f16x2   2: same with fma.f32x2   3: bf16x2 MUFU
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters, float a, float b) {
  float x[16];
  uint32_t h[8];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = -1.0f - 0.01f * j - 0.001f * threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = pack_h2(x[2 * j], x[2 * j + 1]);
  uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  uint64_t a2, b2;
  { float2 t = make_float2(a, a); a2 = *reinterpret_cast<uint64_t*>(&t); t = make_float2(b, b); b2 = *reinterpret_cast<uint64_t*>(&t); }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      if (MODE == 0) {
        acc0 ^= ex2_h2(h[j]); acc1 ^= ex2_h2(h[j + 1]); acc2 ^= ex2_h2(h[j + 2]); acc3 ^= ex2_h2(h[j + 3]);
      } else if (MODE == 3) {
        acc0 ^= ex2_bf2(h[j]); acc1 ^= ex2_bf2(h[j + 1]); acc2 ^= ex2_bf2(h[j + 2]); acc3 ^= ex2_bf2(h[j + 3]);
      } else if (MODE == 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t r = ex2_h2(pack_h2(fmaf(x[2 * (j + u)], a, b), fmaf(x[2 * (j + u) + 1], a, b)));
          if (u == 0) acc0 ^= r; else if (u == 1) acc1 ^= r; else if (u == 2) acc2 ^= r; else acc3 ^= r;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float2 xv = make_float2(x[2 * (j + u)], x[2 * (j + u) + 1]);
          uint64_t y = ffma2(*reinterpret_cast<uint64_t*>(&xv), a2, b2);
          float2 yv = *reinterpret_cast<float2*>(&y);
          uint32_t r = ex2_h2(pack_h2(yv.x, yv.y));
          if (u == 0) acc0 ^= r; else if (u == 1) acc1 ^= r; else if (u == 2) acc2 ^= r; else acc3 ^= r;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] += 1;
#pragma unroll
    for (int j = 0; j < 16; j += 4) x[j] += 1e-7f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 ^ acc1 ^ acc2 ^ acc3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f, 0.f);
  k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f, 0.f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double elems = (double)iters * 16 * warps * 32;
  printf("%-12s warps/SM %2d: %8.0f cycles, %6.2f exps/clk/SM  err=%s\n", name, warps, c, elems / c, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) run<0>("ex2.f16x2", w);
  for (int w : {4, 8, 16}) run<3>("ex2.bf16x2", w);
  for (int w : {4, 8, 16}) run<1>("mix16", w);
  for (int w : {4, 8, 16}) run<2>("mix16+ffma2", w);
  return 0;
}
