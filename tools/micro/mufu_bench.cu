// Microbenchmark: ex2.approx throughput per SM vs warps/SMSP, and a softmax-like mix (FFMA + EX2 + FADD), and an
// FMA-pipe polynomial exp2.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = -1.0f - 0.01f * j - 0.001f * threadIdx.x;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      if (MODE == 0) {            // pure MUFU
        acc0 += ex2(x[j]); acc1 += ex2(x[j + 1]); acc2 += ex2(x[j + 2]); acc3 += ex2(x[j + 3]);
      } else if (MODE == 1) {     // softmax mix: FFMA, EX2, FADD
        acc0 += ex2(fmaf(x[j], a, b)); acc1 += ex2(fmaf(x[j + 1], a, b));
        acc2 += ex2(fmaf(x[j + 2], a, b)); acc3 += ex2(fmaf(x[j + 3], a, b));
      } else {                    // polynomial exp2 on the FMA/ALU pipes (Cody-Waite + degree-3)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float t = fmaf(x[j + u], a, b);
          t = fmaxf(t, -126.f);
          float r = t + 12582912.f;            // round to nearest int (magic)
          float fl = r - 12582912.f;
          float f = t - fl;                    // in [-0.5, 0.5]
          float pl = fmaf(f, 0.05550411f, 0.24022651f);
          pl = fmaf(pl, f, 0.69314718f);
          pl = fmaf(pl, f, 1.0f);
          float res = __int_as_float(__float_as_int(pl) + (__float_as_int(r) << 23));
          if (u == 0) acc0 += res; else if (u == 1) acc1 += res; else if (u == 2) acc2 += res; else acc3 += res;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] += 1e-7f;   // keep the loads loop-variant (cheap FADDs, counted in the mix)
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f, 0.f);
  k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f, 0.f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double elems = (double)iters * 16 * warps * 32;
  printf("%-10s warps/SM %2d (per SMSP %4.1f): %8.0f cycles, %6.2f elem/clk/SM  err=%s\n", name, warps, warps / 4.0, c, elems / c,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 12, 16, 32}) run<0>("mufu", w);
  for (int w : {4, 8, 12, 16, 32}) run<1>("mix", w);
  for (int w : {4, 8, 12, 16, 32}) run<2>("poly", w);
  return 0;
}
