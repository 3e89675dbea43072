// Microbenchmark: tcgen05.mma issue/throughput for the attention kernel's shapes (one CTA per SM, one issuer lane).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../outeffhop_b200/csrc -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdint>
#include "sm100_ptx.cuh"
using namespace oeh;

// MODE 0: QK-like SS  M128 N64  K16 x4 (A,B smem)       1: SS M128 N128 K16 x4
// MODE 2: PV-like TS  M128 N64  K16 x4 (A tmem, B smem)  3: mix: 4x SS N64 + 4x TS N64 (one attention step of one warpgroup)
// MODE 4: mix of two warpgroups interleaved (16 MMAs)    5: TS N64 with two alternating accumulators
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int iters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(ptx::smem_u32(&slot));
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = slot;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    const uint32_t idQK64 = ptx::make_idesc_f16(false, 128, 64, false, false);
    const uint32_t idQK128 = ptx::make_idesc_f16(false, 128, 128, false, false);
    const uint32_t idPV = ptx::make_idesc_f16(false, 128, 64, false, true);
    const uint64_t dq = ptx::make_smem_desc_sw128(smem0, 16, 1024);
    const uint64_t dk = ptx::make_smem_desc_sw128(smem0 + 32768, 16, 1024);
    const uint64_t dv = ptx::make_smem_desc_sw128(smem0 + 65536, 1024, 1024);
    long long t0 = clock64();
    if (ptx::elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t st = (it & 3) * 512;   // rotate smem stages like the ring
        if (MODE == 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::mma_ss(tmem + (it & 1) * 64, dq + kk * 2, dk + st + kk * 2, idQK64, kk > 0);
        } else if (MODE == 1) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::mma_ss(tmem + (it & 1) * 128, dq + kk * 2, dk + st + kk * 2, idQK128, kk > 0);
        } else if (MODE == 2) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::mma_ts(tmem + 384, tmem + (it & 1) * 64 + kk * 8, dv + st + kk * 128, idPV, 1);
        } else if (MODE == 3) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::mma_ts(tmem + 384, tmem + (it % 3) * 64 + kk * 8, dv + st + kk * 128, idPV, 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::mma_ss(tmem + (it % 3) * 64, dq + kk * 2, dk + st + kk * 2, idQK64, kk > 0);
        } else if (MODE == 4) {
#pragma unroll
          for (int w = 0; w < 2; ++w) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ptx::mma_ts(tmem + 384 + w * 64, tmem + (w * 3 + it % 3) * 64 + kk * 8, dv + st + kk * 128, idPV, 1);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ptx::mma_ss(tmem + (w * 3 + it % 3) * 64, dq + w * 1024 + kk * 2, dk + st + kk * 2, idQK64, kk > 0);
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::mma_ts(tmem + 384 + (it & 1) * 64, tmem + (it & 1) * 64 + kk * 8, dv + st + kk * 128, idPV, 1);
        }
      }
      ptx::mma_commit(ptx::smem_u32(&bar));
    }
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

template <int MODE>
void run(const char* name, int mmas_per_iter, double flop_per_iter) {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  int iters = 4000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<MODE><<<148, 128, 100 * 1024>>>(cyc, iters);
  k<MODE><<<148, 128, 100 * 1024>>>(cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-34s %8.1f cycles/iter  %6.1f cycles/MMA  %7.0f flop/clk/SM  (%s)\n", name, c / iters, c / iters / mmas_per_iter,
         flop_per_iter * iters / c, cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  run<0>("SS M128 N64  K16 x4 (QK)", 4, 4 * 2.0 * 128 * 64 * 16);
  run<1>("SS M128 N128 K16 x4", 4, 4 * 2.0 * 128 * 128 * 16);
  run<2>("TS M128 N64  K16 x4 (PV)", 4, 4 * 2.0 * 128 * 64 * 16);
  run<5>("TS M128 N64 x4, 2 accumulators", 4, 4 * 2.0 * 128 * 64 * 16);
  run<3>("PV x4 + QK x4 (one wg step)", 8, 8 * 2.0 * 128 * 64 * 16);
  run<4>("two wg steps interleaved (16)", 16, 16 * 2.0 * 128 * 64 * 16);
  return 0;
}
