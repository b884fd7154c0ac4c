// Micro-benchmark (development aid): how much do warps spinning on mbarrier.try_wait slow down a working warp?
#include <cstdio>
#include <cuda_runtime.h>
#include "ld_tc_common.cuh"
using namespace ld::tc;

template <int MODE>  // 0: no hint, 1: suspend hint, 2: nanosleep backoff
__global__ void __launch_bounds__(448) k(long long* out, int spinners, int chain) {
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(b, 1); fence_barrier_init(); }
  __syncthreads();
  if (warp == 0) {
    long long t0 = clock64();
    unsigned x = threadIdx.x + 1;
    for (int i = 0; i < chain; ++i) x = x * 1664525u + 1013904223u;   // dependent IMAD chain
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = x; mbar_arrive(b); }
  } else if (warp <= spinners) {
    if (MODE == 0) {
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
    } else if (MODE == 1) {
      mbar_wait(b, 0);
    } else {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
        if (!ok) __nanosleep(100);
      }
    }
  }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[2];
  for (int mode = 0; mode < 3; ++mode)
    for (int sp : {0, 3, 13}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, 448>>>(d, sp, 4000); else if (mode == 1) k<1><<<1, 448>>>(d, sp, 4000); else k<2><<<1, 448>>>(d, sp, 4000);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("mode %d spinners %2d: 4000-instr dependent chain = %lld clk (%.1f clk/instr)\n", mode, sp, h[0], h[0] / 4000.0);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
