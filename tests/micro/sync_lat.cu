// Micro-benchmark (development aid): latency of the hand-shake primitives used by the tcgen05 kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../localdiffusion_hallucination_b200/csrc sync_lat.cu -o sync_lat
#include <cstdio>
#include <cuda_runtime.h>
#include "ld_tc_common.cuh"
using namespace ld::tc;

__global__ void __launch_bounds__(128) k(long long* out, int iters) {
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t slot;
  __shared__ __align__(128) uint8_t buf[4096];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b0 = smem_u32(bars), b1 = b0 + 8, b2 = b0 + 16, b3 = b0 + 24;
  if (threadIdx.x == 0) { mbar_init(b0, 1); mbar_init(b1, 1); mbar_init(b2, 1); mbar_init(b3, 32); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  long long t0, t1;
  // (a) ping-pong between warp 0 and warp 1 through two mbarriers (lane 0 arrives, whole warp waits)
  if (warp < 2) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (warp == 0) { if (lane == 0) mbar_arrive(b0); __syncwarp(); mbar_wait(b1, i & 1); }
      else { mbar_wait(b0, i & 1); if (lane == 0) mbar_arrive(b1); __syncwarp(); }
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;   // round trip = 2 hops
  }
  __syncthreads();
  // (b) tcgen05.commit with nothing outstanding -> wait on the barrier (one warp)
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { if (elect_one()) umma_commit(b2); __syncwarp(); mbar_wait(b2, i & 1); }
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / iters;
  }
  __syncthreads();
  // (c) fence.proxy.async after one 16-byte shared store
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { *reinterpret_cast<uint4*>(buf + lane * 16) = make_uint4(i, i, i, i); fence_proxy_async(); }
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / iters;
  }
  __syncthreads();
  // (d) one MMA (M=128,N=32,K=16) + commit -> wait
  if (warp == 0) {
    const uint32_t a_lo = desc_lo(smem_u32(buf), 2048), hi = desc_hi(128);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) { umma_bf16_lh(tm, a_lo, hi, a_lo, hi, make_idesc(128, 32), 0); umma_commit(b2); }
      __syncwarp(); mbar_wait(b2, (iters + i) & 1);
    }
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / iters;
  }
  __syncthreads();
  // (e) 32-thread arrive (count 32) + wait by the same warp
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { mbar_arrive(b3); mbar_wait(b3, i & 1); }
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / iters;
  }
  __syncthreads();
  // (f) tcgen05.ld 32x32b.x16 + wait
  if (warp == 0) {
    uint32_t r[16]; uint32_t acc = 0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { tmem_ld16(tm, r); tmem_ld_wait(); acc += r[0]; }
    t1 = clock64();
    if (lane == 0) { out[5] = (t1 - t0) / iters; out[7] = acc; }
  }
  // (g) 18 MMAs N=32 back to back + commit -> wait
  __syncthreads();
  if (warp == 0) {
    const uint32_t a_lo = desc_lo(smem_u32(buf), 2048), hi = desc_hi(128);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 18; ++j) umma_bf16_lh(tm, a_lo, hi, a_lo, hi, make_idesc(128, 32), j > 0);
        umma_commit(b2);
      }
      __syncwarp(); mbar_wait(b2, i & 1);
    }
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / iters;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  k<<<1, 128>>>(d, 2000); cudaDeviceSynchronize();
  k<<<1, 128>>>(d, 2000);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("err=%s\n(a) mbarrier ping-pong round trip (2 hops): %lld clk\n(b) empty tcgen05.commit -> wait: %lld clk\n(c) STS.128 + fence.proxy.async: %lld clk\n"
         "(d) 1 MMA + commit -> wait: %lld clk\n(e) 32-lane arrive + wait: %lld clk\n(f) tcgen05.ld x16 + wait: %lld clk\n(g) 18 MMAs (N=32) + commit -> wait: %lld clk\n",
         cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  return 0;
}
