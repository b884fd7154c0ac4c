// Micro-benchmark (development aid): sustained rate of tcgen05.mma M=128, N=NN, K=16 (bf16, operands in shared memory, no-swizzle
// K-major) issued the way the conv kernel issues them: TILES tiles of 18 MMAs (9 shifted tap views x 2 k-steps) per CTA, a commit per
// tile, accumulators alternating between two TMEM stages; 1 or 2 co-resident CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../localdiffusion_hallucination_b200/csrc mma_rate.cu -o mma_rate
//   ./mma_rate            -> clocks per MMA for several configurations
#include <cstdio>
#include <cuda_runtime.h>
#include "ld_tc_common.cuh"
using namespace ld::tc;
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

// mode bit 0: shifted tap views (else all taps read the same aligned view); bit 1: wait for every tile's commit before the next tile
// bit 2: two accumulator stages alternate per tile (else one); bit 3: consecutive MMAs alternate between two accumulators
// side traffic from warps 4-7 while the MMAs run (they stop when warp 0 raises a flag): bit 4 tcgen05.ld 32x32b.x32 of the accumulator,
// bit 5 st.shared.v4 of 8 KB (a staging tile), bit 6 ld.shared.v4 broadcast (a bias row), bit 7 ld.shared.v4 + st.shared.v4 of a patch
template <int NN>
__global__ void __launch_bounds__(256) k(long long* out, int tiles, int mode, int pitch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t slot;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) { stop = 0; mbar_init(b0, 1); mbar_init(b0 + 8, 1); mbar_init(b0 + 16, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  // A: patch of 180 pixels x 32 channels as [c8][pixel][16 B] (LBO = 2880 B, SBO = pitch * 16); B: 9 taps x [k8][NN][16 B]
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 16384;
  const uint32_t a_lo0 = desc_lo(a_base, 2880), a_hi = desc_hi(pitch * 16), b_lo0 = desc_lo(b_base, NN * 16), b_hi = desc_hi(128);
  constexpr uint32_t idesc = make_idesc(128, NN);
  if (warp == 0) {
    long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tm + ((mode & 4) ? (t & 1) * 64 : 0);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint32_t a_t = a_lo0 + ((mode & 1) ? (uint32_t)(ky * pitch + kx) : 0u);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t dd = d + ((mode & 8) ? ((tap * 2 + kk) & 1) * 128 : 0);
            umma_bf16_lh(dd, a_t + (uint32_t)(2 * kk) * (2880 >> 4), a_hi, b_lo0 + (uint32_t)(tap * (NN * 32 * 2 >> 4) + 2 * kk * NN), b_hi, idesc,
                         (tap | kk) ? 1u : 0u);
          }
        }
        umma_commit(b0 + 8 * (t & 1));
      }
      __syncwarp();
      if (mode & 2) mbar_wait(b0 + 8 * (t & 1), (t >> 1) & 1);
    }
    // drain: one more commit covers every MMA issued so far
    if (elect_one()) umma_commit(b0 + 16);
    __syncwarp();
    mbar_wait(b0 + 16, 0);
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; *reinterpret_cast<volatile int*>(&stop) = 1; }
  }
  else if (warp >= 4 && (mode & 0xf0)) {
    volatile int* flag = reinterpret_cast<volatile int*>(&stop);
    const uint32_t side = smem_u32(smem) + 40 * 1024;                      // 16 KB of scratch behind the operands (N <= 32 only)
    uint32_t r[32];
    uint4 acc4 = make_uint4(0, 0, 0, 0);
    long long n = 0;
    while (!*flag) {
      if (mode & 16) { tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16), r); tmem_ld_wait(); acc4.x += r[lane & 31]; }
      if (mode & 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sts128(side + ((warp & 3) * 32 + lane) * 64 + (((i ^ (lane >> 1)) & 3) << 4), acc4);
      }
      if (mode & 64) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const uint4 v = lds128(side + 8192 + i * 16); acc4.y += v.x; }
      }
      if (mode & 128) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const uint32_t q = side + (((warp & 3) * 6 + i) * 32 + lane) * 16;
          uint4 v = lds128(q); v.x += acc4.x; sts128(q, v);
        }
      }
      ++n;
    }
    if (acc4.x == 0x12345 && acc4.y == 77) out[1000] = n;   // keep the side work alive
    if (lane == 0 && warp == 4) out[512 + blockIdx.x] = n;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

template <int NN>
void run(const char* name, int ctas_per_sm, int mode, int pitch, long long* d_out) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int tiles = 200;
  // dynamic smem chosen so that exactly `ctas_per_sm` CTAs fit an SM
  const int smem = ctas_per_sm == 1 ? 120 * 1024 : 64 * 1024;
  cudaFuncSetAttribute(k<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = sms * ctas_per_sm;
  k<NN><<<grid, 256, smem>>>(d_out, tiles, mode, pitch);
  cudaDeviceSynchronize();
  k<NN><<<grid, 256, smem>>>(d_out, tiles, mode, pitch);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
  static long long h[1024];
  cudaMemcpy(h, d_out, 1024 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0, side = 0; for (int i = 0; i < grid; ++i) { avg += (double)h[i]; side += (double)h[512 + i]; } avg /= grid; side /= grid;
  // per SM the tensor pipe served ctas_per_sm * tiles * 18 MMAs in `avg` clocks
  printf("%-58s N=%3d ctas/SM=%d: %7.1f clk per MMA per CTA, %6.1f clk per MMA per SM", name, NN, ctas_per_sm, avg / (tiles * 18.0),
         avg / (tiles * 18.0 * ctas_per_sm));
  if (mode & 0xf0) printf("   (side loop: %.1f iterations per tile)", side / tiles);
  printf("\n");
}

int main() {
  long long* d; cudaMalloc(&d, 1024 * sizeof(long long)); cudaMemset(d, 0, 1024 * sizeof(long long));
  for (int c = 1; c <= 2; ++c) {
    run<32>("same view, no wait, one accumulator", c, 0, 10, d);
    run<32>("shifted views, no wait, one accumulator", c, 1, 10, d);
    run<32>("shifted views, no wait, two accumulator stages", c, 1 | 4, 10, d);
    run<32>("shifted views, wait per tile, two stages", c, 1 | 2 | 4, 10, d);
    run<32>("shifted views, MMAs alternate between two accumulators", c, 1 | 8, 10, d);
    run<32>("shifted views, pitch 8 (SBO = 128 B)", c, 1 | 4, 8, d);
    run<32>("shifted views, pitch 18 (8 x 16 tile)", c, 1 | 4, 18, d);
    run<32>("+ tcgen05.ld of the accumulator (4 warps, free running)", c, 1 | 4 | 16, 10, d);
    run<32>("+ st.shared.v4 staging tile (4 warps, free running)", c, 1 | 4 | 32, 10, d);
    run<32>("+ ld.shared.v4 broadcast (4 warps, free running)", c, 1 | 4 | 64, 10, d);
    run<32>("+ ld/st.shared.v4 patch rewrite (4 warps, free running)", c, 1 | 4 | 128, 10, d);
    run<64>("shifted views, no wait, two stages", c, 1 | 4, 10, d);
    if (c == 1) run<128>("shifted views, no wait, one accumulator", c, 1, 10, d);
  }
  return 0;
}
