// Micro-benchmark (development aid): rate of the TMA halo-patch loads of the 3x3 convolution, alone on the machine.
// Tensor [N][H][W][C] bf16 (C = 32).  Every CTA walks tiles of 16 x 8 pixels and loads the (18 x 10)-pixel halo patch of each into a
// 4-stage ring, re-issuing as soon as a stage lands.  Variants of the tensor map / box:
//   0: (8 ch, W, H, C/8, N) box (8, 10, 18, 4, 1), no swizzle      -- the kernel's K-major operand image, 720 rows of 16 B
//   1: (C, W, H, N) box (32, 10, 18, 1), no swizzle                -- 180 rows of 64 B
//   2: same box, SWIZZLE_64B
// A second part times cp.async.bulk (1-D) copies of 2 .. 32 KB out of an L2-resident buffer: ~367 clk per copy and issuing CTA whatever the size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../localdiffusion_hallucination_b200/csrc tma_rate.cu -o tma_rate
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ld_tc_common.cuh"
using namespace ld::tc;

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map, int variant, int ntiles, int tiles_x, int tiles_y, int stages,
                                         long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(b0 + 8 * i, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t stage_bytes = 180 * 64, stage_pitch = 12288;
    const long long t0 = clock64();
    int issued = 0, done = 0;
    const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto issue = [&](int i) {
      const int tile = blockIdx.x + i * gridDim.x, tpi = tiles_x * tiles_y, img = tile / tpi, r = tile - img * tpi, ty = r / tiles_x, tx = r - ty * tiles_x;
      const int s = i % stages;
      mbar_arrive_expect_tx(b0 + 8 * s, stage_bytes);
      if (variant == 0) tma_load_5d(smem_u32(smem) + s * stage_pitch, &map, 0, tx * 8 - 1, ty * 16 - 1, 0, img, b0 + 8 * s);
      else tma_load_4d(smem_u32(smem) + s * stage_pitch, &map, 0, tx * 8 - 1, ty * 16 - 1, img, b0 + 8 * s);
    };
    for (; issued < mine && issued < stages; ++issued) issue(issued);
    for (; done < mine; ++done) {
      mbar_wait(b0 + 8 * (done % stages), (done / stages) & 1);
      if (issued < mine) { issue(issued); ++issued; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

// cp.async.bulk (1-D) of `chunk`-byte pieces out of an L2-resident buffer of `total` bytes, ring of `stages` buffers
__global__ void __launch_bounds__(128) kb(const uint8_t* src, int total, int chunk, int stages, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(b0 + 8 * i, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int n = total / chunk * reps;
    const long long t0 = clock64();
    int issued = 0, done = 0;
    auto issue = [&](int i) {
      const int s = i % stages;
      mbar_arrive_expect_tx(b0 + 8 * s, chunk);
      bulk_g2s(smem_u32(smem) + s * chunk, src + (size_t)(i % (total / chunk)) * chunk, chunk, b0 + 8 * s);
    };
    for (; issued < n && issued < stages; ++issued) issue(issued);
    for (; done < n; ++done) {
      mbar_wait(b0 + 8 * (done % stages), (done / stages) & 1);
      if (issued < n) { issue(issued); ++issued; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int N = 32, H = 256, W = 256, C = 32;
  void* x; cudaMalloc(&x, (size_t)N * H * W * C * 2); cudaMemset(x, 0, (size_t)N * H * W * C * 2);
  long long* d; cudaMalloc(&d, 1024 * 8);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int tiles_x = W / 8, tiles_y = H / 16, ntiles = N * tiles_x * tiles_y;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int variant = 0; variant < 3; ++variant) {
    CUtensorMap m;
    CUresult r;
    if (variant == 0) {
      const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)N};
      const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
      const cuuint32_t box[5] = {8, 10, 18, 4, 1}, es[5] = {1, 1, 1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      const cuuint32_t box[4] = {32, 10, 18, 1}, es[4] = {1, 1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              variant == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("variant %d: encode failed %d\n", variant, (int)r); continue; }
    for (int cps = 1; cps <= 2; ++cps)
      for (int stages = 4; stages <= 8; stages += 4) {
        const int grid = sms * cps, smem = cps == 1 ? 100 * 1024 : 98 * 1024 / 1;   // 2 CTAs of 98 KB fit the 228 KB of an SM
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<grid, 128, smem>>>(m, variant, ntiles, tiles_x, tiles_y, stages, d);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<<<grid, 128, smem>>>(m, variant, ntiles, tiles_x, tiles_y, stages, d);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        static long long h[1024]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
        printf("variant %d  ctas/SM=%d stages=%d: %7.1f us  (%.0f clk per tile per SM, %.0f GB/s of tensor bytes)\n", variant, cps, stages, ms * 1e3,
               avg / ((double)ntiles / sms), (double)N * H * W * C * 2 / (ms * 1e-3) * 1e-9);
      }
  }
  // bulk copies: the weight stream of the 256 -> 256 convolution (1.18 MB per CTA in 32 KB stages, every CTA reads the same bytes)
  {
    const int total = 36 * 32768;
    uint8_t* wsrc; cudaMalloc(&wsrc, total); cudaMemset(wsrc, 1, total);
    cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int chunk : {65536, 32768, 16384, 8192, 4096, 2048})
      for (int stages : {4})
        for (int ctas : {148, 296}) {               // 296: two co-resident CTAs per SM (<= 100 KB each)
          if (stages * chunk > (ctas == 296 ? 100 : 200) * 1024) { if (ctas == 296) continue; }
          if (stages * chunk > 200 * 1024) continue;
          kb<<<ctas, 128, stages * chunk>>>(wsrc, total, chunk, stages, 2, d);
          cudaDeviceSynchronize();
          kb<<<ctas, 128, stages * chunk>>>(wsrc, total, chunk, stages, 2, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("bulk: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          static long long h[1024]; cudaMemcpy(h, d, ctas * 8, cudaMemcpyDeviceToHost);
          double avg = 0; for (int i = 0; i < ctas; ++i) avg += (double)h[i]; avg /= ctas;
          printf("bulk copy chunk %5d B  stages=%d  CTAs=%3d: %6.1f B/clk per CTA, %6.1f clk per copy per CTA\n", chunk, stages, ctas,
                 2.0 * total / avg, avg / (2.0 * total / chunk));
        }
  }
  return 0;
}
