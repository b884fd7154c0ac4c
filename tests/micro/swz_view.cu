// Experiment (development aid): can tcgen05.mma read SHIFTED tap views out of a TMA-written, hardware-swizzled halo patch?
// A patch of 18 x 10 pixels x C channels (C = 32: 64-byte rows, SWIZZLE_64B; C = 64: 128-byte rows, SWIZZLE_128B) is loaded by ONE TMA box
// (rows of 64 / 128 B instead of the 16-byte rows of the no-swizzle K-major image).  For every tap (ky, kx) the A descriptor starts at
// patch pixel ky*10 + kx, 8-pixel core groups are 10 pixels apart (SBO = 10 rows), K advances by 32 B inside the swizzled row.
// B is an identity, so D[m][n] must equal the patch value A[m][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../localdiffusion_hallucination_b200/csrc swz_view.cu -o swz_view
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "ld_tc_common.cuh"
using namespace ld::tc;

__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const void* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst_smem),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
               : "memory");
}

// C channels per pixel; stage base offset `boff` bytes (multiple of 128) tests alignment sensitivity
template <int C>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map, int x0, int y0, int boff, int base_offset_mode, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b0 = smem_u32(bars);
  constexpr int ROW = C * 2;                       // bytes per pixel row
  uint8_t* a_s = smem + boff;                      // patch
  uint8_t* b_s = smem + 48 * 1024;                 // identity B: no-swizzle K-major [C/8][C rows][16 B]
  if (threadIdx.x == 0) { mbar_init(b0, 1); mbar_init(b0 + 8, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < (C / 8) * C * 8; i += 128) {
    const int k8 = i / (C * 8), n = (i / 8) % C, e = i % 8;
    reinterpret_cast<__nv_bfloat16*>(b_s)[i] = __float2bfloat16_rn((k8 * 8 + e) == n ? 1.f : 0.f);
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(smem_u32(&slot), 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(b0, 180 * ROW);
    tma_load_4d(smem_u32(a_s), &map, 0, x0 - 1, y0 - 1, 0, b0);
  }
  mbar_wait(b0, 0);
  tc_fence_after();
  constexpr uint32_t idesc = make_idesc(128, C);
  const uint32_t swz = C == 32 ? (4u << 29) : (2u << 29);             // layout type: SWIZZLE_64B / SWIZZLE_128B
  const uint32_t b_lo0 = desc_lo(smem_u32(b_s), C * 16), b_hi = desc_hi(128);
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t start = smem_u32(a_s) + (uint32_t)(ky * 10 + kx) * ROW;
        uint32_t a_hi = desc_hi(10 * ROW) | swz;
        if (base_offset_mode) a_hi |= ((start >> 7) & 7u) << 17;         // descriptor bits [49,52)
#pragma unroll
        for (int kk = 0; kk < C / 16; ++kk)
          umma_bf16_lh(tm, desc_lo(start + kk * 32, 16), a_hi, b_lo0 + (uint32_t)(2 * kk * C), b_hi, idesc, kk ? 1u : 0u);
        umma_commit(b0 + 8);
      }
      __syncwarp();
    }
    mbar_wait(b0 + 8, tap & 1);
    tc_fence_after();
    uint32_t r[32];
    for (int j = 0; j < C; j += 32) {
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + j, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[((size_t)tap * 128 + warp * 32 + lane) * C + j + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tm, 64);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int C>
int run(EncodeFn enc, int x0, int y0, int boff, int bom) {
  const int H = 64, W = 64;
  std::vector<__nv_bfloat16> h((size_t)H * W * C);
  auto val = [&](int y, int x, int c) { return (float)(((y * W + x) * 7 + c * 3) % 251); };
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c) h[((size_t)y * W + x) * C + c] = __float2bfloat16_rn(val(y, x, c));
  void* g; cudaMalloc(&g, h.size() * 2); cudaMemcpy(g, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 9 * 128 * C * 4); cudaMemset(out, 0xff, 9 * 128 * C * 4);
  CUtensorMap m;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)C, 10, 18, 1}, es[4] = {1, 1, 1, 1};
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return -1; }
  cudaFuncSetAttribute(k<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<C><<<1, 128, 64 * 1024>>>(m, x0, y0, boff, bom, out);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("C=%d: %s\n", C, cudaGetErrorString(cudaGetLastError())); return -1; }
  std::vector<float> o(9 * 128 * C); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int tap = 0; tap < 9; ++tap)
    for (int mrow = 0; mrow < 128; ++mrow)
      for (int c = 0; c < C; ++c) {
        const int r = mrow / 8, x = mrow % 8, ky = tap / 3, kx = tap % 3;
        const int gy = y0 - 1 + r + ky, gx = x0 - 1 + x + kx;
        const float want = (gy < 0 || gy >= H || gx < 0 || gx >= W) ? 0.f : val(gy, gx, c);
        if (o[((size_t)tap * 128 + mrow) * C + c] != want) { if (first < 0) first = (tap * 128 + mrow) * C + c; ++bad; }
      }
  printf("C=%2d origin (%2d,%2d) stage offset %4d base_offset_mode %d: %d mismatches of %d", C, x0, y0, boff, bom, bad, 9 * 128 * C);
  if (bad) printf("  (first at tap %d row %d ch %d: got %.0f)", first / (128 * C), (first / C) % 128, first % C, o[first]);
  printf("\n");
  cudaFree(g); cudaFree(out);
  return bad;
}

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  for (int bom = 0; bom < 2; ++bom)
    for (int boff : {0, 128, 512, 11520, 11520 + 128}) {
      run<32>(enc, 8, 16, boff, bom);
      run<64>(enc, 8, 16, boff, bom);
    }
  run<32>(enc, 0, 0, 0, 0);      // zero fill at the image corner
  run<64>(enc, 56, 48, 0, 0);
  return 0;
}
