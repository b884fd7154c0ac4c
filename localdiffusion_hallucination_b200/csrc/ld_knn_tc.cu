// PatchCore nearest-neighbour search (models.py:179-217: `euclidean_dist` + `nearest_neighbors(n_neighbors=1)`) on tcgen05 / TMEM
// for sm_100a -- the dense part of the anomaly-map stage in front of the sampler:
//
//   d(i, j) = sqrt(clamp(|x_i|^2 - 2 x_i . y_j + |y_j|^2, 0)),   score_i = min_j d(i, j),   location_i = argmin_j d(i, j)
//
// for embeddings x [M][D] against a memory bank y [Nb][D] (fp32).  The [M x Nb] distance matrix the reference materialises never
// exists: the Gram tile x . y^T lives in TMEM and only a running (min, argmin) per embedding row leaves the SM.
//
// Accuracy: the distance is a difference of large numbers, so plain bf16 operands (8 mantissa bits) are not enough.  Both operands
// are split into bf16 hi + bf16 lo parts (x = xh + xl exactly to 16 bits) and the product is accumulated in fp32 as
// xh.yh + xh.yl + xl.yh (the dropped xl.yl term is 2^-16 relative): three tcgen05.mma per k-step, fp32-level results.
//
//   knn_prep_kernel    fp32 rows -> the exact shared-memory images of the UMMA operands: per (128-row block, 64-wide k chunk) a
//                      [hi | lo] pair of K-major tiles [8 chunks of 8][128 rows][8] (rows / columns past the end are zero),
//                      so every pipeline stage is two contiguous 32 KB cp.async.bulk copies; plus the fp32 squared norms.
//   knn_tc_kernel      CTA = (128 embeddings) x (a contiguous range of 128-entry bank blocks); warp 0 loads, warp 1 issues the MMAs
//                      (M=128, N=128, K=16; two accumulator stages), warps 2-5 (one per TMEM lane quarter, thread = embedding row)
//                      turn the finished Gram tile into distances and keep the running minimum; the per-CTA results are merged with
//                      one 64-bit atomicMin per row on (distance^2 bits << 32 | bank index): ties resolve to the lowest index.
//   knn_finish_kernel  sqrt + unpack.
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

#include "ld_kernels.h"
#include "ld_launch.cuh"
#include "ld_tc_common.cuh"

namespace ld {

using namespace tc;

namespace {

constexpr int kThreads = 6 * 32;      // warp 0 loader, 1 MMA, 2-5 distance / min
constexpr int kPart = 128 * 64 * 2;   // bytes of one operand part tile (128 rows x 64 k, bf16)
constexpr int kStageBytes = 4 * kPart;   // x hi | x lo | y hi | y lo
constexpr int kStages = 3;
constexpr int kSmem = kStages * kStageBytes + 16 * 8 + 16;

// grid (k chunks, row blocks), 128 threads: thread = row of the block
__global__ void __launch_bounds__(128) knn_prep_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ img, int R, int D, int nkc) {
  const int kc = blockIdx.x, rb = blockIdx.y, r = threadIdx.x, row = rb * 128 + r;
  __nv_bfloat16* hi = img + ((size_t)rb * nkc + kc) * (2 * 128 * 64);
  __nv_bfloat16* lo = hi + 128 * 64;
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = kc * 64 + c8 * 8 + 2 * j + e;
        v[e] = (row < R && k < D) ? src[(size_t)row * D + k] : 0.f;
      }
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[0]), h1 = __float2bfloat16_rn(v[1]);
      h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      l[j] = pack_bf16x2(v[0] - __bfloat162float(h0), v[1] - __bfloat162float(h1));
    }
    *reinterpret_cast<uint4*>(hi + (size_t)c8 * 1024 + r * 8) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + (size_t)c8 * 1024 + r * 8) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// |row|^2 in fp32, one warp per row (x.pow(2).sum(dim=-1), models.py:192-193)
__global__ void __launch_bounds__(256) knn_norm_kernel(const float* __restrict__ src, float* __restrict__ norms, int R, int D) {
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) { const float v = src[(size_t)row * D + k]; s = fmaf(v, v, s); }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) norms[row] = s;
}

struct KnnParams {
  const __nv_bfloat16* ximg; const __nv_bfloat16* yimg;   // operand images of knn_prep_kernel
  const float* xn; const float* yn;                       // squared norms
  unsigned long long* best;                               // [M] (distance^2 bits << 32 | index), initialised to ~0
  int M, Nb, nkc, nbb, bb_per_cta;
};

__global__ void __launch_bounds__(kThreads, 1) knn_tc_kernel(const KnnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t full = smem_u32(bars), empty = full + 8 * kStages, acc_full = empty + 8 * kStages, acc_empty = acc_full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = blockIdx.x;
  const int bb0 = blockIdx.y * p.bb_per_cta, bb1 = min(p.nbb, bb0 + p.bb_per_cta);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // accumulator stage a at columns a * 128
  pdl_wait();
  if (bb0 >= bb1) {
    if (warp == 1) tmem_dealloc(tmem_base, 256);
    return;
  }

  if (warp == 0) {
    // ---------------------------------------------------------------- bulk loads -----------------
    if (lane == 0) {
      int it = 0;
      for (int bb = bb0; bb < bb1; ++bb)
        for (int kc = 0; kc < p.nkc; ++kc, ++it) {
          const int s = it % kStages;
          mbar_wait(empty + 8 * s, ((it / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(full + 8 * s, kStageBytes);
          const uint32_t dst = smem_u32(smem + (size_t)s * kStageBytes);
          bulk_g2s(dst, p.ximg + ((size_t)rb * p.nkc + kc) * (2 * 128 * 64), 2 * kPart, full + 8 * s);
          bulk_g2s(dst + 2 * kPart, p.yimg + ((size_t)bb * p.nkc + kc) * (2 * 128 * 64), 2 * kPart, full + 8 * s);
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issue ------------------
    constexpr uint32_t idesc = make_idesc(128, 128);
    const uint32_t hi128 = desc_hi(128);
    const uint32_t base_lo = desc_lo(smem_u32(smem), 2048);
    int it = 0, q = 0;
    for (int bb = bb0; bb < bb1; ++bb, ++q) {
      const int as = q & 1;
      mbar_wait(acc_empty + 8 * as, ((q >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kc = 0; kc < p.nkc; ++kc, ++it) {
        const int s = it % kStages;
        mbar_wait(full + 8 * s, (it / kStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = base_lo + (uint32_t)(s * (kStageBytes >> 4));
          const uint32_t xh = st, xl = st + (kPart >> 4), yh = st + 2 * (kPart >> 4), yl = st + 3 * (kPart >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 64 = 4 x 16; x.y ~= xh.yh + xh.yl + xl.yh
            const uint32_t o = (uint32_t)(2 * k * 128);
            umma_bf16_lh(tmem_base + (uint32_t)(as * 128), xh + o, hi128, yh + o, hi128, idesc, (kc > 0 || k > 0) ? 1u : 0u);
            umma_bf16_lh(tmem_base + (uint32_t)(as * 128), xh + o, hi128, yl + o, hi128, idesc, 1u);
            umma_bf16_lh(tmem_base + (uint32_t)(as * 128), xl + o, hi128, yh + o, hi128, idesc, 1u);
          }
          umma_commit(empty + 8 * s);
          if (kc == p.nkc - 1) umma_commit(acc_full + 8 * as);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- distances + running min ----
    const int qd = warp & 3;                            // TMEM lane quarter this warp may read
    const int row = rb * 128 + qd * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);
    const float xn = row < p.M ? p.xn[row] : 0.f;
    float best = INFINITY; int best_j = 0x7fffffff;
    int q = 0;
    for (int bb = bb0; bb < bb1; ++bb, ++q) {
      const int as = q & 1;
      mbar_wait(acc_full + 8 * as, (q >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t g[32];
        tmem_ld32(lane_base + (uint32_t)(as * 128 + c0), g);
        tmem_ld_wait();
        if (c0 == 96) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(acc_empty + 8 * as); }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int j = bb * 128 + c0 + c;
          if (j < p.Nb) {
            // res = x_norm - 2 * (x @ y^T) + y_norm^T, clamp_min(0)  (models.py:195-196), same evaluation order
            float d2 = __fadd_rn(__fsub_rn(xn, __fmul_rn(2.f, __uint_as_float(g[c]))), __ldg(p.yn + j));
            d2 = fmaxf(d2, 0.f);
            if (d2 < best) { best = d2; best_j = j; }
          }
        }
      }
    }
    if (row < p.M && best_j != 0x7fffffff)
      atomicMin(p.best + row, ((unsigned long long)__float_as_uint(best) << 32) | (unsigned int)best_j);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

__global__ void knn_finish_kernel(const unsigned long long* __restrict__ best, float* __restrict__ score, long long* __restrict__ loc, int M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const unsigned long long b = best[i];
  score[i] = sqrtf(__uint_as_float((unsigned int)(b >> 32)));
  loc[i] = (long long)(unsigned int)(b & 0xffffffffull);
}

int g_sms_knn = 0;
bool g_knn_configured = false;

}  // namespace

size_t knn_scratch_bytes(int M, int Nb, int D) {
  const size_t nkc = (size_t)(D + 63) / 64, mrb = (size_t)(M + 127) / 128, nbb = (size_t)(Nb + 127) / 128;
  return (mrb + nbb) * nkc * 2 * kPart + ((size_t)mrb * 128 + nbb * 128) * 4 + (size_t)mrb * 128 * 8 + 256;
}

int knn_tc_launch(const float* x, const float* bank, int M, int Nb, int D, float* score, long long* loc, void* scratch, cudaStream_t s) {
  if (!g_knn_configured) {
    if (cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return -1;
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_knn, cudaDevAttrMultiProcessorCount, dev);
    g_knn_configured = true;
  }
  const int nkc = (D + 63) / 64, mrb = (M + 127) / 128, nbb = (Nb + 127) / 128;
  uint8_t* sc = (uint8_t*)scratch;
  __nv_bfloat16* ximg = (__nv_bfloat16*)sc; sc += (size_t)mrb * nkc * 2 * kPart;
  __nv_bfloat16* yimg = (__nv_bfloat16*)sc; sc += (size_t)nbb * nkc * 2 * kPart;
  float* xn = (float*)sc; sc += (size_t)mrb * 128 * 4;
  float* yn = (float*)sc; sc += (size_t)nbb * 128 * 4;
  unsigned long long* best = (unsigned long long*)(((uintptr_t)sc + 7) & ~(uintptr_t)7);
  cudaMemsetAsync(best, 0xff, (size_t)M * 8, s);
  knn_prep_kernel<<<dim3(nkc, mrb), 128, 0, s>>>(x, ximg, M, D, nkc);
  knn_prep_kernel<<<dim3(nkc, nbb), 128, 0, s>>>(bank, yimg, Nb, D, nkc);
  knn_norm_kernel<<<(M * 32 + 255) / 256, 256, 0, s>>>(x, xn, M, D);
  knn_norm_kernel<<<(Nb * 32 + 255) / 256, 256, 0, s>>>(bank, yn, Nb, D);
  // split the bank so that (row blocks x splits) covers the SMs about twice
  int splits = (2 * g_sms_knn + mrb - 1) / mrb;
  if (splits > nbb) splits = nbb;
  if (splits < 1) splits = 1;
  const int per = (nbb + splits - 1) / splits;
  splits = (nbb + per - 1) / per;
  KnnParams p{ximg, yimg, xn, yn, best, M, Nb, nkc, nbb, per};
  knn_tc_kernel<<<dim3(mrb, splits), kThreads, kSmem, s>>>(p);
  knn_finish_kernel<<<(M + 255) / 256, 256, 0, s>>>(best, score, loc, M);
  return 7;
}

}  // namespace ld
