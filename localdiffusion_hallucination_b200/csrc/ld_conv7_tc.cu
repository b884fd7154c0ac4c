// init_conv (ddpm.py:319): 7x7, pad 3, ONE fp32 input channel -> 32..64 bf16 channels, on tcgen05 for sm_100a.
//
// The sampler state x_t is fp32 and stays fp32 in HBM; the im2col operand is built on the fly in shared memory:
//   builders (4 warps) stage the 22 x 14 halo patch of x ONCE per tile as two bf16 patches, hi and lo (x = hi + lo to
//   ~2^-17, so the fp32 state is not rounded to 8 bits on its way into the network); then thread = output pixel
//   copies, per filter row ky, the 8-element window [px, px+8) of patch row py+ky (five 32-bit loads + funnel
//   shifts, one 16-byte store) into its row of the K-major A operand: K chunk ky = taps (ky, 0..6) + one zero-weight
//   slot, chunk 7 = padding; [K = 64 hi | 64 lo].  (The first version split all 49 taps per pixel: 6272 splits and
//   scalar loads per tile instead of 308 -- the builders, not the tensor pipe, set the pace.)
//   one lane issues 8 tcgen05.mma (M = 128 pixels, N = Cout, K = 16) against the resident weights [w | w];
//   4 epilogue warps add the bias and store bf16 NHWC.  Persistent CTAs, 2 A stages, 2 TMEM accumulators.
//
// Toeplitz form (conv7_toeplitz_kernel, the product path whenever W % 4 == 0): no im2col operand at all.  A tile is 128 image ROWS x XO
// output columns (XO = 256 / Cout).  The A operand of filter row ky is the bf16 patch itself -- row m of the MMA = image row y0 + m + ky - 3,
// K = the 16 patch columns x0 - 4 .. x0 + 11, stored as two 8-column planes [plane][row][16 B] (the canonical no-swizzle K-major image with
// image rows as matrix rows), so a filter row is a view shifted by ky rows -- and the filter row becomes a banded (Toeplitz) B matrix:
// B_ky[n = (xo, c)][k] = w[ky][k - xo - 1][c].  7 filter rows x (hi, lo) = 14 tcgen05.mma of M = 128, N = 256, K = 16 per 128 x XO pixels;
// per tile the builders write 8.6 KB of shared memory (the im2col form wrote 32 KB per 128 pixels and was bound by exactly that, ncu: LSU
// wavefronts 46 %), the accumulator row of a thread is XO whole pixels = 512 contiguous output bytes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include <vector>

#include "ld_conv7_tc.h"
#include "ld_launch.cuh"
#include "ld_tc_common.cuh"

namespace ld {

using namespace tc;

namespace {

constexpr int kThreads = 9 * 32;     // warps 0-3 builders, 4-7 epilogue, 8 MMA
constexpr int kMmaWarp = 8;
constexpr int PH = 22, PW = 14;      // halo patch of a 16 x 8 tile
constexpr int A_STAGE = 16 * 2048;   // [16 chunks of 8 K][128 rows][16 B]
constexpr int PATCH_W = PH * 8;      // 32-bit words of one bf16 patch (row pitch 16 elements)

struct Params {
  const float* x; const __nv_bfloat16* w; const float* bias; __nv_bfloat16* out;
  int N, H, W, tiles_x, tiles_y, ntiles;
};

template <int NT>
__global__ void __launch_bounds__(kThreads, 2) conv7_tc_kernel(const Params p) {
  constexpr int W_BYTES = 16 * NT * 16;               // [16 chunks][NT rows][16 B]
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* a_s = smem;                                // 2 stages
  uint8_t* w_s = a_s + 2 * A_STAGE;
  uint32_t* patch = reinterpret_cast<uint32_t*>(w_s + W_BYTES);   // [2 stages][hi, lo][PH rows][8 words = 16 bf16]
  float* bias_s = reinterpret_cast<float*>(patch + 4 * PATCH_W);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + NT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const uint32_t w_full = smem_u32(bars), a_full = w_full + 8, a_empty = a_full + 16, acc_full = a_empty + 16, acc_empty = acc_full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full + 8 * i, 4); mbar_init(a_empty + 8 * i, 1);
      mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 128);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NT; i += kThreads) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 2 * NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_trigger();   // persistent grid: the next kernel's CTAs may be scheduled (ld_launch.cuh)
  pdl_wait();                            // x is written by the previous timestep's update kernel
  const int tpi = p.tiles_x * p.tiles_y;

  if (warp < 4) {
    // ---------------------------------------------------------------- builders -------------------
    const int m = threadIdx.x;                         // output pixel of the tile: (m >> 3, m & 7)
    const int py = m >> 3, px = m & 7;
    int it = 0;
    // the patch values of the NEXT tile are requested before this tile's window copies: the global-load latency (the patch is
    // read straight from HBM/L2, 176 element pairs per tile) otherwise sits on every tile's critical path
    float pv[4];
    auto load_patch = [&](int tile, float (&v)[4]) {
      const int img = tile / tpi, r = tile - img * tpi;
      const int ty0 = (r / p.tiles_x) * 16, tx0 = (r % p.tiles_x) * 8;
      const float* ximg = p.x + (size_t)img * p.H * p.W;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = m + 128 * k;
        v[2 * k] = v[2 * k + 1] = 0.f;
        if (i < PH * 8) {
          const int hy = i >> 3, hx = (i & 7) * 2;
          const int gy = ty0 + hy - 3, gx = tx0 + hx - 3;
          if ((unsigned)gy < (unsigned)p.H) {
            if (hx < PW && (unsigned)gx < (unsigned)p.W) v[2 * k] = __ldg(ximg + (size_t)gy * p.W + gx);
            if (hx + 1 < PW && (unsigned)(gx + 1) < (unsigned)p.W) v[2 * k + 1] = __ldg(ximg + (size_t)gy * p.W + gx + 1);
          }
        }
      }
    };
    if ((int)blockIdx.x < p.ntiles) load_patch(blockIdx.x, pv);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      uint32_t* ph = patch + s * 2 * PATCH_W;          // hi patch, then lo patch
      uint32_t* pl = ph + PATCH_W;
      // halo patch split into bf16 hi + lo (zero outside the image = the conv's zero padding)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = m + 128 * k;
        if (i < PH * 8) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[2 * k], pv[2 * k + 1]);
          const float2 hf = __bfloat1622float2(h2);
          ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
          pl[i] = pack_bf16x2(pv[2 * k] - hf.x, pv[2 * k + 1] - hf.y);
        }
      }
      if (tile + (int)gridDim.x < p.ntiles) load_patch(tile + gridDim.x, pv);
      named_bar(1, 128);                               // patch complete (and the previous tile's readers are done: 2 patch buffers)
      mbar_wait(a_empty + 8 * s, ((it >> 1) & 1) ^ 1);
      uint8_t* row = a_s + s * A_STAGE + m * 16;
      const int sh = (px & 1) * 16;                    // odd px: the window starts in the upper half of its first word
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {                 // K chunk ky = taps (ky, kx = 0..6) + one slot whose weight is zero
        const uint32_t* rh = ph + (py + ky) * 8 + (px >> 1);
        const uint32_t* rl = pl + (py + ky) * 8 + (px >> 1);
        uint32_t a[5], b[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) { a[j] = rh[j]; b[j] = rl[j]; }
        *reinterpret_cast<uint4*>(row + ky * 2048) = make_uint4(__funnelshift_r(a[0], a[1], sh), __funnelshift_r(a[1], a[2], sh),
                                                                __funnelshift_r(a[2], a[3], sh), __funnelshift_r(a[3], a[4], sh));
        *reinterpret_cast<uint4*>(row + (8 + ky) * 2048) = make_uint4(__funnelshift_r(b[0], b[1], sh), __funnelshift_r(b[1], b[2], sh),
                                                                      __funnelshift_r(b[2], b[3], sh), __funnelshift_r(b[3], b[4], sh));
      }
      if (it < 2) {                                    // the padding chunks of a stage never change
        *reinterpret_cast<uint4*>(row + 7 * 2048) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(row + 15 * 2048) = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8 * s);
    }
  } else if (warp < kMmaWarp) {
    // ---------------------------------------------------------------- epilogue -------------------
    const int ew = warp - 4, m = ew * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const int img = tile / tpi, r = tile - img * tpi;
      const int gy = (r / p.tiles_x) * 16 + (m >> 3), gx = (r % p.tiles_x) * 8 + (m & 7);
      const bool ok = gy < p.H && gx < p.W;
      mbar_wait(acc_full + 8 * as, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * NT);
      __nv_bfloat16* dst = p.out + (((size_t)img * p.H + gy) * p.W + gx) * NT;
#pragma unroll
      for (int j0 = 0; j0 < NT; j0 += 32) {
        uint32_t rr[32];
        tmem_ld32(trow + j0, rr);
        tmem_ld_wait();
        if (j0 + 32 == NT) { tc_fence_before(); mbar_arrive(acc_empty + 8 * as); }
        if (ok) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = pack_bf16x2(__uint_as_float(rr[8 * c + 2 * j]) + bias_s[j0 + 8 * c + 2 * j],
                                 __uint_as_float(rr[8 * c + 2 * j + 1]) + bias_s[j0 + 8 * c + 2 * j + 1]);
            *reinterpret_cast<uint4*>(dst + j0 + 8 * c) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- MMA issue ------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, W_BYTES);
      bulk_g2s(smem_u32(w_s), p.w, W_BYTES, w_full);
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(128, NT);
    const uint32_t hi128 = desc_hi(128);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_s), 2048), b_lo = desc_lo(smem_u32(w_s), NT * 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(acc_empty + 8 * s, ((it >> 1) & 1) ^ 1);
      mbar_wait(a_full + 8 * s, (it >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo0 + (uint32_t)(s * (A_STAGE >> 4));
#pragma unroll
        for (int k = 0; k < 8; ++k)   // K = 128 = 8 x 16
          umma_bf16_lh(tmem_base + (uint32_t)(s * NT), a_lo + (uint32_t)(2 * k * 128), hi128, b_lo + (uint32_t)(2 * k * NT), hi128, idesc,
                       k > 0 ? 1u : 0u);
        umma_commit(a_empty + 8 * s);
        umma_commit(acc_full + 8 * s);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 2 * NT);
}

// ------------------------------------------------------------------------------------------------------------------
// Toeplitz form
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTThreads = 7 * 32;    // warps 0-3 epilogue (TMEM lane quarters), 4 MMA, 5-6 builders
constexpr int kTMma = 4, kTBuild0 = 5;
constexpr int PR = 136;              // patch rows: 128 + 6 halo rows (+2 never read)
constexpr int PLANE = PR * 16;       // one 8-column plane of a bf16 patch
constexpr int TPATCH = 4 * PLANE;    // hi plane 0, hi plane 1, lo plane 0, lo plane 1
constexpr int TW_BYTES = 7 * 2 * 256 * 16;   // [7 ky][2 planes][256 n][16 B]

struct alignas(64) TParams {
  CUtensorMap map_out;               // out as (W * Cout, H, N): box (32 channels-wide column block, 128 rows, 1), 64B swizzle
  const float* x; const __nv_bfloat16* w; __nv_bfloat16* out;
  int N, H, W, tiles_x, tiles_y, ntiles;
  float bias[64];                    // read as constant-bank operands
};

template <int NT>
__global__ void __launch_bounds__(kTThreads, 2) conv7_toeplitz_kernel(const __grid_constant__ TParams p) {
  constexpr int XO = 256 / NT;                         // output columns of a tile
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w_s = smem;
  uint8_t* patch = w_s + TW_BYTES;                     // 2 stages
  uint8_t* o_s = patch + 2 * TPATCH;                   // 2 output staging buffers [128 rows][64 B], 64B-swizzled (1024-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_s + 2 * 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t w_full = smem_u32(bars), p_full = w_full + 8, p_empty = p_full + 16, acc_full = p_empty + 16, acc_empty = acc_full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(p_full + 8 * i, 2); mbar_init(p_empty + 8 * i, 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == kTMma) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_trigger();
  if (warp != kTMma) pdl_wait();                       // x is written by the previous timestep's update kernel (weights are constants)
  const int tpi = p.tiles_x * p.tiles_y;

  if (warp >= kTBuild0) {
    // ---------------------------------------------------------------- builders -------------------
    // item i = (patch row r = i >> 2, column quad c4 = i & 3): one aligned float4 of x (columns x0 - 4 + 4 c4 ..), split into bf16 hi + lo
    const int bt = threadIdx.x - kTBuild0 * 32;        // 0..63
    constexpr int ITEMS = 134 * 4, PER = (ITEMS + 63) / 64;
    float4 pv[PER];
    auto load_patch = [&](int tile) {
      const int img = tile / tpi, r0 = tile - img * tpi;
      const int y0 = (r0 / p.tiles_x) * 128, x0 = (r0 % p.tiles_x) * XO;
      const float* ximg = p.x + (size_t)img * p.H * p.W;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = bt + 64 * k;
        pv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < ITEMS) {
          const int gy = y0 - 3 + (i >> 2), gx = x0 - 4 + 4 * (i & 3);
          if ((unsigned)gy < (unsigned)p.H && gx >= 0 && gx + 4 <= p.W) pv[k] = __ldg(reinterpret_cast<const float4*>(ximg + (size_t)gy * p.W + gx));
        }
      }
    };
    int it = 0;
    if ((int)blockIdx.x < p.ntiles) load_patch(blockIdx.x);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(p_empty + 8 * s, ((it >> 1) & 1) ^ 1);
      uint8_t* ps = patch + s * TPATCH;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = bt + 64 * k;
        if (i < ITEMS) {
          const __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[k].x, pv[k].y), h1 = __floats2bfloat162_rn(pv[k].z, pv[k].w);
          const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
          uint8_t* q = ps + ((i >> 1) & 1) * PLANE + (i >> 2) * 16 + (i & 1) * 8;
          *reinterpret_cast<uint2*>(q) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
          *reinterpret_cast<uint2*>(q + 2 * PLANE) = make_uint2(pack_bf16x2(pv[k].x - f0.x, pv[k].y - f0.y), pack_bf16x2(pv[k].z - f1.x, pv[k].w - f1.y));
        }
      }
      if (tile + (int)gridDim.x < p.ntiles) load_patch(tile + gridDim.x);   // next tile's values in flight behind this tile's MMAs
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + 8 * s);
    }
  } else if (warp == kTMma) {
    // ---------------------------------------------------------------- MMA issue ------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, TW_BYTES);
      bulk_g2s(smem_u32(w_s), p.w, TW_BYTES, w_full);
    }
    __syncwarp();
    pdl_wait();
    mbar_wait(w_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(128, 256);
    const uint32_t hi128 = desc_hi(128);
    const uint32_t a_lo0 = desc_lo(smem_u32(patch), PLANE), b_lo0 = desc_lo(smem_u32(w_s), 256 * 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(acc_empty, (it & 1) ^ 1);
      mbar_wait(p_full + 8 * s, (it >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo0 + (uint32_t)(s * (TPATCH >> 4));
#pragma unroll
        for (int half = 0; half < 2; ++half)             // hi patch, lo patch: the same banded weights
#pragma unroll
          for (int ky = 0; ky < 7; ++ky)                 // filter row ky = the patch viewed ky rows further down
            umma_bf16_lh(tmem_base, a_lo + (uint32_t)(half * (2 * PLANE >> 4) + ky), hi128, b_lo0 + (uint32_t)(ky * (2 * 256 * 16 >> 4)), hi128,
                         idesc, (half | ky) ? 1u : 0u);
        umma_commit(p_empty + 8 * s);
        umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue -------------------
    // 32 accumulator columns at a time (one 64-byte piece of every row): bias, bf16, swizzled staging, ONE TMA store of the 128 x 64 B
    // block (a thread owns an image row: stores straight from registers touch 32 different lines per instruction and ran the launch
    // at 2.3 TB/s).  Rows beyond H and columns beyond W are clipped by the TMA unit.
    const int m = warp * 32 + lane;                      // image row of the tile = TMEM lane
    const int et = threadIdx.x;                          // 0..127
    int it = 0, nst = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int img = tile / tpi, r0 = tile - img * tpi;
      const int y0 = (r0 / p.tiles_x) * 128, x0 = (r0 % p.tiles_x) * XO;
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int j0 = 0; j0 < 256; j0 += 32, ++nst) {
        uint32_t rr[32];
        tmem_ld32(trow + j0, rr);
        tmem_ld_wait();
        if (j0 == 224) {                                   // the accumulator is read: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        const int c0 = j0 % NT;
        uint8_t* ob = o_s + (nst & 1) * 8192;
        if (et == 0) bulk_wait_group_read<1>();            // the store that read this buffer two pieces ago is done with it
        named_bar(1, 128);
        uint8_t* orow = ob + m * 64;
        const int sw = (m >> 1) & 3;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = pack_bf16x2(__uint_as_float(rr[8 * c + 2 * j]) + p.bias[c0 + 8 * c + 2 * j],
                               __uint_as_float(rr[8 * c + 2 * j + 1]) + p.bias[c0 + 8 * c + 2 * j + 1]);
          *reinterpret_cast<uint4*>(orow + ((c ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        named_bar(1, 128);
        if (et == 0) {
          tma_store_3d(&p.map_out, x0 * NT + j0, y0, img, smem_u32(ob));
          bulk_commit_group();
        }
      }
    }
    if (et == 0) bulk_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTMma) tmem_dealloc(tmem_base, 256);
}

constexpr int t_smem_bytes() { return TW_BYTES + 2 * TPATCH + 2 * 8192 + 8 * 8 + 16; }

template <int NT>
constexpr int smem_bytes() { return 2 * A_STAGE + 16 * NT * 16 + (4 * PATCH_W + NT) * 4 + 9 * 8 + 16; }

int g_sms = 0;

}  // namespace

int conv7_tc_pack(const float* w_tap_cout, const float* bias, int Cout, Conv7TcW* out) {
  out->ready = false;
  if (Cout != 32 && Cout != 64) return 0;
  // B operand [Cout rows][K = 128]: K chunk c8 < 7 = filter row ky = c8 (hi part), elements kx = 0..6, element 7 zero; chunk 7 zero;
  // chunks 8..15 repeat the same weights for the lo part
  std::vector<__nv_bfloat16> pk((size_t)16 * Cout * 8);
  for (int c8 = 0; c8 < 16; ++c8)
    for (int n = 0; n < Cout; ++n)
      for (int e = 0; e < 8; ++e) {
        const int ky = c8 & 7;
        const bool live = ky < 7 && e < 7;
        pk[((size_t)c8 * Cout + n) * 8 + e] = __float2bfloat16_rn(live ? w_tap_cout[(size_t)(ky * 7 + e) * Cout + n] : 0.f);
      }
  if (cudaMalloc(&out->w, pk.size() * 2) != cudaSuccess) return -1;
  if (cudaMemcpy(out->w, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  out->bias = nullptr;
  if (bias) {
    if (cudaMalloc(&out->bias, Cout * 4) != cudaSuccess) return -1;
    if (cudaMemcpy(out->bias, bias, Cout * 4, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  }
  // Toeplitz form: B_ky[n = xo * Cout + c][k] = w[ky][k - xo - 1][c] (K = patch columns x0 - 4 + k), [7 ky][2 planes of 8 k][256 n][8]
  {
    std::vector<__nv_bfloat16> tk((size_t)7 * 2 * 256 * 8);
    for (int ky = 0; ky < 7; ++ky)
      for (int k8 = 0; k8 < 2; ++k8)
        for (int n = 0; n < 256; ++n)
          for (int e = 0; e < 8; ++e) {
            const int k = k8 * 8 + e, xo = n / Cout, c = n % Cout, kx = k - xo - 1;
            tk[(((size_t)ky * 2 + k8) * 256 + n) * 8 + e] = __float2bfloat16_rn((kx >= 0 && kx < 7) ? w_tap_cout[(size_t)(ky * 7 + kx) * Cout + c] : 0.f);
          }
    if (cudaMalloc(&out->wt, tk.size() * 2) != cudaSuccess) return -1;
    if (cudaMemcpy(out->wt, tk.data(), tk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    for (int i = 0; i < 64; ++i) out->bias_h[i] = (bias && i < Cout) ? bias[i] : 0.f;
    if (cudaFuncSetAttribute(conv7_toeplitz_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem_bytes()) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(conv7_toeplitz_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem_bytes()) != cudaSuccess) return -1;
  }
  out->Cout = Cout;
  if (Cout == 32) { if (cudaFuncSetAttribute(conv7_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<32>()) != cudaSuccess) return -1; }
  else { if (cudaFuncSetAttribute(conv7_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<64>()) != cudaSuccess) return -1; }
  out->ready = true;
  return 0;
}

void conv7_tc_free(Conv7TcW* w) { cudaFree(w->w); cudaFree(w->wt); cudaFree(w->bias); *w = Conv7TcW(); }

int conv7_tc_launch(const Conv7TcW& w, const float* x, void* out, int N, int H, int W, cudaStream_t s) {
  if (!w.ready) return -1;
  if (!g_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev); }
  typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncFn enc = nullptr;
  static std::mutex mu;
  static std::unordered_map<uint64_t, CUtensorMap> maps;   // output view per (pointer, shape)
  static int toep = -1;   // env LD_CONV7_TOEPLITZ=0: the im2col form everywhere (A/B aid)
  if (toep < 0) { const char* e = getenv("LD_CONV7_TOEPLITZ"); toep = e ? atoi(e) : 1; }
  if (toep && W % 4 == 0 && ((uintptr_t)x & 15) == 0 && w.wt) {
    const int XO = 256 / w.Cout;
    TParams tp{};
    tp.x = x; tp.w = (const __nv_bfloat16*)w.wt; tp.out = (__nv_bfloat16*)out; tp.N = N; tp.H = H; tp.W = W;
    tp.tiles_x = (W + XO - 1) / XO; tp.tiles_y = (H + 127) / 128;
    tp.ntiles = N * tp.tiles_x * tp.tiles_y;
    {
      std::lock_guard<std::mutex> lock(mu);
      if (!enc) {
        void* f = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return -1;
        enc = (EncFn)f;
      }
      uint64_t key = (uint64_t)(uintptr_t)out * 1000003ull ^ ((uint64_t)N << 48) ^ ((uint64_t)H << 32) ^ ((uint64_t)W << 16) ^ (uint64_t)w.Cout;
      auto itm = maps.find(key);
      if (itm == maps.end()) {
        if (maps.size() > 1024) maps.clear();
        CUtensorMap m;
        const cuuint64_t dims[3] = {(cuuint64_t)W * w.Cout, (cuuint64_t)H, (cuuint64_t)N};
        const cuuint64_t strides[2] = {(cuuint64_t)W * w.Cout * 2, (cuuint64_t)H * W * w.Cout * 2};
        const cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
        if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -1;
        itm = maps.emplace(key, m).first;
      }
      tp.map_out = itm->second;
    }
    memcpy(tp.bias, w.bias_h, sizeof tp.bias);
    int grid = 2 * g_sms;
    if (grid > tp.ntiles) grid = tp.ntiles;
    if (w.Cout == 32) launch_k(conv7_toeplitz_kernel<32>, dim3(grid), dim3(kTThreads), t_smem_bytes(), s, true, tp);
    else launch_k(conv7_toeplitz_kernel<64>, dim3(grid), dim3(kTThreads), t_smem_bytes(), s, true, tp);
    return 1;
  }
  Params p{x, (const __nv_bfloat16*)w.w, w.bias, (__nv_bfloat16*)out, N, H, W, (W + 7) / 8, (H + 15) / 16, 0};
  p.ntiles = N * p.tiles_x * p.tiles_y;
  int grid = 2 * g_sms;
  if (grid > p.ntiles) grid = p.ntiles;
  if (w.Cout == 32) launch_k(conv7_tc_kernel<32>, dim3(grid), dim3(kThreads), smem_bytes<32>(), s, true, p);
  else launch_k(conv7_tc_kernel<64>, dim3(grid), dim3(kThreads), smem_bytes<64>(), s, true, p);
  return 1;
}

}  // namespace ld
