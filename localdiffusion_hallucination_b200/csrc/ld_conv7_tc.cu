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
#include <cuda_bf16.h>
#include <stdint.h>

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
  out->Cout = Cout;
  if (Cout == 32) { if (cudaFuncSetAttribute(conv7_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<32>()) != cudaSuccess) return -1; }
  else { if (cudaFuncSetAttribute(conv7_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<64>()) != cudaSuccess) return -1; }
  out->ready = true;
  return 0;
}

void conv7_tc_free(Conv7TcW* w) { cudaFree(w->w); cudaFree(w->bias); *w = Conv7TcW(); }

int conv7_tc_launch(const Conv7TcW& w, const float* x, void* out, int N, int H, int W, cudaStream_t s) {
  if (!w.ready) return -1;
  if (!g_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev); }
  Params p{x, (const __nv_bfloat16*)w.w, w.bias, (__nv_bfloat16*)out, N, H, W, (W + 7) / 8, (H + 15) / 16, 0};
  p.ntiles = N * p.tiles_x * p.tiles_y;
  int grid = 2 * g_sms;
  if (grid > p.ntiles) grid = p.ntiles;
  if (w.Cout == 32) launch_k(conv7_tc_kernel<32>, dim3(grid), dim3(kThreads), smem_bytes<32>(), s, true, p);
  else launch_k(conv7_tc_kernel<64>, dim3(grid), dim3(kThreads), smem_bytes<64>(), s, true, p);
  return 1;
}

}  // namespace ld
