// Full soft-max self-attention (Attend.forward math path, attend.py:98-113; Attention.forward ddpm.py:272-282)
// on tcgen05 / TMEM for sm_100a.  dim_head = 32, any number of heads, any token count n.
//
//   attn_prep_kernel   qkv [N][n][3*hid] (bf16, output of to_qkv) -> per (image, head, 128-token block) the exact
//                      shared-memory images of the UMMA operands: Q (pre-scaled by dim_head^-0.5 * log2 e) and K as
//                      K-major [4 chunks of 8 d][128 tokens][8], V transposed as [16 chunks of 8 tokens][32 d][8]
//                      (tokens past n are zero).  The attention kernel then needs no staging warps: every operand
//                      tile is one contiguous 8 KB cp.async.bulk (TMA engine).
//   attn_tc_kernel     one CTA per (128-query block, head, image), flash-style loop over 128-key blocks:
//                        S = Q K^T              tcgen05.mma  M=128 N=128 K=32, fp32 in TMEM
//                        p = exp2(S - m_new)    thread = query row (TMEM lane), running max / sum in registers
//                        P -> bf16 -> smem      (A operand of the second MMA)
//                        O_blk = P V            tcgen05.mma  M=128 N=32 K=128, fp32 in TMEM
//                        o = o * alpha + O_blk  in registers (one block behind, so it overlaps the next QK^T)
//                      Two CTAs are resident per SM (256 TMEM columns, ~73 KB shared memory each) and interleave
//                      their MMA / TMEM-read / exp phases.  No [n x n] matrix is materialised (the reference
//                      materialises [B, h, n, n] fp32, attend.py:102-109).
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

#include "ld_attn_tc.h"
#include "ld_launch.cuh"
#include "ld_tc_common.cuh"

namespace ld {

using namespace tc;

namespace {

constexpr int kThreads = 6 * 32;     // warps 0-3 soft-max (TMEM lane quarters), 4 MMA, 5 loader
constexpr int kMmaWarp = 4;
constexpr int kLoadWarp = 5;
constexpr int kTile = 8192;          // bytes of one operand tile (128 x 32 bf16)
constexpr int KV_STAGES = 2;
constexpr int kSmem = kTile + KV_STAGES * 2 * kTile + 128 * 128 * 2 + 16 * 8 + 16;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// grid (nblk, heads, N), 128 threads: thread = token of the block
__global__ void __launch_bounds__(128) attn_prep_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ Qp,
                                                        __nv_bfloat16* __restrict__ Kp, __nv_bfloat16* __restrict__ Vt, int n,
                                                        int heads) {
  pdl_wait();
  const int blk = blockIdx.x, h = blockIdx.y, img = blockIdx.z, nblk = gridDim.x;
  const int hid = heads * 32;
  const int r = threadIdx.x, tok = blk * 128 + r;
  const size_t tile = (((size_t)img * heads + h) * nblk + blk) * (kTile / 2);
  uint4 q[4], k[4], v[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) q[c] = k[c] = v[c] = make_uint4(0u, 0u, 0u, 0u);
  if (tok < n) {
    const __nv_bfloat16* src = qkv + ((size_t)img * n + tok) * 3 * hid + h * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      q[c] = *reinterpret_cast<const uint4*>(src + c * 8);
      k[c] = *reinterpret_cast<const uint4*>(src + hid + c * 8);
      v[c] = *reinterpret_cast<const uint4*>(src + 2 * hid + c * 8);
    }
  }
  const float qs = 0.17677669529663687f * 1.4426950408889634f;   // dim_head^-0.5 (attend.py:100) and log2(e)
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t w[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(w[j]); w[j] = pack_bf16x2(f.x * qs, f.y * qs); }
    *reinterpret_cast<uint4*>(Qp + tile + (size_t)c * 1024 + r * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(Kp + tile + (size_t)c * 1024 + r * 8) = k[c];
  }
  // V^T: element (d, token r) at chunk (r/8): [(r/8)][d][r%8]
  const __nv_bfloat16* vb = reinterpret_cast<const __nv_bfloat16*>(v);
  __nv_bfloat16* vt = Vt + tile + (size_t)(r >> 3) * 256 + (r & 7);
#pragma unroll
  for (int d = 0; d < 32; ++d) vt[d * 8] = vb[d];
}

struct AttnParams {
  const __nv_bfloat16* Qp; const __nv_bfloat16* Kp; const __nv_bfloat16* Vt;
  __nv_bfloat16* out;   // [N][n][hid], channel = h*32 + d (ddpm.py:281)
  int n, heads, nblk;
};

__global__ void __launch_bounds__(kThreads, 2) attn_tc_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* q_s = smem;
  uint8_t* kv_s = q_s + kTile;                      // [stage][K | V^T]
  uint8_t* p_s = kv_s + KV_STAGES * 2 * kTile;      // P: [16 chunks of 8 keys][128 rows][8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 128 * 128 * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t q_full = smem_u32(bars), kv_full = q_full + 8, kv_empty = kv_full + 8 * KV_STAGES, s_full = kv_empty + 8 * KV_STAGES,
                 s_empty = s_full + 8, p_full = s_empty + 8, p_empty = p_full + 8, o_full = p_empty + 8, o_empty = o_full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, img = blockIdx.z;
  const int nblk = p.nblk;
  const size_t head_tiles = ((size_t)img * p.heads + h) * nblk;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(kv_full + 8 * i, 1); mbar_init(kv_empty + 8 * i, 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, 4);      // one arrival per soft-max warp
    mbar_init(p_full, 4); mbar_init(p_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_empty, 4);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // S at columns 0..127, O_blk at 128..159
  pdl_wait();   // the operand images come from attn_prep_kernel (ld_launch.cuh)

  if (warp == kLoadWarp) {
    // ---------------------------------------------------------------- TMA bulk loads -------------
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTile);
      bulk_g2s(smem_u32(q_s), p.Qp + (head_tiles + qb) * (kTile / 2), kTile, q_full);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % KV_STAGES;
        mbar_wait(kv_empty + 8 * s, ((j / KV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full + 8 * s, 2 * kTile);
        bulk_g2s(smem_u32(kv_s + (size_t)s * 2 * kTile), p.Kp + (head_tiles + j) * (kTile / 2), kTile, kv_full + 8 * s);
        bulk_g2s(smem_u32(kv_s + (size_t)s * 2 * kTile + kTile), p.Vt + (head_tiles + j) * (kTile / 2), kTile, kv_full + 8 * s);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issue ------------------
    constexpr uint32_t idesc_s = make_idesc(128, 128), idesc_o = make_idesc(128, 32);
    const uint32_t hi128 = desc_hi(128);
    const uint32_t q_lo = desc_lo(smem_u32(q_s), 2048), k_lo0 = desc_lo(smem_u32(kv_s), 2048), v_lo0 = desc_lo(smem_u32(kv_s + kTile), 512),
                   p_lo = desc_lo(smem_u32(p_s), 2048);
    mbar_wait(q_full, 0);
    for (int j = 0; j < nblk; ++j) {
      const int s = j % KV_STAGES;
      mbar_wait(kv_full + 8 * s, (j / KV_STAGES) & 1);
      mbar_wait(s_empty, (j & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 2; ++k)   // d = 32 = 2 x 16
          umma_bf16_lh(tmem_base, q_lo + (uint32_t)(2 * k * 128), hi128, k_lo0 + (uint32_t)(s * (2 * kTile >> 4) + 2 * k * 128), hi128,
                       idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      mbar_wait(o_empty, (j & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 128 keys = 8 x 16
          umma_bf16_lh(tmem_base + 128u, p_lo + (uint32_t)(2 * k * 128), hi128, v_lo0 + (uint32_t)(s * (2 * kTile >> 4) + 2 * k * 32), hi128,
                       idesc_o, k > 0 ? 1u : 0u);
        umma_commit(kv_empty + 8 * s);
        umma_commit(p_empty);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- soft-max + accumulate ------
    const int row = warp * 32 + lane;                 // query row of the block = TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int kvalid = p.n - j * 128;               // keys of this block that exist
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // the whole 128-key row of S goes to registers ONCE (TMEM reads run at 64 B/clk per SM and each costs a round trip);
      // the stage is handed back to the MMA warp right after
      uint32_t sr[128];
      tmem_ld32(lane_base, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld32(lane_base + 32u, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      tmem_ld32(lane_base + 64u, *reinterpret_cast<uint32_t(*)[32]>(&sr[64]));
      tmem_ld32(lane_base + 96u, *reinterpret_cast<uint32_t(*)[32]>(&sr[96]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      if (kvalid < 128) {                              // ragged last block: keys past n carry no weight
#pragma unroll
        for (int c = 0; c < 128; ++c) if (c >= kvalid) sr[c] = 0xff800000u;   // -inf
      }
      float mx4[4] = {m, m, m, m};
#pragma unroll
      for (int c = 0; c < 128; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(sr[c]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float alpha = ex2_approx(m - mx);          // m = -inf on the first block: alpha = 0, o = l = 0 anyway
      m = mx;
      mbar_wait(p_empty, (j & 1) ^ 1);                 // the previous P has been consumed
      // p = exp2(s - m), row sum, P -> smem
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c8 = 0; c8 < 16; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float e0 = ex2_approx(__uint_as_float(sr[8 * c8 + 2 * c]) - mx), e1 = ex2_approx(__uint_as_float(sr[8 * c8 + 2 * c + 1]) - mx);
          s4[c] += e0 + e1;
          pk[c] = pack_bf16x2(e0, e1);
        }
        *reinterpret_cast<uint4*>(p_s + (size_t)c8 * 2048 + row * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      const float su = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      l = l * alpha + su;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // consume O_blk of the PREVIOUS block while the tensor core works on this one's P V
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        uint32_t orr[32];
        tmem_ld32(lane_base + 128u, orr);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(orr[d]));
      }
      alpha_prev = alpha;
    }
    {
      mbar_wait(o_full, (nblk - 1) & 1);
      tc_fence_after();
      uint32_t orr[32];
      tmem_ld32(lane_base + 128u, orr);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(orr[d]));
    }
    const int tok = qb * 128 + row;
    if (tok < p.n) {
      const float inv = 1.0f / l;
      __nv_bfloat16* dst = p.out + ((size_t)img * p.n + tok) * (p.heads * 32) + h * 32;
#pragma unroll
      for (int c = 0; c < 2; ++c) {   // 32-byte stores
        uint32_t w[8];
#pragma unroll
        for (int jx = 0; jx < 8; ++jx) w[jx] = pack_bf16x2(o[16 * c + 2 * jx] * inv, o[16 * c + 2 * jx + 1] * inv);
        st_global_v8(dst + 16 * c, w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 256);
}

bool g_configured = false;

}  // namespace

size_t attn_tc_scratch_bytes(int N, int n, int heads) {
  const size_t nblk = (size_t)(n + 127) / 128;
  return 3 * (size_t)N * heads * nblk * kTile;
}

int attn_tc_launch(const void* qkv, void* out, void* scratch, int N, int n, int heads, cudaStream_t s) {
  if (!g_configured) {
    if (cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return -1;
    g_configured = true;
  }
  const int nblk = (n + 127) / 128;
  const size_t tiles = (size_t)N * heads * nblk * (kTile / 2);
  __nv_bfloat16* Qp = (__nv_bfloat16*)scratch;
  __nv_bfloat16* Kp = Qp + tiles;
  __nv_bfloat16* Vt = Kp + tiles;
  launch_k(attn_prep_kernel, dim3(nblk, heads, N), dim3(128), 0, s, true, (const __nv_bfloat16*)qkv, Qp, Kp, Vt, n, heads);
  AttnParams p{Qp, Kp, Vt, (__nv_bfloat16*)out, n, heads, nblk};
  launch_k(attn_tc_kernel, dim3(nblk, heads, N), dim3(kThreads), kSmem, s, true, p);
  return 2;
}

int attn_tc_configure() {
  if (g_configured) return 0;
  if (cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return -1;
  g_configured = true;
  return 0;
}

}  // namespace ld
