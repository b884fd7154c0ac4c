// LinearAttention (ddpm.py:214-251) fused into two tcgen05 kernels for sm_100a (dim_head = 32; heads = 4, or 8 as two head groups of
// four: pass A runs once per group, pass B walks (tile, group) pairs and accumulates both groups into one output accumulator).
//
// The reference materialises qkv = to_qkv(RMSNorm(x)) ([384, H*W] per image, 768 B per pixel in bf16), two
// soft-maxes and two einsums.  Here x is read twice and the result written once; qkv never leaves the SM:
//
//   pass A  `la_ctx_kernel`   per 64-pixel half tile (two transform warp-groups ping-pong on alternate half tiles):
//        producers   x -> x/|x| (RMSNorm, g*sqrt(C) folded into the weights) -> bf16 -> smem, as [64 px][C] AND [C][64 px]
//        MMA1        K^T[128 (h,d)][64 px] = W'_k . xhat^T                                   (tcgen05, TMEM)
//        transform   thread = row (h,d): ek = exp(k - bound_d) -> bf16 -> smem P[(h,d)][px], row sums in registers
//        MMA2        Z[128 (h,d)][C] += P . xhat     accumulated in TMEM over all tiles of the CTA
//        end         Z and the row sums of ek are added to Z[n], ksum[n]
//      v = W'_v xhat is linear, so context[d][e] = sum_p ek[p,d] v[p,e] = sum_c W'_v[e,c] Z[d][c]: v is never formed
//      per pixel (half the TMEM read-out, a 4x smaller second MMA), W'_v moves into the per-image fold.
//      Soft-max over the pixel axis is shift invariant; instead of a separate max pass the shift is the
//      analytic bound |k_d| <= |W'_k[d,:]| (|xhat| = 1).  `la_fold_kernel` raises a flag if a row sum underflowed.
//   fold   `la_fold_kernel`   Mn[n][c'][(h,d)] = 32^-0.5 / ksum[n][(h,d)] * sum_c U[h][c'][c] Z[n][(h,d)][c],
//                             U[h] = Wout[:, h] . W'_v[h] (C x C, constant)                  (bf16, UMMA layout)
//   pass B  `la_out_kernel`   per 128-pixel tile (two transform warp-groups ping-pong on alternate tiles):
//        MMA1        Q[128 px][128 (h,d)] = xhat . W'_q^T
//        transform   thread = pixel: per-head soft-max over d -> bf16 -> smem P[px][(h,d)]
//        MMA2        O[128 px][C] = P . Mn^T
//        epilogue    + bias -> RMSNorm(g2) -> + x (residual of ddpm.py:425) -> bf16 -> global
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "ld_launch.cuh"
#include "ld_linattn_tc.h"
#include "ld_tc_common.cuh"

namespace ld {

using namespace tc;

namespace {

constexpr int kThreads = 13 * 32;  // warps 0-3 producers, 4 MMA, 5-12 transform/epilogue
constexpr int kMmaWarp = 4;
constexpr int kXfWarp0 = 5;
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct CtxParams {
  const __nv_bfloat16* x;   // [N][HW][C]
  const __nv_bfloat16* wk;  // packed [HG][C/8 + 2][128][8]
  const float* kb2;         // [HG * 128] log2(e) * bound of |k_d|
  float* Z;                 // [N][HG * 128 (h,d)][C]
  float* ksum;              // [N][HG * 128]
  int HW, slices;
  int hgs;                  // head groups (blockIdx.z)
};
struct OutParams {
  const __nv_bfloat16* x;   // [N][HW][C]
  const __nv_bfloat16* wq;  // packed [HG][C/8][128][8]
  const __nv_bfloat16* Mn;  // [N] packed [HG * 16][C][8]
  const float* bout; const float* g2;
  __nv_bfloat16* out;
  int HW, slices;
  int use_max;              // 0: |q| is provably small, soft-max over d needs no max subtraction
  int flat, N;              // flat: the CTAs split the N * tiles-per-image tile list evenly (a CTA may cross ONE image boundary)
};

// Producer side: ROWS pixels of x/|x| staged as a K-major operand (16-byte chunk (pixel p, channels 8*c8..) at
// c8*ROWS*16 + p*16).  Global loads run `XDEPTH` stages ahead of the shared-memory ring through a register FIFO, so
// that ~25-50 KB per SM are in flight (what HBM latency x bandwidth asks for) with only four producer warps.
template <int C, int ROWS>
struct XStage {
  static constexpr int LP = C / 8;                 // lanes per pixel
  static constexpr int ITEMS = ROWS * LP / 128;    // 16-byte loads per thread per stage
  static constexpr int DEPTH = ITEMS >= 16 ? 1 : ((12 / ITEMS) < 2 ? 2 : (12 / ITEMS));
  uint4 v[ITEMS];
  __device__ __forceinline__ void load(const __nv_bfloat16* __restrict__ ximg, int px0, int HW, int tid) {
    const int c8 = tid % LP;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int p = (k * 128 + tid) / LP;
      v[k] = make_uint4(0u, 0u, 0u, 0u);
      if (px0 + p < HW) v[k] = __ldg(reinterpret_cast<const uint4*>(ximg + (size_t)(px0 + p) * C + c8 * 8));
    }
  }
  // WITH_T: also write the transposed image [C rows][ROWS pixels] (K-major over pixels: 8-pixel chunk kc at
  // kc*C*16 + c*16 + (p%8)*2) right behind the first one -- the B operand of Z += P . xhat
  template <bool WITH_T>
  __device__ __forceinline__ void store(uint8_t* stage, int tid) const {
    const int c8 = tid % LP;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int p = (k * 128 + tid) / LP;
      const uint32_t in[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
      float f[8];
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16x2(in[j]);
        f[2 * j] = t.x; f[2 * j + 1] = t.y;
        ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss);
      }
#pragma unroll
      for (int o = 1; o < LP; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);   // F.normalize(x, dim=1) (ddpm.py:132)
      uint32_t o4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o4[j] = pack_bf16x2(f[2 * j] * inv, f[2 * j + 1] * inv);
      *reinterpret_cast<uint4*>(stage + (size_t)c8 * (ROWS * 16) + p * 16) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      if (WITH_T) {
        uint8_t* t = stage + ROWS * C * 2 + (size_t)(p >> 3) * (C * 16) + (size_t)(c8 * 8) * 16 + (p & 7) * 2;
#pragma unroll
        for (int e = 0; e < 8; ++e) *reinterpret_cast<uint16_t*>(t + e * 16) = (uint16_t)(o4[e >> 1] >> (16 * (e & 1)));
      }
    }
  }
};

// the producer loop shared by both passes: stage i covers pixels [(first + i) * ROWS, +ROWS) of one image
// EXT (pass A): two more 8-channel chunks behind the C/8 real ones: channel C is 1 for pixels that exist (0 past the end of
// the image), channels C+1 .. C+15 are 0.  MMA1 uses it to add the per-row soft-max shift, MMA2 to sum P over pixels.
template <int ROWS>
__device__ __forceinline__ void write_ext(uint8_t* stage, int chunks, int valid, int tid) {
  if (tid < 2 * ROWS) {
    const int which = tid / ROWS, p = tid - which * ROWS;
    const uint32_t one = (which == 0 && p < valid) ? 0x3F80u : 0u;   // bf16 1.0 in element 0
    *reinterpret_cast<uint4*>(stage + (size_t)(chunks + which) * (ROWS * 16) + p * 16) = make_uint4(one, 0u, 0u, 0u);
  }
}

template <int C, int ROWS, int XS, bool WITH_T, bool EXT = false>
__device__ __forceinline__ void produce_x(const __nv_bfloat16* __restrict__ ximg, int first, int count, int HW, uint8_t* x_s,
                                          int x_stage_bytes, uint32_t x_full, uint32_t x_empty, int tid, int lane) {
  using X = XStage<C, ROWS>;
  X fifo[X::DEPTH];
#pragma unroll
  for (int d = 0; d < X::DEPTH; ++d)
    if (d < count) fifo[d].load(ximg, (first + d) * ROWS, HW, tid);
  for (int i0 = 0; i0 < count; i0 += X::DEPTH) {
#pragma unroll
    for (int d = 0; d < X::DEPTH; ++d) {
      const int i = i0 + d;
      if (i < count) {
        const int s = i % XS;
        mbar_wait(x_empty + 8 * s, ((i / XS) & 1) ^ 1);
        fifo[d].template store<WITH_T>(x_s + (size_t)s * x_stage_bytes, tid);
        if (EXT) write_ext<ROWS>(x_s + (size_t)s * x_stage_bytes, C / 8, HW - (first + i) * ROWS, tid);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(x_full + 8 * s);
        if (i + X::DEPTH < count) fifo[d].load(ximg, (first + i + X::DEPTH) * ROWS, HW, tid);
      }
    }
  }
}

// Same producer fed by the bulk-copy engine: a stage of x is ROWS*C*2 contiguous bytes in global memory, so one lane
// issues cp.async.bulk copies RS stages ahead into a raw ring (no registers, no scoreboards tied up: the register
// FIFO above cannot keep more than ~6 load batches in flight per warp), and the four warps normalise smem -> smem.
// Thread = pixel: the whole channel row (C * 2 bytes) is summed and scaled by one thread -- no shuffles, C/8 independent 16-byte
// items in flight -- and for 64-pixel stages the four warps split into two teams that normalise alternate stages concurrently
// (ncu source view: with four lanes per pixel and one stage at a time the MMA warp waited on x_full > 50 % of the time in both
// passes; the producers, not the exponentials, bounded LinearAttention).  kXArrivals = warps that arrive per stage.
template <int ROWS> struct XTeams { static constexpr int TEAMS = 128 / ROWS, kArrivals = ROWS / 32; };
template <int C, int ROWS, int XS, int RS, bool EXT = false>
__device__ __forceinline__ void produce_x_raw(const __nv_bfloat16* __restrict__ ximg, int first, int count, int HW, uint8_t* x_s,
                                              int x_stage_bytes, uint32_t x_full, uint32_t x_empty, uint8_t* raw_s, uint32_t raw_full,
                                              uint32_t raw_empty, int tid, int lane) {
  constexpr int LPX = C / 8, RAWB = ROWS * C * 2, TEAMS = XTeams<ROWS>::TEAMS;
  static_assert(RS % TEAMS == 0 && XS % TEAMS == 0 || TEAMS == 1, "a team must own fixed ring slots (it has to observe every barrier phase)");
  const int team = tid / ROWS, p = tid - team * ROWS;
  int issued = 0;
  auto issue = [&](int j) {
    const int slot = j % RS;
    mbar_wait(raw_empty + 8 * slot, ((j / RS) & 1) ^ 1);
    const int px0 = (first + j) * ROWS;
    const int valid = HW - px0 < ROWS ? HW - px0 : ROWS;
    const uint32_t bytes = (uint32_t)valid * C * 2;
    mbar_arrive_expect_tx(raw_full + 8 * slot, bytes);
    bulk_g2s(smem_u32(raw_s + (size_t)slot * RAWB), ximg + (size_t)px0 * C, bytes, raw_full + 8 * slot);
  };
  for (int i = team; i < count; i += TEAMS) {
    if (tid == 0)   // one lane keeps RS raw stages in flight for both teams
      while (issued < count && issued < i + RS) issue(issued++);
    const int slot = i % RS, s = i % XS;
    const int valid = HW - (first + i) * ROWS;               // pixels of this stage that exist (may exceed ROWS)
    mbar_wait(raw_full + 8 * slot, (i / RS) & 1);
    const uint8_t* raw = raw_s + (size_t)slot * RAWB + (size_t)p * (C * 2);
    constexpr bool KEEP = LPX <= 4;                          // wider rows are read twice instead of held in registers
    uint4 v[KEEP ? LPX : 1];
    float ssp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < LPX; ++j) {
      const int cj = (j + p) & (LPX - 1);                    // rotated chunk order: 64/128-byte row strides would conflict 4-way
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if (p < valid) w = *reinterpret_cast<const uint4*>(raw + cj * 16);
      if (KEEP) v[j] = w;
      const uint32_t in[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_bf16x2(in[e]);
        ssp[e] = fmaf(t.x, t.x, ssp[e]); ssp[e] = fmaf(t.y, t.y, ssp[e]);
      }
    }
    const float ss = (ssp[0] + ssp[1]) + (ssp[2] + ssp[3]);
    const float inv = rsqrtf(fmaxf(ss, 1e-24f));             // 1 / max(|x|, 1e-12): F.normalize(x, dim=1) (ddpm.py:132)
    mbar_wait(x_empty + 8 * s, ((i / XS) & 1) ^ 1);
    uint8_t* stage = x_s + (size_t)s * x_stage_bytes + p * 16;
#pragma unroll
    for (int j = 0; j < LPX; ++j) {
      const int cj = (j + p) & (LPX - 1);
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if (KEEP) w = v[j];
      else if (p < valid) w = *reinterpret_cast<const uint4*>(raw + cj * 16);
      const uint32_t in[4] = {w.x, w.y, w.z, w.w};
      uint32_t o4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_bf16x2(in[e]);
        o4[e] = pack_bf16x2(t.x * inv, t.y * inv);
      }
      *reinterpret_cast<uint4*>(stage + (size_t)cj * (ROWS * 16)) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
    }
    if (EXT) {   // ones channel (0 past the end of the image) + 7 zero channels, then a zero chunk: see write_ext
      *reinterpret_cast<uint4*>(stage + (size_t)LPX * (ROWS * 16)) = make_uint4(p < valid ? 0x3F80u : 0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(stage + (size_t)(LPX + 1) * (ROWS * 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) { mbar_arrive(x_full + 8 * s); mbar_arrive(raw_empty + 8 * slot); }
  }
}

// ================================================================================================
// pass A: context
// ================================================================================================
template <int C>
struct CtxCfg {
  static constexpr int CE = C + 16;                   // channels incl. the ones / zero extension (see write_ext)
  static constexpr int CTAS = C >= 128 ? 1 : 2;       // co-resident CTAs per SM (C <= 64: 256 TMEM columns and < 113 KB each)
  static constexpr int NB = C >= 128 ? 4 : 2;         // K^T (TMEM) and P (smem) buffers: NB / 2 per transform warp-group
  static constexpr int LOGNB = C >= 128 ? 2 : 1;
  static constexpr int XS = C >= 128 ? 6 : (C >= 64 ? 4 : 6);   // xhat stages: released only after MMA2 (even, like RS: producer teams)
  static constexpr int X_STAGE = 64 * CE * 2;         // xhat [CE/8][64 px][16 B]: K-major for MMA1, MN-major for MMA2
  static constexpr int W_BYTES = 128 * CE * 2;
  static constexpr int P_BYTES = 128 * 64 * 2;        // one P buffer
  // raw x ring fed by cp.async.bulk (C = 128: register FIFO).  EVEN: the two producer teams take alternate stages, and a team must see
  // every phase of the raw_full barriers it waits on (with 3 slots a team met each slot every other use and could pass a stale phase)
  static constexpr int RS = C >= 128 ? 0 : (C >= 64 ? 2 : 4);
  static constexpr int RAW_STAGE = 64 * C * 2;
  static constexpr int ZCOL = NB * 64;                // TMEM: K^T buffer b at b*64 (64 pixels each); Z at ZCOL .. ZCOL+CE
  static constexpr int TMEM_COLS = C >= 128 ? 512 : 256;
  static constexpr int SMEM = W_BYTES + XS * X_STAGE + NB * P_BYTES + RS * RAW_STAGE + 48 * 8 + 16;
};

template <int C>
__global__ void __launch_bounds__(kThreads, CtxCfg<C>::CTAS) la_ctx_kernel(const CtxParams p) {
  using K = CtxCfg<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* w_s = smem;
  uint8_t* x_s = w_s + K::W_BYTES;
  uint8_t* p_s = x_s + K::XS * K::X_STAGE;            // [buffer][128 rows (h,d)][64 px]
  uint8_t* raw_s = p_s + K::NB * K::P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_s + K::RS * K::RAW_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);
  const uint32_t w_full = smem_u32(bars), x_full = w_full + 8, x_empty = x_full + 64, d1_full = x_empty + 64,
                 d1_empty = d1_full + 32, p_full = d1_empty + 32, p_empty = p_full + 32, z_full = p_empty + 32,
                 raw_full = z_full + 8, raw_empty = raw_full + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y, hg = blockIdx.z;
  const int HT = (p.HW + 63) / 64;
  const int h0 = (int)((long long)HT * blockIdx.x / p.slices), h1 = (int)((long long)HT * (blockIdx.x + 1) / p.slices);
  const int nh = h1 - h0;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < K::XS; ++i) { mbar_init(x_full + 8 * i, K::RS > 0 ? XTeams<64>::kArrivals : 4); mbar_init(x_empty + 8 * i, 1); }
    for (int i = 0; i < K::NB; ++i) {
      mbar_init(d1_full + 8 * i, 1); mbar_init(d1_empty + 8 * i, 4);
      mbar_init(p_full + 8 * i, 4); mbar_init(p_empty + 8 * i, 1);
    }
    mbar_init(z_full, 1);
    for (int i = 0; i < K::RS; ++i) { mbar_init(raw_full + 8 * i, 1); mbar_init(raw_empty + 8 * i, XTeams<64>::kArrivals); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), K::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_trigger();   // one resident wave: the next kernel's CTAs may be scheduled (ld_launch.cuh)
  pdl_wait();                            // x comes from the previous kernel
  // TMEM columns: K^T buffer b at b*64 (64 pixels each);  Z at ZCOL .. ZCOL+C, its column C = sum of P over pixels (ksum)
  if (nh <= 0) {
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, K::TMEM_COLS);
    return;
  }

  if (warp < 4) {
    // ---------------------------------------------------------------- producers ------------------
    const __nv_bfloat16* ximg = p.x + (size_t)n * p.HW * C;
    if constexpr (K::RS > 0)
      produce_x_raw<C, 64, K::XS, (K::RS > 0 ? K::RS : 1), true>(ximg, h0, nh, p.HW, x_s, K::X_STAGE, x_full, x_empty, raw_s, raw_full,
                                                                raw_empty, threadIdx.x, lane);
    else
      produce_x<C, 64, K::XS, false, true>(ximg, h0, nh, p.HW, x_s, K::X_STAGE, x_full, x_empty, threadIdx.x, lane);
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issue ------------------
    // The whole warp runs the loop (barrier waits), one elected lane issues tcgen05.mma / commit.  MMA1 runs NB
    // half tiles ahead of MMA2, so a transform warp-group always finds its next K^T tile ready.
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, K::W_BYTES);
      bulk_g2s(smem_u32(w_s), reinterpret_cast<const uint8_t*>(p.wk) + (size_t)hg * K::W_BYTES, K::W_BYTES, w_full);
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    tc_fence_after();
    // MMA2 reads xhat as an MN-major B operand (N = channels contiguous in 16-byte groups, K = pixels at 16 B stride):
    // the same shared-memory image MMA1 reads K-major, no transposed copy.  idesc bit 16 = B is MN-major.
    constexpr uint32_t idesc1 = make_idesc(128, 64), idesc2 = make_idesc(128, K::CE) | (1u << 16);
    const uint32_t hi128 = desc_hi(128), hi_xt = desc_hi(64 * 16);
    const uint32_t w_lo = desc_lo(smem_u32(w_s), 2048), x_lo0 = desc_lo(smem_u32(x_s), 1024), xt_lo0 = desc_lo(smem_u32(x_s), 128),
                   p_lo0 = desc_lo(smem_u32(p_s), 2048);
    auto mma1 = [&](int i) {            // K^T(i) = W'_k . xhat(i)^T
      const int s = i % K::XS, b = i & (K::NB - 1);
      mbar_wait(x_full + 8 * s, (i / K::XS) & 1);
      mbar_wait(d1_empty + 8 * b, ((i >> K::LOGNB) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t x_lo = x_lo0 + (uint32_t)(s * (K::X_STAGE >> 4));
#pragma unroll
        for (int k = 0; k < K::CE / 16; ++k)   // the last k-step multiplies the ones channel with the (negative) row shift
          umma_bf16_lh(tmem_base + (uint32_t)(b * 64), w_lo + (uint32_t)(2 * k * 128), hi128, x_lo + (uint32_t)(2 * k * 64), hi128, idesc1,
                       k > 0 ? 1u : 0u);
        umma_commit(d1_full + 8 * b);
      }
      __syncwarp();
    };
    for (int i = 0; i < K::NB && i < nh; ++i) mma1(i);
    for (int j = 0; j < nh; ++j) {      // Z += P(j) . xhat(j)
      const int b = j & (K::NB - 1), sj = j % K::XS;
      // K^T of the tile that reuses this buffer first: its transform group released the buffer right after reading it, long before
      // it delivers P(j) -- issuing it here (not after MMA2(j)) keeps that group's next tile ready when it comes back
      if (j + K::NB < nh) mma1(j + K::NB);
      mbar_wait(p_full + 8 * b, (j >> K::LOGNB) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_lo = p_lo0 + (uint32_t)(b * (K::P_BYTES >> 4)), xt_lo = xt_lo0 + (uint32_t)(sj * (K::X_STAGE >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k)   // K = 64 pixels = 4 x 16
          umma_bf16_lh(tmem_base + (uint32_t)K::ZCOL, p_lo + (uint32_t)(2 * k * 128), hi128, xt_lo + (uint32_t)(k * 16), hi_xt, idesc2,
                       (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(p_empty + 8 * b);
        umma_commit(x_empty + 8 * sj);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(z_full);
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- transform ------------------
    const int q = warp & 3, g = (warp - kXfWarp0) >> 2;      // TMEM lane quarter, warp-group (alternate half tiles)
    const int r = q * 32 + lane;                             // row (h,d)
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int i = g; i < nh; i += 2) {
      const int b = i & (K::NB - 1);
      const uint32_t par = (uint32_t)(i >> K::LOGNB) & 1u;
      uint8_t* pbuf = p_s + (size_t)b * K::P_BYTES + r * 16;
      mbar_wait(d1_full + 8 * b, par);
      tc_fence_after();
      // K^T arrives as log2(e) * k - shift_d (log2 e folded into W'_k, the shift added by MMA1 through the ones channel), so
      // the weight is one ex2 per element; its sum over pixels comes back from MMA2 (ones channel again): no FMA, no adds.
      // Pixels past the end of the image have xhat = 0 and ones = 0: their weight 2^0 multiplies zeros.
      uint32_t kr0[32], kr1[32];
      tmem_ld32(lane_base + (uint32_t)(b * 64), kr0);
      tmem_ld32(lane_base + (uint32_t)(b * 64 + 32), kr1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d1_empty + 8 * b);
      // exponentiate in registers first (packed in place), then wait for the P buffer: the MMA2 round trip of the tile that
      // last used it hides behind the ex2 work
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        kr0[j] = pack_bf16x2(ex2_approx(__uint_as_float(kr0[2 * j])), ex2_approx(__uint_as_float(kr0[2 * j + 1])));
        kr1[j] = pack_bf16x2(ex2_approx(__uint_as_float(kr1[2 * j])), ex2_approx(__uint_as_float(kr1[2 * j + 1])));
      }
      mbar_wait(p_empty + 8 * b, par ^ 1);                   // MMA2 of the tile that last used this P buffer has consumed it
#pragma unroll
      for (int c = 0; c < 4; ++c) {     // chunks of 8 pixels along K
        *reinterpret_cast<uint4*>(pbuf + (size_t)c * 2048) = make_uint4(kr0[4 * c], kr0[4 * c + 1], kr0[4 * c + 2], kr0[4 * c + 3]);
        *reinterpret_cast<uint4*>(pbuf + (size_t)(4 + c) * 2048) = make_uint4(kr1[4 * c], kr1[4 * c + 1], kr1[4 * c + 2], kr1[4 * c + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + 8 * b);
    }
    if (g == 0) {
      mbar_wait(z_full, 0);
      tc_fence_after();
      float* dst = p.Z + (((size_t)n * p.hgs + hg) * 128 + r) * C;
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t zr[32];
        tmem_ld32(lane_base + (uint32_t)(K::ZCOL + c0), zr);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) atomicAdd(dst + c0 + e, __uint_as_float(zr[e]));
      }
      uint32_t ks[16];
      tmem_ld16(lane_base + (uint32_t)(K::ZCOL + C), ks);
      tmem_ld_wait();
      atomicAdd(p.ksum + ((size_t)n * p.hgs + hg) * 128 + r, __uint_as_float(ks[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

// ================================================================================================
// fold: Mn[n][c'][(h,d)] = 32^-0.5 / ksum[n][(h,d)] * sum_c U[h][c'][c] * Z[n][(h,d)][c]
//       (bf16, K-major UMMA image [16 chunks of 8 (h,d)][C][8]);  Ut is U transposed: [h][c][c']
// ================================================================================================
// grid (4 heads, N), 256 threads, dynamic smem 32 * C floats
__global__ void __launch_bounds__(256) la_fold_kernel(const float* __restrict__ Z, const float* __restrict__ ksum,
                                                      const float* __restrict__ Ut, __nv_bfloat16* __restrict__ Mn, int C,
                                                      unsigned int* __restrict__ flag) {
  extern __shared__ float zs[];      // [32 d][C]: Z[(h,d)][c] * 32^-0.5 / ksum[(h,d)]; then the 32-column slice of U_h this block needs, [C c][32 c']
  const int h = blockIdx.x, n = blockIdx.y, cz = blockIdx.z, hid = (int)gridDim.x * 32;
  const float* u = Ut + (size_t)h * C * C + cz * 32;
  float* us = zs + 32 * C;
  const int lane = threadIdx.x & 31, dq = threadIdx.x >> 5;     // thread = (c' = 32 cz + lane, 4 consecutive d)
  for (int c = dq; c < C; c += 8) us[c * 32 + lane] = u[(size_t)c * C + lane];   // constants: before the wait on pass A
  if (threadIdx.x == 0) pdl_trigger();
  pdl_wait();
  __shared__ float inv[32];
  if (threadIdx.x < 32) {
    const float ks = ksum[(size_t)n * hid + h * 32 + threadIdx.x];
    if (cz == 0 && !(ks > 1e-30f) && flag) atomicAdd(flag, 1u);
    inv[threadIdx.x] = 0.17677669529663687f / ks;
  }
  __syncthreads();
  {
    const float* zsrc = Z + ((size_t)n * hid + h * 32) * C;
#pragma unroll 4
    for (int i = threadIdx.x; i < 32 * C; i += 256) zs[i] = zsrc[i] * inv[i / C];   // independent loads: the scale no longer sits between them
  }
  __syncthreads();
  // One block per (head, image, 32 output channels c'): 256 threads = 32 c' x 8 groups of 4 d.  Rows of the U slice are read
  // conflict-free, Z is a broadcast.  (The first version gave one block all C output channels with U read from global memory inside
  // the c loop: 34 us per launch at C = 128, 15 us at C = 32.)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* zr = zs + (dq * 4) * C;
#pragma unroll 8
  for (int c = 0; c < C; ++c) {
    const float uu = us[c * 32 + lane];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = fmaf(uu, zr[k * C + c], acc[k]);
  }
  // (h,d) = j: chunk j>>3 = h*4 + (dq >> 1), positions (dq & 1) * 4 + k  -> 8 contiguous bytes
  __nv_bfloat16* dst = Mn + (size_t)n * hid * C;
  *reinterpret_cast<uint2*>(dst + (size_t)(h * 4 + (dq >> 1)) * (C * 8) + (cz * 32 + lane) * 8 + (dq & 1) * 4) =
      make_uint2(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]));
}

// ================================================================================================
// pass B: output
// ================================================================================================
template <int C, int HG>
struct OutCfg {
  static constexpr int XS = C >= 64 ? 2 : 3;
  static constexpr int X_STAGE = 128 * C * 2;
  static constexpr int WQ_BYTES = HG * 128 * C * 2;   // W'_q of every head group
  static constexpr int MN_BYTES = HG * 128 * C * 2;   // Mn of every head group
  static constexpr int P_BYTES = 128 * 128 * 2;
  static constexpr int NP = (C >= 128 || (C >= 64 && HG > 1)) ? 2 : 4;   // P buffers: two per transform warp-group (one when smem is full)
  static constexpr int RS = C >= 128 ? 0 : (C >= 64 ? 2 : 3);   // raw x ring fed by cp.async.bulk
  static constexpr int MN2 = C <= 32 ? MN_BYTES : 0;   // second Mn buffer (flat tile lists cross one image boundary); no room at C >= 64
  static constexpr int SMEM_MIN = WQ_BYTES + MN_BYTES + MN2 + XS * X_STAGE + NP * P_BYTES + RS * X_STAGE + 2 * C * 4 + 32 * 8 + 16;
  static constexpr int SMEM = SMEM_MIN > 117 * 1024 ? SMEM_MIN : 117 * 1024;   // one CTA per SM: it owns all 512 TMEM columns
};

// HG head groups of four heads: the tile list is walked as VIRTUAL tiles j = tile * HG + hg.  Virtual tile j belongs to transform
// warp-group j & 1 (its Q buffer, its P buffers) exactly like a tile does for HG = 1; the second MMA of every virtual tile of a tile
// accumulates into the SAME output accumulator (O buffer = tile & 1), which is complete after the last head group; the epilogue of
// tile t is run by warp-group t & 1.
template <int C, int HG>
__global__ void __launch_bounds__(kThreads, 1) la_out_kernel(const OutParams p) {
  using K = OutCfg<C, HG>;
  constexpr int LHG = HG == 2 ? 1 : 0;
  constexpr int PPG = K::NP / 2;                      // P buffers per warp-group
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* wq_s = smem;
  uint8_t* mn_s = wq_s + K::WQ_BYTES;
  uint8_t* x_s = mn_s + K::MN_BYTES + K::MN2;           // Mn of the CTA's first image and (flat mode) of the next one
  uint8_t* p_s = x_s + K::XS * K::X_STAGE;
  uint8_t* raw_s = p_s + K::NP * K::P_BYTES;
  float* bg_s = reinterpret_cast<float*>(raw_s + K::RS * K::X_STAGE);   // bias[C], g2*sqrt(C)[C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bg_s + 2 * C);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const uint32_t w_full = smem_u32(bars), x_full = w_full + 8, x_empty = x_full + 8 * 3, d1_full = x_empty + 8 * 3,
                 d1_empty = d1_full + 16, p_full = d1_empty + 16, p_empty = p_full + 32, d2_full = p_empty + 32,
                 d2_empty = d2_full + 16, raw_full = d2_empty + 16, raw_empty = raw_full + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NTL = (p.HW + 127) / 128;
  // per-image slices (grid.y = image), or -- HW a multiple of 128 -- one flat tile list split evenly over ALL CTAs: with 32 images
  // and 148 SMs per-image slicing leaves 20 SMs idle.  In flat mode tile indices / pixels simply run on into image n + 1.
  int n, t0, nt, hw_lim;
  if (p.flat) {
    const long long T = (long long)p.N * NTL;
    const int f0 = (int)(T * blockIdx.x / gridDim.x), f1 = (int)(T * (blockIdx.x + 1) / gridDim.x);
    n = f0 / NTL; t0 = f0 - n * NTL; nt = f1 - f0; hw_lim = (p.N - n) * p.HW;
  } else {
    n = blockIdx.y;
    t0 = (int)((long long)NTL * blockIdx.x / p.slices);
    nt = (int)((long long)NTL * (blockIdx.x + 1) / p.slices) - t0;
    hw_lim = p.HW;
  }

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < K::XS; ++i) { mbar_init(x_full + 8 * i, 4); mbar_init(x_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(d1_full + 8 * i, 1); mbar_init(d1_empty + 8 * i, 4);     // one arrival per transform warp
      mbar_init(d2_full + 8 * i, 1); mbar_init(d2_empty + 8 * i, 128);
    }
    for (int i = 0; i < K::NP; ++i) { mbar_init(p_full + 8 * i, 4); mbar_init(p_empty + 8 * i, 1); }
    for (int i = 0; i < K::RS; ++i) { mbar_init(raw_full + 8 * i, 1); mbar_init(raw_empty + 8 * i, 4); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < C; i += kThreads) { bg_s[i] = p.bout[i]; bg_s[C + i] = p.g2[i] * sqrtf((float)C); }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_trigger();   // one resident wave (ld_launch.cuh)
  pdl_wait();                            // Mn comes from la_fold_kernel, x from the kernel before pass A
  // TMEM columns: Q of warp-group g at g*128;  O of warp-group g at 256 + g*C
  if (nt <= 0) {
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
    return;
  }
  const __nv_bfloat16* ximg = p.x + (size_t)n * p.HW * C;
  // P buffer of tile j: warp-group g = j & 1, its (j >> 1) % PPG -th buffer; completion parity of its k-th use
  auto pidx = [](int j) { return (j & 1) * PPG + ((j >> 1) % PPG); };
  auto ppar = [](int j) { return (uint32_t)(((j >> 1) / PPG) & 1); };

  if (warp < 4) {
    // ---------------------------------------------------------------- producers ------------------
    if constexpr (K::RS > 0)
      produce_x_raw<C, 128, K::XS, (K::RS > 0 ? K::RS : 1)>(ximg, t0, nt, hw_lim, x_s, K::X_STAGE, x_full, x_empty, raw_s, raw_full, raw_empty,
                                                           threadIdx.x, lane);
    else
      produce_x<C, 128, K::XS, false>(ximg, t0, nt, hw_lim, x_s, K::X_STAGE, x_full, x_empty, threadIdx.x, lane);
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issue ------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, K::WQ_BYTES + K::MN_BYTES + K::MN2);
      bulk_g2s(smem_u32(wq_s), p.wq, K::WQ_BYTES, w_full);
      bulk_g2s(smem_u32(mn_s), p.Mn + (size_t)n * HG * 128 * C, K::MN_BYTES, w_full);
      const int n2 = (p.flat && n + 1 < p.N) ? n + 1 : n;
      if (K::MN2) bulk_g2s(smem_u32(mn_s + K::MN_BYTES), p.Mn + (size_t)n2 * HG * 128 * C, K::MN_BYTES, w_full);
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc1 = make_idesc(128, 128), idesc2 = make_idesc(128, C);
    const uint32_t hi128 = desc_hi(128);
    const uint32_t wq_lo = desc_lo(smem_u32(wq_s), 2048), mn_lo = desc_lo(smem_u32(mn_s), C * 16), x_lo0 = desc_lo(smem_u32(x_s), 2048),
                   p_lo0 = desc_lo(smem_u32(p_s), 2048);
    auto mma2 = [&](int j) {
      const int t = j >> LHG, hg = j & (HG - 1), ob = t & 1, pi = pidx(j);
      mbar_wait(p_full + 8 * pi, ppar(j));
      if (hg == 0) mbar_wait(d2_empty + 8 * ob, ((t >> 1) & 1) ^ 1);   // the epilogue of tile t - 2 has read this accumulator
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_lo = p_lo0 + (uint32_t)(pi * (K::P_BYTES >> 4));
        const uint32_t mn_t = mn_lo + (uint32_t)((K::MN2 && t0 + t >= NTL) ? (K::MN_BYTES >> 4) : 0) + (uint32_t)(hg * 16 * C);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // K = 128 (h,d) of this head group = 8 x 16
          umma_bf16_lh(tmem_base + 256u + (uint32_t)(ob * C), p_lo + (uint32_t)(2 * k * 128), hi128, mn_t + (uint32_t)(2 * k * C), hi128,
                       idesc2, (hg > 0 || k > 0) ? 1u : 0u);
        umma_commit(p_empty + 8 * pi);
        if (hg == HG - 1) umma_commit(d2_full + 8 * ob);
      }
      __syncwarp();
    };
    auto mma1 = [&](int j) {
      const int t = j >> LHG, hg = j & (HG - 1);
      const int s = t % K::XS, g = j & 1, u = j >> 1;
      mbar_wait(x_full + 8 * s, (t / K::XS) & 1);
      mbar_wait(d1_empty + 8 * g, (u & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t x_lo = x_lo0 + (uint32_t)(s * (K::X_STAGE >> 4));
        const uint32_t wq_g = wq_lo + (uint32_t)(hg * (128 * C * 2 >> 4));
#pragma unroll
        for (int k = 0; k < C / 16; ++k)
          umma_bf16_lh(tmem_base + (uint32_t)(g * 128), x_lo + (uint32_t)(2 * k * 128), hi128, wq_g + (uint32_t)(2 * k * 128), hi128, idesc1,
                       k > 0 ? 1u : 0u);
        if (hg == HG - 1) umma_commit(x_empty + 8 * s);   // the stage is read by every head group of the tile
        umma_commit(d1_full + 8 * g);
      }
      __syncwarp();
    };
    // Q of virtual tile j+1 before the second MMA of j-1: its group frees the Q buffer three quarters into the soft-max of j-1
    // and delivers P(j-1) only at the end -- this order has Q(j+1) waiting when the group comes back
    const int nv = nt * HG;
    mma1(0);
    for (int j = 0; j < nv; ++j) {
      if (j + 1 < nv) mma1(j + 1);
      if (j >= 1) mma2(j - 1);
    }
    mma2(nv - 1);
  } else {
    // ---------------------------------------------------------------- transform + epilogue -------
    // Warp-group g owns tiles g, g+2, ...  Order: T(i), T(i+2), E(i), T(i+4), E(i+2), ... so that the epilogue of a
    // tile never waits for its own second MMA (with a single P buffer per group the order is T(i), E(i)).
    const int q = warp & 3, g = (warp - kXfWarp0) >> 2;
    const int m = q * 32 + lane;                    // pixel row inside the tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    auto transform = [&](int i) {
      const int u = i >> 1, pi = pidx(i);
      uint8_t* pbuf = p_s + (size_t)pi * K::P_BYTES + m * 16;
      mbar_wait(d1_full + 8 * g, u & 1);
      tc_fence_after();
      mbar_wait(p_empty + 8 * pi, ppar(i) ^ 1);     // the MMA2 that last read this P buffer is complete
      // one head = 32 TMEM columns of this pixel row; the next head's columns are in flight while this one is exponentiated
      auto head = [&](const uint32_t (&qr)[32], int h) {
        // q arrives pre-multiplied by log2(e) (folded into W_q): softmax_d(q) = 2^(q' - max) / sum
        float mx = 0.f;
        if (p.use_max) {
          mx = __uint_as_float(qr[0]);
#pragma unroll
          for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(qr[j]));
        }
        float e[32], s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) { e[j] = ex2_approx(__uint_as_float(qr[j]) - mx); s4[j & 3] += e[j]; }
        const float sc = 1.0f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));   // softmax(dim=-2) (ddpm.py:242); dim_head^-0.5 (ddpm.py:245) is in Mn
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o4[j] = pack_bf16x2(e[8 * c + 2 * j] * sc, e[8 * c + 2 * j + 1] * sc);
          *reinterpret_cast<uint4*>(pbuf + (size_t)(h * 4 + c) * 2048) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      };
      const uint32_t qbase = lane_base + (uint32_t)(g * 128);
      uint32_t qa[32], qb[32];
      tmem_ld32(qbase, qa);
      tmem_ld_wait();
      tmem_ld32(qbase + 32, qb);
      head(qa, 0);
      tmem_ld_wait();
      tmem_ld32(qbase + 64, qa);
      head(qb, 1);
      tmem_ld_wait();
      tmem_ld32(qbase + 96, qb);
      head(qa, 2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d1_empty + 8 * g);
      head(qb, 3);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + 8 * pi);
    };
    auto epilogue = [&](int i) {   // i = tile; run by warp-group i & 1 == g
      const int u = i >> 1;
      const int px = (t0 + i) * 128 + m;
      const __nv_bfloat16* xr = ximg + (size_t)px * C;
      __nv_bfloat16* orow = p.out + ((size_t)n * p.HW + px) * C;
      if constexpr (C <= 64) {
        // one sweep: the whole row stays in registers; the residual row of x is requested before the accumulator is read
        uint32_t xw[C / 2];                                   // the residual row, 32 bytes per load
        if (px < hw_lim) {
#pragma unroll
          for (int j = 0; j < C / 16; ++j) ld_global_nc_v8(xr + 16 * j, xw + 8 * j);
        }
        mbar_wait(d2_full + 8 * g, u & 1);
        tc_fence_after();
        float o[C];
        float ss = 0.f;
#pragma unroll
        for (int j0 = 0; j0 < C; j0 += 32) {
          uint32_t rr[32];
          tmem_ld32(lane_base + 256u + (uint32_t)(g * C + j0), rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { o[j0 + j] = __uint_as_float(rr[j]) + bg_s[j0 + j]; ss = fmaf(o[j0 + j], o[j0 + j], ss); }
        }
        tc_fence_before();
        mbar_arrive(d2_empty + 8 * g);
        if (px < hw_lim) {
          const float inv = rsqrtf(fmaxf(ss, 1e-24f));            // to_out RMSNorm (ddpm.py:231,251)
#pragma unroll
          for (int j0 = 0; j0 < C; j0 += 16) {   // 32-byte stores: a lane fills a whole sector of its pixel row
            uint32_t o8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 xf = unpack_bf16x2(xw[j0 / 2 + j]);
              o8[j] = pack_bf16x2(fmaf(o[j0 + 2 * j] * inv, bg_s[C + j0 + 2 * j], xf.x), fmaf(o[j0 + 2 * j + 1] * inv, bg_s[C + j0 + 2 * j + 1], xf.y));
            }
            st_global_v8(orow + j0, o8);
          }
        }
      } else {
        // two sweeps over the TMEM row (sum of squares, then normalise + store)
        mbar_wait(d2_full + 8 * g, u & 1);
        tc_fence_after();
        float ss = 0.f;
#pragma unroll 1
        for (int j0 = 0; j0 < C; j0 += 32) {
          uint32_t rr[32];
          tmem_ld32(lane_base + 256u + (uint32_t)(g * C + j0), rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float v = __uint_as_float(rr[j]) + bg_s[j0 + j]; ss = fmaf(v, v, ss); }
        }
        const float inv = rsqrtf(fmaxf(ss, 1e-24f));              // to_out RMSNorm (ddpm.py:231,251)
#pragma unroll 1
        for (int j0 = 0; j0 < C; j0 += 32) {
          uint32_t rr[32];
          tmem_ld32(lane_base + 256u + (uint32_t)(g * C + j0), rr);
          tmem_ld_wait();
          if (j0 + 32 == C) { tc_fence_before(); mbar_arrive(d2_empty + 8 * g); }
          if (px < hw_lim) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 xv = *reinterpret_cast<const uint4*>(xr + j0 + 8 * c);
              const uint32_t xi[4] = {xv.x, xv.y, xv.z, xv.w};
              uint32_t o4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int cc = 8 * c + 2 * j;
                const float2 xf = unpack_bf16x2(xi[j]);
                const float a = fmaf((__uint_as_float(rr[cc]) + bg_s[j0 + cc]) * inv, bg_s[C + j0 + cc], xf.x);
                const float b = fmaf((__uint_as_float(rr[cc + 1]) + bg_s[j0 + cc + 1]) * inv, bg_s[C + j0 + cc + 1], xf.y);
                o4[j] = pack_bf16x2(a, b);
              }
              *reinterpret_cast<uint4*>(orow + j0 + 8 * c) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            }
          }
        }
      }
    };
    // virtual tiles j = g, g + 2, ...: tile t = j / HG; this group runs the epilogue of the tiles with t & 1 == g
    const int nv = nt * HG;
    if constexpr (PPG == 2) {
      if (g < nv) transform(g);
      for (int j = g; j < nv; j += 2) {
        if (j + 2 < nv) transform(j + 2);
        const int t = j >> LHG;
        if (HG == 1 || (t & 1) == g) epilogue(t);
      }
    } else {
      for (int j = g; j < nv; j += 2) {
        transform(j);
        const int t = j >> LHG;
        if (HG == 1 || (t & 1) == g) epilogue(t);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

int g_sms = 0;
int sms() {
  if (!g_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev); }
  return g_sms;
}

template <int C, int HG>
int configure_c() {
  if (cudaFuncSetAttribute(la_ctx_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtxCfg<C>::SMEM) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(la_out_kernel<C, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, OutCfg<C, HG>::SMEM) != cudaSuccess) return -1;
  return 0;
}

template <int C, int HG>
int launch_c(const LinAttnTcW& w, const LinAttnTcArgs& a, cudaStream_t s) {
  // slices per image: ONE wave of CTAs (each owns all 512 TMEM columns of its SM), at least 4 half tiles per CTA
  const int HT = (a.HW + 63) / 64, NTL = (a.HW + 127) / 128;
  int sl = sms() / a.N; if (sl < 1) sl = 1;
  int slA = CtxCfg<C>::CTAS * sms() / (a.N * HG); if (slA > HT / 4) slA = HT / 4; if (slA < 1) slA = 1;
  int slB = sl; if (slB > NTL / 2) slB = NTL / 2; if (slB < 1) slB = 1;
  // pass B: when tiles never straddle images, all SMs share one flat tile list (a CTA crosses at most one image boundary)
  static int flat_env = -1; if (flat_env < 0) { const char* e = getenv("LD_LA_FLAT"); flat_env = e ? atoi(e) : 1; }   // LD_LA_FLAT=0: per-image slices (A/B aid)
  const long long Tb = (long long)a.N * NTL;
  int ctasB = sms(); if (ctasB > Tb / 2) ctasB = (int)(Tb / 2);
  const bool flatB = flat_env && OutCfg<C, HG>::MN2 > 0 && a.HW % 128 == 0 && ctasB >= 1 && Tb / ctasB + 1 <= NTL;
  CtxParams cp{(const __nv_bfloat16*)a.x, (const __nv_bfloat16*)w.wk, w.kb2, a.Z, a.ksum, a.HW, slA, HG};
  static int only = -1; if (only < 0) { const char* e = getenv("LD_LA_ONLY"); only = e ? atoi(e) : 0; }   // debug: 1 = pass A only, 2 = pass B only
  if (only != 2) launch_k(la_ctx_kernel<C>, dim3(slA, a.N, HG), dim3(kThreads), CtxCfg<C>::SMEM, s, true, cp);
  launch_k(la_fold_kernel, dim3(4 * HG, a.N, C / 32), dim3(256), 64 * C * sizeof(float), s, true, (const float*)a.Z, (const float*)a.ksum, (const float*)w.Ut,
           (__nv_bfloat16*)a.Mn, C, a.flag);
  pdl_after_small() = 1;
  OutParams op{(const __nv_bfloat16*)a.x, (const __nv_bfloat16*)w.wq, (const __nv_bfloat16*)a.Mn, w.bout, w.g2,
               (__nv_bfloat16*)a.out, a.HW, slB, w.q_use_max, flatB ? 1 : 0, a.N};
  if (only != 1) launch_k(la_out_kernel<C, HG>, flatB ? dim3(ctasB, 1) : dim3(slB, a.N), dim3(kThreads), OutCfg<C, HG>::SMEM, s, true, op);
  return 3;
}

}  // namespace

int linattn_tc_pack(const float* wqkv, const float* g, const float* wout, const float* bout, const float* g2, int C, int heads,
                    LinAttnTcW* out) {
  out->ready = false;
  // heads = 4, or 8 = two head groups (C <= 64: pass B keeps W'_q and Mn of both groups in shared memory)
  if (!((heads == 4 && (C == 32 || C == 64 || C == 128)) || (heads == 8 && (C == 32 || C == 64)))) return 0;
  out->C = C; out->heads = heads;
  const int HG = heads / 4, hid = heads * 32;
  const float sq = sqrtf((float)C);
  // fold RMSNorm's g * sqrt(C) (ddpm.py:131-132) into the 1x1 to_qkv weights (ddpm.py:227); rows: q 0..hid-1, k hid..2hid-1, v 2hid..3hid-1
  // a block = the 128 rows of one head group, packed [chunks][128][8] with `chunks` >= C/8 (extra chunks zero)
  auto pack_rows = [&](int row0, int chunks, std::vector<__nv_bfloat16>& dst, float extra) {
    dst.assign((size_t)HG * chunks * 128 * 8, __float2bfloat16_rn(0.f));
    for (int b = 0; b < HG; ++b)
      for (int c8 = 0; c8 < C / 8; ++c8)
        for (int r = 0; r < 128; ++r)
          for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            dst[(((size_t)b * chunks + c8) * 128 + r) * 8 + e] = __float2bfloat16_rn(wqkv[(size_t)(row0 + b * 128 + r) * C + c] * g[c] * sq * extra);
          }
  };
  std::vector<__nv_bfloat16> q, k;
  const int KCH = C / 8 + 2;
  pack_rows(0, C / 8, q, 1.4426950408889634f);   // q rows carry log2(e): the soft-max over d uses ex2 directly
  pack_rows(hid, KCH, k, 1.4426950408889634f);   // k rows carry log2(e) as well
  // Soft-max over the pixel axis is shift invariant: row d is shifted by the analytic bound |k'_d| <= |W'_k[d,:]| * |xhat|
  // (|xhat| <= 1 + 2^-8).  The (bf16) negative bound sits in the weight of the ones channel the producers append to xhat
  // (two extra 8-channel chunks: [C/8] = {-bound, 0 x 7}, [C/8 + 1] = 0), so MMA1 delivers k' - bound directly.
  std::vector<float> kb((size_t)hid);
  for (int b = 0; b < HG; ++b)
    for (int r = 0; r < 128; ++r) {
      double ss = 0;
      for (int c8 = 0; c8 < C / 8; ++c8)
        for (int e = 0; e < 8; ++e) { const double v = (double)__bfloat162float(k[(((size_t)b * KCH + c8) * 128 + r) * 8 + e]); ss += v * v; }
      kb[b * 128 + r] = (float)(sqrt(ss) * 1.01);
      k[(((size_t)b * KCH + C / 8) * 128 + r) * 8] = __float2bfloat16_rn(-kb[b * 128 + r]);
    }
  // |q'_d| <= |W'_q[d,:]|: when every bound is small the soft-max over d needs no max subtraction (2^q' cannot overflow)
  double qb = 0;
  for (int b = 0; b < HG; ++b)
    for (int r = 0; r < 128; ++r) {
      double ss = 0;
      for (int c8 = 0; c8 < C / 8; ++c8)
        for (int e = 0; e < 8; ++e) { const double v = (double)__bfloat162float(q[(((size_t)b * (C / 8) + c8) * 128 + r) * 8 + e]); ss += v * v; }
      if (sqrt(ss) > qb) qb = sqrt(ss);
    }
  out->q_use_max = qb * 1.01 > 60.0 ? 1 : 0;
  // Ut[h][c][c'] = sum_e Wout[c'][(h,e)] * W'_v[(h,e)][c],  W'_v = W_v * g * sqrt(C)   (to_out.0 o v-projection, per head)
  std::vector<float> ut((size_t)heads * C * C);
  for (int h = 0; h < heads; ++h)
    for (int c = 0; c < C; ++c)
      for (int cp = 0; cp < C; ++cp) {
        double a = 0;
        for (int e = 0; e < 32; ++e)
          a += (double)wout[(size_t)cp * hid + h * 32 + e] * (double)wqkv[(size_t)(2 * hid + h * 32 + e) * C + c] * (double)g[c] * (double)sq;
        ut[((size_t)h * C + c) * C + cp] = (float)a;
      }
  auto up = [](const void* h, size_t bytes, void** d) {
    if (cudaMalloc(d, bytes) != cudaSuccess) return -1;
    return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
  };
  if (up(q.data(), q.size() * 2, &out->wq) || up(k.data(), k.size() * 2, &out->wk) || up(kb.data(), kb.size() * 4, (void**)&out->kb2) ||
      up(ut.data(), ut.size() * 4, (void**)&out->Ut) || up(bout, C * 4, (void**)&out->bout) || up(g2, C * 4, (void**)&out->g2))
    return -1;
  int rc;
  if (HG == 1) rc = C == 32 ? configure_c<32, 1>() : C == 64 ? configure_c<64, 1>() : configure_c<128, 1>();
  else rc = C == 32 ? configure_c<32, 2>() : configure_c<64, 2>();
  if (rc) return -1;
  out->ready = true;
  return 0;
}

void linattn_tc_free(LinAttnTcW* w) {
  cudaFree(w->wq); cudaFree(w->wk); cudaFree(w->kb2); cudaFree(w->Ut); cudaFree(w->bout); cudaFree(w->g2);
  *w = LinAttnTcW();
}

int linattn_tc_launch(const LinAttnTcW& w, const LinAttnTcArgs& a, cudaStream_t s) {
  if (!w.ready) return -1;
  if (w.heads == 8) {
    if (w.C == 32) return launch_c<32, 2>(w, a, s);
    if (w.C == 64) return launch_c<64, 2>(w, a, s);
    return -1;
  }
  switch (w.C) {
    case 32: return launch_c<32, 1>(w, a, s);
    case 64: return launch_c<64, 1>(w, a, s);
    case 128: return launch_c<128, 1>(w, a, s);
  }
  return -1;
}

}  // namespace ld
