// tcgen05 / TMEM implicit-GEMM convolution for sm_100a: persistent, warp-specialised, with the
// GroupNorm that surrounds every convolution of the denoiser fused into it.
//
//   D[128 pixels, NT couts] (fp32, TMEM)  +=  A[128 pixels, 16 cin] (bf16, smem) * B[NT couts, 16 cin] (bf16, smem)
//
// One CTA loops over tiles of 128 output pixels x NT output channels (tile = blockIdx.x + i * gridDim.x):
//   * 3x3 convs: a tile is a 16 x 8 pixel patch; its 18 x 10 halo patch is staged in shared memory ONCE per
//     64- (or 32-) channel chunk and all nine filter taps are issued as tcgen05.mma on *shifted views* of that
//     patch: the operand uses the canonical no-swizzle K-major layout ([8-channel chunk][pixel][16 B]), in which
//     moving by one pixel is a 16-byte move of the descriptor start address and the 8-row core-matrix groups of
//     the tile sit at a constant stride (SBO = halo pitch * 16 B).  Activations are read once, not nine times.
//   * 1x1 convs: a tile is 128 consecutive pixels of the flattened [N*H*W] axis, no halo.
//   * warp roles (14 warps): warps 0-7 are two producer teams that alternate pipeline stages (global -> registers
//     -> [GroupNorm affine + FiLM + SiLU/ReLU of the PREVIOUS layer, "normalise on load"] -> shared memory; direct
//     loads make zero padding, the virtual channel concat of two sources (torch.cat, ddpm.py:435-448) and the
//     nearest x2 up-sampling (ddpm.py:116) free); warps 8-11 run the epilogue (TMEM -> registers -> bias /
//     residual / GroupNorm statistics of THIS layer's output -> bf16 -> global); warp 12 issues the MMAs from one
//     lane; warp 13 drives the weight pipeline.  The accumulator is double-buffered in TMEM so the epilogue of
//     tile i overlaps the MMAs of tile i+1, and the activation ring is 3 deep.
//   * weights are pre-packed on the host into the exact shared-memory image of every (chunk, tap) stage and
//     copied with the TMA bulk-copy engine (cp.async.bulk -> UBLKCP): once per CTA when the whole filter fits in
//     shared memory (all high-resolution layers), else streamed through a 4-deep mbarrier ring per tile.
//
// Reference call sites this kernel serves: nn.Conv2d at ddpm.py:117,173,198,227,230,268,269,372,391 and
// unet_model.py:20,24,30 (every conv with Cin >= 32 and Cout >= 32); nn.GroupNorm + scale/shift + SiLU of
// Block.forward (ddpm.py:174-185) and GroupNorm + ReLU of BasicBlock (unet_model.py:21-25) as prologue/epilogue.
#include <cuda_bf16.h>
#include <stdint.h>

#include <vector>

#include "ld_conv_tc.h"
#include "ld_tc_common.cuh"

namespace ld {

using namespace tc;

namespace {

constexpr int kTeamThreads = 128;      // one producer team = 4 warps
constexpr int kTeams = 2;
constexpr int kEpiWarp0 = 8;           // warps 8..11 (warp % 4 == TMEM lane quarter)
constexpr int kMmaWarp = 12;
constexpr int kWWarp = 13;
constexpr int kThreads = 14 * 32;
constexpr int SA = 3;                  // activation stages
constexpr int SB = 4;                  // weight stages (streaming mode)
constexpr int kBatch = 6;              // 16-byte loads in flight per producer thread

struct KParams {
  const __nv_bfloat16* src0; const __nv_bfloat16* src1;
  int C0, C1;
  int N, H, W, Hin, Win, up;
  int tiles_x, tiles_y, ntiles;
  int nchunks;
  const __nv_bfloat16* w; const float* bias;
  int Cout;
  __nv_bfloat16* dst; const __nv_bfloat16* res;
  long long M;
  int resident, nb_stages, coef_floats;
  // normalise-on-load prologue (GroupNorm affine [+ FiLM] + activation of the source tensor)
  const double* pro_stats; const float* pro_gamma; const float* pro_beta; const float* pro_film;
  int pro_film_stride, pro_G, pro_act; float pro_eps;
  // GroupNorm statistics of the output
  double* stats; int stats_G;
};

template <int KS, int KC>
struct Geo {
  static constexpr int TH = KS == 3 ? 16 : 1;
  static constexpr int TW = KS == 3 ? 8 : 128;
  static constexpr int PITCH = TW + KS - 1;
  static constexpr int HPIX = (TH + KS - 1) * PITCH;                   // staged pixels per chunk
  static constexpr int CH = KC / 8;                                    // 16-byte channel groups per pixel
  static constexpr int LBO = ((HPIX * 16 + 127) / 128) * 128 + 16;     // == 16 (mod 128): conflict-free staging stores
  static constexpr int A_STAGE = CH * LBO;
  static constexpr int ITEMS = (HPIX * CH + kTeamThreads - 1) / kTeamThreads;
  static constexpr int SBO = (KS == 3 ? PITCH : 8) * 16;               // stride between 8-pixel core-matrix groups
};

// y = act(a * x + b) on 8 bf16 channels
__device__ __forceinline__ uint4 pro_apply(uint4 v, const float (&a)[8], const float (&b)[8], int act) {
  uint32_t in[4] = {v.x, v.y, v.z, v.w}, out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = unpack_bf16x2(in[j]);
    float y0 = fmaf(x.x, a[2 * j], b[2 * j]), y1 = fmaf(x.y, a[2 * j + 1], b[2 * j + 1]);
    if (act == 1) { y0 = silu_fast(y0); y1 = silu_fast(y1); }
    else if (act == 2) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
    out[j] = pack_bf16x2(y0, y1);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

// per-warp partial GroupNorm sums of 16 consecutive channels held by each lane (one pixel per lane)
template <int CPG>
__device__ __forceinline__ void stats_chunk(const float (&f)[16], bool valid, float* sacc, int grp0, int lane) {
  constexpr int NG = CPG >= 16 ? 1 : 16 / CPG;   // groups touched by this 16-channel chunk
  constexpr int W = CPG >= 16 ? 16 : CPG;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) { const float v = valid ? f[g * W + j] : 0.f; s += v; q = fmaf(v, v, q); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) { atomicAdd(sacc + 2 * (grp0 + g), s); atomicAdd(sacc + 2 * (grp0 + g) + 1, q); }
  }
}

template <int NT, int KS, int KC>
__global__ void __launch_bounds__(kThreads, 2) conv_tc_kernel(const KParams p) {
  using G = Geo<KS, KC>;
  constexpr int TAPS = KS * KS;
  constexpr int B_STAGE = NT * KC * 2;
  constexpr uint32_t TM_COLS = 2 * NT;   // two accumulator stages; NT in {32,64,128,256} -> power of two >= 64
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + SA * G::A_STAGE;
  float* coef = reinterpret_cast<float*>(b_s + (size_t)p.nb_stages * B_STAGE);
  float* sacc = coef + p.coef_floats;                 // [256] per-tile GroupNorm partial sums
  float* bias_s = sacc + 256;                         // [NT] bias of this CTA's output channels
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + NT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 4);
  const uint32_t a_full = smem_u32(bars), a_empty = a_full + 8 * SA, b_full = a_empty + 8 * SA, b_empty = b_full + 8 * SB,
                 acc_full = b_empty + 8 * SB, acc_empty = acc_full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(a_full + 8 * i, 4); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 128); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += kThreads) sacc[i] = 0.f;
  for (int i = threadIdx.x; i < NT; i += kThreads) bias_s[i] = p.bias ? p.bias[blockIdx.y * NT + i] : 0.f;
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tile = blockIdx.y;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp < kTeams * 4) {
    // ================================================================== producers =====================
    const int team = warp >> 2, tid = threadIdx.x & (kTeamThreads - 1);
    const int ch = tid % G::CH;                       // constant per thread: 128 % CH == 0
    float* cA = coef + team * 2 * (p.coef_floats / 4);  // [Cin] scale, then [Cin] shift
    float* cB = cA + p.coef_floats / 4;
    int cur_img = -1;
    int it_tile = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it_tile) {
      int img = 0, ty0 = 0, tx0 = 0;
      long long pix0 = 0;
      if (KS == 3) {
        img = tile / tiles_per_img;
        const int r = tile - img * tiles_per_img;
        const int ty = r / p.tiles_x;
        ty0 = ty * G::TH; tx0 = (r - ty * p.tiles_x) * G::TW;
      } else {
        pix0 = (long long)tile * 128;
      }
      if (p.pro_stats && img != cur_img) {
        // GroupNorm coefficients of the source tensor for this image (ddpm.py:174-185 folded to y = a x + b)
        named_bar(1 + team, kTeamThreads);            // nobody of this team still reads the old coefficients
        const int Cin = p.C0, cpg = Cin / p.pro_G;
        const double cnt = (double)p.Hin * p.Win * cpg;
        for (int c = tid; c < Cin; c += kTeamThreads) {
          const int g = c / cpg;
          const double su = p.pro_stats[((size_t)img * p.pro_G + g) * 2], sq = p.pro_stats[((size_t)img * p.pro_G + g) * 2 + 1];
          const double mean = su / cnt;
          double var = sq / cnt - mean * mean;
          if (var < 0) var = 0;
          const float rstd = (float)(1.0 / sqrt(var + (double)p.pro_eps));
          float a = rstd * p.pro_gamma[c], b = p.pro_beta[c] - (float)mean * a;
          if (p.pro_film) {
            const float sc = p.pro_film[(size_t)img * p.pro_film_stride + c] + 1.0f;
            const float sf = p.pro_film[(size_t)img * p.pro_film_stride + Cin + c];
            a *= sc; b = b * sc + sf;
          }
          cA[c] = a; cB[c] = b;
        }
        named_bar(1 + team, kTeamThreads);
        cur_img = img;
      }
      for (int c = 0; c < p.nchunks; ++c) {
        const int g = it_tile * p.nchunks + c;
        if ((g & 1) != team) continue;
        const int sa = g % SA;
        mbar_wait(a_empty + 8 * sa, ((g / SA) & 1) ^ 1);
        const int cbase = c * KC;
        const __nv_bfloat16* src; int cs, cb;
        if (cbase < p.C0) { src = p.src0; cs = p.C0; cb = cbase; } else { src = p.src1; cs = p.C1; cb = cbase - p.C0; }
        float pa[8], pb[8];
        if (p.pro_stats) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { pa[j] = cA[cbase + ch * 8 + j]; pb[j] = cB[cbase + ch * 8 + j]; }
        }
        uint8_t* stage = a_s + sa * G::A_STAGE + ch * G::LBO;
#pragma unroll 1
        for (int it0 = 0; it0 < G::ITEMS; it0 += kBatch) {
          uint4 v[kBatch];
          bool ok[kBatch];
#pragma unroll
          for (int k = 0; k < kBatch; ++k) {
            const int hp = ((it0 + k) * kTeamThreads + tid) / G::CH;
            v[k] = make_uint4(0u, 0u, 0u, 0u);
            ok[k] = false;
            if (it0 + k < G::ITEMS && hp < G::HPIX) {
              long long goff = -1;
              if (KS == 3) {
                const int hy = hp / G::PITCH, hx = hp - hy * G::PITCH;
                int gy = ty0 + hy - 1, gx = tx0 + hx - 1;
                if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
                  if (p.up) { gy >>= 1; gx >>= 1; }
                  goff = ((long long)img * p.Hin + gy) * p.Win + gx;
                }
              } else {
                const long long gp = pix0 + hp;
                if (gp < p.M) goff = gp;
              }
              if (goff >= 0) { v[k] = __ldg(reinterpret_cast<const uint4*>(src + goff * cs + cb + ch * 8)); ok[k] = true; }
            }
          }
#pragma unroll
          for (int k = 0; k < kBatch; ++k) {
            const int hp = ((it0 + k) * kTeamThreads + tid) / G::CH;
            if (it0 + k < G::ITEMS && hp < G::HPIX) {
              if (p.pro_stats && ok[k]) v[k] = pro_apply(v[k], pa, pb, p.pro_act);   // padding stays exactly zero
              *reinterpret_cast<uint4*>(stage + hp * 16) = v[k];
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + 8 * sa);
      }
    }
  } else if (warp < kMmaWarp) {
    // ================================================================== epilogue ======================
    const int ew = warp - kEpiWarp0, etid = threadIdx.x - kEpiWarp0 * 32;
    const int m = ew * 32 + lane;
    const int nbase = n_tile * NT;
    const int cpg = p.stats ? p.Cout / p.stats_G : 1;
    int it_tile = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it_tile) {
      const int as = it_tile & 1;
      int img = 0;
      long long opix = -1;
      if (KS == 3) {
        img = tile / tiles_per_img;
        const int r = tile - img * tiles_per_img;
        const int ty = r / p.tiles_x;
        const int gy = ty * G::TH + (m >> 3), gx = (r - ty * p.tiles_x) * G::TW + (m & 7);
        if (gy < p.H && gx < p.W) opix = ((long long)img * p.H + gy) * p.W + gx;
      } else {
        const long long gp = (long long)tile * 128 + m;
        if (gp < p.M) opix = gp;
      }
      mbar_wait(acc_full + 8 * as, (it_tile >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * NT);
#pragma unroll 1
      for (int j0 = 0; j0 < NT; j0 += 16) {
        uint32_t r[16];
        tmem_ld16(trow + j0, r);
        tmem_ld_wait();
        if (j0 + 16 == NT) {              // every TMEM read of this thread is complete: hand the stage back
          tc_fence_before();
          mbar_arrive(acc_empty + 8 * as);
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j0 + j);
          f[j] = __uint_as_float(r[j]) + b4.x; f[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
          f[j + 2] = __uint_as_float(r[j + 2]) + b4.z; f[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
        }
        if (p.stats) {
          const int grp0 = (nbase + j0) / cpg - nbase / cpg;
          const bool valid = opix >= 0;
          switch (cpg) {
            case 2: stats_chunk<2>(f, valid, sacc, grp0, lane); break;
            case 4: stats_chunk<4>(f, valid, sacc, grp0, lane); break;
            case 8: stats_chunk<8>(f, valid, sacc, grp0, lane); break;
            default: stats_chunk<16>(f, valid, sacc, grp0, lane); break;   // cpg >= 16: chunk inside one group
          }
        }
        if (opix >= 0) {
          const size_t o = (size_t)opix * p.Cout + nbase + j0;
          if (p.res) {
            const uint4 r0 = *reinterpret_cast<const uint4*>(p.res + o), r1 = *reinterpret_cast<const uint4*>(p.res + o + 8);
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 t2 = unpack_bf16x2(rr[j]);
              f[2 * j] += t2.x; f[2 * j + 1] += t2.y;
            }
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
          *reinterpret_cast<uint4*>(p.dst + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(p.dst + o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      if (p.stats) {
        // flush the tile's partial sums: one double atomic per (group, statistic)
        named_bar(3, 128);
        const int ng2 = 2 * (NT / cpg > 0 ? NT / cpg : 1);
        if (etid < ng2) {
          const float v = sacc[etid];
          sacc[etid] = 0.f;
          const int g = nbase / cpg + (etid >> 1);
          atomicAdd(p.stats + ((size_t)img * p.stats_G + g) * 2 + (etid & 1), (double)v);
        }
        named_bar(3, 128);
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================== MMA issue =====================
    // One thread issues every tcgen05.mma of the CTA, so its instruction stream is kept minimal: the descriptors
    // are (lo, hi) register pairs and every tap / k-step only adds a compile-time constant to `lo`.
    {
      constexpr uint32_t idesc = make_idesc(128, NT);
      const uint32_t a_hi = desc_hi(G::SBO), b_hi = desc_hi(128);
      const uint32_t a_lo0 = desc_lo(smem_u32(a_s), G::LBO), b_lo0 = desc_lo(smem_u32(b_s), NT * 16);
      int sb = 0; uint32_t pb = 0;
      int g = 0, it_tile = 0;
      if (p.resident) { mbar_wait(b_full, 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it_tile) {
        const int as = it_tile & 1;
        mbar_wait(acc_empty + 8 * as, ((it_tile >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(as * NT);
        uint32_t acc = 0;
        for (int c = 0; c < p.nchunks; ++c, ++g) {
          const int sa = g % SA;
          mbar_wait(a_full + 8 * sa, (g / SA) & 1);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)(sa * (G::A_STAGE >> 4));
          if (p.resident) {
            const uint32_t b_lo = b_lo0 + (uint32_t)(c * TAPS * (B_STAGE >> 4));
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < TAPS; ++tap) {
                const int ky = tap / KS, kx = tap - ky * KS;
#pragma unroll
                for (int k = 0; k < KC / 16; ++k) {
                  umma_bf16_lh(dcol, a_lo + (uint32_t)(ky * G::PITCH + kx + 2 * k * (G::LBO >> 4)), a_hi,
                               b_lo + (uint32_t)(tap * (B_STAGE >> 4) + 2 * k * NT), b_hi, idesc, acc);
                  acc = 1;
                }
              }
              umma_commit(a_empty + 8 * sa);
              if (c == p.nchunks - 1) umma_commit(acc_full + 8 * as);
            }
            acc = 1;
            __syncwarp();
          } else {
#pragma unroll 1
            for (int tap = 0; tap < TAPS; ++tap) {
              mbar_wait(b_full + 8 * sb, pb);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + (uint32_t)(sb * (B_STAGE >> 4));
              const int ky = tap / KS, kx = tap - ky * KS;
              const uint32_t a_t = a_lo + (uint32_t)(ky * G::PITCH + kx);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < KC / 16; ++k) {
                  umma_bf16_lh(dcol, a_t + (uint32_t)(2 * k * (G::LBO >> 4)), a_hi, b_lo + (uint32_t)(2 * k * NT), b_hi, idesc, acc);
                  acc = 1;
                }
                umma_commit(b_empty + 8 * sb);
                if (tap == TAPS - 1) {
                  umma_commit(a_empty + 8 * sa);
                  if (c == p.nchunks - 1) umma_commit(acc_full + 8 * as);
                }
              }
              acc = 1;
              __syncwarp();
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else {
    // ================================================================== weight pipeline ===============
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)n_tile * p.nchunks * TAPS * B_STAGE;
      const int total = p.nchunks * TAPS;
      if (p.resident) {
        if (blockIdx.x < p.ntiles) {
          mbar_arrive_expect_tx(b_full, (uint32_t)total * B_STAGE);
          for (int i = 0; i < total; ++i) bulk_g2s(smem_u32(b_s + (size_t)i * B_STAGE), wsrc + (size_t)i * B_STAGE, B_STAGE, b_full);
        }
      } else {
        int sb = 0; uint32_t pb = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
          for (int i = 0; i < total; ++i) {
            mbar_wait(b_empty + 8 * sb, pb ^ 1);
            mbar_arrive_expect_tx(b_full + 8 * sb, B_STAGE);
            bulk_g2s(smem_u32(b_s + sb * B_STAGE), wsrc + (size_t)i * B_STAGE, B_STAGE, b_full + 8 * sb);
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, TM_COLS);
}

constexpr size_t kResidentBudget = 200 * 1024;

template <int KS, int KC>
size_t smem_bytes(int NT, int nb_stages, int coef_floats) {
  return (size_t)SA * Geo<KS, KC>::A_STAGE + (size_t)nb_stages * NT * KC * 2 + (size_t)(coef_floats + 256 + NT) * 4 +
         (2 * SA + 2 * SB + 4) * 8 + 16;
}

struct Cfg { int max_smem = 0; int max_smem_sm = 0; int sms = 0; };
Cfg& cfg() {
  static Cfg c;
  if (!c.sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&c.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&c.max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  }
  return c;
}

// opt in to the maximum dynamic shared memory once per instantiation (done at pack time, outside any graph capture)
template <int NT, int KS, int KC>
int configure_one() {
  return cudaFuncSetAttribute(conv_tc_kernel<NT, KS, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg().max_smem) == cudaSuccess ? 0 : -1;
}
template <int KS, int KC>
int configure_nt(int NT) {
  switch (NT) {
    case 32: return configure_one<32, KS, KC>();
    case 64: return configure_one<64, KS, KC>();
    case 128: return configure_one<128, KS, KC>();
    case 256: return configure_one<256, KS, KC>();
  }
  return -1;
}

template <int NT, int KS, int KC>
int launch_one(KParams& p, int ntiles_y, size_t smem, cudaStream_t s) {
  // persistent grid: as many CTAs as can be resident (shared memory, 2*NT TMEM columns of 512 each)
  // (registers allow two CTAs per SM, see __launch_bounds__)
  int occ = (int)((size_t)cfg().max_smem_sm / (smem + 1024));
  if (occ > 2) occ = 2;
  const int tm = 512 / (2 * NT);
  if (occ > tm) occ = tm;
  if (occ < 1) occ = 1;
  const int occ_cache = occ;
  int gx = cfg().sms * occ_cache / ntiles_y;
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  conv_tc_kernel<NT, KS, KC><<<dim3((unsigned)gx, (unsigned)ntiles_y), kThreads, smem, s>>>(p);
  return 1;
}

template <int KS, int KC>
int launch_nt(int NT, KParams& p, int ntiles_y, cudaStream_t s) {
  // weights resident in shared memory when everything fits
  const int total = p.nchunks * KS * KS;
  const size_t res_bytes = smem_bytes<KS, KC>(NT, total, p.coef_floats);
  const size_t limit = (size_t)cfg().max_smem < kResidentBudget ? (size_t)cfg().max_smem : kResidentBudget;
  p.resident = res_bytes <= limit ? 1 : 0;
  p.nb_stages = p.resident ? total : SB;
  const size_t smem = smem_bytes<KS, KC>(NT, p.nb_stages, p.coef_floats);
  if (smem > (size_t)cfg().max_smem) return -1;
  switch (NT) {
    case 32: return launch_one<32, KS, KC>(p, ntiles_y, smem, s);
    case 64: return launch_one<64, KS, KC>(p, ntiles_y, smem, s);
    case 128: return launch_one<128, KS, KC>(p, ntiles_y, smem, s);
    case 256: return launch_one<256, KS, KC>(p, ntiles_y, smem, s);
  }
  return -1;
}

}  // namespace

static int pick_ntile(int Cout) {
  if (Cout <= 256) return (Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256) ? Cout : 0;
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  return 0;
}

int conv_tc_pack(const float* w, const float* bias, int Cin, int Cout, int ks, int stride, int pad, ConvTcW* out) {
  out->ready = false;
  if (!((ks == 3 && pad == 1) || (ks == 1 && pad == 0)) || stride != 1) return 0;
  const int nt = pick_ntile(Cout);
  if (!nt || Cin % 32) return 0;
  out->Cin = Cin; out->Cout = Cout; out->ks = ks; out->stride = stride; out->pad = pad; out->ntile = nt;
  const int taps = ks * ks;
  // two packings are kept: KC = 64 (when every concat source is a multiple of 64 channels) and KC = 32
  for (int v = 0; v < 2; ++v) {
    const int kc = v == 0 ? 64 : 32;
    void** slot = v == 0 ? &out->w : &out->w32;
    *slot = nullptr;
    if (Cin % kc) continue;
    const int nch = Cin / kc, ntiles = Cout / nt;
    std::vector<__nv_bfloat16> pk((size_t)Cin * Cout * taps);
    for (int nti = 0; nti < ntiles; ++nti)
      for (int c = 0; c < nch; ++c)
        for (int t = 0; t < taps; ++t)
          for (int k8 = 0; k8 < kc / 8; ++k8)
            for (int n = 0; n < nt; ++n)
              for (int e = 0; e < 8; ++e) {
                const int cin = c * kc + k8 * 8 + e, co = nti * nt + n;
                const size_t dst = (((((size_t)nti * nch + c) * taps + t) * (kc / 8) + k8) * nt + n) * 8 + e;
                pk[dst] = __float2bfloat16_rn(w[((size_t)t * Cin + cin) * Cout + co]);
              }
    if (cudaMalloc(slot, pk.size() * 2) != cudaSuccess) return -1;
    if (cudaMemcpy(*slot, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  }
  out->bias = nullptr;
  if (bias) {
    if (cudaMalloc(&out->bias, Cout * sizeof(float)) != cudaSuccess) return -1;
    if (cudaMemcpy(out->bias, bias, Cout * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  }
  if (ks == 3) { if (configure_nt<3, 64>(nt) || configure_nt<3, 32>(nt)) return -1; }
  else { if (configure_nt<1, 64>(nt) || configure_nt<1, 32>(nt)) return -1; }
  out->ready = true;
  return 0;
}

bool conv_tc_supports(const ConvTcW& w, const ConvTcArgs& a) {
  if (!w.ready) return false;
  if (a.C0 + a.C1 != w.Cin || a.C0 % 32 || a.C1 % 32) return false;
  if (a.up && (w.ks != 3 || a.src1)) return false;
  if (a.pro_stats && (w.ks != 3 || a.src1 || a.up || a.pro_G < 1 || a.C0 % a.pro_G)) return false;
  if (a.stats) {
    if (w.ks != 3 || a.stats_G < 1 || w.Cout % a.stats_G) return false;
    const int cpg = w.Cout / a.stats_G;
    if (!(cpg == 2 || cpg == 4 || cpg == 8 || (cpg >= 16 && cpg % 16 == 0))) return false;
    if (w.ntile % cpg || 2 * (w.ntile / cpg) > 128) return false;   // a group never straddles two n-tiles
  }
  return true;
}

int conv_tc_launch(const ConvTcW& w, const ConvTcArgs& a, cudaStream_t s) {
  if (!conv_tc_supports(w, a)) return -1;
  const bool k64 = w.w && a.C0 % 64 == 0 && a.C1 % 64 == 0;
  KParams p{};
  p.src0 = (const __nv_bfloat16*)a.src0; p.src1 = (const __nv_bfloat16*)a.src1; p.C0 = a.C0; p.C1 = a.C1;
  p.N = a.N; p.H = a.H; p.W = a.W; p.Hin = a.Hin; p.Win = a.Win; p.up = a.up;
  p.nchunks = w.Cin / (k64 ? 64 : 32);
  p.w = (const __nv_bfloat16*)(k64 ? w.w : w.w32); p.bias = w.bias; p.Cout = w.Cout;
  p.dst = (__nv_bfloat16*)a.dst; p.res = (const __nv_bfloat16*)a.res;
  p.M = (long long)a.N * a.H * a.W;
  p.pro_stats = a.pro_stats; p.pro_gamma = a.pro_gamma; p.pro_beta = a.pro_beta; p.pro_film = a.pro_film;
  p.pro_film_stride = a.pro_film_stride; p.pro_G = a.pro_G; p.pro_act = a.pro_act; p.pro_eps = a.pro_eps;
  p.coef_floats = a.pro_stats ? 4 * a.C0 : 0;   // two teams x (scale, shift)
  p.stats = a.stats; p.stats_G = a.stats_G;
  const int ny = w.Cout / w.ntile;
  if (w.ks == 3) {
    p.tiles_x = (a.W + 7) / 8; p.tiles_y = (a.H + 15) / 16;
    p.ntiles = a.N * p.tiles_x * p.tiles_y;
    return k64 ? launch_nt<3, 64>(w.ntile, p, ny, s) : launch_nt<3, 32>(w.ntile, p, ny, s);
  }
  p.tiles_x = 1; p.tiles_y = 1;
  p.ntiles = (int)((p.M + 127) / 128);
  return k64 ? launch_nt<1, 64>(w.ntile, p, ny, s) : launch_nt<1, 32>(w.ntile, p, ny, s);
}

}  // namespace ld
