// tcgen05 / TMEM implicit-GEMM convolution for sm_100a: persistent, warp-specialised, TMA-fed, with the
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
//   * activations arrive by TMA (cp.async.bulk.tensor, one elected thread): the NHWC tensor is described to the
//     TMA unit as the 5-D view (8 ch, W, H, C/8, N), so one box copy of (8, 10, 18, KC/8, 1) lands the halo patch
//     directly in the UMMA operand layout, with hardware zero fill outside the image (conv padding) -- no staging
//     warps, no address arithmetic, several stages in flight.  Two cases keep a register staging path through two
//     producer teams (warps 0-7): the nearest x2 up-sampling read (ddpm.py:116) and the "normalise on load"
//     prologue, where GroupNorm affine + FiLM + SiLU/ReLU of the PREVIOUS layer is applied on the way to smem.
//   * warps 8-11 run the epilogue (TMEM -> registers -> bias / residual / GroupNorm statistics of THIS layer's
//     output -> bf16); for NT <= 64 the tile is written swizzled to shared memory and leaves by one TMA tensor
//     store (full-sector writes, hardware clipping at the image border), else by direct 16-byte stores.
//     Warp 12 issues the MMAs from one lane; warp 13 drives the weight pipeline.  The accumulator is
//     double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * weights are pre-packed on the host into the exact shared-memory image of every (chunk, tap) stage and
//     copied with the bulk-copy engine (cp.async.bulk): once per CTA when the whole filter fits in shared memory
//     (all high-resolution layers), else streamed through a 4-deep mbarrier ring per tile.
//
// Reference call sites this kernel serves: nn.Conv2d at ddpm.py:117,173,198,227,230,268,269,372,391 and
// unet_model.py:20,24,30 (every conv with Cin >= 32 and Cout >= 32); nn.GroupNorm + scale/shift + SiLU of
// Block.forward (ddpm.py:174-185) and GroupNorm + ReLU of BasicBlock (unet_model.py:21-25) as prologue/epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "ld_conv_tc.h"
#include "ld_launch.cuh"
#include "ld_tc_common.cuh"

namespace ld {

using namespace tc;

namespace {

#ifndef LD_CONV_MX
#define LD_CONV_MX 0                   // merged-x geometry for 3x3 / 32-cout layers (see Geo): parity-green, not faster yet (epilogue bound); -DLD_CONV_MX=1 for A/B builds
#endif
#ifndef LD_CONV_EG
#ifndef LD_CONV_SPLITLD
#define LD_CONV_SPLITLD 0             // narrow tiles: accumulator read-out in 16-column halves (see the epilogue)
#endif
#define LD_CONV_EG 2                   // epilogue warp-groups of the lean kernel (alternate tiles)
#endif
constexpr bool kUseMX = LD_CONV_MX != 0;
// LD_EXP (bit mask, A/B builds for TIMING ONLY -- results are wrong when set): 32 no transform work, 64 no proxy fence after the transform,
// 128 no register statistics; 1 no fence.proxy.async in the epilogue, 2 no named barriers,
// 4 no staging st.shared, 8 no TMEM read, 16 no cp.async.bulk.wait_group
#ifndef LD_EXP
#define LD_EXP 0
#endif
constexpr int kTeamThreads = 128;      // one producer team = 4 warps
constexpr int kTeams = 2;
// Warp roles.  Full kernel (register-staging producers): warps 0-7 producers, 8-11 epilogue, 12 MMA, 13 weights.
// Lean kernel (TMA activation loads only): warp 0 TMA, 1 MMA, 2 weights, 4-7 and 8-11 two epilogue warp-groups that
// drain alternate tiles (one accumulator stage each): a tile's epilogue is a ~1.4 k-clock dependent instruction chain of
// one warp per TMEM lane quarter (measured with LD_CONV_DBG=7), so what counts is how many of them run per SM.
template <bool LEAN> struct Roles {
  static constexpr int kProdWarps = LEAN ? 1 : 8;
  static constexpr int kEpiWarp0 = LEAN ? 4 : 8;       // warp % 4 == TMEM lane quarter
  static constexpr int kEpiGroups = LEAN ? LD_CONV_EG : 1;   // epilogue warp-groups; group g drains accumulator stage g
  static constexpr int kMmaWarp = LEAN ? 1 : 12;
  static constexpr int kWWarp = LEAN ? 2 : 13;
  static constexpr int kThreads = LEAN ? (4 + 4 * kEpiGroups) * 32 : 14 * 32;
  static constexpr int kMinCtas = LEAN ? (kEpiGroups == 2 ? 2 : 3) : 2;
};
constexpr int SA_MAX = 6;              // activation stages (runtime: 3..6)
constexpr int SB = 4;                  // weight stages (streaming mode)
constexpr int kBatch = 3;              // 16-byte loads in flight per producer thread (register staging path)

struct alignas(64) KParams {
  CUtensorMap map_a0, map_a1, map_out;   // TMA views of the two sources and of the output (see map_in_* / map_out)
  const __nv_bfloat16* src0; const __nv_bfloat16* src1;
  int C0, C1;
  int N, H, W, Hin, Win, up;
  int tiles_x, tiles_y, ntiles;
  int step_img, step_ty, step_tx;   // gridDim.x = step_img * tiles_x * tiles_y + step_ty * tiles_x + step_tx
  int nchunks;
  const __nv_bfloat16* w; const float* bias;
  int Cout;
  __nv_bfloat16* dst; const __nv_bfloat16* res;
  __nv_bfloat16* dst2; const float* bias2; int dual;   // fused 1x1 convolution of the same input (second accumulator)
  long long M;
  int resident, nb_stages, coef_floats;
  int tma_in, tma_out, sa;           // TMA activation loads / TMA output store / number of activation stages
  int xf;                            // lean kernel, warps 8-11 transform instead of draining tiles: 1 normalise the TMA-loaded patch in
                                     // place (pro_ab), 2 expand the low-resolution patch of a nearest x2 up-sampled source
  int off_raw, raw_stage;            // xf == 2: ring of raw low-resolution patches (bytes)
  int obuf;                          // output staging buffers (TMA store path): 2, or 4 = two per epilogue warp-group
  int bias_const;                    // 1: single n-tile of <= 64 channels, bias (and the fused 1x1's) read from bias_c (constant bank operands:
                                     // ncu r3q: the broadcast LDS.128 of the bias cost ~6 shared-memory wavefronts each next to the operand reads)
  float bias_c[128];
  int rot;                           // streamed weights: CTA b walks the taps of a chunk starting at tap b % 9 -- the CTAs of a launch no longer
                                     // ask the L2 for the same 32 KB weight stage at the same moment (the ring is latency x depth bound)
  int xhelp;                         // xf == 1 with resident weights: warps 2 and 3 (weights issued once / idle) join the four transform warps
  int regstats;                      // 1: narrow tiles keep the GroupNorm partial sums in registers (stats_acc); bit 1 (env LD_CONV_REGSTATS) also for dual launches
  int mt, nacc;                      // mt = 2: every streamed weight stage serves TWO consecutive tiles of the CTA (two accumulators);
                                     // nacc = accumulator stages (2, or 1 when two NT-wide accumulators already fill TMEM)
  int ds, ds_cs;                     // 2x2 stride-2 (pixel-unshuffle) conv run as a 1x1 over four strided TMA gathers; ds_cs = source channels
  int ps;                            // > 0: pixel-shuffle output of the folded nearest-x2 + 3x3 convolution (ConvTcArgs::ps = real Cout)
  int a_stage, lbo16;                // activation stage pitch (bytes) and chunk stride (16-byte units)
  int swz;                           // 1: TMA-fed stage = dense pixel rows of KC * 2 bytes, hardware-swizzled (64B for KC = 32, 128B for KC = 64):
                                     // the TMA unit moves 64 / 128-byte rows instead of the 16-byte rows of the no-swizzle K-major image (which
                                     // cost one shared-memory write cycle EACH: 720 per tile, as many as the operand reads of the 18 MMAs) and
                                     // tcgen05.mma reads the shifted tap views through the same swizzle (absolute address bits; tests/micro/swz_view.cu)
  int off_a, off_b, off_coef;        // shared memory carve-up (bytes)
  // normalise-on-load prologue (GroupNorm affine [+ FiLM] + activation of the source tensor)
  const float* pro_ab;   // [N][2][C0]: scale then shift (gn_coef_kernel), or null
  int pro_act;
  // GroupNorm statistics of the output
  double* stats; int stats_G;
  int contig;                        // contiguous tile walk (see TileWalk)
  int dbg;   // development aid (env LD_CONV_DBG): 1 no activation loads, 2 no MMAs, 4 no output stores
};

// MX ("merged x", 3x3 with 32 output channels): D[128, 32] tiles make every MMA read 4 KB of A for 1 KB of B, and the
// tensor pipe is bound by those shared-memory operand reads.  MX widens N to 96 = (kx, cout): the three horizontal taps
// share one A view, a tile is 8 rows x 16 patch columns (14 outputs + the 2 halo columns, so the 128 MMA rows are one
// linear run of patch pixels) and only the three vertical taps are issued as shifted views; the epilogue combines
// out[c] = D[c][kx=0] + D[c+1][kx=1] + D[c+2][kx=2] with two warp shuffles per channel.  A reads drop 3x.
template <int KS, int KC, bool MX = false>
struct Geo {
  static constexpr int TH = KS == 3 ? (MX ? 8 : 16) : 1;
  static constexpr int TW = KS == 3 ? (MX ? 14 : 8) : 128;
  static constexpr int PITCH = TW + KS - 1;
  static constexpr int HPIX = (TH + KS - 1) * PITCH;                   // staged pixels per chunk
  static constexpr int CH = KC / 8;                                    // 16-byte channel groups per pixel
  static constexpr int LBO_REG = ((HPIX * 16 + 127) / 128) * 128 + 16; // register path: == 16 (mod 128), conflict-free stores
  static constexpr int LBO_TMA = HPIX * 16;                            // TMA path: dense box image
  static constexpr int ITEMS = (HPIX * CH + kTeamThreads - 1) / kTeamThreads;
  static constexpr int SBO = MX ? 128 : (KS == 3 ? PITCH : 8) * 16;    // stride between 8-pixel core-matrix groups
  static constexpr int RP = TW / 2 + 2, RR = TH / 2 + 2;               // low-resolution patch behind an up-sampled halo patch
};

// walks tile = blockIdx.x + k * gridDim.x and keeps its (image, tile row, tile column) without integer divisions
struct TileWalk {
  int tile, end, img, ty, tx;
  // strided walk: tile = blockIdx.x + k * gridDim.x.  Contiguous walk (p.contig): CTA b owns one run of consecutive tiles,
  // so its tiles stay inside one or two images and per-image partial sums can live in the CTA across tiles.
  __device__ __forceinline__ void init(const KParams& p, bool k3) {
    if (p.contig) {
      const int base = p.ntiles / (int)gridDim.x, rem = p.ntiles - base * (int)gridDim.x, b = (int)blockIdx.x;
      tile = b * base + (b < rem ? b : rem);
      end = tile + base + (b < rem ? 1 : 0);
    } else { tile = blockIdx.x; end = p.ntiles; }
    img = 0; ty = 0; tx = 0;
    if (k3) {
      const int tpi = p.tiles_x * p.tiles_y;
      img = tile / tpi;
      const int r = tile - img * tpi;
      ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
    }
  }
  __device__ __forceinline__ void next(const KParams& p) {
    tile += p.contig ? 1 : (int)gridDim.x;
    tx += p.step_tx;
    if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
    ty += p.step_ty;
    if (ty >= p.tiles_y) { ty -= p.tiles_y; ++img; }
    img += p.step_img;
  }
};
// number of tiles this CTA walks
__device__ __forceinline__ int my_tile_count(const KParams& p) {
  const int G = (int)gridDim.x, b = (int)blockIdx.x;
  if (p.contig) { const int base = p.ntiles / G; return base + (b < p.ntiles - base * G ? 1 : 0); }
  return b < p.ntiles ? (p.ntiles - b + G - 1) / G : 0;
}

// ring position: stage index and phase parity, advanced without divisions
struct Ring {
  int s = 0; uint32_t ph = 0;
  __device__ __forceinline__ void advance(int n) { if (++s == n) { s = 0; ph ^= 1; } }
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

// same with the SiLU coefficients pre-halved (a / 2, b / 2): h = a' x + b' = y / 2, silu(y) = h + h tanh(h) -- one FMUL less per element
__device__ __forceinline__ uint4 pro_apply_h(uint4 v, const float (&a)[8], const float (&b)[8], int act) {
  uint32_t in[4] = {v.x, v.y, v.z, v.w}, out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = unpack_bf16x2(in[j]);
    float y0 = fmaf(x.x, a[2 * j], b[2 * j]), y1 = fmaf(x.y, a[2 * j + 1], b[2 * j + 1]);
    if (act == 1) { y0 = fmaf(y0, tanh_approx(y0), y0); y1 = fmaf(y1, tanh_approx(y1), y1); }
    else if (act == 2) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
    out[j] = pack_bf16x2(y0, y1);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

// The whole packed filter of a CTA in as few bulk copies as possible: a cp.async.bulk costs its issuing thread ~367 clk whatever its size
// (tests/micro/tma_rate.cu: 2 KB .. 32 KB pieces all take 367 clk each, 5.6 .. 89 B/clk) -- one copy per 2 KB tap stage kept the first
// MMA of every 32-channel launch waiting for 18 .. 40 of them (3 .. 7 us)
__device__ __forceinline__ void resident_weights(uint32_t dst, const uint8_t* src, uint32_t bytes, uint32_t bar) {
  constexpr uint32_t kPiece = 128 * 1024;
  for (uint32_t o = 0; o < bytes; o += kPiece) bulk_g2s(dst + o, src + o, bytes - o < kPiece ? bytes - o : kPiece, bar);
}

// SiLU variant with the coefficients pre-halved, bf16 pairs unpacked with one ALU op per element (low half: shift, high half: mask)
__device__ __forceinline__ uint4 pro_apply_silu_h(uint4 v, const float (&a)[8], const float (&b)[8]) {
  uint32_t in[4] = {v.x, v.y, v.z, v.w}, out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float h0 = fmaf(__uint_as_float(in[j] << 16), a[2 * j], b[2 * j]);
    const float h1 = fmaf(__uint_as_float(in[j] & 0xffff0000u), a[2 * j + 1], b[2 * j + 1]);
    out[j] = pack_bf16x2(fmaf(h0, tanh_approx(h0), h0), fmaf(h1, tanh_approx(h1), h1));
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

// In-place "normalise on load" of one landed activation stage by a team of NXW warps.  The stage is cut into pairs of 32-pixel blocks
// of one 8-channel group; warp xw takes PP consecutive pairs (same channel group as long as possible: its coefficients stay in
// registers).  Both 16-byte loads of a pair are issued before the arithmetic and nothing branches around them (clamped address,
// predicated store), so the two dependency chains (LDS -> FFMA -> MUFU.TANH -> FFMA -> pack -> STS) overlap: ncu r3i showed the old
// block-per-iteration loop, whose `if (inside)` bodies could not be interleaved, stalled on the LDS latency for 30 % of its samples.
// Padding pixels (hardware zero fill) are not rewritten: they stay exactly zero.
// SWZ: the stage holds dense pixel rows of CH * 16 bytes, hardware-swizzled (KParams::swz): the 16-byte chunk of channel group c8 of the
// pixel at absolute shared-memory address A sits at chunk c8 ^ ((A >> 7) & (CH - 1)) of its row.
template <int NXW, class G, bool SWZ>
__device__ __forceinline__ void xf_stage(uint8_t* stage, const float* ca, const float* cb, int xw, int lane, bool interior, int ty0,
                                         int tx0, int H, int W) {
  constexpr int NBLK = (G::HPIX + 31) / 32, HB = NBLK / 2, P = G::CH * HB, PP = P / NXW;
  static_assert(NBLK % 2 == 0 && P % NXW == 0 && (HB - 1) * 64 + 31 < G::HPIX, "pair split of the halo patch");
  float pa[8], pb[8];
  int cur = -1;
#pragma unroll 1
  for (int i = 0; i < PP; ++i) {
    const int pi = xw * PP + i, c8 = pi / HB, hp0 = (pi - c8 * HB) * 64 + lane, hp1 = hp0 + 32;
    if (c8 != cur) {
      const float4 a0 = *reinterpret_cast<const float4*>(ca + c8 * 8), a1 = *reinterpret_cast<const float4*>(ca + c8 * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(cb + c8 * 8), b1 = *reinterpret_cast<const float4*>(cb + c8 * 8 + 4);
      pa[0] = a0.x; pa[1] = a0.y; pa[2] = a0.z; pa[3] = a0.w; pa[4] = a1.x; pa[5] = a1.y; pa[6] = a1.z; pa[7] = a1.w;
      pb[0] = b0.x; pb[1] = b0.y; pb[2] = b0.z; pb[3] = b0.w; pb[4] = b1.x; pb[5] = b1.y; pb[6] = b1.z; pb[7] = b1.w;
      cur = c8;
    }
    bool ok0 = true, ok1 = hp1 < G::HPIX;
    if (!interior) {
      const int hy0 = hp0 / G::PITCH, hx0 = hp0 - hy0 * G::PITCH, hy1 = hp1 / G::PITCH, hx1 = hp1 - hy1 * G::PITCH;
      ok0 = (unsigned)(ty0 + hy0 - 1) < (unsigned)H && (unsigned)(tx0 + hx0 - 1) < (unsigned)W;
      ok1 = ok1 && (unsigned)(ty0 + hy1 - 1) < (unsigned)H && (unsigned)(tx0 + hx1 - 1) < (unsigned)W;
    }
    const int hp1c = hp1 < G::HPIX ? hp1 : hp0;
    uint4 *q0, *q1;
    if (SWZ) {
      constexpr int ROW = G::CH * 16;
      const uint32_t sa = smem_u32(stage);
      q0 = reinterpret_cast<uint4*>(stage + hp0 * ROW + ((c8 ^ (int)(((sa + hp0 * ROW) >> 7) & (G::CH - 1))) << 4));
      q1 = reinterpret_cast<uint4*>(stage + hp1c * ROW + ((c8 ^ (int)(((sa + hp1c * ROW) >> 7) & (G::CH - 1))) << 4));
    } else {
      uint8_t* col = stage + (size_t)c8 * G::LBO_TMA;
      q0 = reinterpret_cast<uint4*>(col + hp0 * 16);
      q1 = reinterpret_cast<uint4*>(col + hp1c * 16);
    }
    uint4 v0 = *q0, v1 = *q1;
    v0 = pro_apply_silu_h(v0, pa, pb);   // SiLU only (every ResnetBlock): other activations take the generic loop of the caller
    v1 = pro_apply_silu_h(v1, pa, pb);
    if (ok0) *q0 = v0;
    if (ok1) *q1 = v1;
  }
}

// per-warp partial GroupNorm sums of 16 consecutive channels held by each lane (one pixel per lane).  The NV = 2 * groups
// values of a lane are reduced together: every butterfly step halves the number of live values (a lane keeps the half its
// lane bit selects and sends the other), so NV values cost NV - 1 + log2(32 / NV) shuffles instead of 5 * NV.
template <int CPG>
__device__ __forceinline__ void stats_chunk(const float (&f)[16], bool valid, float* sacc, int grp0, int lane) {
  constexpr int NG = CPG >= 16 ? 1 : 16 / CPG;   // groups touched by this 16-channel chunk
  constexpr int W = CPG >= 16 ? 16 : CPG;
  constexpr int NV = 2 * NG;                     // 2, 4, 8 or 16 values
  constexpr int LOG = NV == 2 ? 1 : NV == 4 ? 2 : NV == 8 ? 3 : 4;
  float v[NV];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) { const float x = valid ? f[g * W + j] : 0.f; s += x; q = fmaf(x, x, q); }
    v[2 * g] = s; v[2 * g + 1] = q;
  }
#pragma unroll
  for (int st = 0; st < LOG; ++st) {
    const int off = 16 >> st, half = NV >> (st + 1);
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half], keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (int off = 16 >> LOG; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  // lane now holds value index (lane >> (5 - LOG)) = 2 * g + statistic, summed over the warp
  if ((lane & ((32 >> LOG) - 1)) == 0) atomicAdd(sacc + 2 * grp0 + (lane >> (5 - LOG)), v[0]);
}

// Register variant for narrow tiles (NT <= 64, at most 8 groups in the tile): every thread keeps the running sums of ITS pixel row over
// all the tiles of an image; the warp reduces them once, when the tile walk leaves the image (stats_flush_regs).  ncu r3i: the per-tile
// butterfly + shared-memory CAS loops of stats_chunk were ~250 of the ~500 instructions per warp and tile, and the single epilogue
// warp-group of the normalise-on-load launches was busy 83 % of the time.  `j0` is a constant after unrolling: the indices are static.
template <int CPG>
__device__ __forceinline__ void stats_acc(const float (&f)[16], float (&racc)[16], int j0) {
  constexpr int NG = 16 / CPG;                   // CPG in {4, 8, 16}
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) { const float x = f[g * CPG + j]; s += x; q = fmaf(x, x, q); }
    const int gi = j0 / CPG + g;
    racc[2 * gi] += s; racc[2 * gi + 1] += q;
  }
}
// 16 values per lane -> value (lane >> 1), summed over the warp, in the even lanes; added to sacc[0 .. nv)
__device__ __forceinline__ void stats_flush_regs(float (&racc)[16], float* sacc, int nv, int lane) {
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    const int off = 16 >> st, half = 8 >> st;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? racc[i] : racc[i + half], keep = up ? racc[i + half] : racc[i];
      racc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  const float v = racc[0] + __shfl_xor_sync(0xffffffffu, racc[0], 1);
  if (!(lane & 1) && (lane >> 1) < nv) atomicAdd(sacc + (lane >> 1), v);
#pragma unroll
  for (int i = 0; i < 16; ++i) racc[i] = 0.f;
}

template <int NT, int KS, int KC, bool LEAN, int MT = 1, bool RS = false>
__global__ void __launch_bounds__(Roles<LEAN>::kThreads, Roles<LEAN>::kMinCtas) conv_tc_kernel(const __grid_constant__ KParams p) {
  constexpr bool MX = kUseMX && KS == 3 && NT == 32;
  using G = Geo<KS, KC, MX>;
  using R = Roles<LEAN>;
  constexpr int kThreads = R::kThreads, kEpiWarp0 = R::kEpiWarp0, kMmaWarp = R::kMmaWarp, kWWarp = R::kWWarp;
  constexpr int TAPS = MX ? 3 : KS * KS;  // MMA taps per channel chunk (MX: vertical taps only)
  constexpr int NMMA = MX ? 3 * NT : NT;  // UMMA N
  constexpr int B_STAGE = NMMA * KC * 2;
  // two accumulator stages of NT (dual: 2 NT) fp32 columns; NT in {32,64,128,256} -> power of two >= 64
  const int NACC = p.nacc;   // accumulator stages: 2, 4 (narrow tiles: the MMA warp may run further ahead of the epilogue), or 1 (MT == 2, NT = 256)
  const int LNACC = NACC == 4 ? 2 : (NACC == 2 ? 1 : 0);
  const uint32_t acc_cols = MX ? 4 * NT : (p.dual ? 2 * NT : NT);   // MX: 96 + 32 (fused 1x1)
  const uint32_t acc_stride = (uint32_t)MT * acc_cols, TM_COLS = (uint32_t)NACC * acc_stride;
  const int TAPSW = TAPS + (p.dual ? 1 : 0);   // weight stages per channel chunk
  constexpr int O_ROW = NT * 2;          // bytes of one staged output row (TMA store path, NT <= 64)
  constexpr int DUALC = MX ? 3 * NT : NT; // first TMEM column of the fused-1x1 accumulator inside a stage
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* o_s = smem;                                // [2][128 rows][O_ROW] swizzled output tiles (TMA store path)
  uint8_t* a_s = smem + p.off_a;
  uint8_t* b_s = smem + p.off_b;
  float* coef = reinterpret_cast<float*>(smem + p.off_coef);
  float* sacc_all = coef + p.coef_floats;             // [2][256] GroupNorm partial sums of each epilogue warp-group
  float* bias_s = sacc_all + 512;                         // [NT] bias of this CTA's output channels
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 2 * NT);   // bias_s[NT..2NT): bias of the fused 1x1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * SA_MAX + 2 * SB + 8);
  const uint32_t a_full = smem_u32(bars), a_empty = a_full + 8 * SA_MAX, b_full = a_empty + 8 * SA_MAX, b_empty = b_full + 8 * SB,
                 acc_full = b_empty + 8 * SB, acc_empty = acc_full + 32, raw_full = acc_empty + 32;
  const bool xf = LEAN && p.xf;      // TMA -> raw_full -> transform warps -> a_full -> MMA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SA = p.sa;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(a_full + 8 * i, (p.tma_in && !xf) ? 1 : (xf && p.xhelp ? 6 : 4)); mbar_init(a_empty + 8 * i, 1); mbar_init(raw_full + 8 * i, 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, (MT == 2 && R::kEpiGroups == 2 && !xf) ? 256 : 128); }
    fence_barrier_init();
    if (p.tma_in) { tma_prefetch_desc(&p.map_a0); if (p.C1) tma_prefetch_desc(&p.map_a1); }
    if (p.tma_out) tma_prefetch_desc(&p.map_out);
  }
  for (int i = threadIdx.x; i < 512; i += kThreads) sacc_all[i] = 0.f;
  for (int i = threadIdx.x; i < NT; i += kThreads) {
    bias_s[i] = p.bias ? p.bias[blockIdx.y * NT + i] : 0.f;
    bias_s[NT + i] = (p.dual && p.bias2) ? p.bias2[blockIdx.y * NT + i] : 0.f;
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tile = blockIdx.y;
  // PDL (ld_launch.cuh): the grid is persistent (every CTA resident), so the next kernel's CTAs may be scheduled from now on; every
  // role that touches tensors of earlier kernels (activations, residual, GroupNorm coefficients / statistics) waits for them here --
  // the weight warp does not: weights and bias are constants, its loads overlap the previous kernel's tail
  if (threadIdx.x == 0) pdl_trigger();
  if (warp != kWWarp) pdl_wait();

  if (warp < R::kProdWarps) {
    if (LEAN || p.tma_in) {
      // ================================================================ TMA producer (one thread) =======
      if (warp == 0 && lane == 0) {
        Ring ra;
        TileWalk tw0;
        tw0.init(p, KS == 3 || p.ds);
        const uint32_t stage_bytes = p.xf == 2 ? (uint32_t)(G::CH * G::RP * G::RR * 16) : (uint32_t)(G::CH * G::LBO_TMA);
        // mt == 2: stages are filled in the order the MMA warp consumes them: (tile A, chunk c), (tile B, chunk c), ...
        // (two named walkers + scalar selects: an indexed pair would live in local memory)
        TileWalk twA = tw0, twB = tw0;
        while (twA.tile < twA.end) {
          int nm = 1;
          if (MT == 2) { twB = twA; twB.next(p); if (twB.tile < twB.end) nm = 2; }
          for (int cm = 0; cm < p.nchunks * nm; ++cm) {
            const int c = nm == 2 ? (cm >> 1) : cm;
            const bool second = nm == 2 && (cm & 1);
            struct { int tile, img, ty, tx; } tw = {second ? twB.tile : twA.tile, second ? twB.img : twA.img, second ? twB.ty : twA.ty,
                                                    second ? twB.tx : twA.tx};
            mbar_wait(a_empty + 8 * ra.s, ra.ph ^ 1);
            const int cbase = c * KC;
            const void* map = (p.ds || cbase < p.C0) ? (const void*)&p.map_a0 : (const void*)&p.map_a1;
            const int cb8 = (cbase < p.C0 ? cbase : cbase - p.C0) >> 3;
            const uint32_t dst = smem_u32(a_s + (size_t)ra.s * p.a_stage);
            const uint32_t fbar = (xf ? raw_full : a_full) + 8 * ra.s;
            mbar_arrive_expect_tx(fbar, stage_bytes);
            if (!(p.dbg & 1)) {
              if (KS == 3 && p.xf == 2)   // nearest x2 source (ddpm.py:116): the (TH/2+2) x (TW/2+2) low-resolution pixels under the halo patch
                tma_load_5d(smem_u32(smem + p.off_raw + (size_t)ra.s * p.raw_stage), map, 0, tw.tx * (G::TW / 2) - 1, tw.ty * (G::TH / 2) - 1, cb8,
                            tw.img, fbar);
              else if (KS == 3) {
                if (p.swz) tma_load_4d(dst, map, cb8 << 3, tw.tx * G::TW - 1, tw.ty * G::TH - 1, tw.img, fbar);
                else tma_load_5d(dst, map, 0, tw.tx * G::TW - 1, tw.ty * G::TH - 1, cb8, tw.img, fbar);
              } else if (p.ds) {
                // chunk c covers channels [cbase % Cs, +KC) of unshuffle tap q = cbase / Cs = (p1, p2): every other pixel
                // of the source starting at (2 ty + p1, 2 tx + p2) -- one strided TMA gather of 16 x 8 pixels
                const int q = cbase / p.ds_cs, cq = cbase - q * p.ds_cs;
                if (p.swz) tma_load_4d(dst, map, cq, 2 * (tw.tx * 8) + (q & 1), 2 * (tw.ty * 16) + (q >> 1), tw.img, fbar);
                else tma_load_5d(dst, map, 0, 2 * (tw.tx * 8) + (q & 1), 2 * (tw.ty * 16) + (q >> 1), cq >> 3, tw.img, fbar);
              } else if (p.swz) tma_load_2d(dst, map, cb8 << 3, tw.tile * 128, fbar);
              else tma_load_3d(dst, map, 0, tw.tile * 128, cb8, fbar);
            } else {
              // development aid: complete the transaction without data
              asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(fbar), "r"(stage_bytes) : "memory");
            }
            ra.advance(SA);
          }
          if (nm == 2) twA = twB;
          twA.next(p);
        }
      }
    } else if constexpr (!LEAN) {
      // ================================================================ register-staging producers ======
      const int team = warp >> 2, tid = threadIdx.x & (kTeamThreads - 1);
      const int ch = tid % G::CH;                       // constant per thread: 128 % CH == 0
      const int hp0 = tid / G::CH;                      // my first staged pixel; item `it` stages pixel hp0 + it * (128 / CH)
      constexpr int HSTEP = kTeamThreads / G::CH;
      Ring ra;                                          // position of global stage g
      int g = 0;
      TileWalk tw;
      tw.init(p, KS == 3);
      for (; tw.tile < tw.end; tw.next(p)) {
        if (p.nchunks == 1 && (g & 1) != team) { ++g; ra.advance(SA); continue; }
        const int img = tw.img, ty0 = tw.ty * G::TH, tx0 = tw.tx * G::TW;
        const long long pix0 = (long long)tw.tile * 128;
        // the whole halo patch lies inside the image: no bounds checks
        const bool interior = KS == 3 && ty0 > 0 && tx0 > 0 && ty0 + G::TH < p.H && tx0 + G::TW < p.W && !p.up;
        const long long tbase = ((long long)img * p.Hin + (ty0 - 1)) * p.Win + (tx0 - 1);   // halo origin (3x3, no up-sampling)
        for (int c = 0; c < p.nchunks; ++c, ++g, ra.advance(SA)) {
          if ((g & 1) != team) continue;
          mbar_wait(a_empty + 8 * ra.s, ra.ph ^ 1);
          const int cbase = c * KC;
          const __nv_bfloat16* src; int cs, cb;
          if (cbase < p.C0) { src = p.src0; cs = p.C0; cb = cbase; } else { src = p.src1; cs = p.C1; cb = cbase - p.C0; }
          src += cb + ch * 8;
          float pa[8], pb[8];
          if (p.pro_ab) {   // y = a x + b per (image, channel), precomputed by gn_coef_kernel
            const float4* ab = reinterpret_cast<const float4*>(p.pro_ab + ((size_t)img * 2) * p.C0 + cbase + ch * 8);
            const float4 a0 = __ldg(ab), a1 = __ldg(ab + 1);
            const float4* bb = reinterpret_cast<const float4*>(p.pro_ab + ((size_t)img * 2 + 1) * p.C0 + cbase + ch * 8);
            const float4 b0 = __ldg(bb), b1 = __ldg(bb + 1);
            pa[0] = a0.x; pa[1] = a0.y; pa[2] = a0.z; pa[3] = a0.w; pa[4] = a1.x; pa[5] = a1.y; pa[6] = a1.z; pa[7] = a1.w;
            pb[0] = b0.x; pb[1] = b0.y; pb[2] = b0.z; pb[3] = b0.w; pb[4] = b1.x; pb[5] = b1.y; pb[6] = b1.z; pb[7] = b1.w;
          }
          uint8_t* stage = a_s + (size_t)ra.s * p.a_stage + ch * G::LBO_REG + hp0 * 16;
#pragma unroll 1
          for (int it0 = 0; it0 < G::ITEMS; it0 += kBatch) {
            uint4 v[kBatch];
            bool ok[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
              const int hp = hp0 + (it0 + k) * HSTEP;
              v[k] = make_uint4(0u, 0u, 0u, 0u);
              ok[k] = false;
              if (it0 + k < G::ITEMS && hp < G::HPIX) {
                long long goff = -1;
                if (KS == 3) {
                  const int hy = hp / G::PITCH, hx = hp - hy * G::PITCH;
                  if (interior) {
                    goff = tbase + hy * p.Win + hx;
                  } else {
                    int gy = ty0 + hy - 1, gx = tx0 + hx - 1;
                    if ((unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W) {
                      if (p.up) { gy >>= 1; gx >>= 1; }
                      goff = ((long long)img * p.Hin + gy) * p.Win + gx;
                    }
                  }
                } else {
                  const long long gp = pix0 + hp;
                  if (gp < p.M) goff = gp;
                }
                if (goff >= 0 && !(p.dbg & 1)) { v[k] = __ldg(reinterpret_cast<const uint4*>(src + goff * cs)); ok[k] = true; }
              }
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
              const int hp = hp0 + (it0 + k) * HSTEP;
              if (it0 + k < G::ITEMS && hp < G::HPIX) {
                if (p.pro_ab && ok[k]) v[k] = pro_apply(v[k], pa, pb, p.pro_act);   // padding stays exactly zero
                *reinterpret_cast<uint4*>(stage + (it0 + k) * (HSTEP * 16)) = v[k];
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full + 8 * ra.s);
        }
      }
    }
  } else if (xf && ((warp >= kEpiWarp0 + 4 && warp < kEpiWarp0 + 8) || (p.xhelp && (warp == kWWarp || warp == kWWarp + 1)))) {
    // ================================================================== in-place transform ============
    // "normalise on load" for TMA-fed launches: the raw patch landed in the operand layout.  The stage is cut into units of (8-channel
    // group c8, block of 32 patch pixels); transform warp xw of nxw takes the units xw, xw + nxw, ..: coefficients of the unit's channel
    // group in registers, lanes on consecutive pixels (conflict-free 16-byte accesses).  With resident weights the weight warp (after
    // issuing its one bulk copy) and the spare warp join the four transform warps (ncu r3i: the transform warps were busy 80 % of the
    // launch and, once the GroupNorm sums of the epilogue moved into registers, the stage that bounds it).
    // Padding pixels (hardware zero fill) stay exactly zero, like the reference's conv padding of the normalised tensor.
    if (warp == kWWarp && lane == 0) {   // helper mode implies resident weights: the whole filter, once
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)n_tile * p.nchunks * TAPSW * B_STAGE;
      const int total = p.nchunks * TAPSW;
      mbar_arrive_expect_tx(b_full, (uint32_t)total * B_STAGE);
      resident_weights(smem_u32(b_s), wsrc, (uint32_t)total * B_STAGE, b_full);
    }
    if (warp == kWWarp) pdl_wait();      // (skipped above for the weight warp) the coefficient table comes from an earlier kernel
    __syncwarp();
    const int nxw = p.xhelp ? 6 : 4;
    const int xw = warp >= kEpiWarp0 + 4 ? warp - (kEpiWarp0 + 4) : 4 + (warp - kWWarp);
    const int xtid = xw * 32 + lane;                       // 0 .. 32 * nxw - 1 inside the transform team
    int coef_img = -1;                                     // image whose coefficient table sits in shared memory (xf == 1)
    Ring ra;
    TileWalk twA, twB;
    twA.init(p, true);
    twB = twA;
    while (twA.tile < twA.end) {
      int nm = 1;
      if (MT == 2) { twB = twA; twB.next(p); if (twB.tile < twB.end) nm = 2; }
      for (int cm = 0; cm < p.nchunks * nm; ++cm) {
        const int c = nm == 2 ? (cm >> 1) : cm;
        const bool second = nm == 2 && (cm & 1);
        const int img = second ? twB.img : twA.img, ty0 = (second ? twB.ty : twA.ty) * G::TH, tx0 = (second ? twB.tx : twA.tx) * G::TW;
        const bool interior = ty0 > 0 && tx0 > 0 && ty0 + G::TH < p.H && tx0 + G::TW < p.W;
        if (p.xf == 1 && img != coef_img) {
          // y = a x + b table of this image -> shared memory (ncu: fetched from global memory per stage, the loads were exposed for an L2
          // round trip on every tile).  For SiLU the table holds a / 2, b / 2: silu(y) = h + h tanh(h) with h = y / 2.
          named_bar(5, 32 * nxw);                          // every transform warp is done with the previous image's table
          const float sc = p.pro_act == 1 ? 0.5f : 1.0f;
          for (int i = xtid; i < 2 * p.C0; i += 32 * nxw) coef[i] = __ldg(p.pro_ab + (size_t)img * 2 * p.C0 + i) * sc;
          named_bar(5, 32 * nxw);
          coef_img = img;
        }
        mbar_wait(raw_full + 8 * ra.s, ra.ph);
        uint8_t* stage = a_s + (size_t)ra.s * p.a_stage;
        if (p.xf == 2) {
          // halo pixel (hy, hx) of the up-sampled image <- low-resolution pixel ((ty0-1+hy) >> 1, (tx0-1+hx) >> 1); the raw patch
          // starts at (ty0/2 - 1, tx0/2 - 1), so (hy+1) >> 1 and (hx+1) >> 1 index it (tile origins are even); pixels outside the
          // up-sampled image map to rows / columns the TMA unit zero-filled
          const uint8_t* raw = smem + p.off_raw + (size_t)ra.s * p.raw_stage;
          for (int c8 = xw; c8 < G::CH; c8 += 4) {
            const uint8_t* rc = raw + (size_t)c8 * (G::RP * G::RR * 16);
            uint8_t* col = stage + (size_t)c8 * G::LBO_TMA;
#pragma unroll 2
            for (int hp = lane; hp < G::HPIX; hp += 32) {
              const int hy = hp / G::PITCH, hx = hp - hy * G::PITCH;
              *reinterpret_cast<uint4*>(col + hp * 16) =
                  *reinterpret_cast<const uint4*>(rc + (((hy + 1) >> 1) * G::RP + ((hx + 1) >> 1)) * 16);
            }
          }
        } else if (LD_EXP & 32) {   // timing experiment: no transform work
        } else if (KS == 3 && !MX && p.swz && p.pro_act == 1) {
          if constexpr (KS == 3 && !MX) {
            if (p.xhelp) xf_stage<6, G, true>(stage, coef + c * KC, coef + p.C0 + c * KC, xw, lane, interior, ty0, tx0, p.H, p.W);
            else xf_stage<4, G, true>(stage, coef + c * KC, coef + p.C0 + c * KC, xw, lane, interior, ty0, tx0, p.H, p.W);
          }
        } else {   // other activations (the once-per-call conditional encoder), no-swizzle image, 1x1: channel group per warp, lanes on pixels
          const uint32_t sa = smem_u32(stage);
          for (int c8 = xw; c8 < G::CH; c8 += nxw) {
            const float4* ab = reinterpret_cast<const float4*>(coef + c * KC + c8 * 8);
            const float4 a0 = ab[0], a1 = ab[1];
            const float4* bb = reinterpret_cast<const float4*>(coef + p.C0 + c * KC + c8 * 8);
            const float4 b0 = bb[0], b1 = bb[1];
            const float pa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float pb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint8_t* col = stage + (size_t)c8 * G::LBO_TMA;
            for (int hp = lane; hp < G::HPIX; hp += 32) {
              bool ok = interior;
              if (!ok) {
                const int hy = hp / G::PITCH, hx = hp - hy * G::PITCH;
                ok = (unsigned)(ty0 + hy - 1) < (unsigned)p.H && (unsigned)(tx0 + hx - 1) < (unsigned)p.W;
              }
              if (ok) {
                uint4* q = reinterpret_cast<uint4*>(col + hp * 16);
                if (p.swz) q = reinterpret_cast<uint4*>(stage + hp * (G::CH * 16) + ((c8 ^ (int)(((sa + hp * (G::CH * 16)) >> 7) & (G::CH - 1))) << 4));
                *q = pro_apply_h(*q, pa, pb, p.pro_act);
              }
            }
          }
        }
        if (!(LD_EXP & 64)) fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + 8 * ra.s);
        ra.advance(SA);
      }
      if (nm == 2) twA = twB;
      twA.next(p);
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4 * R::kEpiGroups) {
    // ================================================================== epilogue ======================
    const bool two_groups = R::kEpiGroups == 2 && !xf;
    // warp-group eg drains the tiles whose accumulator stage is eg (every tile when there is one group)
    const int eg = (warp - kEpiWarp0) >> 2;
    const int ew = warp & 3, etid = threadIdx.x - (kEpiWarp0 + 4 * eg) * 32;
    const int ebar = 3 + eg;                           // named barrier of this warp-group
    float* sacc = sacc_all + 256 * eg;
    const int m = ew * 32 + lane;
    const int nbase = n_tile * NT;
    const int cpg = p.stats ? p.Cout / p.stats_G : 1;
    int it_tile = 0, n_mine = 0;
    TileWalk tw;
    const bool tile2d = KS == 3 || p.ds;
    tw.init(p, tile2d);
    // GroupNorm partial sums of the image being walked live in shared memory (sacc) and go to global memory, one double
    // atomic per (group, statistic), when the walk leaves the image
    int stat_img = -1;
    // RS instantiation (host: p.regstats): narrow tile, GroupNorm statistics with 4 / 8 / 16 channels per group, at most 8 groups
    constexpr bool kRegStats = RS && NT <= 64 && !MX;
    constexpr bool reg_stats = kRegStats;
    float racc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) racc[i] = 0.f;
    auto flush_stats = [&](int im) {
      if (kRegStats && reg_stats) stats_flush_regs(racc, sacc, 2 * (NT / cpg), lane);
      named_bar(ebar, 128);
      const int ng2 = 2 * (NT / cpg > 0 ? NT / cpg : 1);
      if (etid < ng2) {
        const float v = sacc[etid];
        sacc[etid] = 0.f;
        const int gi = nbase / cpg + (etid >> 1);
        atomicAdd(p.stats + ((size_t)im * p.stats_G + gi) * 2 + (etid & 1), (double)v);
      }
      named_bar(ebar, 128);
    };
    const int my_tiles = my_tile_count(p);
    for (; tw.tile < tw.end; tw.next(p), ++it_tile) {
      // accumulator of this tile: pair q = it_tile / MT, member mi; stage as = q % NACC, completion parity of its use
      const int q = MT == 2 ? it_tile >> 1 : it_tile, mi = MT == 2 ? (it_tile & 1) : 0;
      const int as = q & (NACC - 1);
      const uint32_t aph = (uint32_t)(q >> LNACC) & 1u;
      // two warp-groups: group eg drains the accumulator stages of its parity (mt == 1) or member eg of every pair (mt == 2)
      if (two_groups && (MT == 2 ? mi : (as & 1)) != eg) continue;
      // mt == 2 with one warp-group: the stage goes back to the MMA warp after the last member of the pair
      const bool arrive_here = MT == 1 || two_groups || mi == 1 || it_tile == my_tiles - 1;
      const int img = tw.img;
      if (p.stats && img != stat_img) {
        if (stat_img >= 0) flush_stats(stat_img);
        stat_img = img;
      }
      long long opix = -1;
      int ogy = 0, ogx = 0;
      if (MX) {
        const int gy = tw.ty * G::TH + (m >> 4), gx = tw.tx * G::TW + (m & 15);
        if ((m & 15) < G::TW && gy < p.H && gx < p.W) opix = ((long long)img * p.H + gy) * p.W + gx;
      } else if (tile2d) {
        const int gy = tw.ty * 16 + (m >> 3), gx = tw.tx * 8 + (m & 7);
        ogy = gy; ogx = gx;
        if (gy < p.H && gx < p.W) opix = ((long long)img * p.H + gy) * p.W + gx;
      } else {
        const long long gp = (long long)tw.tile * 128 + m;
        if (gp < p.M) opix = gp;
      }
      mbar_wait(acc_full + 8 * as, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)as * acc_stride + (uint32_t)mi * acc_cols;
      // TMA store path: the store that read this staging buffer two tiles ago must have finished reading it
      // (ncu source view: with ONE staging buffer per warp-group 31 % of the stall samples sat here, waiting for the previous tile's
      // store to drain; every group now alternates between two buffers)
      const int ob = two_groups ? (p.obuf == 4 ? 2 * eg + (n_mine & 1) : eg) : (as & 1);
      ++n_mine;
      if (NT <= 64 && p.tma_out) {
        if (etid == 0 && !(LD_EXP & 16)) { if (two_groups && p.obuf != 4) bulk_wait_group_read<0>(); else bulk_wait_group_read<1>(); }
        if (!(LD_EXP & 2)) named_bar(ebar, 128);
      }
      // staged output row: MX tiles are 8 x 14 pixels dense (patch columns 14, 15 of every row produce nothing)
      const int orow_i = MX ? (m >> 4) * G::TW + (m & 15) : m;
      uint8_t* orow = o_s + (size_t)ob * (128 * O_ROW) + (size_t)orow_i * O_ROW;
#pragma unroll (NT <= 64 ? 2 : 1)
      for (int j1 = 0; j1 < NT; j1 += 32) {
        // narrow tiles read the accumulator 16 columns at a time: the GroupNorm sums kept in registers across tiles (stats_acc) plus 32
        // accumulator columns do not fit the 80-register budget of two co-resident CTAs (spills landed in the MMA warp's loop)
        constexpr bool kSplitLd = LD_CONV_SPLITLD && NT <= 64 && !MX;
        uint32_t r32[kSplitLd ? 16 : 32];
        if (!MX && !kSplitLd) {
          if (!(LD_EXP & 8)) { tmem_ld32(trow + j1, reinterpret_cast<uint32_t(&)[32]>(r32)); tmem_ld_wait(); }
          else { for (int j = 0; j < 32; ++j) r32[j] = (uint32_t)(m + j); }
          if (j1 + 32 == NT && !p.dual && arrive_here) {   // every TMEM read of this thread is complete: hand the stage back
            tc_fence_before();
            mbar_arrive(acc_empty + 8 * as);
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j0 = j1 + 16 * h;
          float f[16];
          if (kSplitLd) {
            tmem_ld16(trow + j0, reinterpret_cast<uint32_t(&)[16]>(r32)); tmem_ld_wait();
            if (h == 1 && j1 + 32 == NT && !p.dual && arrive_here) { tc_fence_before(); mbar_arrive(acc_empty + 8 * as); }
          }
          constexpr int rb = kSplitLd ? 0 : 16;   // offset of half 1 inside r32
          if (MX) {
            // out[lane] = D[lane][kx=0] + D[lane+1][kx=1] + D[lane+2][kx=2]: neighbours along the patch row are the next lanes
            uint32_t a0[16], a1[16], a2[16];
            tmem_ld16(trow + j0, a0);
            tmem_ld16(trow + NT + j0, a1);
            tmem_ld16(trow + 2 * NT + j0, a2);
            tmem_ld_wait();
            if (h == 1 && !p.dual) { tc_fence_before(); mbar_arrive(acc_empty + 8 * as); }
#pragma unroll
            for (int j = 0; j < 16; ++j)
              f[j] = __uint_as_float(a0[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(a1[j]), 1) +
                     __shfl_down_sync(0xffffffffu, __uint_as_float(a2[j]), 2) + bias_s[j0 + j];
          } else {
#pragma unroll
            if (NT <= 64 && p.bias_const) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(r32[rb * h + j]) + p.bias_c[j0 + j];
            } else
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j0 + j);
              f[j] = __uint_as_float(r32[rb * h + j]) + b4.x; f[j + 1] = __uint_as_float(r32[rb * h + j + 1]) + b4.y;
              f[j + 2] = __uint_as_float(r32[rb * h + j + 2]) + b4.z; f[j + 3] = __uint_as_float(r32[rb * h + j + 3]) + b4.w;
            }
          }
          if (p.stats) {
            const int grp0 = (nbase + j0) / cpg - nbase / cpg;
            const bool valid = opix >= 0;
            if (kRegStats && reg_stats) {
              if (valid && !(LD_EXP & 128)) {
                if (cpg == 4) stats_acc<4>(f, racc, j0);
                else if (cpg == 8) stats_acc<8>(f, racc, j0);
                else stats_acc<16>(f, racc, j0);
              }
            } else
            switch (cpg) {
              case 2: stats_chunk<2>(f, valid, sacc, grp0, lane); break;
              case 4: stats_chunk<4>(f, valid, sacc, grp0, lane); break;
              case 8: stats_chunk<8>(f, valid, sacc, grp0, lane); break;
              default: stats_chunk<16>(f, valid, sacc, grp0, lane); break;   // cpg >= 16: chunk inside one group
            }
          }
          if (p.res && opix >= 0) {
            const size_t o = (size_t)opix * p.Cout + nbase + j0;
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
          if (NT <= 64 && p.tma_out) {
            // 16-byte chunk index XOR row bits = the TMA 64B / 128B swizzle pattern: conflict-free row-per-thread stores
            const int c0 = j0 >> 3;
            const int sw = NT == 32 ? ((orow_i >> 1) & 3) : (orow_i & 7);
            if ((!MX || (m & 15) < G::TW) && !((LD_EXP & 4) && pk[0] != 0x12345678u)) {
              *reinterpret_cast<uint4*>(orow + ((c0 ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(orow + (((c0 + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          } else if (opix >= 0 && !(p.dbg & 4)) {
            size_t o = (size_t)opix * p.Cout + nbase + j0;
            if (p.ps) {   // virtual channel (py, px, c) of low-resolution pixel (gy, gx) -> channel c of output pixel (2 gy + py, 2 gx + px)
              const int cidx = nbase + j0, q = cidx / p.ps, c = cidx - q * p.ps;
              o = ((((size_t)img * (2 * p.H) + 2 * ogy + (q >> 1)) * (2 * p.W)) + 2 * ogx + (q & 1)) * p.ps + c;
            }
            // one 32-byte store per 16 channels: a lane fills a whole sector (two 16-byte stores left every sector half written per
            // instruction -- 1024 sector operations per 64-channel tile, the epilogue-only knock-out ran at 2245 clk per tile)
            // (the register-statistics instantiation keeps 16-byte stores: it takes this path only under LD_CONV_DIRECT=2, and its
            //  register allocation -- 80-register budget -- measured 3 % slower on the TMA store path with the wide store in its code)
            if constexpr (!RS) st_global_v8(p.dst + o, pk);
            else {
              *reinterpret_cast<uint4*>(p.dst + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(p.dst + o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
      if (!RS && p.dual) {   // (the register-statistics instantiation never runs dual launches)
        // second accumulator: the 1x1 res_conv of the same input tile (ddpm.py:198,212), direct 32-byte stores
#pragma unroll (NT <= 64 ? 2 : 1)
        for (int j1 = 0; j1 < NT; j1 += 32) {
          uint32_t r32[32];
          tmem_ld32(trow + DUALC + j1, r32);
          tmem_ld_wait();
          if (j1 + 32 == NT) { tc_fence_before(); mbar_arrive(acc_empty + 8 * as); }
          if (opix >= 0 && !(p.dbg & 4)) {
            __nv_bfloat16* o2 = p.dst2 + (size_t)opix * p.Cout + nbase + j1;
#pragma unroll
            for (int c = 0; c < 4; c += 2) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int cj = j1 + 8 * c + 2 * j;
                const float b0 = (NT <= 64 && p.bias_const) ? p.bias_c[64 + (NT <= 64 ? cj : 0)] : bias_s[NT + cj];
                const float b1 = (NT <= 64 && p.bias_const) ? p.bias_c[64 + (NT <= 64 ? cj + 1 : 0)] : bias_s[NT + cj + 1];
                pk[j] = pack_bf16x2(__uint_as_float(r32[8 * c + 2 * j]) + b0, __uint_as_float(r32[8 * c + 2 * j + 1]) + b1);
              }
              st_global_v8(o2 + 8 * c, pk);
            }
          }
        }
      }
      if (NT <= 64 && p.tma_out) {
        if (!(LD_EXP & 1)) fence_proxy_async();               // my generic-proxy writes are visible to the TMA unit
        if (!(LD_EXP & 2)) named_bar(ebar, 128);
        if (etid == 0) {
          if (!(p.dbg & 4)) {
            const uint32_t src = smem_u32(o_s + (size_t)ob * (128 * O_ROW));
            if (MX) tma_store_4d(&p.map_out, nbase, tw.tx * G::TW, tw.ty * G::TH, img, src);
            else if (tile2d) tma_store_4d(&p.map_out, nbase, tw.tx * 8, tw.ty * 16, img, src);
            else tma_store_2d(&p.map_out, nbase, tw.tile * 128, src);
            bulk_commit_group();
          }
        }
      }
    }
    if (MT == 2 && two_groups && eg == 1 && (my_tiles & 1)) {
      // the last pair has one member (drained by group 0): this group still owes its 128 arrivals on the stage
      const int q = my_tiles >> 1, as = q & (NACC - 1);
      mbar_wait(acc_full + 8 * as, (uint32_t)(q >> LNACC) & 1u);
      mbar_arrive(acc_empty + 8 * as);
    }
    if (p.stats && stat_img >= 0) flush_stats(stat_img);
    if (NT <= 64 && p.tma_out && etid == 0) bulk_wait_group<0>();
  } else if (warp == kMmaWarp) {
    // ================================================================== MMA issue =====================
    // The whole warp runs the loop (barrier waits); one elected lane issues tcgen05.mma / commit.  Descriptors
    // are (lo, hi) register pairs and every tap / k-step only adds a small constant to `lo`.
    {
      constexpr uint32_t idesc = make_idesc(128, NMMA), idesc_d = make_idesc(128, NT);
      // swizzled stage: pixel rows of KC * 2 bytes; 8-pixel core groups PITCH rows apart (3x3) or dense (1x1); a tap view starts
      // (ky * PITCH + kx) rows into the patch, a k-step advances 32 bytes inside the swizzled row
      constexpr uint32_t ROW16 = KC * 2 / 16;
      const uint32_t a_hi = p.swz ? (desc_hi((KS == 3 ? G::PITCH : 8) * KC * 2) | (KC == 32 ? (4u << 29) : (2u << 29))) : desc_hi(G::SBO);
      const uint32_t b_hi = desc_hi(128);
      const uint32_t lbo16 = p.swz ? 1u : (uint32_t)p.lbo16;     // distance of the two 8-channel halves of a k-step is 2 * lbo16 (16-byte units)
      const uint32_t tap16 = p.swz ? ROW16 : 1u;                 // one patch pixel in 16-byte units
      const uint32_t a_lo0 = desc_lo(smem_u32(a_s), p.swz ? 16u : (lbo16 << 4)), b_lo0 = desc_lo(smem_u32(b_s), NMMA * 16),
                     bd_lo0 = desc_lo(smem_u32(b_s), NT * 16);   // fused-1x1 stage: [kc/8][NT][8] at the head of its slot
      const uint32_t a_stage16 = (uint32_t)p.a_stage >> 4;
      Ring ra, rb;
      int it_tile = 0;
      if (p.resident) { mbar_wait(b_full, 0); tc_fence_after(); }
      const int my_tiles = my_tile_count(p);
      for (int q = 0; it_tile < my_tiles; it_tile += MT, ++q) {
        // one iteration = one tile, or (mt == 2) a pair of consecutive tiles that share every streamed weight stage
        const int nm = (MT == 2 && it_tile + 1 < my_tiles) ? 2 : 1;
        const int as = q & (NACC - 1);
        mbar_wait(acc_empty + 8 * as, ((uint32_t)(q >> LNACC) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)as * acc_stride;
        uint32_t acc = 0;
        for (int c = 0; c < p.nchunks; ++c) {
          Ring ra1 = ra;                       // stage of the pair's second tile
          if (nm == 2) ra1.advance(SA);
          mbar_wait(a_full + 8 * ra.s, ra.ph);
          if (nm == 2) mbar_wait(a_full + 8 * ra1.s, ra1.ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)ra.s * a_stage16;
          const uint32_t a_lo1 = a_lo0 + (uint32_t)ra1.s * a_stage16;
          if (p.resident) {
            const uint32_t b_lo = b_lo0 + (uint32_t)(c * TAPSW * (B_STAGE >> 4));
            if (elect_one()) {
              if (!(p.dbg & 2)) {
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                  const int ky = MX ? tap : tap / KS, kx = MX ? 0 : tap - ky * KS;
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k) {
                    umma_bf16_lh(dcol, a_lo + (uint32_t)(ky * G::PITCH + kx) * tap16 + (uint32_t)(2 * k) * lbo16, a_hi,
                                 b_lo + (uint32_t)(tap * (B_STAGE >> 4) + 2 * k * NMMA), b_hi, idesc, acc);
                    acc = 1;
                  }
                }
                if (p.dual) {   // fused 1x1: centre-tap view of the same patch, last weight stage, second accumulator
                  const uint32_t bd_lo = bd_lo0 + (uint32_t)((c * TAPSW + TAPS) * (B_STAGE >> 4));
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k)
                    umma_bf16_lh(dcol + DUALC, a_lo + (uint32_t)(G::PITCH + 1) * tap16 + (uint32_t)(2 * k) * lbo16, a_hi,
                                 bd_lo + (uint32_t)(2 * k * NT), b_hi, idesc_d, (c > 0 || k > 0) ? 1u : 0u);
                }
              }
              umma_commit(a_empty + 8 * ra.s);
              if (c == p.nchunks - 1) umma_commit(acc_full + 8 * as);
            }
            acc = 1;
            __syncwarp();
          } else {
#pragma unroll 1
            for (int tap = 0; tap < TAPSW; ++tap) {
              mbar_wait(b_full + 8 * rb.s, rb.ph);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + (uint32_t)(rb.s * (B_STAGE >> 4));
              const bool extra = tap == TAPS;   // fused 1x1 on the centre-tap view, second accumulator
              int tapr = tap;
              if (KS == 3 && !MX && p.rot && !extra) { tapr = tap + (int)(blockIdx.x % TAPS); if (tapr >= TAPS) tapr -= TAPS; }
              const int ky = extra ? KS / 2 : (MX ? tapr : tapr / KS), kx = extra ? KS / 2 : (MX ? 0 : tapr - ky * KS);
              const uint32_t a_t = a_lo + (uint32_t)(ky * G::PITCH + kx) * tap16;
              if (elect_one()) {
                if (extra) {
                  const uint32_t bd_lo = bd_lo0 + (uint32_t)(rb.s * (B_STAGE >> 4));
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k)
                    umma_bf16_lh(dcol + DUALC, a_t + (uint32_t)(2 * k) * lbo16, a_hi, bd_lo + (uint32_t)(2 * k * NT), b_hi, idesc_d,
                                 (c > 0 || k > 0) ? 1u : 0u);
                } else {
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k)
                    umma_bf16_lh(dcol, a_t + (uint32_t)(2 * k) * lbo16, a_hi, b_lo + (uint32_t)(2 * k * NMMA), b_hi, idesc, (acc | (uint32_t)k) ? 1u : 0u);
                  if (nm == 2) {   // same weight stage, second tile of the pair, second accumulator
                    const uint32_t a_t1 = a_lo1 + (uint32_t)(ky * G::PITCH + kx) * tap16;
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k)
                      umma_bf16_lh(dcol + acc_cols, a_t1 + (uint32_t)(2 * k) * lbo16, a_hi, b_lo + (uint32_t)(2 * k * NMMA), b_hi, idesc,
                                   (acc | (uint32_t)k) ? 1u : 0u);
                  }
                  acc = 1;
                }
                umma_commit(b_empty + 8 * rb.s);
                if (tap == TAPSW - 1) {
                  umma_commit(a_empty + 8 * ra.s);
                  if (nm == 2) umma_commit(a_empty + 8 * ra1.s);
                  if (c == p.nchunks - 1) umma_commit(acc_full + 8 * as);
                }
              }
              if (!extra) acc = 1;
              __syncwarp();
              rb.advance(SB);
            }
          }
          ra.advance(SA);
          if (nm == 2) ra.advance(SA);
        }
      }
    }
  } else if (warp == kWWarp) {
    // ================================================================== weight pipeline ===============
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)n_tile * p.nchunks * TAPSW * B_STAGE;
      const int total = p.nchunks * TAPSW;
      if (p.resident) {
        mbar_arrive_expect_tx(b_full, (uint32_t)total * B_STAGE);
        resident_weights(smem_u32(b_s), wsrc, (uint32_t)total * B_STAGE, b_full);
      } else {
        Ring rb;
        const int my_tiles = my_tile_count(p);
        for (int it = 0; it < my_tiles; it += MT) {
          for (int c = 0, i = 0; c < p.nchunks; ++c)
            for (int tap = 0; tap < TAPSW; ++tap, ++i) {
              int src = i;
              if (KS == 3 && !MX && p.rot && tap < TAPS) { int tr = tap + (int)(blockIdx.x % TAPS); if (tr >= TAPS) tr -= TAPS; src = c * TAPSW + tr; }
              mbar_wait(b_empty + 8 * rb.s, rb.ph ^ 1);
              mbar_arrive_expect_tx(b_full + 8 * rb.s, B_STAGE);
              bulk_g2s(smem_u32(b_s + rb.s * B_STAGE), wsrc + (size_t)src * B_STAGE, B_STAGE, b_full + 8 * rb.s);
              rb.advance(SB);
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

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
constexpr size_t kResidentBudget = 200 * 1024;

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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (getenv("LD_CONV_NO_TMA")) return nullptr;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

// NHWC bf16 activation [N][H][W][C] as the view (8 ch, W, H, C/8, N): a box of (8, PITCH, ROWS, KC/8, 1) is the
// UMMA K-major operand image [KC/8 chunks][ROWS*PITCH pixels][16 B] of a halo patch, zero filled outside the image.
bool map_in_3x3(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int kc, int pitch, int rows) {
  EncodeTiledFn f = encode_fn();
  if (!f) return false;
  const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)N};
  const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)pitch, (cuuint32_t)rows, (cuuint32_t)(kc / 8), 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// Swizzled variants (KParams::swz): the same boxes with the channels of a pixel as ONE inner row of kc * 2 = 64 / 128 bytes, written
// to shared memory through the 64B / 128B hardware swizzle.  [N][H][W][C] as (C, W, H, N), box (kc, PITCH, ROWS, 1).
bool map_in_3x3_sw(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int kc, int pitch, int rows, int estride = 1) {
  EncodeTiledFn f = encode_fn();
  if (!f || (kc != 32 && kc != 64)) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)pitch, (cuuint32_t)rows, 1};
  const cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  return f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// flattened pixels [M][C] as (C, M): box (kc, 128)
bool map_in_1x1_sw(CUtensorMap* m, const void* ptr, long long M, int C, int kc) {
  EncodeTiledFn f = encode_fn();
  if (!f || (kc != 32 && kc != 64)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kc, 128};
  const cuuint32_t es[2] = {1, 1};
  return f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// pixel-unshuffle gather: every other pixel of [N][H][W][C] in both directions -> 16 x 8 pixels per box
bool map_in_ds(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int kc) {
  EncodeTiledFn f = encode_fn();
  if (!f) return false;
  const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)N};
  const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[5] = {8, 16, 32, (cuuint32_t)(kc / 8), 1};   // traversal stride 2: ceil(16/2) x ceil(32/2) pixels are loaded
  const cuuint32_t es[5] = {1, 2, 2, 1, 1};
  return f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// flattened pixels [M][C] as (8 ch, M, C/8): box (8, 128, KC/8)
bool map_in_1x1(CUtensorMap* m, const void* ptr, long long M, int C, int kc) {
  EncodeTiledFn f = encode_fn();
  if (!f) return false;
  const cuuint64_t dims[3] = {8, (cuuint64_t)M, (cuuint64_t)(C / 8)};
  const cuuint64_t strides[2] = {(cuuint64_t)C * 2, 16};
  const cuuint32_t box[3] = {8, 128, (cuuint32_t)(kc / 8)};
  const cuuint32_t es[3] = {1, 1, 1};
  return f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// output [N][H][W][Cout]: box (NT, 8, 16, 1) (3x3) or [M][Cout]: box (NT, 128) (1x1), 64B / 128B swizzle for NT = 32 / 64
bool map_out(CUtensorMap* m, void* ptr, int N, int H, int W, int Cout, int nt, int ks, bool mx = false) {
  EncodeTiledFn f = encode_fn();
  if (!f || (nt != 32 && nt != 64)) return false;
  const CUtensorMapSwizzle sw = nt == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  const cuuint32_t es[4] = {1, 1, 1, 1};
  if (ks == 3) {
    const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2};
    const cuuint32_t box[4] = {(cuuint32_t)nt, mx ? 14u : 8u, mx ? 8u : 16u, 1};   // MX tiles: 8 rows x 14 pixels
    return f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)N * H * W};
  const cuuint64_t strides[1] = {(cuuint64_t)Cout * 2};
  const cuuint32_t box[2] = {(cuuint32_t)nt, 128};
  return f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// opt in to the maximum dynamic shared memory once per instantiation (done at pack time, outside any graph capture)
template <int NT, int KS, int KC>
int configure_one() {
  if (cudaFuncSetAttribute(conv_tc_kernel<NT, KS, KC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg().max_smem) != cudaSuccess) return -1;
  if constexpr (NT >= 128) {
    if (cudaFuncSetAttribute(conv_tc_kernel<NT, KS, KC, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg().max_smem) != cudaSuccess) return -1;
  } else {
    if (cudaFuncSetAttribute(conv_tc_kernel<NT, KS, KC, true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg().max_smem) != cudaSuccess) return -1;
  }
  return cudaFuncSetAttribute(conv_tc_kernel<NT, KS, KC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg().max_smem) == cudaSuccess ? 0 : -1;
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

// shared memory carve-up; returns total bytes
template <int NT, int KS, int KC>
size_t layout(KParams& p, int sa, int nb_stages) {
  constexpr bool MX = kUseMX && KS == 3 && NT == 32;
  using G = Geo<KS, KC, MX>;
  size_t off = 0;
  // output staging (1024-byte aligned, first): two buffers per epilogue warp-group
  if (p.tma_out) off += (size_t)(p.obuf == 4 ? 4 : 2) * 128 * NT * 2;
  p.off_a = (int)off;
  p.a_stage = p.tma_in ? G::CH * G::LBO_TMA : G::CH * G::LBO_REG;
  p.a_stage = (p.a_stage + 127) & ~127;
  p.lbo16 = (p.tma_in ? G::LBO_TMA : G::LBO_REG) >> 4;
  off += (size_t)sa * p.a_stage;
  off = (off + 127) & ~(size_t)127;
  p.off_b = (int)off;
  off += (size_t)nb_stages * (MX ? 3 * NT : NT) * KC * 2;
  p.off_coef = (int)off;
  off += (size_t)(p.coef_floats + 512 + 2 * NT) * 4 + (3 * SA_MAX + 2 * SB + 8) * 8 + 16;
  off = (off + 127) & ~(size_t)127;
  p.off_raw = (int)off;
  p.raw_stage = p.xf == 2 ? ((G::CH * G::RP * G::RR * 16 + 127) & ~127) : 0;
  off += (size_t)sa * p.raw_stage;
  return off;
}

template <int NT, int KS, int KC>
int launch_one(KParams& p, int ntiles_y, cudaStream_t s) {
  constexpr bool MX = kUseMX && KS == 3 && NT == 32;
  const int total = p.nchunks * ((MX ? 3 : KS * KS) + (p.dual ? 1 : 0));
  const size_t limit = (size_t)cfg().max_smem < kResidentBudget ? (size_t)cfg().max_smem : kResidentBudget;
  // weights resident in shared memory when everything fits; more activation stages when fed by TMA
  int sa = p.tma_in ? 4 : 3;
  p.obuf = 2;
  p.resident = layout<NT, KS, KC>(p, sa, total) <= limit ? 1 : 0;
  p.nb_stages = p.resident ? total : SB;
  {
    static int rot_env = -1;   // env LD_CONV_ROT=1: rotated tap order per CTA.  MEASURED: no effect (38.9 vs 39.1 us, 256 -> 256 at 32 x 32 x 32): the
                               // L2 is not hot-spotted by 148 CTAs reading the same stage; off by default (canonical accumulation order)
    if (rot_env < 0) { const char* e = getenv("LD_CONV_ROT"); rot_env = e ? atoi(e) : 0; }
    p.rot = (!p.resident && rot_env) ? 1 : 0;
  }
  {
    static int xh_env = -1;   // env LD_CONV_XHELP=0: four transform warps only (A/B aid)
    if (xh_env < 0) { const char* e = getenv("LD_CONV_XHELP"); xh_env = e ? atoi(e) : 1; }
    p.xhelp = (p.tma_in && p.xf == 1 && p.swz && p.resident && KS == 3 && !MX && xh_env) ? 1 : 0;
  }
  // Streamed weights: a stage (NT x KC) feeds KC/16 MMAs of NT/2 clocks each, i.e. the ring must take in 64 B/clk per SM and its
  // four stages cover ~1 us of MMA work -- less than the L2 latency under load (tensor pipe 58 % busy, ncu).  With mt = 2 every
  // stage serves two consecutive tiles of the CTA (two accumulators): half the weight traffic per FLOP, twice the cover.
  static int mt_env = -1;   // env LD_CONV_MT=1 disables (A/B aid)
  if (mt_env < 0) { const char* e = getenv("LD_CONV_MT"); mt_env = e ? atoi(e) : 2; }
  // (not with the in-place normalise mode: its single epilogue warp-group would drain both accumulators back to back -- measured slower)
  // (NT = 256: with the cheap swizzled activation loads a single tile per weight pass is faster again -- 35.3 vs 38.8 us for 256 -> 256 at
  //  32 x 32 x 32: the epilogue of one tile overlaps the MMAs of the next, and the bulk copies of the filter (90 B/clk per SM in 32 KB pieces,
  //  tests/micro/tma_rate.cu) hide behind the MMAs either way; NT = 128 keeps the pairs: 38.8 vs 45.4 us; env LD_CONV_MT=3 forces pairs)
  p.mt = (p.tma_in && !p.resident && (NT == 128 || (NT >= 128 && mt_env == 3)) && KS == 3 && !p.dual && p.xf != 1 && mt_env >= 2) ? 2 : 1;
  size_t smem = layout<NT, KS, KC>(p, sa, p.nb_stages);
  if (smem > (size_t)cfg().max_smem) { sa = 3; p.mt = 1; smem = layout<NT, KS, KC>(p, sa, p.nb_stages); }   // pairs need four activation stages
  p.nacc = (p.mt == 2 && NT == 256) ? 1 : 2;
  {
    static int rs_env = -1;   // 0 = per-tile butterfly everywhere, 1 = registers except dual launches (default: the dual launch measured 4 % slower with the
                              // 16 extra live registers), 3 = registers everywhere
    if (rs_env < 0) { const char* e = getenv("LD_CONV_REGSTATS"); rs_env = e ? atoi(e) : 1; }
    const int cpg = p.stats ? p.Cout / p.stats_G : 0;
    const bool fits = p.stats && p.tma_in && NT <= 64 && !MX && p.mt == 1 && (cpg == 4 || cpg == 8 || cpg == 16) && NT / cpg <= 8;
    p.regstats = (fits && (p.dual ? (rs_env & 2) : (rs_env & 1))) ? 1 : 0;
  }
  {
    // narrow tiles (NT <= 64) leave TMEM to spare: four accumulator stages let the MMA warp run two tiles ahead of each epilogue
    // warp-group.  MEASURED: no effect on any variant of the 32-channel convolution (81 / 88 / 136 / 178 us with 2 or 4 stages,
    // profiles/r3_conv_pipeline_ab.md) -- the launch is bound by shared-memory bandwidth, not by pipeline depth.  env LD_CONV_NACC=4 enables.
    static int nacc_env = -1;
    if (nacc_env < 0) { const char* e = getenv("LD_CONV_NACC"); nacc_env = e ? atoi(e) : 2; }
    if (nacc_env == 4 && !MX && p.mt == 1 && NT <= 64 && (!p.dual || NT == 32) && p.tma_in && LD_CONV_EG == 2) p.nacc = 4;
  }
  const int tm_cols_raw = MX ? 8 * NT : (p.mt == 2 ? p.nacc * 2 * NT : (p.dual ? 2 : 1) * p.nacc * NT);
  int tm_cols = 32; while (tm_cols < tm_cols_raw) tm_cols <<= 1;   // TMEM columns per CTA (allocations are powers of two >= 32)
  if (smem > (size_t)cfg().max_smem) return -1;
  if (p.tma_in && sa == 4) {   // a third co-resident CTA is worth more than a fourth activation stage
    const size_t smem3 = layout<NT, KS, KC>(p, 3, p.nb_stages);
    const size_t want = (size_t)Roles<true>::kMinCtas;   // co-resident CTAs the launch bounds allow
    if ((size_t)cfg().max_smem_sm / (smem3 + 1024) >= want && (size_t)cfg().max_smem_sm / (smem + 1024) < want && (size_t)(512 / tm_cols) >= want) { sa = 3; smem = smem3; }
    else layout<NT, KS, KC>(p, sa, p.nb_stages);
  }
  p.sa = sa;
  if (p.tma_out && p.tma_in && !p.xf && LD_CONV_EG == 2) {
    // a second staging buffer per epilogue warp-group (the store of the previous tile drains while the next tile is staged) --
    // unless the extra shared memory costs a co-resident CTA
    const int occ2 = (int)((size_t)cfg().max_smem_sm / (smem + 1024));
    p.obuf = 4;
    const size_t smem4 = layout<NT, KS, KC>(p, sa, p.nb_stages);
    const int occ4 = (int)((size_t)cfg().max_smem_sm / (smem4 + 1024));
    const int cap = Roles<true>::kMinCtas < 512 / tm_cols ? Roles<true>::kMinCtas : 512 / tm_cols;
    if (smem4 <= (size_t)cfg().max_smem && (occ4 >= occ2 || occ4 >= cap)) smem = smem4;
    else {
      // next best: give up the fourth activation stage for it (the staging wait was 23 % of the stall samples of the dual launch)
      const size_t smem43 = sa == 4 && p.mt == 1 ? layout<NT, KS, KC>(p, 3, p.nb_stages) : 0;
      const int occ43 = smem43 ? (int)((size_t)cfg().max_smem_sm / (smem43 + 1024)) : 0;
      if (smem43 && (occ43 >= occ2 || occ43 >= cap)) { sa = 3; p.sa = 3; smem = smem43; }
      else { p.obuf = 2; layout<NT, KS, KC>(p, sa, p.nb_stages); }
    }
  }
  {
    // deeper activation ring while it costs no co-resident CTA (env LD_CONV_SA = 5 or 6).  MEASURED: 4, 5 and 6 stages give the same
    // times on every variant (profiles/r3_conv_pipeline_ab.md), so the default stays at 4.
    static int sa_cap = -1;
    if (sa_cap < 0) { const char* e = getenv("LD_CONV_SA"); sa_cap = e ? atoi(e) : 4; if (sa_cap > SA_MAX) sa_cap = SA_MAX; }
    const int cap = Roles<true>::kMinCtas < 512 / tm_cols ? Roles<true>::kMinCtas : 512 / tm_cols;
    auto occ_of = [&](size_t b) { int o = (int)((size_t)cfg().max_smem_sm / (b + 1024)); return o > cap ? cap : o; };
    if (p.tma_in && p.mt == 1) {
      const int occ_now = occ_of(smem);
      while (sa < sa_cap) {
        const size_t s2 = layout<NT, KS, KC>(p, sa + 1, p.nb_stages);
        if (s2 <= (size_t)cfg().max_smem && occ_of(s2) >= occ_now) { ++sa; smem = s2; }
        else { layout<NT, KS, KC>(p, sa, p.nb_stages); break; }
      }
      p.sa = sa;
    }
  }
  // persistent grid: as many CTAs as are co-resident (registers allow two per SM; 2*NT of 512 TMEM columns each)
  const bool lean = p.tma_in != 0;
  int occ = (int)((size_t)cfg().max_smem_sm / (smem + 1024));
  const int occ_max = lean ? Roles<true>::kMinCtas : Roles<false>::kMinCtas;
  if (occ > occ_max) occ = occ_max;
  const int tm = 512 / tm_cols;
  if (occ > tm) occ = tm;
  if (occ < 1) occ = 1;
  int gx = cfg().sms * occ / ntiles_y;
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  {
    static int walk = -1;   // env LD_CONV_WALK: 0 strided, 1 contiguous when statistics are fused (default), 2 always contiguous
    if (walk < 0) { const char* e = getenv("LD_CONV_WALK"); walk = e ? atoi(e) : 1; }
    p.contig = (walk == 2 || (walk == 1 && p.stats)) ? 1 : 0;
  }
  if (p.contig) { p.step_img = 0; p.step_ty = 0; p.step_tx = 1; }
  else {
    const int tpi = p.tiles_x * p.tiles_y;
    p.step_img = gx / tpi;
    const int r = gx - p.step_img * tpi;
    p.step_ty = r / p.tiles_x; p.step_tx = r - p.step_ty * p.tiles_x;
  }
  if constexpr (NT >= 128) {
    if (lean && p.mt == 2) { launch_k(conv_tc_kernel<NT, KS, KC, true, 2>, dim3((unsigned)gx, (unsigned)ntiles_y), Roles<true>::kThreads, smem, s, true, p); return 1; }
  }
  if constexpr (NT <= 64) {
    if (lean && p.regstats) { launch_k(conv_tc_kernel<NT, KS, KC, true, 1, true>, dim3((unsigned)gx, (unsigned)ntiles_y), Roles<true>::kThreads, smem, s, true, p); return 1; }
  }
  if (lean) launch_k(conv_tc_kernel<NT, KS, KC, true>, dim3((unsigned)gx, (unsigned)ntiles_y), Roles<true>::kThreads, smem, s, true, p);
  else launch_k(conv_tc_kernel<NT, KS, KC, false>, dim3((unsigned)gx, (unsigned)ntiles_y), Roles<false>::kThreads, smem, s, true, p);
  return 1;
}

template <int KS, int KC>
int launch_nt(int NT, KParams& p, int ntiles_y, cudaStream_t s) {
  switch (NT) {
    case 32: return launch_one<32, KS, KC>(p, ntiles_y, s);
    case 64: return launch_one<64, KS, KC>(p, ntiles_y, s);
    case 128: return launch_one<128, KS, KC>(p, ntiles_y, s);
    case 256: return launch_one<256, KS, KC>(p, ntiles_y, s);
  }
  return -1;
}

}  // namespace

namespace {
// GroupNorm (+FiLM) folded to y = a x + b per (image, channel) (ddpm.py:174-185): ab[n][0][c] = a, ab[n][1][c] = b
__global__ void gn_coef_kernel(const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ film, int film_stride, int G, int C, double cnt, float eps, float* __restrict__ ab) {
  const int n = blockIdx.x;
  if (threadIdx.x == 0) pdl_trigger();   // N small blocks, all resident: let the consuming convolution set itself up (ld_launch.cuh)
  pdl_wait();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / (C / G);
    const double su = stats[((size_t)n * G + g) * 2], sq = stats[((size_t)n * G + g) * 2 + 1];
    const double mean = su / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float a = rstd * gamma[c], b = beta[c] - (float)mean * a;
    if (film) {
      const float sc = film[(size_t)n * film_stride + c] + 1.0f, sf = film[(size_t)n * film_stride + C + c];
      a *= sc; b = b * sc + sf;
    }
    ab[((size_t)n * 2) * C + c] = a;
    ab[((size_t)n * 2 + 1) * C + c] = b;
  }
}
}  // namespace

int gn_coef_launch(const double* stats, const float* gamma, const float* beta, const float* film, int film_stride, int G, int C, int N,
                   long long HW, float eps, float* ab, cudaStream_t s) {
  launch_k(gn_coef_kernel, dim3(N), dim3(C < 256 ? C : 256), 0, s, true, stats, gamma, beta, film, film_stride, G, C, (double)HW * (C / G), eps, ab);
  pdl_after_small() = 1;
  return 1;
}

static int pick_ntile(int Cout) {
  if (Cout <= 256) return (Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256) ? Cout : 0;
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  return 0;
}

int conv_tc_pack(const float* w, const float* bias, int Cin, int Cout, int ks, int stride, int pad, ConvTcW* out, const float* w1,
                 const float* bias1) {
  out->ready = false;
  out->dual = false; out->bias2 = nullptr;
  if (w1 && ks != 3) return 0;
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
    // MX (3x3, 32 couts): a stage is one vertical tap with N = (kx, cout) = 96 rows; else one (ky, kx) tap with nt rows.
    // dual: the last stage of every chunk holds the 1x1 weights as [kc/8][nt][8] at the head of a full-size slot.
    const bool mx = kUseMX && ks == 3 && nt == 32;
    const int mtaps = mx ? 3 : taps, nrow = mx ? 3 * nt : nt;
    const int tapsw = mtaps + (w1 ? 1 : 0);
    const size_t slot_elems = (size_t)nrow * kc;
    std::vector<__nv_bfloat16> pk((size_t)ntiles * nch * tapsw * slot_elems, __float2bfloat16_rn(0.f));
    for (int nti = 0; nti < ntiles; ++nti)
      for (int c = 0; c < nch; ++c)
        for (int t = 0; t < tapsw; ++t) {
          const size_t base = (((size_t)nti * nch + c) * tapsw + t) * slot_elems;
          const bool extra = t == mtaps;
          const int rows = extra ? nt : nrow;
          for (int k8 = 0; k8 < kc / 8; ++k8)
            for (int r = 0; r < rows; ++r)
              for (int e = 0; e < 8; ++e) {
                const int cin = c * kc + k8 * 8 + e, co = nti * nt + (r % nt);
                float v;
                if (extra) v = w1[(size_t)cin * Cout + co];
                else {
                  const int tap = mx ? t * 3 + r / nt : t;   // MX: stage = ky, row block = kx
                  v = w[((size_t)tap * Cin + cin) * Cout + co];
                }
                pk[base + ((size_t)k8 * rows + r) * 8 + e] = __float2bfloat16_rn(v);
              }
        }
    if (cudaMalloc(slot, pk.size() * 2) != cudaSuccess) return -1;
    if (cudaMemcpy(*slot, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  }
  for (int i = 0; i < 64; ++i) { out->bias_h[i] = (bias && i < Cout) ? bias[i] : 0.f; out->bias2_h[i] = (bias1 && i < Cout) ? bias1[i] : 0.f; }
  out->bias = nullptr;
  if (bias) {
    if (cudaMalloc(&out->bias, Cout * sizeof(float)) != cudaSuccess) return -1;
    if (cudaMemcpy(out->bias, bias, Cout * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  }
  if (w1) {
    out->dual = true;
    if (bias1) {
      if (cudaMalloc(&out->bias2, Cout * sizeof(float)) != cudaSuccess) return -1;
      if (cudaMemcpy(out->bias2, bias1, Cout * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    }
  }
  if (ks == 3) { if (configure_nt<3, 64>(nt) || configure_nt<3, 32>(nt)) return -1; }
  else { if (configure_nt<1, 64>(nt) || configure_nt<1, 32>(nt)) return -1; }
  out->ready = true;
  return 0;
}

int conv_tc_pack_up2(const float* w, const float* bias, int Cin, int Cout, ConvTcW* out) {
  out->ready = false;
  if (Cout % 16 || Cin % 32 || !pick_ntile(4 * Cout) || pick_ntile(4 * Cout) < 128) return 0;
  // which low-resolution offset d in {-1, 0, +1} (index d + 1) a filter tap k in {0, 1, 2} lands on, per output parity
  static const int off[2][3] = {{0, 1, 1}, {1, 1, 2}};
  const int Co4 = 4 * Cout;
  std::vector<float> wv((size_t)9 * Cin * Co4, 0.f), bv;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const int tap = off[py][ky] * 3 + off[px][kx], q = py * 2 + px;
          for (int c = 0; c < Cin; ++c)
            for (int o = 0; o < Cout; ++o)
              wv[((size_t)tap * Cin + c) * Co4 + q * Cout + o] += w[((size_t)(ky * 3 + kx) * Cin + c) * Cout + o];
        }
  if (bias) { bv.resize(Co4); for (int i = 0; i < Co4; ++i) bv[i] = bias[i % Cout]; }
  return conv_tc_pack(wv.data(), bias ? bv.data() : nullptr, Cin, Co4, 3, 1, 1, out);
}

bool conv_tc_supports(const ConvTcW& w, const ConvTcArgs& a) {
  if (!w.ready) return false;
  if (a.ds) return w.ks == 1 && w.Cin == 4 * a.C0 && a.C0 % 32 == 0 && !a.src1 && !a.up && !a.pro_ab && !a.stats && !a.res &&
                  a.Hin == 2 * a.H && a.Win == 2 * a.W && encode_fn() != nullptr;
  if (a.C0 + a.C1 != w.Cin || a.C0 % 32 || a.C1 % 32) return false;
  if (a.ps) return w.ks == 3 && !w.dual && w.Cout == 4 * a.ps && a.ps % 16 == 0 && w.ntile >= 128 && !a.src1 && !a.up && !a.res && !a.stats && !a.pro_ab &&
                   !a.dst2 && a.Hin == a.H && a.Win == a.W;
  if (w.dual != (a.dst2 != nullptr)) return false;
  if (w.dual && (w.ks != 3 || w.ntile > 64 || w.Cout != w.ntile || a.up || a.res)) return false;   // 4 NT TMEM columns, two CTAs per SM
  if (a.up && (w.ks != 3 || a.src1)) return false;
  if (a.pro_ab && (w.ks != 3 || a.src1 || a.up)) return false;
  if (a.stats) {
    if (w.ks != 3 || a.stats_G < 1 || w.Cout % a.stats_G) return false;
    const int cpg = w.Cout / a.stats_G;
    if (!(cpg == 2 || cpg == 4 || cpg == 8 || (cpg >= 16 && cpg % 16 == 0))) return false;
    if (w.ntile % cpg || 2 * (w.ntile / cpg) > 128) return false;   // a group never straddles two n-tiles
  }
  return true;
}

long long* conv_tc_trace() { return nullptr; }

// Kernel parameters (incl. the three encoded tensor maps) are cached per distinct (weights, arguments): plans replay
// the same launches every timestep and cuTensorMapEncodeTiled is a host-side driver call.
static bool build_params(const ConvTcW& w, const ConvTcArgs& a, KParams& p) {
  bool k64 = w.w && a.C0 % 64 == 0 && a.C1 % 64 == 0;
  // folded up-sampling convolution with a filter of <= 150 KB: 32-channel chunks halve the activation stages, which lets the whole
  // filter stay resident in shared memory instead of being streamed for every tile
  if (a.ps && w.w32 && (size_t)9 * w.Cin * w.Cout * 2 <= 150 * 1024) k64 = false;
  const int kc = k64 ? 64 : 32;
  const bool mx = kUseMX && w.ks == 3 && w.ntile == 32;   // merged-x geometry (see Geo): 8 x 14 output tiles
  p = KParams{};
  p.ds = a.ds; p.ds_cs = a.C0; p.ps = a.ps;
  p.src0 = (const __nv_bfloat16*)a.src0; p.src1 = (const __nv_bfloat16*)a.src1; p.C0 = a.C0; p.C1 = a.C1;
  p.N = a.N; p.H = a.H; p.W = a.W; p.Hin = a.Hin; p.Win = a.Win; p.up = a.up;
  p.nchunks = w.Cin / kc;
  p.w = (const __nv_bfloat16*)(k64 ? w.w : w.w32); p.bias = w.bias; p.Cout = w.Cout;
  p.dst = (__nv_bfloat16*)a.dst; p.res = (const __nv_bfloat16*)a.res;
  p.dst2 = (__nv_bfloat16*)a.dst2; p.bias2 = w.bias2; p.dual = w.dual ? 1 : 0;
  p.bias_const = (w.Cout == w.ntile && w.ntile <= 64) ? 1 : 0;
  if (p.bias_const) { memcpy(p.bias_c, w.bias_h, 64 * sizeof(float)); memcpy(p.bias_c + 64, w.bias2_h, 64 * sizeof(float)); }
  p.M = (long long)a.N * a.H * a.W;
  p.pro_ab = a.pro_ab; p.pro_act = a.pro_act;
  p.coef_floats = 0;   // set below for the in-place normalise mode (xf == 1): the y = a x + b table of one image
  p.stats = a.stats; p.stats_G = a.stats_G;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("LD_CONV_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  // TMA activation loads whenever the source is read as stored (no up-sampling, no normalise-on-load)
  p.tma_in = 0;
  p.swz = 0;
  static int noswz = -1;   // env LD_CONV_NO_SWZ=1: the no-swizzle K-major stage image (16-byte TMA rows) everywhere (A/B aid)
  if (noswz < 0) { const char* e = getenv("LD_CONV_NO_SWZ"); noswz = e ? atoi(e) : 0; }
  const bool want_swz = !noswz && !mx;
  if (a.ds) {
    if (want_swz && map_in_3x3_sw(&p.map_a0, a.src0, a.N, a.Hin, a.Win, a.C0, kc, 16, 32, 2)) { p.tma_in = 1; p.swz = 1; }
    else p.tma_in = map_in_ds(&p.map_a0, a.src0, a.N, a.Hin, a.Win, a.C0, kc) ? 1 : 0;
  } else if (a.up) {
    // nearest x2 source: TMA fetches the low-resolution pixels under the halo patch, transform warps replicate them
    static int noup = -1;   // env LD_CONV_NO_XF=1: register-staging kernel instead (A/B aid)
    if (noup < 0) { const char* e = getenv("LD_CONV_NO_XF"); noup = e ? atoi(e) : 0; }
    const int rp = (mx ? 14 : 8) / 2 + 2, rr = (mx ? 8 : 16) / 2 + 2;
    if (!noup && a.H == 2 * a.Hin && a.W == 2 * a.Win && map_in_3x3(&p.map_a0, a.src0, a.N, a.Hin, a.Win, a.C0, kc, rp, rr)) { p.tma_in = 1; p.xf = 2; }
  } else {
    bool ok = false;
    if (want_swz && (w.ks == 3 || !a.pro_ab)) {   // (the in-place normalise rewrite knows the swizzled rows for 3x3 only)
      ok = w.ks == 3 ? map_in_3x3_sw(&p.map_a0, a.src0, a.N, a.Hin, a.Win, a.C0, kc, 10, 18) : map_in_1x1_sw(&p.map_a0, a.src0, p.M, a.C0, kc);
      if (ok && a.src1)
        ok = w.ks == 3 ? map_in_3x3_sw(&p.map_a1, a.src1, a.N, a.Hin, a.Win, a.C1, kc, 10, 18) : map_in_1x1_sw(&p.map_a1, a.src1, p.M, a.C1, kc);
      p.swz = ok ? 1 : 0;
    }
    if (!ok) {
      ok = w.ks == 3 ? map_in_3x3(&p.map_a0, a.src0, a.N, a.Hin, a.Win, a.C0, kc, mx ? 16 : 10, mx ? 10 : 18) : map_in_1x1(&p.map_a0, a.src0, p.M, a.C0, kc);
      if (ok && a.src1)
        ok = w.ks == 3 ? map_in_3x3(&p.map_a1, a.src1, a.N, a.Hin, a.Win, a.C1, kc, mx ? 16 : 10, mx ? 10 : 18) : map_in_1x1(&p.map_a1, a.src1, p.M, a.C1, kc);
    }
    p.tma_in = ok ? 1 : 0;
    static int noxf = -1;   // env LD_CONV_NO_XF=1: normalise-on-load through the register-staging kernel (A/B aid)
    if (noxf < 0) { const char* e = getenv("LD_CONV_NO_XF"); noxf = e ? atoi(e) : 0; }
    if (a.pro_ab && noxf) { p.tma_in = 0; p.swz = 0; }
    p.xf = (a.pro_ab && p.tma_in) ? 1 : 0;
    if (p.xf == 1) p.coef_floats = 2 * a.C0;
  }
  static int no_tma_out = -1;   // env LD_CONV_NO_TMA_OUT=1: direct 16-byte stores from registers instead of smem staging + TMA store (A/B aid)
  if (no_tma_out < 0) { const char* e = getenv("LD_CONV_NO_TMA_OUT"); no_tma_out = e ? atoi(e) : 0; }
  // Output path of narrow tiles: staging in shared memory + one TMA store per tile, or 16-byte stores straight from the accumulator
  // registers.  Shared-memory bandwidth is what bounds these launches (operand reads of the MMAs + every other access), so the direct
  // path wins where the epilogue has registers to spare: launches without statistics and dual launches (65 vs 71 us, 132 vs 146 us at
  // 32x256x256); with the register statistics it loses (74 vs 69 us, 121 vs 111 us).  env LD_CONV_DIRECT: 0 never, 1 (default) that rule, 2 always.
  static int direct = -1;
  if (direct < 0) { const char* e = getenv("LD_CONV_DIRECT"); direct = e ? atoi(e) : 1; }
  const bool go_direct = direct == 2 || (direct == 1 && (!a.stats || w.dual));
  p.tma_out = (!a.ps && !no_tma_out && !go_direct && map_out(&p.map_out, a.dst, a.N, a.H, a.W, w.Cout, w.ntile, a.ds ? 3 : w.ks, mx)) ? 1 : 0;
  if (mx) {
    p.tiles_x = (a.W + 13) / 14; p.tiles_y = (a.H + 7) / 8;
    p.ntiles = a.N * p.tiles_x * p.tiles_y;
  } else if (w.ks == 3 || a.ds) {
    p.tiles_x = (a.W + 7) / 8; p.tiles_y = (a.H + 15) / 16;
    p.ntiles = a.N * p.tiles_x * p.tiles_y;
  } else {
    p.tiles_x = 1; p.tiles_y = 1;
    p.ntiles = (int)((p.M + 127) / 128);
  }
  return k64;
}

// The cache is shared by every handle of the process (plans of different handles may launch concurrently from different
// threads): guarded by a mutex, keyed on an explicit field-by-field serialisation (no struct padding in the key).
namespace {
struct ParamKey {
  uint64_t v[20];
  bool operator==(const ParamKey& o) const { return memcmp(v, o.v, sizeof v) == 0; }
};
struct ParamKeyHash {
  size_t operator()(const ParamKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
ParamKey make_key(const ConvTcW& w, const ConvTcArgs& a) {
  ParamKey k{};
  int i = 0;
  k.v[i++] = (uint64_t)(uintptr_t)a.src0; k.v[i++] = (uint64_t)(uintptr_t)a.src1; k.v[i++] = (uint64_t)(uintptr_t)a.dst;
  k.v[i++] = (uint64_t)(uintptr_t)a.res; k.v[i++] = (uint64_t)(uintptr_t)a.dst2; k.v[i++] = (uint64_t)(uintptr_t)a.pro_ab;
  k.v[i++] = (uint64_t)(uintptr_t)a.stats;
  k.v[i++] = ((uint64_t)(uint32_t)a.C0 << 32) | (uint32_t)a.C1;
  k.v[i++] = ((uint64_t)(uint32_t)a.N << 32) | (uint32_t)a.H;
  k.v[i++] = ((uint64_t)(uint32_t)a.W << 32) | (uint32_t)a.Hin;
  k.v[i++] = ((uint64_t)(uint32_t)a.Win << 32) | (uint32_t)((a.up & 1) | ((a.ds & 1) << 1) | ((a.pro_act & 3) << 2));
  k.v[i++] = ((uint64_t)(uint32_t)a.ps << 32) | (uint32_t)a.stats_G;
  k.v[i++] = (uint64_t)(uintptr_t)w.w; k.v[i++] = (uint64_t)(uintptr_t)w.w32; k.v[i++] = (uint64_t)(uintptr_t)w.bias;
  k.v[i++] = (uint64_t)(uintptr_t)w.bias2;
  k.v[i++] = ((uint64_t)(uint32_t)w.Cin << 32) | (uint32_t)w.Cout;
  k.v[i++] = ((uint64_t)(uint32_t)w.ks << 32) | (uint32_t)w.ntile;
  int dev = 0; cudaGetDevice(&dev);
  k.v[i++] = (uint64_t)dev;
  return k;
}
}  // namespace

int conv_tc_launch(const ConvTcW& w, const ConvTcArgs& a, cudaStream_t s) {
  if (!conv_tc_supports(w, a)) return -1;
  struct Cached { KParams p; bool k64; };
  static std::unordered_map<ParamKey, Cached, ParamKeyHash> cache;
  static std::mutex mu;
  KParams p;
  bool k64;
  {
    const ParamKey key = make_key(w, a);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
      if (cache.size() > 4096) cache.clear();
      Cached c;
      c.k64 = build_params(w, a, c.p);
      it = cache.emplace(key, c).first;
    }
    p = it->second.p;
    k64 = it->second.k64;
  }
  // the key holds pointers only: the bias VALUES carried in the parameters always come from the weights of this call
  if (p.bias_const) { memcpy(p.bias_c, w.bias_h, 64 * sizeof(float)); memcpy(p.bias_c + 64, w.bias2_h, 64 * sizeof(float)); }
  if (a.ds && !p.tma_in) return -1;
  const int ny = w.Cout / w.ntile;
  if (w.ks == 3) return k64 ? launch_nt<3, 64>(w.ntile, p, ny, s) : launch_nt<3, 32>(w.ntile, p, ny, s);
  return k64 ? launch_nt<1, 64>(w.ntile, p, ny, s) : launch_nt<1, 32>(w.ntile, p, ny, s);
}

}  // namespace ld
