// tcgen05 / TMEM implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels, NT couts] (fp32, TMEM)  +=  A[128 pixels, 16 cin] (bf16, smem) * B[NT couts, 16 cin] (bf16, smem)
//
// One CTA computes one tile of 128 output pixels x NT output channels:
//   * 3x3 convs: the tile is a 16 x 8 pixel patch; its 18 x 10 halo patch is staged in shared memory ONCE per
//     64- (or 32-) channel chunk and all nine filter taps are issued as tcgen05.mma on *shifted views* of that
//     patch: the operand uses the canonical no-swizzle K-major layout ([8-channel chunk][pixel][16 B]), in which
//     moving by one pixel is a 16-byte move of the descriptor start address and the 8-row core-matrix groups of
//     the tile sit at a constant stride (SBO = halo pitch * 16 B).  Activations are therefore read once from
//     L2/HBM instead of nine times.
//   * 1x1 convs: the tile is 128 consecutive pixels of the flattened [N*H*W] axis, no halo.
//   * the staging warps read global memory directly (predicated 16-byte loads), which is what makes zero padding,
//     the virtual channel concat of two sources (torch.cat in ddpm.py:435-448) and the nearest x2 up-sampling
//     (ddpm.py:116) free, and leaves room for a normalise-on-load prologue.
//   * weights are pre-packed on the host into the exact shared-memory image of every (chunk, tap) stage and
//     streamed with the TMA bulk-copy engine (cp.async.bulk -> UBLKCP) through a 4-deep mbarrier ring.
//   * warp roles: warps 0-3 stage A and later run the epilogue (TMEM -> registers -> bias/residual -> bf16 ->
//     global), warp 4 allocates TMEM and issues the MMAs from one lane, warp 5 drives the weight ring.
//
// Reference call sites this kernel serves: nn.Conv2d at ddpm.py:117,173,198,227,230,268,269,372,391 and
// unet_model.py:20,24,30 (every conv with Cin >= 32 and Cout >= 32).
#include <cuda_bf16.h>
#include <stdint.h>

#include <vector>

#include "ld_conv_tc.h"

namespace ld {

namespace {

constexpr int kProducerThreads = 128;
constexpr int kThreads = 192;
constexpr int SA = 2;  // activation stages
constexpr int SB = 4;  // weight stages

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE
}

struct KParams {
  const __nv_bfloat16* src0; const __nv_bfloat16* src1;
  int C0, C1;
  int N, H, W, Hin, Win, up;
  int tiles_x, tiles_y;
  int nchunks;
  const __nv_bfloat16* w; const float* bias;
  int Cout;
  __nv_bfloat16* dst; const __nv_bfloat16* res;
  long long M;
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
  static constexpr int ITEMS = (HPIX * CH + kProducerThreads - 1) / kProducerThreads;
  static constexpr int SBO = (KS == 3 ? PITCH : 8) * 16;               // stride between 8-pixel core-matrix groups
};

template <int NT, int KS, int KC>
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(KParams p) {
  using G = Geo<KS, KC>;
  constexpr int TAPS = KS * KS;
  constexpr int B_STAGE = NT * KC * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + SA * G::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + SB * B_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 1);
  const uint32_t a_full = smem_u32(bars), a_empty = a_full + 8 * SA, b_full = a_empty + 8 * SA, b_empty = b_full + 8 * SB,
                 acc_full = b_empty + 8 * SB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(a_full + 8 * i, 4); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile coordinates
  int img = 0, ty0 = 0, tx0 = 0;
  long long pix0 = 0;
  if (KS == 3) {
    int t = blockIdx.x;
    tx0 = (t % p.tiles_x) * G::TW; t /= p.tiles_x;
    ty0 = (t % p.tiles_y) * G::TH; img = t / p.tiles_y;
  } else {
    pix0 = (long long)blockIdx.x * 128;
  }
  const int n_tile = blockIdx.y;

  if (warp < 4) {
    // ------------------------------------------------------------------ A staging -------------------
    const int tid = threadIdx.x;
    long long goff[G::ITEMS];   // element offset of the source pixel (per source: multiplied by C later), -1 = zero fill
    int soff[G::ITEMS];         // byte offset inside the stage
    int coff[G::ITEMS];         // channel offset inside the chunk (elements)
#pragma unroll
    for (int it = 0; it < G::ITEMS; ++it) {
      const int i = it * kProducerThreads + tid;
      const int hp = i / G::CH, ch = i - hp * G::CH;
      goff[it] = -1; soff[it] = ch * G::LBO + hp * 16; coff[it] = ch * 8;
      if (hp < G::HPIX) {
        if (KS == 3) {
          const int hy = hp / G::PITCH, hx = hp - hy * G::PITCH;
          int gy = ty0 + hy - 1, gx = tx0 + hx - 1;
          if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
            if (p.up) { gy >>= 1; gx >>= 1; }
            goff[it] = ((long long)img * p.Hin + gy) * p.Win + gx;
          }
        } else {
          const long long g = pix0 + hp;
          if (g < p.M) goff[it] = g;
        }
      } else {
        soff[it] = -1;
      }
    }
    for (int c = 0; c < p.nchunks; ++c) {
      const int sa = c % SA;
      mbar_wait(a_empty + 8 * sa, ((c / SA) & 1) ^ 1);
      const int cbase = c * KC;
      const __nv_bfloat16* src; int cs, cb;
      if (cbase < p.C0) { src = p.src0; cs = p.C0; cb = cbase; } else { src = p.src1; cs = p.C1; cb = cbase - p.C0; }
      uint4 v[G::ITEMS];
#pragma unroll
      for (int it = 0; it < G::ITEMS; ++it) {
        v[it] = make_uint4(0u, 0u, 0u, 0u);
        if (goff[it] >= 0) v[it] = __ldg(reinterpret_cast<const uint4*>(src + goff[it] * cs + cb + coff[it]));
      }
      uint8_t* stage = a_s + sa * G::A_STAGE;
#pragma unroll
      for (int it = 0; it < G::ITEMS; ++it)
        if (soff[it] >= 0) *reinterpret_cast<uint4*>(stage + soff[it]) = v[it];
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8 * sa);
    }
    // ------------------------------------------------------------------ epilogue --------------------
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    long long opix = -1;
    if (KS == 3) {
      const int gy = ty0 + (m >> 3), gx = tx0 + (m & 7);
      if (gy < p.H && gx < p.W) opix = ((long long)img * p.H + gy) * p.W + gx;
    } else {
      if (pix0 + m < p.M) opix = pix0 + m;
    }
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int nbase = n_tile * NT;
#pragma unroll 1
    for (int j0 = 0; j0 < NT; j0 += 16) {
      uint32_t r[16];
      tmem_ld16(trow + j0, r);
      tmem_ld_wait();
      if (opix >= 0) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + nbase + j0 + j) : 0.f);
        const size_t o = (size_t)opix * p.Cout + nbase + j0;
        if (p.res) {
          const uint4 r0 = *reinterpret_cast<const uint4*>(p.res + o), r1 = *reinterpret_cast<const uint4*>(p.res + o + 8);
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[j]));
            f[2 * j] += t2.x; f[2 * j + 1] += t2.y;
          }
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(p.dst + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(p.dst + o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issue -------------------
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = NT, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int sb = 0; uint32_t pb = 0;
      uint32_t acc = 0;
      for (int c = 0; c < p.nchunks; ++c) {
        const int sa = c % SA;
        mbar_wait(a_full + 8 * sa, (c / SA) & 1);
        tc_fence_after();
        const uint32_t abase = smem_u32(a_s + sa * G::A_STAGE);
#pragma unroll 1
        for (int tap = 0; tap < TAPS; ++tap) {
          mbar_wait(b_full + 8 * sb, pb);
          tc_fence_after();
          const uint32_t bbase = smem_u32(b_s + sb * B_STAGE);
          const int ky = tap / KS, kx = tap - ky * KS;
          const uint32_t ashift = (uint32_t)(ky * G::PITCH + kx) * 16u;
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            const uint64_t ad = make_desc(abase + ashift + (uint32_t)(2 * k) * G::LBO, G::LBO, G::SBO);
            const uint64_t bd = make_desc(bbase + (uint32_t)(2 * k) * (NT * 16), NT * 16, 128);
            umma_bf16(tmem_base, ad, bd, idesc, acc);
            acc = 1;
          }
          umma_commit(b_empty + 8 * sb);
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
        umma_commit(a_empty + 8 * sa);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ weight ring ------------------
    if (lane == 0) {
      int sb = 0; uint32_t pb = 0;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)n_tile * p.nchunks * TAPS * B_STAGE;
      const int total = p.nchunks * TAPS;
      for (int i = 0; i < total; ++i) {
        mbar_wait(b_empty + 8 * sb, pb ^ 1);
        mbar_arrive_expect_tx(b_full + 8 * sb, B_STAGE);
        bulk_g2s(smem_u32(b_s + sb * B_STAGE), wsrc + (size_t)i * B_STAGE, B_STAGE, b_full + 8 * sb);
        if (++sb == SB) { sb = 0; pb ^= 1; }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, NT);
}

template <int KS, int KC>
size_t smem_bytes(int NT) {
  return (size_t)SA * Geo<KS, KC>::A_STAGE + (size_t)SB * NT * KC * 2 + (2 * SA + 2 * SB + 1) * 8 + 16;
}

// opt in to > 48 KB dynamic shared memory once per instantiation (done at pack time, outside any graph capture)
template <int NT, int KS, int KC>
int configure_one() {
  const size_t sm = smem_bytes<KS, KC>(NT);
  return cudaFuncSetAttribute(conv_tc_kernel<NT, KS, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) == cudaSuccess ? 0 : -1;
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
int launch_one(const KParams& p, dim3 grid, cudaStream_t s) {
  conv_tc_kernel<NT, KS, KC><<<grid, kThreads, smem_bytes<KS, KC>(NT), s>>>(p);
  return 1;
}

template <int KS, int KC>
int launch_nt(int NT, const KParams& p, dim3 grid, cudaStream_t s) {
  switch (NT) {
    case 32: return launch_one<32, KS, KC>(p, grid, s);
    case 64: return launch_one<64, KS, KC>(p, grid, s);
    case 128: return launch_one<128, KS, KC>(p, grid, s);
    case 256: return launch_one<256, KS, KC>(p, grid, s);
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
  dim3 grid;
  if (w.ks == 3) {
    p.tiles_x = (a.W + 7) / 8; p.tiles_y = (a.H + 15) / 16;
    grid = dim3((unsigned)(a.N * p.tiles_x * p.tiles_y), (unsigned)(w.Cout / w.ntile));
    return k64 ? launch_nt<3, 64>(w.ntile, p, grid, s) : launch_nt<3, 32>(w.ntile, p, grid, s);
  }
  grid = dim3((unsigned)((p.M + 127) / 128), (unsigned)(w.Cout / w.ntile));
  return k64 ? launch_nt<1, 64>(w.ntile, p, grid, s) : launch_nt<1, 32>(w.ntile, p, grid, s);
}

}  // namespace ld
