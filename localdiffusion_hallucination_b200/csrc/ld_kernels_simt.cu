// CUDA-core (fp32 math) kernels of the LocalDiffusion denoiser: the fp32 parity path uses all of
// them; the bf16 path uses everything here except the generic convolution, which is replaced by
// the tcgen05 implicit GEMM in ld_conv_tc.cu.  All tensors are NHWC, storage type T.
#include "ld_kernels.h"
#include "ld_launch.cuh"

namespace ld {

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// =================================================================================================
// generic implicit-GEMM convolution, CUDA cores.  Tile 64 pixels x (16*TN) output channels,
// K step 16 input channels of one filter tap; 256 threads, 4 x TN outputs each.
// Reference ops: nn.Conv2d call sites ddpm.py:117,123,173,198,227,230,268,269,372,391 and
// unet_model.py:20,24,30.
// =================================================================================================
template <typename T, int TN>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvP p) {
  constexpr int BM = 64, BK = 16, BN = 16 * TN;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lm = tid >> 2, lc = (tid & 3) * 4;
  const long long gm = m0 + lm;
  const bool mvalid = gm < p.M;
  int ln = 0, ly = 0, lx = 0;
  if (mvalid) {
    long long r = gm;
    lx = (int)(r % p.W); r /= p.W;
    ly = (int)(r % p.H); ln = (int)(r / p.H);
  }
  const int Hv = p.up ? p.Hin * 2 : p.Hin, Wv = p.up ? p.Win * 2 : p.Win;
  const int Ctot = p.C0 + p.C1;
  constexpr int BROW_T = BN / 4;          // threads per B row
  const int bk = tid / BROW_T, bc = (tid % BROW_T) * 4;
  const bool bload = bk < BK;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const T* s0 = reinterpret_cast<const T*>(p.src0);
  const T* s1 = reinterpret_cast<const T*>(p.src1);
  const int taps = p.ks * p.ks;
  for (int tap = 0; tap < taps; ++tap) {
    const int ky = tap / p.ks, kx = tap - ky * p.ks;
    int iy = ly * p.stride + ky - p.pad, ix = lx * p.stride + kx - p.pad;
    const bool v = mvalid && iy >= 0 && iy < Hv && ix >= 0 && ix < Wv;
    if (p.up) { iy >>= 1; ix >>= 1; }
    const size_t pix = ((size_t)ln * p.Hin + iy) * p.Win + ix;
    const float* wt = p.w + (size_t)tap * Ctot * p.Cout;
    for (int c0 = 0; c0 < Ctot; c0 += BK) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      if (v) {
        const int c = c0 + lc;
        if (c < p.C0) load4(s0 + pix * p.C0 + c, a);
        else load4(s1 + pix * p.C1 + (c - p.C0), a);
      }
      float b[4] = {0.f, 0.f, 0.f, 0.f};
      if (bload && n0 + bc < p.Cout) load4(wt + (size_t)(c0 + bk) * p.Cout + n0 + bc, b);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) As[lc + i][lm] = a[i];
      if (bload) *reinterpret_cast<float4*>(&Bs[bk][bc]) = make_float4(b[0], b[1], b[2], b[3]);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float bv[TN];
        if constexpr (TN == 4) {
          float4 t4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          bv[0] = t4.x; bv[1] = t4.y; bv[2] = t4.z; bv[3] = t4.w;
        } else {
          float2 t2 = *reinterpret_cast<const float2*>(&Bs[k][tx * 2]);
          bv[0] = t2.x; bv[1] = t2.y;
        }
        const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(aa[i], bv[j], acc[i][j]);
      }
    }
  }
  const int co = n0 + tx * TN;
  if (co >= p.Cout) return;
  T* dst = reinterpret_cast<T*>(p.dst);
  const T* res = reinterpret_cast<const T*>(p.res);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float v = acc[i][j] + (p.bias ? p.bias[co + j] : 0.f);
      if (res) v += to_f(res[(size_t)m * p.Cout + co + j]);
      from_f(dst[(size_t)m * p.Cout + co + j], v);
    }
  }
}

int launch_conv_simt(const ConvP& p, bool bf, cudaStream_t s) {
  const bool narrow = p.Cout <= 32;
  dim3 grid(cdiv(p.M, 64), cdiv(p.Cout, narrow ? 32 : 64));
  if (bf) {
    if (narrow) conv_simt_kernel<bf16, 2><<<grid, 256, 0, s>>>(p);
    else conv_simt_kernel<bf16, 4><<<grid, 256, 0, s>>>(p);
  } else {
    if (narrow) conv_simt_kernel<float, 2><<<grid, 256, 0, s>>>(p);
    else conv_simt_kernel<float, 4><<<grid, 256, 0, s>>>(p);
  }
  return 1;
}

// =================================================================================================
// Cin == 1 direct convolution (init_conv 7x7 ddpm.py:319; cond encoder first convs unet_model.py:20,30)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) conv_c1_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                      const float* __restrict__ bias, T* __restrict__ out, int N, int H,
                                                      int W, int Cout, int ks) {
  extern __shared__ float sw[];  // [ks*ks][Cout] + [Cout]
  const int taps = ks * ks;
  for (int i = threadIdx.x; i < taps * Cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[taps * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int cg = Cout >> 2;
  const long long total = (long long)N * H * W * cg;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c = (int)(gid % cg) * 4;
  long long pix = gid / cg;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  const int pad = ks >> 1;
  float acc[4] = {sw[taps * Cout + c], sw[taps * Cout + c + 1], sw[taps * Cout + c + 2], sw[taps * Cout + c + 3]};
  const float* img = in + (size_t)n * H * W;
  for (int ky = 0; ky < ks; ++ky) {
    const int iy = y + ky - pad;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < ks; ++kx) {
      const int ix = x + kx - pad;
      if (ix < 0 || ix >= W) continue;
      const float v = __ldg(img + (size_t)iy * W + ix);
      const float* wp = sw + (ky * ks + kx) * Cout + c;
      acc[0] = fmaf(v, wp[0], acc[0]); acc[1] = fmaf(v, wp[1], acc[1]);
      acc[2] = fmaf(v, wp[2], acc[2]); acc[3] = fmaf(v, wp[3], acc[3]);
    }
  }
  store4(out + (size_t)pix * Cout + c, acc);
}

int launch_conv_c1(const float* in, const float* w, const float* bias, void* out, int N, int H, int W, int Cout, int ks,
                   bool bf, cudaStream_t s) {
  const long long total = (long long)N * H * W * (Cout / 4);
  const size_t sm = (size_t)(ks * ks + 1) * Cout * sizeof(float);
  if (bf) conv_c1_kernel<bf16><<<cdiv(total, 256), 256, sm, s>>>(in, w, bias, (bf16*)out, N, H, W, Cout, ks);
  else conv_c1_kernel<float><<<cdiv(total, 256), 256, sm, s>>>(in, w, bias, (float*)out, N, H, W, Cout, ks);
  return 1;
}

// =================================================================================================
// final 1x1 convolution to one fp32 channel (ddpm.py:398)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) conv_cout1_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         long long P, int C) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= P) return;
  float acc = bias ? bias[0] : 0.f;
  const T* px = in + (size_t)pix * C;
  for (int c = 0; c < C; c += 4) {
    float v[4];
    load4(px + c, v);
    acc = fmaf(v[0], __ldg(w + c), acc); acc = fmaf(v[1], __ldg(w + c + 1), acc);
    acc = fmaf(v[2], __ldg(w + c + 2), acc); acc = fmaf(v[3], __ldg(w + c + 3), acc);
  }
  out[pix] = acc;
}

int launch_conv_cout1(const void* in, const float* w, const float* bias, float* out, long long P, int C, bool bf,
                      cudaStream_t s) {
  if (bf) conv_cout1_kernel<bf16><<<cdiv(P, 256), 256, 0, s>>>((const bf16*)in, w, bias, out, P, C);
  else conv_cout1_kernel<float><<<cdiv(P, 256), 256, 0, s>>>((const float*)in, w, bias, out, P, C);
  return 1;
}

// =================================================================================================
// GroupNorm (ddpm.py:174, unet_model.py:21,25,31): statistics pass + fused apply pass
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, double* __restrict__ sums, int HW, int C,
                                                       int G, int pix_per_block) {
  extern __shared__ float sh[];  // [C] sum, [C] sumsq
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int cv = C >> 2;                     // vec4 per pixel
  const int lanes = blockDim.x / cv;         // pixels processed concurrently (blockDim % cv == 0)
  const int myc = (threadIdx.x % cv) * 4;
  const int myp = threadIdx.x / cv;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (myp < lanes) {
    const T* base = x + (size_t)n * HW * C + myc;
    for (int p = p0 + myp; p < p1; p += lanes) {
      float v[4];
      load4(base + (size_t)p * C, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) { s[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { atomicAdd(&sh[myc + i], s[i]); atomicAdd(&sh[C + myc + i], q[i]); }
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double a = 0, b = 0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += (double)sh[c]; b += (double)sh[C + c]; }
    atomicAdd(&sums[((size_t)n * G + g) * 2], a);
    atomicAdd(&sums[((size_t)n * G + g) * 2 + 1], b);
  }
}

int launch_gn_stats(const void* x, double* sums, int N, int HW, int C, int G, bool bf, cudaStream_t s) {
  const int cv = C / 4;
  int threads = 256;
  if (threads % cv) threads = (256 / cv) * cv;  // C is a multiple of 32 on this path -> cv | 256 or cv = 24, 48, 96
  if (threads == 0) threads = cv;
  int ppb = 4096 * 32 / C;                     // ~128K elements per block
  if (ppb < 64) ppb = 64;
  dim3 grid(cdiv(HW, ppb), N);
  const size_t sm = 2 * (size_t)C * sizeof(float);
  if (bf) gn_stats_kernel<bf16><<<grid, threads, sm, s>>>((const bf16*)x, sums, HW, C, G, ppb);
  else gn_stats_kernel<float><<<grid, threads, sm, s>>>((const float*)x, sums, HW, C, G, ppb);
  return 1;
}

template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(GnApplyP p, int pix_per_block) {
  extern __shared__ float sh[];  // aA[C], bA[C], (aB[C], bB[C])
  const int n = blockIdx.y, C = p.C;
  float* aA = sh; float* bA = sh + C; float* aB = sh + 2 * C; float* bB = sh + 3 * C;
  const double cntA = (double)p.HW * (C / p.GA);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / (C / p.GA);
    const double su = p.statsA[((size_t)n * p.GA + g) * 2], sq = p.statsA[((size_t)n * p.GA + g) * 2 + 1];
    const double mean = su / cntA;
    double var = sq / cntA - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    float a = rstd * p.gA[c], b = p.bA[c] - (float)mean * a;
    if (p.film) {
      const float sc = p.film[(size_t)n * p.film_stride + c] + 1.0f, sf = p.film[(size_t)n * p.film_stride + C + c];
      a *= sc; b = b * sc + sf;
    }
    aA[c] = a; bA[c] = b;
    if (p.modeB == 2) {
      const int g2 = c / (C / p.GB);
      const double cntB = (double)p.HW * (C / p.GB);
      const double su2 = p.statsB[((size_t)n * p.GB + g2) * 2], sq2 = p.statsB[((size_t)n * p.GB + g2) * 2 + 1];
      const double mean2 = su2 / cntB;
      double var2 = sq2 / cntB - mean2 * mean2;
      if (var2 < 0) var2 = 0;
      const float r2 = (float)(1.0 / sqrt(var2 + (double)p.eps));
      aB[c] = r2 * p.gB[c]; bB[c] = p.bB[c] - (float)mean2 * aB[c];
    }
  }
  __syncthreads();
  const int cv = C >> 2;
  const long long base = (long long)blockIdx.x * pix_per_block * cv;
  const long long end = min((long long)p.HW * cv, base + (long long)pix_per_block * cv);
  const T* xa = reinterpret_cast<const T*>(p.xa) + (size_t)n * p.HW * C;
  const T* xb = p.xb ? reinterpret_cast<const T*>(p.xb) + (size_t)n * p.HW * C : nullptr;
  T* out = reinterpret_cast<T*>(p.out) + (size_t)n * p.HW * C;
  for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
    const int c = (int)(i % cv) * 4;
    const size_t off = (size_t)(i / cv) * C + c;
    float v[4], r[4];
    load4(xa + off, v);
    if (xb) load4(xb + off, r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float y = fmaf(v[k], aA[c + k], bA[c + k]);
      if (p.modeB == 2) y += fmaf(r[k], aB[c + k], bB[c + k]);
      if (p.act == 1) y = y / (1.0f + expf(-y));
      else if (p.act == 2) y = fmaxf(y, 0.f);
      if (p.modeB == 1) y += r[k];
      v[k] = y;
    }
    store4(out + off, v);
  }
}


// bf16 fast path of the fused GroupNorm apply: 16-byte loads (8 channels per thread), coefficients in registers,
// SiLU through one MUFU op (tanh.approx), four independent 16-byte loads in flight per thread.
// Handles modeB in {0, 1}; requires C % 8 == 0 and 2048 % C == 0 (every GroupNorm'd tensor of the denoiser).
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// DOT: fold a 1x1 convolution to one fp32 channel into the output pass (GnApplyP::dot_out) -- a separate instantiation, so that the
// plain pass keeps its register count (73) and three blocks per SM
template <bool DOT>
__global__ void __launch_bounds__(256) gn_apply_bf16_fast_kernel(GnApplyP p, int vec_per_block) {
  extern __shared__ float sh[];  // aA[C], bA[C]
  const int n = blockIdx.y, C = p.C;
  float* aA = sh; float* bA = sh + C;
  const double cntA = (double)p.HW * (C / p.GA);
  pdl_wait();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / (C / p.GA);
    const double su = p.statsA[((size_t)n * p.GA + g) * 2], sq = p.statsA[((size_t)n * p.GA + g) * 2 + 1];
    const double mean = su / cntA;
    double var = sq / cntA - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    float a = rstd * p.gA[c], b = p.bA[c] - (float)mean * a;
    if (p.film) {
      const float sc = p.film[(size_t)n * p.film_stride + c] + 1.0f, sf = p.film[(size_t)n * p.film_stride + C + c];
      a *= sc; b = b * sc + sf;
    }
    aA[c] = a; bA[c] = b;
  }
  __syncthreads();
  const int c0 = (threadIdx.x * 8) % C;          // constant per thread: (256 * 8) % C == 0
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = aA[c0 + j]; b[j] = bA[c0 + j]; }
  const long long nvec = (long long)p.HW * C / 8;                       // 16-byte vectors of this image
  const long long v0 = (long long)blockIdx.x * vec_per_block, v1 = min(nvec, v0 + vec_per_block);
  const uint4* xa = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.xa) + (size_t)n * p.HW * C);
  const uint4* xb = p.xb ? reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.xb) + (size_t)n * p.HW * C) : nullptr;
  uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + (size_t)n * p.HW * C);
  constexpr int U = 4;
  const int lpp = C / 8;                         // lanes per pixel
  float dw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dw[j] = DOT ? p.dot_w[c0 + j] : 0.f;
  for (long long i0 = v0 + threadIdx.x; i0 < v1; i0 += 256 * U) {
    uint4 va[U], vb[U];
    float acc_dot[U] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + (long long)u * 256;
      if (i < v1) { va[u] = __ldg(xa + i); if (xb) vb[u] = __ldg(xb + i); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + (long long)u * 256;
      if (i < v1) {
        const uint32_t wa[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
        const uint32_t wb[4] = {vb[u].x, vb[u].y, vb[u].z, vb[u].w};
        uint32_t o[4];
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wa[j]));
          float y0 = fmaf(x.x, a[2 * j], b[2 * j]), y1 = fmaf(x.y, a[2 * j + 1], b[2 * j + 1]);
          if (p.act == 1) { y0 = silu_tanh(y0); y1 = silu_tanh(y1); }
          else if (p.act == 2) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
          if (xb) {
            const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wb[j]));
            y0 += r.x; y1 += r.y;
          }
          __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
          o[j] = *reinterpret_cast<uint32_t*>(&h2);
          if (DOT) { d = fmaf(y0, dw[2 * j], d); d = fmaf(y1, dw[2 * j + 1], d); }
        }
        if (!DOT) out[i] = make_uint4(o[0], o[1], o[2], o[3]);
        else acc_dot[u] = d;
      }
    }
    if (DOT) {
      // the C / 8 threads that hold one pixel are neighbouring lanes (C <= 256): butterfly over them, the first one stores
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float d = acc_dot[u];
        for (int o = 1; o < lpp; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        const long long i = i0 + (long long)u * 256;
        if (i < v1 && (threadIdx.x & (lpp - 1)) == 0) p.dot_out[(size_t)n * p.HW + i / lpp] = d + p.dot_b[0];
      }
    }
  }
}

bool gn_apply_can_dot(int C, bool bf) { return bf && C % 8 == 0 && 2048 % C == 0 && C <= 256; }
int launch_gn_apply(const GnApplyP& p, bool bf, cudaStream_t s) {
  if (p.dot_out && !(bf && p.modeB != 2 && p.C % 8 == 0 && 2048 % p.C == 0 && p.C <= 256)) return -1;   // callers check gn_apply_can_dot()
  if (bf && p.modeB != 2 && p.C % 8 == 0 && 2048 % p.C == 0) {
    const long long nvec = (long long)p.HW * p.C / 8;
    int vpb = 256 * 4 * 4;                       // 16 vectors (256 B) per thread
    if (nvec < vpb) vpb = (int)(((nvec + 255) / 256) * 256);
    dim3 grid(cdiv(nvec, vpb), p.N);
    if (p.dot_out) launch_k(gn_apply_bf16_fast_kernel<true>, grid, dim3(256), 2 * (size_t)p.C * sizeof(float), s, true, p, vpb);
    else launch_k(gn_apply_bf16_fast_kernel<false>, grid, dim3(256), 2 * (size_t)p.C * sizeof(float), s, true, p, vpb);
    return 1;
  }
  int ppb = 2048 * 32 / p.C;
  if (ppb < 32) ppb = 32;
  dim3 grid(cdiv(p.HW, ppb), p.N);
  const size_t sm = 4 * (size_t)p.C * sizeof(float);
  if (bf) gn_apply_kernel<bf16><<<grid, 256, sm, s>>>(p, ppb);
  else gn_apply_kernel<float><<<grid, 256, sm, s>>>(p, ppb);
  return 1;
}

// =================================================================================================
// RMSNorm (ddpm.py:126-132): one sub-warp of L lanes per pixel, each lane holds up to 4 vec4
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const T* __restrict__ x, const float* __restrict__ g,
                                                      const T* __restrict__ res, T* __restrict__ out, long long P, int C,
                                                      int L) {
  pdl_wait();
  const int cv = C >> 2;            // vec4 per pixel
  const int per = cv / L;           // vec4 per lane (1, 2 or 4)
  const int lane = threadIdx.x % L;
  const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const bool valid = pix < P;
  float v[4][4];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < per && valid) {
      load4(x + (size_t)pix * C + (size_t)(k * L + lane) * 4, v[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) ss = fmaf(v[k][i], v[k][i], ss);
    }
  }
  for (int o = L >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (!valid) return;
  const float scale = sqrtf((float)C) / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < per) {
      const int c = (k * L + lane) * 4;
      float r[4] = {0, 0, 0, 0};
      if (res) load4(res + (size_t)pix * C + c, r);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[k][i] = v[k][i] * scale * g[c + i] + r[i];
      store4(out + (size_t)pix * C + c, v[k]);
    }
  }
}

int launch_rmsnorm(const void* x, const float* g, const void* res, void* out, long long P, int C, bool bf, cudaStream_t s) {
  const int cv = C / 4;
  int L = cv;  // C in {32,64,96,128,...}: choose the largest power-of-two lane count <= 32 dividing cv with cv/L <= 4
  L = 1;
  while (L * 2 <= 32 && cv % (L * 2) == 0) L *= 2;
  while (cv / L > 4 && L < 32) L *= 2;
  const long long threads = P * L;
  if (bf) launch_k(rmsnorm_kernel<bf16>, dim3(cdiv(threads, 256)), dim3(256), 0, s, true, (const bf16*)x, g, (const bf16*)res, (bf16*)out, P, C, L);
  else rmsnorm_kernel<float><<<cdiv(threads, 256), 256, 0, s>>>((const float*)x, g, (const float*)res, (float*)out, P, C, L);
  return 1;
}

// =================================================================================================
// MaxPool2d(2) (unet_model.py:120)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_kernel(const T* __restrict__ x, T* __restrict__ out, int N, int H, int W,
                                                       int C) {
  const int cv = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * cv;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c = (int)(gid % cv) * 4;
  long long r = gid / cv;
  const int xo = (int)(r % Wo); r /= Wo;
  const int yo = (int)(r % Ho);
  const int n = (int)(r / Ho);
  const T* b = x + (((size_t)n * H + 2 * yo) * W + 2 * xo) * C + c;
  float a0[4], a1[4], a2[4], a3[4];
  load4(b, a0); load4(b + C, a1); load4(b + (size_t)W * C, a2); load4(b + (size_t)W * C + C, a3);
#pragma unroll
  for (int i = 0; i < 4; ++i) a0[i] = fmaxf(fmaxf(a0[i], a1[i]), fmaxf(a2[i], a3[i]));
  store4(out + (((size_t)n * Ho + yo) * Wo + xo) * C + c, a0);
}

int launch_maxpool2(const void* x, void* out, int N, int H, int W, int C, bool bf, cudaStream_t s) {
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
  if (bf) maxpool2_kernel<bf16><<<cdiv(total, 256), 256, 0, s>>>((const bf16*)x, (bf16*)out, N, H, W, C);
  else maxpool2_kernel<float><<<cdiv(total, 256), 256, 0, s>>>((const float*)x, (float*)out, N, H, W, C);
  return 1;
}

// =================================================================================================
// LinearAttention (ddpm.py:234-251).  qkv: [N,HW,3*hid], hid = heads*32; channel = part*hid + h*32 + d.
// =================================================================================================
// (1) per-chunk column max of k
template <typename T>
__global__ void __launch_bounds__(256) la_kmax_kernel(const T* __restrict__ qkv, float* __restrict__ part, int HW, int hid,
                                                      int chunk_px, int chunks) {
  extern __shared__ float sh[];  // [lanes][hid]
  const int n = blockIdx.y, ch = blockIdx.x;
  const int cv = hid >> 2;
  const int lanes = blockDim.x / cv;
  const int myc = (threadIdx.x % cv) * 4, myp = threadIdx.x / cv;
  const int p0 = ch * chunk_px, p1 = min(HW, p0 + chunk_px);
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  if (myp < lanes) {
    const T* base = qkv + (size_t)n * HW * 3 * hid + hid + myc;
    for (int p = p0 + myp; p < p1; p += lanes) {
      float v[4];
      load4(base + (size_t)p * 3 * hid, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = fmaxf(m[i], v[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) sh[myp * hid + myc + i] = m[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < hid; c += blockDim.x) {
    float mm = -INFINITY;
    for (int l = 0; l < lanes; ++l) mm = fmaxf(mm, sh[l * hid + c]);
    part[((size_t)n * chunks + ch) * hid + c] = mm;
  }
}
__global__ void la_kmax_reduce_kernel(const float* __restrict__ part, float* __restrict__ kmax, int hid, int chunks) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < hid; c += blockDim.x) {
    float mm = -INFINITY;
    for (int k = 0; k < chunks; ++k) mm = fmaxf(mm, part[((size_t)n * chunks + k) * hid + c]);
    kmax[(size_t)n * hid + c] = mm;
  }
}
// (2) ctx[d][e] += sum_p exp(k[p,d]-kmax[d]) * v[p,e];  ksum[d] += sum_p exp(..)
template <typename T>
__global__ void __launch_bounds__(256) la_context_kernel(const T* __restrict__ qkv, const float* __restrict__ kmax,
                                                         float* __restrict__ ctx, float* __restrict__ ksum, int HW,
                                                         int heads, int chunk_px) {
  __shared__ float ek[64][33];
  __shared__ float vv[64][33];
  const int n = blockIdx.z, h = blockIdx.y;
  const int hid = heads * 32;
  const int p0 = blockIdx.x * chunk_px, p1 = min(HW, p0 + chunk_px);
  const int d = threadIdx.x >> 3, e0 = (threadIdx.x & 7) * 4;
  float acc[4] = {0, 0, 0, 0}, sacc = 0.f;
  const T* base = qkv + (size_t)n * HW * 3 * hid;
  const int lp = threadIdx.x >> 3, lc = (threadIdx.x & 7) * 4;  // loader: 32 pixels x 8 vec4 per pass
  for (int t0 = p0; t0 < p1; t0 += 64) {
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int pl = lp + half * 32, p = t0 + pl;
      float kk[4] = {0, 0, 0, 0}, v4[4] = {0, 0, 0, 0};
      if (p < p1) {
        load4(base + (size_t)p * 3 * hid + hid + h * 32 + lc, kk);
        load4(base + (size_t)p * 3 * hid + 2 * hid + h * 32 + lc, v4);
#pragma unroll
        for (int i = 0; i < 4; ++i) kk[i] = expf(kk[i] - kmax[(size_t)n * hid + h * 32 + lc + i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { ek[pl][lc + i] = kk[i]; vv[pl][lc + i] = v4[i]; }
    }
    __syncthreads();
#pragma unroll 8
    for (int p = 0; p < 64; ++p) {
      const float kd = ek[p][d];
      acc[0] = fmaf(kd, vv[p][e0], acc[0]); acc[1] = fmaf(kd, vv[p][e0 + 1], acc[1]);
      acc[2] = fmaf(kd, vv[p][e0 + 2], acc[2]); acc[3] = fmaf(kd, vv[p][e0 + 3], acc[3]);
      if ((threadIdx.x & 7) == 0) sacc += kd;
    }
  }
  float* c = ctx + (((size_t)n * heads + h) * 32 + d) * 32 + e0;
#pragma unroll
  for (int i = 0; i < 4; ++i) atomicAdd(c + i, acc[i]);
  if ((threadIdx.x & 7) == 0) atomicAdd(ksum + (size_t)n * hid + h * 32 + d, sacc);
}
// (3) fold to_out.0 into the context: Mn[n][h*32+d][c] = scale * sum_e Wout[h*32+e][c] * ctx[d][e] / ksum[d]
__global__ void la_fold_kernel(const float* __restrict__ ctx, const float* __restrict__ ksum, const float* __restrict__ wout,
                               float* __restrict__ Mn, int heads, int C) {
  const int n = blockIdx.y, hid = heads * 32;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= hid * C) return;
  const int j = idx / C, c = idx % C;
  const int h = j >> 5, d = j & 31;
  const float* cx = ctx + (((size_t)n * heads + h) * 32 + d) * 32;
  float a = 0.f;
  for (int e = 0; e < 32; ++e) a = fmaf(wout[(size_t)(h * 32 + e) * C + c], cx[e], a);
  Mn[((size_t)n * hid + j) * C + c] = a * 0.17677669529663687f / ksum[(size_t)n * hid + j];
}
// (4) out = RMSNorm(softmax_d(q) @ Mn + bias) * g2 * sqrt(C) + x.  64 pixels per block.
template <typename T, int CPT>  // CPT = C/16 output channels per thread
__global__ void __launch_bounds__(256) la_out_kernel(const T* __restrict__ qkv, const float* __restrict__ Mn,
                                                     const float* __restrict__ bout, const float* __restrict__ g2,
                                                     const T* __restrict__ x, T* __restrict__ out, int HW, int heads) {
  constexpr int C = CPT * 16;
  __shared__ float qs[32][64 + 4];   // [d][pixel]
  extern __shared__ float mh[];      // [32][C]
  const int n = blockIdx.y, hid = heads * 32;
  const int p0 = blockIdx.x * 64;
  const int pg = threadIdx.x >> 4, cg = threadIdx.x & 15;
  float acc[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;
  const int sp = threadIdx.x >> 2, sd = (threadIdx.x & 3) * 8;  // softmax: 4 threads per pixel, 8 d each
  for (int h = 0; h < heads; ++h) {
    __syncthreads();
    {
      float q[8];
      const int p = p0 + sp;
      if (p < HW) {
        const T* qp = qkv + ((size_t)n * HW + p) * 3 * hid + h * 32 + sd;
        load4(qp, q); load4(qp + 4, q + 4);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = 0.f;
      }
      float m = q[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) m = fmaxf(m, q[i]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float su = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { q[i] = expf(q[i] - m); su += q[i]; }
      su += __shfl_xor_sync(0xffffffffu, su, 1);
      su += __shfl_xor_sync(0xffffffffu, su, 2);
      const float inv = 1.0f / su;
#pragma unroll
      for (int i = 0; i < 8; ++i) qs[sd + i][sp] = q[i] * inv;
    }
    for (int i = threadIdx.x; i < 32 * C; i += 256) mh[i] = Mn[((size_t)n * hid + h * 32) * C + i];
    __syncthreads();
#pragma unroll 4
    for (int d = 0; d < 32; ++d) {
      const float4 qv = *reinterpret_cast<const float4*>(&qs[d][pg * 4]);
      const float qq[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float w = mh[d * C + cg + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(qq[i], w, acc[i][j]);
      }
    }
  }
  const float sqc = sqrtf((float)C);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) { acc[i][j] += bout[cg + 16 * j]; ss = fmaf(acc[i][j], acc[i][j], ss); }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4); ss += __shfl_xor_sync(0xffffffffu, ss, 8);
    const int p = p0 + pg * 4 + i;
    if (p >= HW) continue;
    const float sc = sqc / fmaxf(sqrtf(ss), 1e-12f);
    const size_t off = ((size_t)n * HW + p) * C;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = cg + 16 * j;
      from_f(out[off + c], acc[i][j] * sc * g2[c] + to_f(x[off + c]));
    }
  }
}

template <typename T>
static int linear_attention_t(const LinAttnP& p, cudaStream_t s) {
  const int hid = p.heads * 32;
  const int chunk_px = cdiv(p.HW, p.chunks);
  const int cv = hid / 4;
  int th = (256 / cv) * cv;
  if (th == 0) th = cv;
  const int lanes = th / cv;
  la_kmax_kernel<T><<<dim3(p.chunks, p.N), th, (size_t)lanes * hid * sizeof(float), s>>>((const T*)p.qkv, p.kmax_part, p.HW,
                                                                                         hid, chunk_px, p.chunks);
  la_kmax_reduce_kernel<<<p.N, 256, 0, s>>>(p.kmax_part, p.kmax, hid, p.chunks);
  la_context_kernel<T><<<dim3(p.chunks, p.heads, p.N), 256, 0, s>>>((const T*)p.qkv, p.kmax, p.ctx, p.ksum, p.HW, p.heads,
                                                                    chunk_px);
  la_fold_kernel<<<dim3(cdiv((long long)hid * p.C, 256), p.N), 256, 0, s>>>(p.ctx, p.ksum, p.wout, p.Mn, p.heads, p.C);
  dim3 g(cdiv(p.HW, 64), p.N);
  const size_t sm = (size_t)32 * p.C * sizeof(float);
#define LA_OUT(CPT)                                                                                             \
  la_out_kernel<T, CPT><<<g, 256, sm, s>>>((const T*)p.qkv, p.Mn, p.bout, p.g2, (const T*)p.x, (T*)p.out, p.HW, \
                                           p.heads)
  switch (p.C / 16) {
    case 2: LA_OUT(2); break;
    case 4: LA_OUT(4); break;
    case 8: LA_OUT(8); break;
    case 16: LA_OUT(16); break;
    default: return -1;
  }
#undef LA_OUT
  return 5;
}
int launch_linear_attention(const LinAttnP& p, bool bf, cudaStream_t s) {
  return bf ? linear_attention_t<bf16>(p, s) : linear_attention_t<float>(p, s);
}

// =================================================================================================
// Full attention, CUDA cores (attend.py:98-113).  One thread per query, keys/values staged in smem.
// qkv: [N,n,3*hid]; out: [N,n,hid] with channel = h*32 + d (ddpm.py:281).
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(128) attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int n, int heads) {
  __shared__ float Ks[64][32];
  __shared__ float Vs[64][32];
  const int b = blockIdx.z, h = blockIdx.y, hid = heads * 32;
  const int qi = blockIdx.x * 128 + threadIdx.x;
  const bool valid = qi < n;
  const T* base = qkv + (size_t)b * n * 3 * hid;
  float q[32], o[32];
  const float scale = 0.17677669529663687f;  // 32^-0.5
#pragma unroll
  for (int d = 0; d < 32; d += 4) {
    float t[4] = {0, 0, 0, 0};
    if (valid) load4(base + (size_t)qi * 3 * hid + h * 32 + d, t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { q[d + i] = t[i] * scale; o[d + i] = 0.f; }
  }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < n; j0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {
      const int r = i >> 3, c = (i & 7) * 4;
      float kk[4] = {0, 0, 0, 0}, v4[4] = {0, 0, 0, 0};
      if (j0 + r < n) {
        load4(base + (size_t)(j0 + r) * 3 * hid + hid + h * 32 + c, kk);
        load4(base + (size_t)(j0 + r) * 3 * hid + 2 * hid + h * 32 + c, v4);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { Ks[r][c + k] = kk[k]; Vs[r][c + k] = v4[k]; }
    }
    __syncthreads();
    const int jn = min(64, n - j0);
    for (int js = 0; js < jn; js += 16) {
      float sc[16];
      float tm = m;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) a = fmaf(q[d], Ks[js + j][d], a);
        sc[j] = (js + j < jn) ? a : -INFINITY;
        tm = fmaxf(tm, sc[j]);
      }
      const float corr = expf(m - tm);
      l *= corr;
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] *= corr;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pj = expf(sc[j] - tm);
        l += pj;
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] = fmaf(pj, Vs[js + j][d], o[d]);
      }
      m = tm;
    }
  }
  if (!valid) return;
  const float inv = 1.0f / l;
  T* op = out + ((size_t)b * n + qi) * hid + h * 32;
#pragma unroll
  for (int d = 0; d < 32; d += 4) {
    float t[4] = {o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv};
    store4(op + d, t);
  }
}

int launch_attention_simt(const void* qkv, void* out, int N, int n, int heads, bool bf, cudaStream_t s) {
  dim3 g(cdiv(n, 128), heads, N);
  if (bf) attn_simt_kernel<bf16><<<g, 128, 0, s>>>((const bf16*)qkv, (bf16*)out, n, heads);
  else attn_simt_kernel<float><<<g, 128, 0, s>>>((const float*)qkv, (float*)out, n, heads);
  return 1;
}

// =================================================================================================
// time embedding + FiLM vectors (ddpm.py:142-149, 339-344, 191-194)
// =================================================================================================
__global__ void time_embed_kernel(TimeP p) {
  extern __shared__ float sh[];  // emb[dim], h1[4dim]
  const int n = blockIdx.x, dim = p.dim, td = 4 * dim, half = dim / 2;
  float* emb = sh; float* h1 = sh + dim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float tf = p.t_scalar ? (float)(*p.t_scalar) : (float)p.t[n];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf((float)i * p.neg_step);
    const float a = tf * f;
    emb[i] = sinf(a); emb[half + i] = cosf(a);
  }
  __syncthreads();
  // one warp per output row: coalesced weight reads, shuffle reduction (rows are too few for a thread each to hide latency)
  for (int j = warp; j < td; j += nwarps) {
    float a = 0.f;
    for (int k = lane; k < dim; k += 32) a = fmaf(p.w1[(size_t)j * dim + k], emb[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    a += p.b1[j];
    if (lane == 0) h1[j] = 0.5f * a * (1.0f + erff(a * 0.70710678118654752f));  // exact GELU (nn.GELU default)
  }
  __syncthreads();
  for (int j = warp; j < td; j += nwarps) {
    float a = 0.f;
    for (int k = lane; k < td; k += 32) a = fmaf(p.w2[(size_t)j * td + k], h1[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    a += p.b2[j];
    if (lane == 0) p.st[(size_t)n * td + j] = a / (1.0f + expf(-a));  // SiLU in front of every block MLP (ddpm.py:192)
  }
}
// all block MLPs at once (ddpm.py:191-194): one warp per output row j, coalesced weight reads, every image of the batch
__global__ void film_kernel(TimeP p) {
  extern __shared__ float st[];  // [N][4dim]
  const int td = 4 * p.dim;
  for (int i = threadIdx.x; i < p.N * td; i += blockDim.x) st[i] = p.st[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= p.total) return;
  const float* w = p.wf + (size_t)j * td;
  const float bj = p.bf_[j];
  for (int n = 0; n < p.N; ++n) {
    float a = 0.f;
    for (int k = lane; k < td; k += 32) a = fmaf(w[k], st[n * td + k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) p.film[(size_t)n * p.total + j] = a + bj;
  }
}
int launch_time_film(const TimeP& p, cudaStream_t s) {
  // the sampler shares one timestep across the batch (t_scalar): a single row is computed and consumers use stride 0
  time_embed_kernel<<<p.N, 512, (size_t)5 * p.dim * sizeof(float), s>>>(p);
  int launches = 1;
  for (int n0 = 0; n0 < p.N; n0 += 16) {   // 16 images per launch keep the staged embeddings inside 48 KB
    TimeP q = p;
    q.N = p.N - n0 < 16 ? p.N - n0 : 16;
    q.st = p.st + (size_t)n0 * 4 * p.dim; q.film = p.film + (size_t)n0 * p.total;
    film_kernel<<<cdiv(p.total, 8), 256, (size_t)q.N * 4 * p.dim * sizeof(float), s>>>(q);
    ++launches;
  }
  return launches;
}

__global__ void film_gather_kernel(const float* __restrict__ table, int total, const int* __restrict__ t_scalar, float* __restrict__ film) {
  if (threadIdx.x == 0) pdl_trigger();   // first kernel of a timestep (launched without the PDL attribute): lets the init conv set itself up
  pdl_wait();
  const float4* src = reinterpret_cast<const float4*>(table + (size_t)(*t_scalar) * total);
  float4* dst = reinterpret_cast<float4*>(film);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total / 4; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
int launch_film_gather(const float* table, int total, const int* t_scalar, float* film, cudaStream_t s) {
  film_gather_kernel<<<cdiv(total / 4, 256), 256, 0, s>>>(table, total, t_scalar, film);
  return 1;
}

// =================================================================================================
// sampler elementwise kernels
// =================================================================================================
// mask partition of the conditional image (ddpm.py:672-690)
__global__ void prep_cond_kernel(PrepP p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int ones = 0, zeros = 0;
  if (i < p.n) {
    const float bm = p.mask[i] >= 1.0f ? 1.0f : 0.0f;
    const float c = p.cond[i];
    p.bm[i] = bm;
    p.cond_out[i] = __fmul_rn(c, bm);
    const float m2 = fminf(fmaxf(1.0f - bm, p.floor), 1.0f);
    p.cond_in[i] = __fmul_rn(c, m2);
    ones = bm == 1.0f; zeros = bm == 0.0f;
  }
  ones = __reduce_add_sync(0xffffffffu, ones);
  zeros = __reduce_add_sync(0xffffffffu, zeros);
  if ((threadIdx.x & 31) == 0) {
    if (ones) atomicAdd(p.counters, ones);
    if (zeros) atomicAdd(p.counters + 1, zeros);
  }
}
int launch_prep_cond(const PrepP& p, cudaStream_t s) {
  prep_cond_kernel<<<cdiv(p.n, 256), 256, 0, s>>>(p);
  return 1;
}

// One DDPM update (ddpm.py:841-860) with the x0 handling of p_mean_variance / model_predictions
// (ddpm.py:697-708, 775-810).  Separate mul/add roundings (no FMA contraction) so that, given the
// same x0, the update is bit-identical to the reference's chain of elementwise torch ops.
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float post(float c1, float c2, float sg, float x0, float xt, float z) {
  const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xt));
  return __fadd_rn(mean, __fmul_rn(sg, z));
}
// Every thread owns four consecutive pixels (16-byte loads / stores of each of the fp32 planes; n % 4 == 0).  The loop bookkeeping of
// p_sample_loop (`t -= 1`, ddpm.py:951) is folded in: every block reads t when it starts and takes a ticket when it is done; the block
// that draws the last ticket -- by then every block has read t -- writes t - 1 for the next timestep and re-arms the ticket.
__device__ __forceinline__ float4 ld4(const float* p, long long i) { return *reinterpret_cast<const float4*>(p + i); }
__device__ __forceinline__ void st4(float* p, long long i, const float (&v)[4]) { *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void unpack4(const float4 v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ bool last_block_done(unsigned int* ticket) {
  __shared__ unsigned int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int k = atomicAdd(ticket, 1u);
    s_last = (k == gridDim.x - 1) ? 1u : 0u;
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  return s_last != 0u;
}
__global__ void __launch_bounds__(256) step_kernel(StepP p) {
  pdl_wait();
  const int t = *p.t_ptr;
  const float c1 = p.coef1[t], c2 = p.coef2[t], sg = p.sigma[t];
  const float* z = (t > 0 && p.z) ? p.z + (size_t)(p.tloop - t) * p.z_stride : nullptr;
  float* tr = p.x0_trace ? p.x0_trace + (size_t)(p.tloop - 1 - t) * p.trace_stride : nullptr;
  unsigned int zo = 0, zi = 0;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i < p.n) {
    float zz[4] = {0.f, 0.f, 0.f, 0.f}, xo[4], oo[4] = {0.f, 0.f, 0.f, 0.f}, r0[4], r1[4];
    if (z) unpack4(ld4(z, i), zz);
    unpack4(ld4(p.x_out, i), xo);
    if (p.o_out) unpack4(ld4(p.o_out, i), oo);
    if (p.kind == 2) {
      const bool conv = p.ca != nullptr;   // pred_noise / pred_v: x0 from (x_t, model output), separately rounded like the reference
      const float ca = conv ? p.ca[t] : 0.f, cb = conv ? p.cb[t] : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float raw = conv ? __fsub_rn(__fmul_rn(ca, xo[k]), __fmul_rn(cb, oo[k])) : oo[k];
        r0[k] = clampf(raw, p.lo, p.hi);
        xo[k] = post(c1, c2, sg, r0[k], xo[k], zz[k]);
      }
      st4(p.x_out, i, xo);
      if (p.x0_out) st4(p.x0_out, i, r0);
      if (tr) { st4(tr, i, r0); st4(tr + p.n, i, xo); }   // trace slot 1 on single-trajectory steps: the updated image x_{t-1}
    } else {
      float bm[4], xi[4], oi[4], co[4] = {0.f, 0.f, 0.f, 0.f};
      unpack4(ld4(p.bm, i), bm);
      unpack4(ld4(p.x_in, i), xi);
      unpack4(ld4(p.o_in, i), oi);
      if (p.mask_x && p.ood_uses_cond) unpack4(ld4(p.cond_out, i), co);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v;
        if (p.mask_x) {
          if (p.ood_uses_cond) v = co[k];
          else v = (bm[k] == 0.0f) ? p.lo : __fmul_rn(oo[k], bm[k]);
        } else {
          v = oo[k];
        }
        const float x0o = clampf(v, p.lo, p.hi);
        const float x0i = clampf(oi[k], p.lo, p.hi);
        if (p.kind == 0) {
          xo[k] = post(c1, c2, sg, x0o, xo[k], zz[k]);
          xi[k] = post(c1, c2, sg, x0i, xi[k], zz[k]);
          r0[k] = x0o; r1[k] = x0i;
        } else {  // fusion step (ddpm.py:779-810)
          const float im = 1.0f - bm[k];
          float x0 = __fadd_rn(__fmul_rn(x0i, im), x0o);
          const float a = __fmul_rn(xo[k], bm[k]), b = __fmul_rn(xi[k], im);
          zo += a == 0.0f; zi += b == 0.0f;
          const float xt = (a == 0.0f) ? b : a;
          x0 = clampf(x0, p.lo, p.hi);
          xo[k] = post(c1, c2, sg, x0, xt, zz[k]);
          r0[k] = x0;
        }
      }
      st4(p.x_out, i, xo);
      if (p.kind == 0) st4(p.x_in, i, xi);
      if (p.x0_out) st4(p.x0_out, i, r0);
      if (p.kind == 0 && p.x0_in) st4(p.x0_in, i, r1);
      if (tr) { st4(tr, i, r0); if (p.kind == 0) st4(tr + p.n, i, r1); }
    }
  }
  if (p.kind == 1) {
    zo = __reduce_add_sync(0xffffffffu, zo);
    zi = __reduce_add_sync(0xffffffffu, zi);
    if ((threadIdx.x & 31) == 0) {
      if (zo) atomicAdd(p.counters + 2, zo);
      if (zi) atomicAdd(p.counters + 3, zi);
    }
  }
  if (p.ticket && last_block_done(p.ticket) && threadIdx.x == 0) *p.t_ptr = t - 1;
}
int launch_step(const StepP& p, cudaStream_t s) {
  launch_k(step_kernel, dim3(cdiv(cdiv(p.n, 4), 256)), dim3(256), 0, s, true, p);
  return 1;
}
// ---- DDIM (ddpm.py:979-1075): every product and sum is rounded separately, in the reference's evaluation order ----
// predict_noise_from_start (ddpm.py:637-641)
__device__ __forceinline__ float eps_from(float sr, float srm1, float xt, float x0) {
  return __fdiv_rn(__fsub_rn(__fmul_rn(sr, xt), x0), srm1);
}
// x_start * alpha_next.sqrt() + c * pred_noise + sigma * noise (ddpm.py:1041-1043, 1068-1070)
__device__ __forceinline__ float ddim_next(float san, float c, float sg, float x0, float eps, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x0, san), __fmul_rn(c, eps)), __fmul_rn(sg, z));
}
__global__ void __launch_bounds__(256) ddim_step_kernel(DdimP p) {
  pdl_wait();
  const int idx = *p.idx_ptr;
  const bool last = idx >= p.nsteps - 1;   // time_next < 0 (ddpm.py:1009-1012, 1053-1056)
  const float* cf = p.coefs + (size_t)idx * 5;
  const float sr = cf[0], srm1 = cf[1], san = cf[2], c = cf[3], sg = cf[4];
  const float* z = (!last && p.z) ? p.z + (size_t)(1 + idx) * p.z_stride : nullptr;
  unsigned int zo = 0, zi = 0;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i < p.n) {
    float zz[4] = {0.f, 0.f, 0.f, 0.f}, xo[4], oo[4] = {0.f, 0.f, 0.f, 0.f};
    if (z) unpack4(ld4(z, i), zz);
    unpack4(ld4(p.x_out, i), xo);
    if (p.o_out) unpack4(ld4(p.o_out, i), oo);
    if (p.kind == 2) {
      const bool conv = p.ca != nullptr;
      const int tt = conv ? p.times[idx] : 0;
      const float ca = conv ? p.ca[tt] : 0.f, cb = conv ? p.cb[tt] : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // clip_x_start=True, rederive_pred_noise=True (ddpm.py:1049): eps is re-derived from the clamped x0 for every objective
        const float raw = conv ? __fsub_rn(__fmul_rn(ca, xo[k]), __fmul_rn(cb, oo[k])) : oo[k];
        const float x0 = clampf(raw, p.lo, p.hi);
        xo[k] = last ? x0 : ddim_next(san, c, sg, x0, eps_from(sr, srm1, xo[k], x0), zz[k]);
      }
      st4(p.x_out, i, xo);
    } else {
      float bm[4], xi[4], oi[4], co[4] = {0.f, 0.f, 0.f, 0.f};
      unpack4(ld4(p.bm, i), bm);
      unpack4(ld4(p.x_in, i), xi);
      unpack4(ld4(p.o_in, i), oi);
      if (p.mask_x && p.ood_uses_cond) unpack4(ld4(p.cond_out, i), co);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v;
        if (p.mask_x) {   // ddpm.py:697-708
          if (p.ood_uses_cond) v = co[k];
          else v = (bm[k] == 0.0f) ? p.lo : __fmul_rn(oo[k], bm[k]);
        } else {
          v = oo[k];
        }
        const float x0o = clampf(v, p.lo, p.hi), x0i = clampf(oi[k], p.lo, p.hi);
        if (last) {                         // img = [x_start_out, x_start_in]: never fused on the last step
          xo[k] = x0o; xi[k] = x0i;
        } else {
          const float eo = eps_from(sr, srm1, xo[k], x0o), ei = eps_from(sr, srm1, xi[k], x0i);
          if (p.kind == 0) {
            xo[k] = ddim_next(san, c, sg, x0o, eo, zz[k]);
            xi[k] = ddim_next(san, c, sg, x0i, ei, zz[k]);
          } else {                          // fusion: select on x_start_out == 0, composite of the noise predictions
            const float x0 = clampf(x0o == 0.0f ? x0i : x0o, p.lo, p.hi);
            const float a = __fmul_rn(eo, bm[k]), b = __fmul_rn(ei, __fsub_rn(1.0f, bm[k]));
            zo += a == 0.0f; zi += b == 0.0f;
            const float eps = (a == 0.0f) ? b : a;
            xo[k] = ddim_next(san, c, sg, x0, eps, zz[k]);
          }
        }
      }
      st4(p.x_out, i, xo);
      if (last || p.kind == 0) st4(p.x_in, i, xi);
    }
  }
  if (p.kind == 1) {
    zo = __reduce_add_sync(0xffffffffu, zo);
    zi = __reduce_add_sync(0xffffffffu, zi);
    if ((threadIdx.x & 31) == 0) {
      if (zo) atomicAdd(p.counters + 2, zo);
      if (zi) atomicAdd(p.counters + 3, zi);
    }
  }
  // step bookkeeping folded in (see step_kernel): idx += 1; t = times[idx] (ddpm.py:996-998)
  if (p.ticket && last_block_done(p.ticket) && threadIdx.x == 0) {
    *p.idx_ptr = idx + 1;
    if (idx + 1 < p.nsteps) *p.t_ptr = p.times[idx + 1];
  }
}
int launch_ddim_step(const DdimP& p, cudaStream_t s) {
  launch_k(ddim_step_kernel, dim3(cdiv(cdiv(p.n, 4), 256)), dim3(256), 0, s, true, p);
  return 1;
}
// =================================================================================================
// layout helpers
// =================================================================================================
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int HW, int C, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // index into NCHW output
  if (i >= total) return;
  const int p = (int)(i % HW);
  const int c = (int)((i / HW) % C);
  const long long n = i / ((long long)HW * C);
  out[i] = to_f(in[((size_t)n * HW + p) * C + c]);
}
int launch_nhwc_to_nchw_f32(const void* in, float* out, int N, int HW, int C, bool bf, cudaStream_t s) {
  const long long total = (long long)N * HW * C;
  if (bf) nhwc_to_nchw_kernel<bf16><<<cdiv(total, 256), 256, 0, s>>>((const bf16*)in, out, HW, C, total);
  else nhwc_to_nchw_kernel<float><<<cdiv(total, 256), 256, 0, s>>>((const float*)in, out, HW, C, total);
  return 1;
}
template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) from_f(out[i], to_f(in[i]));
}
int launch_convert(const void* in, bool in_bf, void* out, bool out_bf, long long n, cudaStream_t s) {
  const int g = cdiv(n, 256);
  if (in_bf && out_bf) convert_kernel<bf16, bf16><<<g, 256, 0, s>>>((const bf16*)in, (bf16*)out, n);
  else if (in_bf) convert_kernel<bf16, float><<<g, 256, 0, s>>>((const bf16*)in, (float*)out, n);
  else if (out_bf) convert_kernel<float, bf16><<<g, 256, 0, s>>>((const float*)in, (bf16*)out, n);
  else convert_kernel<float, float><<<g, 256, 0, s>>>((const float*)in, (float*)out, n);
  return 1;
}
__global__ void copy_f32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
int launch_copy_f32(const float* in, float* out, long long n, cudaStream_t s) {
  copy_f32_kernel<<<cdiv(n, 256), 256, 0, s>>>(in, out, n);
  return 1;
}

}  // namespace ld
