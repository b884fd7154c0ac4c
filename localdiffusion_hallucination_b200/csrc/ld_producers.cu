// The two stages immediately BEFORE the sampler hot path, as small HBM-bound kernels:
//   * conditional-image producers: MNIST 2x down / bilinear up + [0,2] scaling (data.py:814-836), MRI centre crop +
//     normalise + translate-zero (data.py:380-414);
//   * anomaly map -> soft mask (`mask_pred`, == 1.0 inside the OOD region) and binary mask with the reference's per-dataset
//     threshold rules (test.py:237-381), incl. the bilinear resize to the image size and the manual left-columns override.
// Arithmetic is fp32 with separately rounded operations in the reference's evaluation order (no FMA contraction).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ld_sampler.h"
#include "ld_kernels.h"

namespace ld {

namespace {

// order-preserving float <-> unsigned transform for atomicMax / atomicMin on floats of either sign
__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// torch's upsample_bilinear2d, align_corners=False (`area_pixel_compute_source_index`): src = max(scale * (dst + 0.5) - 0.5, 0)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int* i0, int* i1, float* l1) {
  float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  s = s < 0.f ? 0.f : s;
  int a = (int)s;
  if (a > in_size - 1) a = in_size - 1;
  *i0 = a;
  *i1 = a + (a < in_size - 1 ? 1 : 0);
  *l1 = __fsub_rn(s, (float)a);
}
// `src` is addressed as src[(y * sy) * pitch + x * sx]: (sy, sx) = (2, 2) reads the [::2, ::2] sub-sampled image in place
__device__ __forceinline__ float bilinear(const float* src, int pitch, int sy, int sx, int Hi, int Wi, float rh, float rw, int y, int x) {
  int y0, y1, x0, x1; float ly, lx;
  src_index(y, rh, Hi, &y0, &y1, &ly);
  src_index(x, rw, Wi, &x0, &x1, &lx);
  const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
  const float a = src[(size_t)(y0 * sy) * pitch + x0 * sx], b = src[(size_t)(y0 * sy) * pitch + x1 * sx];
  const float c = src[(size_t)(y1 * sy) * pitch + x0 * sx], d = src[(size_t)(y1 * sy) * pitch + x1 * sx];
  const float top = __fadd_rn(__fmul_rn(hx, a), __fmul_rn(lx, b)), bot = __fadd_rn(__fmul_rn(hx, c), __fmul_rn(lx, d));
  return __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
}

// ---- MNIST (data.py:814-836): hr = 2 * (x / 255); cond = 2 * (bilinear_up(sub-sampled x) / 255).  The reference's `img[:, ::2, ::2]`
// acts on the 4-D tensor [1, 1, S, S] (data.py:822-826): only the rows are sub-sampled, so the up-sampling is vertical only.
__global__ void __launch_bounds__(256) mnist_cond_kernel(const float* __restrict__ raw, float* __restrict__ hr, float* __restrict__ cond,
                                                         int N, int S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * S * S) return;
  const int x = (int)(i % S), y = (int)((i / S) % S);
  const float* img = raw + (i / ((long long)S * S)) * S * S;
  const int Sd = (S + 1) / 2;
  const float scale = (float)Sd / (float)S;
  hr[i] = __fmul_rn(2.f, __fdiv_rn(img[(size_t)y * S + x], 255.f));
  cond[i] = __fmul_rn(2.f, __fdiv_rn(bilinear(img, S, 2, 1, Sd, S, scale, 1.0f, y, x), 255.f));
}

// ---- MRI (data.py:380-414): centre crop, (x - mean) / std, then + |min| of the image when translate_zero ----------------------
__global__ void __launch_bounds__(256) mri_norm_kernel(const float* __restrict__ raw, float* __restrict__ out, unsigned int* __restrict__ mins,
                                                       int Hs, int Ws, int crop, float mean, float stdv, int pass) {
  const int n = blockIdx.y;
  // torchvision CenterCrop: top = round((H - crop) / 2), left = round((W - crop) / 2) (python round: half to even)
  const int top = (int)rintf((float)(Hs - crop) * 0.5f), left = (int)rintf((float)(Ws - crop) * 0.5f);
  const float* img = raw + (size_t)n * Hs * Ws;
  float* o = out + (size_t)n * crop * crop;
  float mn = INFINITY;
  const float shift = pass == 1 ? fabsf(ord2f(mins[n])) : 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < crop * crop; i += gridDim.x * blockDim.x) {
    const int y = i / crop, x = i - y * crop;
    const float v = __fdiv_rn(__fsub_rn(img[(size_t)(top + y) * Ws + left + x], mean), stdv);
    if (pass == 0) { mn = fminf(mn, v); if (!mins) o[i] = v; }
    else o[i] = __fadd_rn(v, shift);
  }
  if (pass == 0 && mins) {
    for (int k = 16; k > 0; k >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, k));
    if ((threadIdx.x & 31) == 0 && mn < INFINITY) atomicMin(mins + n, f2ord(mn));
  }
}

// ---- anomaly map -> masks (test.py:237-381) -----------------------------------------------------------------------------------
struct MaskStats { unsigned int vmax, vmin; double sum, sumsq; };

// (optional bilinear resize to S x S, test.py:254-255) + global max / min / sum / sum of squares over the whole batch
__global__ void __launch_bounds__(256) amap_stats_kernel(const float* __restrict__ amap, float* __restrict__ resized, MaskStats* st, int B, int h,
                                                         int w, int S, int resize) {
  const long long total = (long long)B * S * S;
  float mx = -INFINITY, mn = INFINITY; double su = 0, sq = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v;
    if (resize) {
      const int x = (int)(i % S), y = (int)((i / S) % S);
      const float* img = amap + (i / ((long long)S * S)) * h * w;
      v = bilinear(img, w, 1, 1, h, w, (float)h / (float)S, (float)w / (float)S, y, x);
      resized[i] = v;
    } else v = amap[i];
    mx = fmaxf(mx, v); mn = fminf(mn, v); su += (double)v; sq += (double)v * (double)v;
  }
  for (int k = 16; k > 0; k >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, k)); mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, k));
    su += __shfl_xor_sync(0xffffffffu, su, k); sq += __shfl_xor_sync(0xffffffffu, sq, k);
  }
  if ((threadIdx.x & 31) == 0 && mx > -INFINITY) {
    atomicMax(&st->vmax, f2ord(mx)); atomicMin(&st->vmin, f2ord(mn));
    atomicAdd(&st->sum, su); atomicAdd(&st->sumsq, sq);
  }
}

// the per-dataset threshold rules of test.py:259-365; returns false when the anomaly score is below the gate (masks := 1)
__device__ bool mask_rule(int rule, float amax, float amin, float sd, float* thr, float* lo) {
  float t = 0.f, l = 0.f;
  switch (rule) {
    case LD_MASK_MNIST_8TO3:        // test.py:261-274
      if (!(amax > 37.0f)) return false;
      t = amax > 44.f ? 41.7f : (amax > 40.0f ? 38.2f : 35.0f); l = __fsub_rn(t, sd); break;
    case LD_MASK_MNIST_8TO5:        // test.py:275-289
      if (!(amax > 58.5f)) return false;
      t = amax > 71.0f ? 61.0f : (amax > 65.f ? 57.0f : 55.0f); l = __fsub_rn(t, sd); break;
    case LD_MASK_MRI_T12FLAIR:      // test.py:299-315
      if (!(amax > 43.f)) return false;
      t = amax > 60.f ? __fsub_rn(amax, 12.f) : (amax > 51.f ? 47.f : (amax > 48.5f ? 44.f : 42.f)); l = __fsub_rn(t, sd); break;
    case LD_MASK_MRI_FLAIR2T1:      // test.py:317-331
      if (!(amax > 43.f)) return false;
      t = amax > 60.f ? 47.f : (amax > 50.f ? 43.f : 42.f); l = __fsub_rn(t, sd); break;
    case LD_MASK_MVTEC_TRANSISTOR:  // test.py:337-354
      if (!(amax > 32.f)) return false;
      t = amax > 40.0f ? 33.5f : (amax > 36.8f ? __fsub_rn(amax, __fmul_rn(2.f, sd)) : (amax > 35.0f ? __fsub_rn(amax, sd) : 29.5f));
      l = __fsub_rn(t, __fmul_rn(0.5f, sd)); break;
    case LD_MASK_MVTEC_TOOTHBRUSH:  // test.py:355-366
      if (!(amax > 35.f)) return false;
      t = amax > 49.f ? 40.0f : 28.0f; l = amin; break;
    default:                        // LD_MASK_MVTEC_GRID, test.py:367-381
      if (!(amax > 27.f)) return false;
      t = amax > 40.f ? 35.0f : (amax > 35.0f ? 30.0f : 26.5f); l = amin; break;
  }
  *thr = t; *lo = l;
  return true;
}

__global__ void __launch_bounds__(256) mask_apply_kernel(const float* __restrict__ a, const MaskStats* __restrict__ st, float* __restrict__ mask_pred,
                                                         float* __restrict__ binary, long long total, int S, int rule, int manual_cols) {
  const float amax = ord2f(st->vmax), amin = ord2f(st->vmin);
  const double n = (double)total;
  double var = (st->sumsq - st->sum * st->sum / n) / (n - 1.0);   // torch.std(): unbiased, over every element
  if (var < 0) var = 0;
  const float sd = (float)sqrt(var);
  float thr = 0.f, lo = 0.f;
  const bool active = mask_rule(rule, amax, amin, sd, &thr, &lo);
  // map_pred = clip(a, lo, thr); its minimum is the clipped minimum of a
  const float mn = fminf(fmaxf(amin, lo), thr);
  const float den = __fsub_rn(thr, mn);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float mp = 1.f, bm = 1.f;
    if (manual_cols > 0) {          // test.py:379-381: the manual mask overrides whatever the detector produced
      mp = bm = ((int)(i % S) < manual_cols) ? 1.f : 0.f;
    } else if (active) {
      const float v = a[i];
      bm = v > thr ? 1.f : 0.f;
      const float c = fminf(fmaxf(v, lo), thr);
      const float r = __fdiv_rn(__fsub_rn(c, mn), den);
      mp = __fmul_rn(r, r);
    }
    mask_pred[i] = mp;
    if (binary) binary[i] = bm;
  }
}

}  // namespace

int launch_mnist_cond(const float* raw, float* hr, float* cond, int N, int S, cudaStream_t s) {
  const long long total = (long long)N * S * S;
  mnist_cond_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(raw, hr, cond, N, S);
  return 1;
}

int launch_mri_norm(const float* raw, float* out, unsigned int* mins, int N, int Hs, int Ws, int crop, float mean, float stdv, int translate_zero,
                    cudaStream_t s) {
  int gx = (crop * crop + 256 * 8 - 1) / (256 * 8);
  if (gx < 1) gx = 1;
  if (!translate_zero) { mri_norm_kernel<<<dim3(gx, N), 256, 0, s>>>(raw, out, nullptr, Hs, Ws, crop, mean, stdv, 0); return 1; }
  cudaMemsetAsync(mins, 0xff, (size_t)N * sizeof(unsigned int), s);   // ordered-uint +max
  mri_norm_kernel<<<dim3(gx, N), 256, 0, s>>>(raw, out, mins, Hs, Ws, crop, mean, stdv, 0);
  mri_norm_kernel<<<dim3(gx, N), 256, 0, s>>>(raw, out, mins, Hs, Ws, crop, mean, stdv, 1);
  return 2;
}

size_t mask_scratch_bytes(int B, int S) { return 256 + (size_t)B * S * S * sizeof(float); }

int launch_mask_from_anomaly(const float* amap, int B, int h, int w, int S, int rule, int manual_cols, float* mask_pred, float* binary, void* scratch,
                             cudaStream_t s) {
  MaskStats* st = (MaskStats*)scratch;
  float* resized = (float*)((char*)scratch + 256);
  const int resize = (h != S || w != S) ? 1 : 0;
  const long long total = (long long)B * S * S;
  MaskStats init; init.vmax = 0u; init.vmin = 0xffffffffu; init.sum = 0; init.sumsq = 0;
  cudaMemcpyAsync(st, &init, sizeof init, cudaMemcpyHostToDevice, s);
  int g = (int)((total + 255) / 256); if (g > 148 * 8) g = 148 * 8;
  amap_stats_kernel<<<g, 256, 0, s>>>(amap, resized, st, B, h, w, S, resize);
  mask_apply_kernel<<<g, 256, 0, s>>>(resize ? resized : amap, st, mask_pred, binary, total, S, rule, manual_cols);
  return 2;
}

}  // namespace ld
