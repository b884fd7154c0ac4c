// Shared device helpers for the LocalDiffusion sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ld {

typedef __nv_bfloat16 bf16;

// ---- 4-wide vector access on NHWC channel runs (fp32 math, T storage) -------------------------
__device__ __forceinline__ void load4(const float* p, float v[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float v[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
__device__ __forceinline__ void store4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float v[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void from_f(float& d, float x) { d = x; }
__device__ __forceinline__ void from_f(bf16& d, float x) { d = __float2bfloat16_rn(x); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- parameter blocks (plain structs, passed by value) ----------------------------------------
// Implicit-GEMM convolution over NHWC tensors.  The input is the *virtual concatenation* of up to
// two sources along C (torch.cat in ddpm.py:435,439,442,448 never materialises); `up` reads the
// source through a nearest x2 up-sampling (ddpm.py:116); the pixel-unshuffle down-sampling
// (ddpm.py:122) is expressed as a 2x2 stride-2 convolution with re-ordered weights.
struct ConvP {
  const void* src0; const void* src1;
  int C0, C1;
  int N, H, W;        // output extent
  int Hin, Win;       // extent of the stored sources
  int ks, stride, pad, up;
  const float* w;     // [ks*ks][C0+C1][Cout] fp32
  const float* bias;  // [Cout] or null
  int Cout;
  void* dst;          // [N,H,W,Cout]
  const void* res;    // optional residual, same shape as dst, added after bias
  long long M;        // N*H*W
};

struct GnApplyP {
  const void* xa; const double* statsA; const float* gA; const float* bA; int GA;
  const void* xb; const double* statsB; const float* gB; const float* bB; int GB;
  int modeB;          // 0 none, 1 raw add AFTER the activation, 2 GroupNorm'd add BEFORE it
  const float* film;  // per-image [2C]: scale then shift (ddpm.py:204-206), or null
  int film_stride;
  int act;            // 0 none, 1 SiLU, 2 ReLU
  void* out;
  int N, HW, C;
  float eps;
  // bf16 fast path only: instead of storing the C-channel result, reduce it with a 1x1 convolution to ONE fp32 channel
  // (final_conv of the denoiser, ddpm.py:398, folded into the last ResnetBlock's output pass): dot_out[n*HW + p] = dot_b[0] + sum_c dot_w[c] y[p,c]
  const float* dot_w; const float* dot_b; float* dot_out;
};

}  // namespace ld
