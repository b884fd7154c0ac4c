// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ld {

// weights re-packed for the tensor-core kernel (see ld_conv_tc.cu for the layout)
struct ConvTcW {
  bool ready = false;
  int Cin = 0, Cout = 0, ks = 1, stride = 1, pad = 0;
  int ntile = 0;         // output channels per CTA (UMMA N)
  void* w = nullptr;     // bf16, [n_tile][cchunk][tap][kc/8][ntile][8] with kc = 64 (null if Cin % 64)
  void* w32 = nullptr;   // same with kc = 32 (used when a concat source is not a multiple of 64 channels)
  float* bias = nullptr; // fp32 [Cout] or null
  float bias_h[64] = {}, bias2_h[64] = {};   // host copies for Cout <= 64: they travel in the kernel parameters (constant bank)
  // dual packing: every channel chunk carries a tenth stage, the 1x1 res_conv (ddpm.py:198) of the same input, whose
  // product with the centre-tap view goes to a second accumulator and a second output tensor (ConvTcArgs::dst2)
  bool dual = false;
  float* bias2 = nullptr;
};

struct ConvTcArgs {
  const void* src0 = nullptr; const void* src1 = nullptr;  // bf16 NHWC, virtual concat along C
  int C0 = 0, C1 = 0;
  int N = 0, H = 0, W = 0;     // output extent
  int Hin = 0, Win = 0;        // stored source extent
  int up = 0;                  // read src0 through a nearest x2 up-sampling
  int ds = 0;                  // pixel-unshuffle down-sampling (ddpm.py:122): weights packed as a 1x1 over (p1, p2, c); Hin = 2H
  int ps = 0;                  // nearest x2 up-sampling + 3x3 (ddpm.py:114-118) run as ONE 3x3 convolution over the LOW-resolution source
                               // with 4 * ps "virtual" output channels (output parity (py, px), channel c), see conv_tc_pack_up2: N, H, W are
                               // the low-resolution extent, dst is [N, 2H, 2W, ps] and the epilogue scatters (pixel shuffle); ps = real Cout
  void* dst = nullptr;         // bf16 [N,H,W,Cout]
  const void* res = nullptr;   // optional bf16 residual added in the epilogue
  void* dst2 = nullptr;        // dual weights only: bf16 [N,H,W,Cout] output of the fused 1x1 convolution
  // fused "normalise on load" prologue on src0 (3x3, single source, no up-sampling):
  //   x' = act(a[n,c] * x + b[n,c]) with the GroupNorm (+FiLM) coefficients of gn_coef_launch  (ddpm.py:174-185, unet_model.py:21-22)
  const float* pro_ab = nullptr;       // [N][2][C0] (scale, shift) or null
  int pro_act = 0;                     // 0 none, 1 SiLU, 2 ReLU
  // fused GroupNorm statistics of the output (3x3 only): stats[n][stats_G][2] += {sum, sumsq}; zeroed by the caller
  double* stats = nullptr; int stats_G = 0;
};

// host: pack fp32 [taps][Cin][Cout] weights; leaves `ready == false` for unsupported shapes
// `w_1x1` ([Cin][Cout]) / `bias_1x1`: optional 1x1 convolution of the same input fused as a second output (3x3 only)
int conv_tc_pack(const float* w_tap_cin_cout, const float* bias, int Cin, int Cout, int ks, int stride, int pad, ConvTcW* out,
                 const float* w_1x1 = nullptr, const float* bias_1x1 = nullptr);
// nearest x2 up-sampling followed by a 3x3 convolution == four 2x2 convolutions of the low-resolution image, one per output parity:
// output (2i+py, 2j+px) only sees low-resolution rows {i-1, i} (py = 0) or {i, i+1} (py = 1), with the taps that fall on the same
// source pixel pre-summed.  Packed as a 3x3 convolution with 4 * Cout output channels ordered (py, px, c); 2.25x fewer MACs and
// tensor-core tiles with N = 4 * Cout instead of Cout.  `w` is fp32 [9 taps][Cin][Cout] like conv_tc_pack.
int conv_tc_pack_up2(const float* w_tap_cin_cout, const float* bias, int Cin, int Cout, ConvTcW* out);
bool conv_tc_supports(const ConvTcW& w, const ConvTcArgs& a);
// GroupNorm statistics {sum, sumsq}[N][G][2] (+ optional per-image FiLM [2C] scale, shift) -> coefficient table [N][2][C]
int gn_coef_launch(const double* stats, const float* gamma, const float* beta, const float* film, int film_stride, int G, int C, int N,
                   long long HW, float eps, float* ab, cudaStream_t s);
// returns number of kernels launched, < 0 on error
int conv_tc_launch(const ConvTcW& w, const ConvTcArgs& a, cudaStream_t s);

// development aid: device buffer of per-role clock stamps (null unless env LD_CONV_TRACE is set)
long long* conv_tc_trace();

}  // namespace ld
