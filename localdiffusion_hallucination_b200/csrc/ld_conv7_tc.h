// tcgen05 7x7 single-input-channel convolution (init_conv, ddpm.py:319) for sm_100a: see ld_conv7_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ld {

struct Conv7TcW {
  bool ready = false;
  int Cout = 0;
  void* w = nullptr;      // bf16 [16 chunks][Cout][8]: K = [49 taps + pad | 49 taps + pad]
  float* bias = nullptr;  // fp32 [Cout] or null
  void* wt = nullptr;     // bf16 [7 ky][2][256 n][8]: the banded (Toeplitz) form of every filter row (see ld_conv7_tc.cu)
  float bias_h[64] = {};  // host copy: travels in the kernel parameters
};

// host weights fp32 [49 taps][Cout]; leaves ready == false for unsupported widths
int conv7_tc_pack(const float* w_tap_cout, const float* bias, int Cout, Conv7TcW* out);
void conv7_tc_free(Conv7TcW* w);
// x: fp32 [N][H][W] -> out: bf16 [N][H][W][Cout].  Returns kernels launched (1), < 0 on error.
int conv7_tc_launch(const Conv7TcW& w, const float* x, void* out, int N, int H, int W, cudaStream_t s);

}  // namespace ld
