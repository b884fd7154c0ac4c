// Fused tcgen05 LinearAttention (ddpm.py:214-251) for sm_100a: see ld_linattn_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ld {

struct LinAttnTcW {
  bool ready = false;
  int C = 0, heads = 4;
  void* wq = nullptr;    // bf16 [heads/4][C/8][128][8]: the q rows of to_qkv per group of four heads, RMSNorm g*sqrt(C) folded in
  void* wk = nullptr;    // bf16 [heads/4][C/8 + 2][128][8]: the k rows (+ the soft-max shift in the ones-channel chunk)
  float* kb2 = nullptr;  // [heads*32] log2(e) * upper bound of |k_d| (soft-max shift)
  float* Ut = nullptr;   // fp32 [heads][C][C]: (to_out.0 o v-projection) per head, transposed
  float* bout = nullptr; // fp32 [C]
  float* g2 = nullptr;   // fp32 [C] to_out.1.g
  int q_use_max = 1;     // 0 when |q * log2 e| is provably < 60 for every (h,d): no max pass in the soft-max over d
};

struct LinAttnTcArgs {
  const void* x = nullptr;   // bf16 [N][HW][C] (input of the attention block AND its residual)
  void* out = nullptr;       // bf16 [N][HW][C]
  int N = 0, HW = 0;
  float* Z = nullptr;        // [N][heads*32][C] fp32, zeroed by the caller: sum_p exp(k[p,(h,d)]) * xhat[p,c]
  float* ksum = nullptr;     // [N][heads*32] fp32, zeroed by the caller
  void* Mn = nullptr;        // [N][heads*32*C] bf16 scratch
  unsigned int* flag = nullptr;  // optional: incremented when a soft-max row sum underflowed
};

// host weights in torch layout: wqkv [384][C], g [C], wout [C][128], bout [C], g2 [C]
int linattn_tc_pack(const float* wqkv, const float* g, const float* wout, const float* bout, const float* g2, int C, int heads,
                    LinAttnTcW* out);
void linattn_tc_free(LinAttnTcW* w);
// returns the number of kernels launched (3), < 0 on error
int linattn_tc_launch(const LinAttnTcW& w, const LinAttnTcArgs& a, cudaStream_t s);

}  // namespace ld
