// tcgen05 flash-style full attention (attend.py:98-113) for sm_100a: see ld_attn_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ld {

// bytes of scratch (operand images written by the prep kernel) for N images of n tokens
size_t attn_tc_scratch_bytes(int N, int n, int heads);
// opt in to the kernel's dynamic shared memory (call once outside any stream capture)
int attn_tc_configure();
// qkv: bf16 [N][n][3*heads*32] -> out: bf16 [N][n][heads*32].  Returns kernels launched (2), < 0 on error.
int attn_tc_launch(const void* qkv, void* out, void* scratch, int N, int n, int heads, cudaStream_t s);

}  // namespace ld
