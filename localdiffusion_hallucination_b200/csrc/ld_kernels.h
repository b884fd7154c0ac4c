// Host-side launchers of the hand-written kernels (implemented in ld_kernels_*.cu / ld_conv_tc.cu).
// `bf` selects the storage type of activation tensors: false = fp32, true = bf16.  Math is fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ld_common.cuh"

namespace ld {

// every launcher returns the number of kernels it enqueued (for ld_launch_count)
int launch_conv_simt(const ConvP& p, bool bf, cudaStream_t s);
// Cin == 1 direct convolution, fp32 input image [N,H,W] -> T [N,H,W,Cout]; w is [ks*ks][Cout]
int launch_conv_c1(const float* in, const float* w, const float* bias, void* out, int N, int H, int W,
                   int Cout, int ks, bool bf, cudaStream_t s);
// 1x1 convolution to a single fp32 channel (final_conv, ddpm.py:398): out[p] = b + sum_c w[c] x[p,c]
int launch_conv_cout1(const void* in, const float* w, const float* bias, float* out, long long P, int C,
                      bool bf, cudaStream_t s);
// GroupNorm statistics: sums[n][g] = {sum, sumsq} (double), buffer must be zeroed by the caller
int launch_gn_stats(const void* x, double* sums, int N, int HW, int C, int G, bool bf, cudaStream_t s);
int launch_gn_apply(const GnApplyP& p, bool bf, cudaStream_t s);
// whether launch_gn_apply can fold a 1x1 convolution to one channel into its output pass (GnApplyP::dot_out)
bool gn_apply_can_dot(int C, bool bf);
// RMSNorm over C per pixel (ddpm.py:131-132): out = x / max(|x|,1e-12) * g * sqrt(C) (+ res)
int launch_rmsnorm(const void* x, const float* g, const void* res, void* out, long long P, int C, bool bf,
                   cudaStream_t s);
int launch_maxpool2(const void* x, void* out, int N, int H, int W, int C, bool bf, cudaStream_t s);

// LinearAttention (ddpm.py:234-251) on a qkv tensor [N,HW,3*heads*32]:
//   ctx/ksum accumulate exp(k - kmax)-weighted sums; fold builds the per-image matrix
//   Mn[n][j][c] = sum_e Wout[c][h*32+e] * ctx[n][h][d][e] / ksum[n][h][d] * 32^-0.5,  j = h*32+d
//   and la_out applies softmax_d(q) @ Mn + bias -> RMSNorm(g2) -> + x.
struct LinAttnP {
  const void* qkv; int N, HW, heads, C;
  float* kmax_part; int chunks;   // [N][chunks][hid]
  float* kmax;                    // [N][hid]
  float* ctx;                     // [N][heads][32][32], zeroed by caller
  float* ksum;                    // [N][hid], zeroed by caller
  const float* wout;              // [hid][C]  (1x1 conv packing, tap-major)
  const float* bout;              // [C]
  float* Mn;                      // [N][hid][C]
  const float* g2;                // to_out.1.g [C]
  const void* x;                  // residual input [N,HW,C]
  void* out;                      // [N,HW,C]
};
int launch_linear_attention(const LinAttnP& p, bool bf, cudaStream_t s);

// Full softmax attention (attend.py:98-113) on qkv [N,n,3*heads*32] -> out [N,n,heads*32]
int launch_attention_simt(const void* qkv, void* out, int N, int n, int heads, bool bf, cudaStream_t s);

// time embedding (ddpm.py:142-149, 339-344) and all per-block FiLM vectors (ddpm.py:191-194, 204)
struct TimeP {
  const int64_t* t;   // device [N] (per-sample timesteps) ...
  const int* t_scalar; // ... or one device scalar shared by the whole batch (sampler loop)
  float neg_step;     // -ln(theta)/(dim/2-1), computed in fp64 on the host like ddpm.py:145
  int N, dim;         // dim = Unet dim (sinusoidal width), time_dim = 4*dim
  float theta;
  const float* w1; const float* b1;  // [4dim][dim]
  const float* w2; const float* b2;  // [4dim][4dim]
  float* st;          // [N][4dim]  = SiLU(time_mlp(t))
  const float* wf; const float* bf_; // concatenated block MLPs [total][4dim], [total]
  int total;
  float* film;        // [N][total]
};
int launch_time_film(const TimeP& p, cudaStream_t s);
// film[0][:] = table[*t][:] (the sampler's per-timestep FiLM rows are precomputed once: they depend on t only)
int launch_film_gather(const float* table, int total, const int* t_scalar, float* film, cudaStream_t s);

// sampler elementwise kernels (ddpm.py:672-690, 697-708, 775-810, 852-858)
struct PrepP {
  const float* cond; const float* mask; float* bm; float* cond_out; float* cond_in;
  float floor; long long n; unsigned int* counters;  // counters[0] = #(bm==1), [1] = #(bm==0)
};
int launch_prep_cond(const PrepP& p, cudaStream_t s);

struct StepP {
  int kind;                 // 0 branched, 1 fusion, 2 single
  const float* o_out; const float* o_in;   // raw UNet outputs (o_out may be null when ood_uses_cond)
  float* x_out; float* x_in;               // states, updated in place (fusion/single write x_out)
  float* x0_out; float* x0_in;             // optional x0 records (may be null)
  const float* bm; const float* cond_out; const float* z;  // z null when t == 0
  int* t_ptr;               // device scalar: current timestep (graph-replay friendly); decremented by the kernel when `ticket` is set
  const float* coef1; const float* coef2; const float* sigma;  // [T]
  int mask_x, ood_uses_cond;
  float lo, hi;
  long long n;              // B*H*W
  long long z_stride;       // elements between successive draws of the tape; z = tape + (Tloop-1-t)*stride
  int tloop;                // loop length (index of draw for step t is tloop - t, draw 0 is x_T)
  unsigned int* counters;   // [2]=#(x_out*m==0), [3]=#(x_in*(1-m)==0) at the fusion step
  float* x0_trace; long long trace_stride;  // optional [tloop][2][n]
  unsigned int* ticket;     // optional (zero-initialised, self re-arming): the last block to finish writes *t_ptr = t - 1 (ddpm.py:951)  // single-trajectory steps with objective pred_noise / pred_v (ddpm.py:731-737, 757-761): x0 = ca[t] * x_t - cb[t] * model_output
  // (predict_start_from_noise / predict_start_from_v, ddpm.py:631-653); null for pred_x0 (x0 = model_output)
  const float* ca; const float* cb;
};
int launch_step(const StepP& p, cudaStream_t s);

// DDIM update of the branch sampler (ddpm.py:979-1075).  Step i of the host-built schedule: time = times[i],
// coefs[i] = {sqrt_recip_alphas_cumprod[time], sqrt_recipm1_alphas_cumprod[time], sqrt(alpha_next), c, sigma}; the last
// step (time_next < 0) returns x_start itself.
struct DdimP {
  int kind;                 // 0 branched, 1 fusion (ddpm.py:1022-1043), 2 single
  const float* o_out; const float* o_in;
  float* x_out; float* x_in;
  const float* bm; const float* cond_out;
  const float* z;           // noise tape: draw 0 is x_T, step i uses draw 1 + i (none for the last step)
  long long z_stride;
  int* idx_ptr;             // device scalar: current step index (graph-replay friendly); advanced by the kernel when `ticket` is set
  int nsteps;
  const float* coefs;       // device [nsteps][5]
  int mask_x, ood_uses_cond;
  float lo, hi;
  long long n;
  unsigned int* counters;   // [2]=#(eps_out*m==0), [3]=#(eps_in*(1-m)==0) at the fusion step
  unsigned int* ticket;     // optional: the last block to finish advances idx and loads t = times[idx] (ddpm.py:996-998)
  const int* times; int* t_ptr;   // device [nsteps] schedule of `time` values and the scalar the UNet plans read
  const float* ca; const float* cb;   // [T] tables indexed by time, see StepP (single-trajectory steps only)
};
int launch_ddim_step(const DdimP& p, cudaStream_t s);

// ---- stages in front of the sampler (ld_producers.cu) --------------------------------------------------------------------
int launch_mnist_cond(const float* raw, float* hr, float* cond, int N, int S, cudaStream_t s);
int launch_mri_norm(const float* raw, float* out, unsigned int* mins, int N, int Hs, int Ws, int crop, float mean, float stdv, int translate_zero,
                    cudaStream_t s);
size_t mask_scratch_bytes(int B, int S);
int launch_mask_from_anomaly(const float* amap, int B, int h, int w, int S, int rule, int manual_cols, float* mask_pred, float* binary, void* scratch,
                             cudaStream_t s);

// PatchCore nearest-neighbour search on tcgen05 (ld_knn_tc.cu): x [M][D], bank [Nb][D] fp32 -> score [M], loc [M]
size_t knn_scratch_bytes(int M, int Nb, int D);
int knn_tc_launch(const float* x, const float* bank, int M, int Nb, int D, float* score, long long* loc, void* scratch, cudaStream_t s);

int launch_nhwc_to_nchw_f32(const void* in, float* out, int N, int HW, int C, bool bf, cudaStream_t s);
int launch_convert(const void* in, bool in_bf, void* out, bool out_bf, long long n, cudaStream_t s);
int launch_copy_f32(const float* in, float* out, long long n, cudaStream_t s);

}  // namespace ld
