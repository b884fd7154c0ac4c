/*
 * ld_sampler.h -- C ABI of the B200-native LocalDiffusion sampler (libld_sampler.so).
 *
 * The reference (edshkim98/LocalDiffusion-Hallucination) has no FFI: its boundary for this
 * path is two Python nn.Module APIs, `Unet.forward` (ddpm.py:404) and
 * `GaussianDiffusion.sample` (ddpm.py:1078).  This header is the C-ABI those two calls are
 * re-hosted on; the Python shims in `localdiffusion_hallucination_b200/` bind it with ctypes
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success or a negative ld_status; nothing throws across the
 *     ABI; `ld_last_error()` returns a thread-local, NUL-terminated description.
 *   - tensors are plain pointers + sizes.  Images are dense NCHW fp32 with C == 1 (which is
 *     byte-identical to NHWC), exactly what the reference passes around (ddpm.py:1121-1125).
 *   - device pointers are borrowed for the duration of the call; device work is ordered after the
 *     work already enqueued on the caller's `stream` (a cudaStream_t passed as void*), and the
 *     stream is made to wait for the results; the sampler entry points run their loop on an
 *     engine-owned stream joined to `stream` by events (see ld_sample for host synchronisation).
 *   - one handle per device; a handle is not thread-safe (distinct handles may be used from
 *     distinct threads concurrently).
 *   - there is no CPU fallback: every compute entry point fails with LD_ERR_NO_DEVICE when no
 *     sm_100 device is usable.
 */
#ifndef LD_SAMPLER_H_
#define LD_SAMPLER_H_

#include <stdint.h>

#if defined(__GNUC__)
#define LD_API __attribute__((visibility("default")))
#else
#define LD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ld_handle ld_handle;

typedef enum ld_status {
  LD_OK = 0,
  LD_ERR_INVALID = -1,   /* bad argument / shape contract (ddpm.py:405, 515-516, 536, 561) */
  LD_ERR_NO_DEVICE = -2, /* no usable sm_100 CUDA device */
  LD_ERR_CUDA = -3,      /* a CUDA call failed; see ld_last_error() */
  LD_ERR_STATE = -4,     /* call order violated (weights missing, schedule missing, ...) */
  LD_ERR_KEY = -5,       /* unknown / duplicate / wrongly-shaped state_dict key */
  LD_ERR_MASK = -6       /* "mask should be binary" (ddpm.py:698) or fusion sanity (ddpm.py:790) */
} ld_status;

enum { LD_MAX_LEVELS = 8 };

/* Arithmetic used for the dense contractions.  Statistics, residual stream of the sampler
 * (x_t, x0, posterior) and accumulators are fp32 in both modes. */
typedef enum ld_precision {
  LD_PREC_FP32 = 0, /* fp32 storage + fp32 CUDA-core FMA: the parity path (rel err <= 1e-3) */
  LD_PREC_BF16 = 1  /* bf16 storage + tcgen05 (fp32 accumulate in TMEM): the fast path      */
} ld_precision;

/* Conditional-encoder depth, `ResUnet(data=mode)` (unet_model.py:91-137). */
typedef enum ld_cond_mode {
  LD_COND_MRI = 0,  /* 4 blocks, 3 max-pools, 256 output channels at S/8 ('mri','mvtec','mvtecGray') */
  LD_COND_MNIST = 1 /* 3 blocks, 2 max-pools, 128 output channels at S/4 ('mnist','mvtecSR')         */
} ld_cond_mode;

/* Mirrors the constructor arguments of `Unet` that reach the sampling path (ddpm.py:287-307). */
typedef struct ld_model_desc {
  int32_t dim;                       /* ddpm.py:289 */
  int32_t init_dim;                  /* ddpm.py:290 (resolved, never 0) */
  int32_t n_levels;                  /* len(dim_mults) */
  int32_t dim_mults[LD_MAX_LEVELS];  /* ddpm.py:292 */
  int32_t full_attn[LD_MAX_LEVELS];  /* ddpm.py:304, 0/1 per level */
  int32_t channels;                  /* ddpm.py:293; must be 1 on this path */
  int32_t resnet_groups;             /* ddpm.py:296 */
  int32_t attn_heads;                /* ddpm.py:303 */
  int32_t attn_dim_head;             /* ddpm.py:302; must be 32 */
  float sinusoidal_theta;            /* ddpm.py:301 */
  int32_t cond_mode;                 /* ld_cond_mode */
  int32_t precision;                 /* ld_precision */
} ld_model_desc;

/* Flags of one sampling call, i.e. the state of the reference's `config` dict *after* the
 * per-call fix-ups of `GaussianDiffusion.sample` (ddpm.py:1093-1117), which stay on the host. */
typedef struct ld_sample_desc {
  int32_t batch;            /* B */
  int32_t height, width;    /* image_size (square in the reference, ddpm.py:1121) */
  int32_t num_timesteps;    /* loop length: T, or use_gt_timestep (ddpm.py:944) */
  int32_t branch_out;       /* 1: two-trajectory mode (ddpm.py:671) */
  int32_t start_intermediate; /* 1: fuse at t <= start_timestep (ddpm.py:779) */
  int32_t start_timestep;   /* config['start_timestep'] */
  int32_t mask_x;           /* 1: masked fill of the OOD x0 (ddpm.py:697-703) */
  int32_t ood_uses_cond;    /* 1: non-MRI data, OOD x0 := cond_out (ddpm.py:704-708) */
  float cond_in_floor;      /* 0.95, or 0.5 for data=='mnist' (ddpm.py:683-686) */
  float min_val, max_val;   /* min_max_val[0], [1] (ddpm.py:775-776) */
  int32_t return_pair;      /* 1: output is [2,B,1,H,W] (ddpm.py:965-970) */
  int32_t record_x0;        /* 1: also write every step's x0 to `x0_trace` (return_all_outputs) */
} ld_sample_desc;

LD_API const char* ld_last_error(void);
/* Build id, e.g. "ld_sampler 0.1 sm_100a". */
LD_API const char* ld_version(void);
/* Number of usable sm_100 devices (0 when there is none or no driver). */
LD_API int ld_device_count(void);

/* --- lifetime ----------------------------------------------------------------------------- */
LD_API int ld_create(const ld_model_desc* desc, int device, ld_handle** out);
LD_API int ld_destroy(ld_handle* h);

/* --- weights: replaces `load_state_dict` on the reference `Unet` (ddpm.py:1513-1521) -------
 * `key` is the reference state_dict key without the `model.` prefix,
 * `data` is host fp32 in the reference's layout (conv: [Cout,Cin,kh,kw]).  Every key of the
 * model must be loaded exactly once before `ld_finalize_weights`; the dead
 * `conv_fusion.mlp.1.{weight,bias}` are accepted and ignored (ddpm.py:436). */
LD_API int ld_num_weights(const ld_handle* h);
LD_API int ld_weight_info(const ld_handle* h, int index, const char** key, int64_t shape[4], int* ndim);
LD_API int ld_load_weight(ld_handle* h, const char* key, const float* data, const int64_t* shape, int ndim);
LD_API int ld_finalize_weights(ld_handle* h);

/* --- schedule: the three `[T]` buffers the DDPM update reads (ddpm.py:591-593, 659-666) -----
 * `sigma` (optional) is exp(0.5*posterior_log_variance_clipped) as the host computed it
 * (ddpm.py:853); when NULL it is derived from the log-variance with expf. */
LD_API int ld_set_schedule(ld_handle* h, int T, const float* posterior_mean_coef1,
                    const float* posterior_mean_coef2, const float* posterior_log_variance_clipped,
                    const float* sigma);

/* Objective of the denoiser (ddpm.py:534-536, 731-761).  Default pred_x0 (the shipped config.yaml:45): x0 = model output.  For
 * pred_noise pass (sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod), for pred_v (sqrt_alphas_cumprod,
 * sqrt_one_minus_alphas_cumprod): x0 = a[t] * x_t - b[t] * output (ddpm.py:631-653); a == NULL switches back to pred_x0.
 * Single-trajectory sampling only: ld_sample / ld_sample_ddim fail with LD_ERR_INVALID for sd->branch_out with these objectives,
 * where the reference dies with UnboundLocalError. */
LD_API int ld_set_objective(ld_handle* h, int T, const float* a, const float* b);

/* --- `Unet.forward(x, cond_img, time)` (ddpm.py:404-451) -------------------------------------
 * x, cond, out: device fp32 [N,1,H,W]; t: device int64 [N]. */
LD_API int ld_unet_forward(ld_handle* h, const float* x, const float* cond, const int64_t* t, float* out,
                    int N, int H, int W, void* stream);
/* `ResUnet.forward` (unet_model.py:122-137): cond [N,1,H,W] -> feat fp32 NCHW [N,Cf,H/f,W/f]. */
LD_API int ld_cond_encode(ld_handle* h, const float* cond, float* feat, int N, int H, int W, void* stream);

/* --- `GaussianDiffusion.p_sample_loop` (ddpm.py:930-977) --------------------------------------
 * cond, mask: device fp32 [B,1,H,W] (mask is the *soft* map; >= 1.0 means OOD, ddpm.py:672).
 * noise: device fp32 [num_timesteps, B,1,H,W]; noise[0] is x_T (already q_sample'd by the host
 *        when use_gt), noise[1+i] is the draw of loop iteration i (t = T-1-i), none for t == 0.
 * out:   device fp32 [B,1,H,W] (or [2,B,1,H,W] when return_pair).
 * x0_trace: optional device fp32 [num_timesteps, 2, B,1,H,W]: slot 0 = x0 of the step (OOD branch while branched), slot 1 = x0 of
 *        the IND branch while branched, the UPDATED image x_{t-1} on single-trajectory steps (return_all_timesteps, ddpm.py:964). */
LD_API int ld_sample(ld_handle* h, const ld_sample_desc* sd, const float* cond, const float* mask,
              const float* noise, float* out, float* x0_trace, void* stream);
/* Host synchronisation: the loop itself never touches the host (t lives on the device, one CUDA graph per timestep).  By default
 * ld_sample waits ONCE, at the end, for its internal stream, because the reference's asserts are synchronous: "mask should be
 * binary" (ddpm.py:698) and "x_out and x_in should be masked" (ddpm.py:790) come back as LD_ERR_MASK from this very call.  With
 * option "async" = 1 nothing waits: the call returns after enqueueing, `stream` is made to wait for the result, and
 * ld_sample_finish(h) delivers the deferred status (it must be called before the next ld_sample on the handle). */
LD_API int ld_sample_finish(ld_handle* h);

/* --- `GaussianDiffusion.ddim_sample` (ddpm.py:979-1075): the DDIM variant of the branch sampler -----------------
 * times: host int32 [nsteps], the `time` of every step (ddpm.py:984-986 without the trailing -1).
 * coefs: host fp32 [nsteps][5] = {sqrt_recip_alphas_cumprod[time], sqrt_recipm1_alphas_cumprod[time], sqrt(alpha_next), c, sigma}
 *        (ddpm.py:637-641, 1013-1018); the last three are unused on the last step, which returns x_start.
 * noise: device fp32 [nsteps, B,1,H,W]; noise[0] is x_T, noise[1+i] the draw of step i (none for the last step).
 * fuse_step: index of the step that composites the branches (time <= start_timestep_ddim, ddpm.py:1022), -1 = never;
 *        a fuse_step equal to the last step never fuses, exactly like the reference (its `continue` comes first).
 * out: [B,1,H,W], or [2,B,1,H,W] when sd->return_pair (never fused: the reference returns the list [out, in]).
 * sd->num_timesteps, start_timestep, start_intermediate and record_x0 are ignored here. */
LD_API int ld_sample_ddim(ld_handle* h, const ld_sample_desc* sd, const float* cond, const float* mask, const float* noise,
                          float* out, const int32_t* times, const float* coefs, int nsteps, int fuse_step, void* stream);

/* --- one DDPM update on caller-owned state (ddpm.py:841-860 + 768-838), for parity tests -----
 * kind 0: branched step; kind 1: fusion step (composite, then single update); kind 2: single.
 * x_out/x_in/x0_out/x0_in/z/mask: device fp32 [n]; raw UNet outputs come in through x0_*, the
 * clamped (and, for kind 1, fused) x0 goes back out through them; x_* are updated in place. */
LD_API int ld_posterior_step(ld_handle* h, int kind, int t, float* x_out, float* x_in, float* x0_out,
                      float* x0_in, const float* cond, const float* mask, const float* z,
                      const ld_sample_desc* sd, int64_t n, void* stream);

/* --- the two stages in front of the sampler: conditional-image producers and anomaly map -> masks -------------
 * All pointers are device fp32; work is enqueued on `stream`; nothing synchronises. */
/* `MNIST.__getitem__` (data.py:814-836): raw [N,S,S] in 0..255 -> hr = 2*(x/255) and cond = 2*(up(x[::2 rows])/255), both [N,1,S,S]
 * (the reference's slice sub-samples the rows only, data.py:822-826; bilinear, align_corners=False). */
LD_API int ld_prep_mnist(const float* raw, float* hr, float* cond, int N, int S, void* stream);
/* `MedDataset_png` (data.py:380-414): centre crop of raw [N,Hs,Ws] to crop x crop, (x - mean)/std, + |min| per image when
 * translate_zero -> out [N,1,crop,crop].  `scratch`: N * 4 bytes of device memory (unused without translate_zero). */
LD_API int ld_prep_mri(const float* raw, float* out, void* scratch, int N, int Hs, int Ws, int crop, float mean, float std,
                       int translate_zero, void* stream);
/* Per-dataset threshold rules of test.py:259-375. */
typedef enum ld_mask_rule {
  LD_MASK_MNIST_8TO3 = 0, LD_MASK_MNIST_8TO5 = 1, LD_MASK_MRI_T12FLAIR = 2, LD_MASK_MRI_FLAIR2T1 = 3,
  LD_MASK_MVTEC_TRANSISTOR = 4, LD_MASK_MVTEC_TOOTHBRUSH = 5, LD_MASK_MVTEC_GRID = 6
} ld_mask_rule;
/* test.py:237-381: anomaly map [B,1,h,w] (statistics over the whole batch, as the reference computes them; it runs B = 1) ->
 * mask_pred [B,1,S,S] (soft, exactly 1.0 where the map reaches the threshold) and binary_mask (may be NULL).  h,w != S: bilinear
 * resize first (test.py:246-247).  manual_cols > 0: the manual left-columns mask of test.py:379-381 replaces the detector's.
 * `scratch`: ld_mask_scratch_bytes(B, S) bytes of device memory. */
LD_API int64_t ld_mask_scratch_bytes(int B, int S);
LD_API int ld_mask_from_anomaly(const float* amap, int B, int h, int w, int S, int rule, int manual_cols, float* mask_pred,
                                float* binary_mask, void* scratch, void* stream);

/* PatchCore nearest-neighbour search (models.py:179-217, `euclidean_dist` + `nearest_neighbors(n_neighbors=1)`) on tcgen05:
 * embedding [M][D], memory_bank [Nb][D] (device fp32) -> patch_scores [M] = min_j sqrt(clamp(|x|^2 - 2 x.y + |y|^2, 0)) and
 * locations [M] (int64, lowest index among exact ties).  Operands are split into bf16 hi + lo parts (three MMAs per k-step), so the
 * distances are fp32-accurate.  `scratch`: ld_knn_scratch_bytes(M, Nb, D) bytes of device memory. */
LD_API int64_t ld_knn_scratch_bytes(int M, int Nb, int D);
LD_API int ld_knn_min(const float* embedding, const float* memory_bank, int M, int Nb, int D, float* patch_scores, int64_t* locations,
                      void* scratch, void* stream);

/* --- introspection for bench.py -------------------------------------------------------------- */
/* Kernel launches enqueued by this handle since creation (our own kernels only). */
LD_API int64_t ld_launch_count(const ld_handle* h);
/* Device workspace currently owned by the handle, bytes. */
LD_API int64_t ld_workspace_bytes(const ld_handle* h);
/* Tunables (cached plans are dropped, the new value takes effect on the next call):
 *   "use_graph" (default 1)  replay one CUDA graph per timestep instead of launching the ~100 kernels of a step one by one;
 *   "async"     (default 0)  ld_sample / ld_sample_ddim return without synchronising: all work is enqueued, `stream` waits on it,
 *                            and the reference's asserts (ddpm.py:698, 790) are reported by ld_sample_finish instead;
 *   "la_exact"  (default 0)  LinearAttention through the exact-max kernels (the engine switches to them by itself when the fused
 *                            kernels' analytic soft-max shift underflows, see ld_sample);
 *   "up2"       (default 1)  nearest x2 up-sampling folded into the conv filter; 0 = replicate-on-load kernel;
 *   "pdl"       (default 0)  programmatic dependent launch: 1 between all kernels of a timestep (measured slower), 2 only for the
 *                            launches that follow a tiny kernel (measured equal within noise);
 *   "attn_simt", "debug_keep", "use_tc" (before ld_finalize_weights): test aids. */
LD_API int ld_set_option(ld_handle* h, const char* name, int64_t value);
/* Current value of a tunable ("la_exact" reads 1 once the engine has switched LinearAttention to the exact-max kernels). */
LD_API int ld_get_option(const ld_handle* h, const char* name, int64_t* value);

/* --- test hooks (used by tests/ only) ---------------------------------------------------------
 * Named intermediate activations of the last `ld_unet_forward` (requires option "debug_keep"=1
 * before the first forward): dims = {N, C, H, W}; fetch converts to NCHW fp32. */
LD_API int ld_debug_num_taps(ld_handle* h);
LD_API int ld_debug_tap_info(ld_handle* h, int index, const char** name, int32_t dims[4]);
LD_API int ld_debug_tap_fetch(ld_handle* h, int index, float* out_nchw, void* stream);
/* One convolution through one kernel: kernel 0 = CUDA-core fp32, 1 = CUDA-core with bf16 storage,
 * 2 = tcgen05, 3 = tcgen05 with the nearest x2 up-sampling folded into the filter (up == 1, 3x3 only).
 * x0/x1/res/out: device fp32 NHWC; w_host: host fp32 [Cout, C0+C1, ks, ks]. */
LD_API int ld_debug_conv(int kernel, const float* x0, int C0, const float* x1, int C1, int N, int Hin,
                         int Win, int up, int H, int W, const float* w_host, const float* bias_host,
                         int Cout, int ks, const float* res, float* out, void* stream);
/* Test hook: 3x3 tcgen05 convolution with the fused GroupNorm prologue (normalise-on-load of the source, ddpm.py:174-185)
 * and the fused GroupNorm statistics of its output.  x0/out: fp32 NHWC device; w/bias: host, torch layout. */
LD_API int ld_debug_conv_fused(const float* x0, int C0, int N, int H, int W, const float* w_host, const float* bias_host,
                               int Cout, const double* pro_stats, const float* pro_gamma, const float* pro_beta,
                               const float* pro_film, int pro_film_stride, int pro_G, int pro_act, double* stats_out,
                               int stats_G, float* out, void* stream);
/* Test hook: ResnetBlock block1.proj (3x3, GroupNorm statistics of its output) and res_conv (1x1) of the same virtual
 * concat [x0 | x1] in one tcgen05 launch with two accumulators (ddpm.py:207,212).  x0/x1/out/out2: fp32 NHWC device;
 * w3 [Cout][C0+C1][3][3], w1 [Cout][C0+C1], biases: host fp32. */
LD_API int ld_debug_conv_dual(const float* x0, int C0, const float* x1, int C1, int N, int H, int W, const float* w3_host,
                              const float* b3_host, const float* w1_host, const float* b1_host, int Cout, double* stats_out,
                              int stats_G, float* out, float* out2, void* stream);
/* Test hook: the fused tcgen05 LinearAttention block, attn(x) + x (ddpm.py:214-251, 425).  x/out: fp32 NHWC device;
 * wqkv [384][C], g [C], wout [C][128], bout [C], g2 [C]: host fp32 in the reference's parameter layout. */
LD_API int ld_debug_linattn(const float* x, int C, int N, int HW, const float* wqkv, const float* g, const float* wout,
                            const float* bout, const float* g2, float* out, void* stream);
/* Same with `heads` in {4, 8} (wqkv [3*heads*32][C], wout [C][heads*32]): 8 heads run as two groups of four (C <= 64). */
LD_API int ld_debug_linattn_h(const float* x, int C, int N, int HW, int heads, const float* wqkv, const float* g, const float* wout,
                              const float* bout, const float* g2, float* out, void* stream);
/* Test hook: tcgen05 flash-style soft-max attention (attend.py:98-113).  qkv: fp32 [N][n][3*heads*32] device
 * (channel = part*hid + h*32 + d, ddpm.py:276-277); out: fp32 [N][n][heads*32]. */
LD_API int ld_debug_attention(const float* qkv, int N, int n, int heads, float* out, void* stream);
/* Test hook: tcgen05 7x7 convolution of a single fp32 channel (init_conv, ddpm.py:319).  x: fp32 [N][H][W] device;
 * w_host [Cout][1][7][7], bias_host [Cout]: host; out: fp32 [N][H][W][Cout] device. */
LD_API int ld_debug_conv7(const float* x, int N, int H, int W, const float* w_host, const float* bias_host, int Cout,
                          float* out, void* stream);
/* Average device time (ms, CUDA events on `stream`) of `iters` launches of one convolution kernel on
 * synthetic operands; used by bench.py for the roofline of the dominant kernel. */
LD_API int ld_debug_conv_time(int kernel, int C0, int C1, int N, int H, int W, int up, int Cout, int ks,
                              int iters, float* ms_out, void* stream);

/* Same for the dominant 3x3 tcgen05 convolution in the variants the sampler launches it in (ddpm.py:173-212):
 * variant 0 plain; 1 + fused GroupNorm statistics of the output; 2 + normalise-on-load prologue of the source and statistics;
 * 3 dual: 3x3 + 1x1 res_conv of the same virtual concat [C0 | C1] with statistics (two outputs). */
LD_API int ld_debug_conv_variant_time(int variant, int C0, int C1, int N, int H, int W, int Cout, int iters, float* ms_out,
                                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LD_SAMPLER_H_ */
