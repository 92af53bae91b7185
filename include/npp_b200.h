/*
 * npp_b200.h -- C ABI of libnpp_b200.so, the B200 (sm_100a) implementation of NPP-Net's
 * per-image training hot path:
 *
 *   periodicity-aware positional encoding -> coordinate MLP forward -> masked-MSE loss
 *   -> backward -> Adam
 *
 * The reference (ArmastusChen/Learning-Continuous-Implicit-Representation-for-Near-Periodic-
 * Patterns) has no FFI: its boundary for this path is the Python surface of models/.  Each entry
 * point below names the reference code it replaces (paths relative to the reference root).
 * The Python binding is learning-..._b200/_native.py (ctypes); INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; npp_last_error() returns the
 *     message of the last failure on the calling thread.  Nothing throws across the ABI.
 *   - all pointers marked "device" are CUDA device pointers owned by the caller (torch tensors);
 *     the plan owns only its activation workspace, fp16 shadow weights and TMA descriptors.
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it.
 *   - one plan per (process, device); one process per GPU.  No CPU fallback exists.
 */
#ifndef NPP_B200_H_
#define NPP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct NppPlan NppPlan;

enum { NPP_MODEL_TOPK = 0, /* models/networks.py:8-95   NPP_Net      (p_topk > 1) */
       NPP_MODEL_TOP1 = 1, /* models/networks.py:99-173 NPP_Net_top1 (p_topk == 1) */
       NPP_MODEL_LIGHT = 2 /* models/networks.py:176-263 NPP_Net_light with len(freq_scales) == 1: the search-stage
                              fit of NPP_proposal/search.py:85-148 (create_npp_net(is_search=True), helpers.py:91-103).
                              Encoders in search mode: Embedder_periodic without raw input and without the Fourier
                              expansion (embedder.py:93-95,140-148; 4*n_aug columns) and the 2-D Embedder on
                              normalised coordinates (embedder.py:51-56,76-80; 2 + 4*n_freq columns). */ };

enum { NPP_ACT_SNAKE = 0, /* SnakeActivation, models/activations.py:29-35 */
       NPP_ACT_RELU = 1 };

typedef struct NppConfig {
  int32_t model;          /* NPP_MODEL_* */
  int32_t topk;           /* number of periodicity proposals K (create_npp_net, models/helpers.py:108-116) */
  int32_t depth;          /* netdepth D (options/arg_config.py:55-74), default 8 */
  int32_t width;          /* netwidth W: 512 (options/arg_config.py:57) or 256 (the constructors' default and the
                             search default, options/arg_config.py:116) */
  int32_t skip_layer;     /* skips=[4] (models/helpers.py:90); -1 = none */
  int32_t n_aug;          /* len(freq_scales)*len(freq_offsets)*len(angle_offsets), embedder.py:117-120 */
  int32_t n_freq;         /* multires, number of Gaussian Fourier frequencies (embedder.py:25-26) */
  int32_t include_input;  /* 1 outside search mode (embedder.py:105-109); must be 0 for NPP_MODEL_LIGHT */
  int32_t res_h, res_w;   /* image resolution res=(H,W) (NPP_completion/train.py:64) */
  int32_t wgrad_splits;   /* split-K factor of the weight-gradient GEMM, 0 = auto */
  int32_t activation;     /* NPP_ACT_SNAKE (activation == 'snake', the default of options/arg_config.py:29) or
                             NPP_ACT_RELU (any other value: F.relu, models/networks.py:51-54,66-69) */
  int64_t max_rows;       /* workspace capacity in coordinate rows per step */
  const float* cos_t;     /* host [topk][2][n_aug]  cos(deg2rad(angle+offset))   embedder.py:123-124 */
  const float* sin_t;     /* host [topk][2][n_aug]  sin(...)                                         */
  const float* period;    /* host [topk][2][n_aug]  (period + freq_offset) * freq_scale  embedder.py:121 */
  const float* freq;      /* host [n_freq]          torch.normal(0,1,(n,1))*10            embedder.py:26 */
} NppConfig;

typedef struct NppTensorInfo {
  char name[64];     /* state_dict key of the reference module, e.g. "periodic_linears.0.weight" */
  int64_t offset;    /* float offset inside the parameter arena */
  int32_t rows;      /* out_features (or 1 for a bias) */
  int32_t cols;      /* in_features  (or out_features for a bias) */
  int32_t is_bias;
  int32_t trained;   /* 0 for parameters the reference allocates but never uses (alpha_linear, ...) */
} NppTensorInfo;

const char* npp_last_error(void);
int npp_abi_version(void);

/* Replaces create_npp_net's model construction (models/helpers.py:121-132) for the fused path. */
int npp_plan_create(const NppConfig* cfg, NppPlan** out);
int npp_plan_destroy(NppPlan* plan);

/* Parameter arena layout: trained tensors first (Adam runs flat over [0, trained_floats)). */
int npp_plan_arena_floats(const NppPlan* plan, int64_t* total_floats, int64_t* trained_floats);
int npp_plan_tensor_count(const NppPlan* plan);
int npp_plan_tensor_info(const NppPlan* plan, int index, NppTensorInfo* info);
int npp_plan_encoding_width(const NppPlan* plan); /* K * B * (1+2*n_freq), 1386 at the defaults;
                                                     NPP_MODEL_LIGHT: (2+4*n_freq) + 4*n_aug = 62 */

/* Bind caller-owned device arenas (fp32, `total_floats` each; grads/exp_avg/exp_avg_sq may be NULL
 * for inference-only use). */
int npp_plan_bind(NppPlan* plan, float* params, float* grads, float* exp_avg, float* exp_avg_sq);

/* Re-derive the fp16 shadow weights from the fp32 arena (after init, load_state_dict or any
 * external in-place edit of the parameters). */
int npp_sync_weights(NppPlan* plan, void* stream);

/* Embedder_periodic.embed + Embedder.embed (models/embedder.py:51-56,140-148), materialised in the
 * reference layout [n, K*462] fp32.  coords: device [n,2] fp32 (row y, col x).
 * NPP_MODEL_LIGHT (search mode): [n, 42 + 20] = [embedder.embed(coords) | embedder_periodic.embed(coords)], the two
 * tensors NPP_proposal/search.py:104-108 builds. */
int npp_encode(NppPlan* plan, const float* coords, int64_t n, float* out, void* stream);

/* NPP_Net.forward / NPP_Net_top1.forward on raw coordinates (models/networks.py:56-95,145-173);
 * writes the pre-sigmoid logits [n,3] fp32 and keeps the activations needed by npp_backward. */
int npp_forward(NppPlan* plan, const float* coords, int64_t n, float* logits, void* stream);

/* Full-image inference (NPP_completion/train.py:277-309: render the train / val pixel pools in chunks and scatter them
 * into an image): forward on `n` coordinate rows (any n; processed in max_rows chunks), sigmoid (normalize_type 1) or
 * tanh (2) as models/helpers.py:55-58, written to image[y, x, :] of a device [img_h, img_w, 3] fp32 image.  Pixels not
 * named by a row are left untouched; rows outside the image are skipped. */
int npp_render_into(NppPlan* plan, const float* coords, int64_t n, float* image, int32_t img_h, int32_t img_w,
                    int32_t normalize_type, void* stream);

/* Same network on a materialised encoding: enc device [n, K*462] fp32 in the reference layout (rows of the
 * table built at NPP_completion/train.py:93-105), i.e. NPP_Net.forward(None, x_periodic) verbatim.
 * NPP_MODEL_LIGHT: enc is the [n, 62] layout of npp_encode, i.e. NPP_Net_light.forward(x, x_periodic). */
int npp_forward_encoded(NppPlan* plan, const float* enc, int64_t n, float* logits, void* stream);

/* autograd backward of the forward above: grad_logits device [n,3] fp32 -> bound grads arena
 * (reference: loss.backward(), NPP_completion/train.py:253). */
int npp_backward(NppPlan* plan, int64_t n, const float* grad_logits, void* stream);

/* render()'s sigmoid (models/helpers.py:55-56) + img2mse(...,'l2',...,mask) (models/mse_calculator.py:
 * 13-27) and its gradient w.r.t. the logits.  n_norm is the number of pixel rows the mean runs over
 * (global count under data parallelism).  mask may be NULL (all ones); pred may be NULL.
 * loss is a device float that this call accumulates into (zero it first). */
int npp_mse_fwd_bwd(NppPlan* plan, const float* logits, const float* target, const float* mask, int64_t n,
                    int64_t n_norm, float* pred, float* grad_logits, float* loss, void* stream);

/* torch.optim.Adam single-tensor step (same arithmetic as npp_adam_step) on one caller-owned fp32 tensor of n elements:
 * the small foreign parameters the reference also hands to its optimizer (adaptive_pix latents and the adaptive LPIPS /
 * style heads, models/helpers.py:144-151).  step: 1-based count of the steps in which this tensor had a gradient. */
int npp_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int64_t step, void* stream);

/* img2mse(x, y, 'l2', None, mask) of models/mse_calculator.py:13-27 on the network OUTPUT x [n,3] (after the sigmoid,
 * as NPP_completion/train.py:205-208 calls it) with its gradient: loss (device float, overwritten) = mean over [n,3] of
 * ((x - y)(m + 0.3 (1 - m)))^2, grad_x [n,3] = dL/dx.  mask [n,1] or NULL.  No plan needed. */
int npp_l2_fwd_bwd(const float* x, const float* y, const float* mask, int64_t n, float* loss, float* grad_x,
                   void* stream);

/* Barron's adaptive robust pixel loss, the reference's default --loss_type (models/mse_calculator.py:24-25 ->
 * externel_lib/robust_loss_pytorch/adaptive.py:178-198, distribution.py:173-210, general.py:84-118), forward and
 * backward in one pass over x = network output [n,3] (after the sigmoid), y = target [n,3], mask [n,1] or NULL:
 *   d = (x - y) (m + 0.3 (1 - m));  loss = mean_{n,3} [ rho(d, alpha_c, s_c) + log s_c + log Z(alpha_c) ]
 * latent_alpha / latent_scale: the 3 + 3 parameters of AdaptiveLossFunction (device).  cfg4 (host) = {alpha_lo,
 * alpha_hi, scale_lo, scale_init}.  logz_values / logz_derivs (device, n_knots floats each): log Z and its derivative
 * at uniformly spaced alpha in [0, alpha_max].  scratch9: 9 device floats.  out7 (device) = {loss, dL/dlatent_alpha[3],
 * dL/dlatent_scale[3]};  grad_x [n,3] = dL/dx.  No plan needed. */
int npp_robust_adaptive_fwd_bwd(const float* x, const float* y, const float* mask, int64_t n, const float* latent_alpha,
                                const float* latent_scale, const float* cfg4, const float* logz_values,
                                const float* logz_derivs, int n_knots, float alpha_max, float* scratch9, float* out7,
                                float* grad_x, void* stream);

/* torch.optim.Adam.step over the trained part of the arena (models/helpers.py:164) followed by the
 * shadow-weight refresh.  `step` is the 1-based step count used for bias correction. */
int npp_adam_step(NppPlan* plan, float lr, float beta1, float beta2, float eps, int64_t step, void* stream);

/* One whole train step (NPP_completion/train.py:187-254 with --loss_type l2): encode, forward,
 * loss, backward, Adam.  loss: device float, overwritten. */
int npp_train_step(NppPlan* plan, const float* coords, const float* target, const float* mask, int64_t n,
                   int64_t n_norm, float lr, float beta1, float beta2, float eps, int64_t step, float* loss,
                   void* stream);

/* The same step in three phases, for data parallelism inside one image (reference: nn.DataParallel,
 * models/helpers.py:133-137; here one process per GPU, rows split over the ranks).  The gradient arena
 * exists between the phases so that the caller can sum it over the ranks:
 *   npp_step_forward_backward  encode + forward + RGB head + masked l2 (normalised by the GLOBAL count n_norm) + head
 *                              backward + dgrad chain: one persistent kernel; this rank's share of the loss stays on
 *                              the device until npp_step_finish
 *   npp_step_wgrad             weight + bias gradients of the dense layers [layer_begin, layer_end) (execution order,
 *                              npp_plan_layer_count of them; rgb_linear rides with the last one) into the gradient
 *                              arena, floats [offset, offset + count) of npp_plan_layer_grad_range.  One grouped GEMM
 *                              launch + one reduction launch per call: call it for a few layer groups and all-reduce
 *                              each group's range while the next group's GEMMs run
 *   npp_step_finish            torch.optim.Adam over the arena (models/helpers.py:164), shadow refresh, loss -> *loss
 * n / n_norm must be the values given to npp_step_forward_backward. */
int npp_step_forward_backward(NppPlan* plan, const float* coords, const float* target, const float* mask, int64_t n,
                              int64_t n_norm, void* stream);
int npp_plan_layer_count(const NppPlan* plan);
int npp_plan_layer_grad_range(const NppPlan* plan, int32_t layer_begin, int32_t layer_end, int64_t* offset,
                              int64_t* count);
int npp_step_wgrad(NppPlan* plan, int32_t layer_begin, int32_t layer_end, int64_t n, int64_t n_norm, void* stream);
int npp_step_finish(NppPlan* plan, int64_t n_norm, float lr, float beta1, float beta2, float eps, int64_t step,
                    float* loss, void* stream);

/* A run of `iters` train steps without returning to the caller in between: the loop of
 * NPP_proposal/search.py:110-146 (and of NPP_completion/train.py:150-263 when every batch is known up front).
 * Step i (0-based) reads coords_all[i] ([iters, n, 2]), target_all[i] ([iters, n, 3]) and, if given, mask_all[i]
 * ([iters, n, 1]); its loss goes to losses[i] (device, [iters]).  The learning rate follows the scripts' rewrite
 * (search.py:139-144, train.py:258-263: applied AFTER optimizer.step() with the pre-increment global_step), i.e.
 * Adam step k = first_step + i uses lrate * decay_rate ^ (max(k - 2, 0) / decay_steps).  The calling thread only
 * enqueues kernels: several plans can be driven from several host threads onto several streams at once. */
int npp_fit_run(NppPlan* plan, const float* coords_all, const float* target_all, const float* mask_all, int64_t n,
                int64_t iters, float lrate, float decay_rate, float decay_steps, float beta1, float beta2, float eps,
                int64_t first_step, float* losses, void* stream);

/* The same loop for k plans at once, advanced in lock step: ONE captured step of every plan (parallel branches of a single
 * CUDA graph) launched `iters` times -- the candidate loop of NPP_proposal/search.py:85-148, whose candidates all see the
 * same batches (search.py:91-92 reseeds per candidate).  Arrays of k pointers (host arrays of device pointers; entries may
 * repeat for coords / targets / masks); first_steps[i] is plan i's first Adam step; losses[i] is a device array [iters].
 * NPP_MODEL_LIGHT plans only.  Same arithmetic as k calls of npp_fit_run. */
int npp_multi_fit_run(NppPlan* const* plans, int32_t k, const float* const* coords_all, const float* const* target_all,
                      const float* const* mask_all, int64_t n, int64_t iters, float lrate, float decay_rate,
                      float decay_steps, float beta1, float beta2, float eps, const int64_t* first_steps,
                      float* const* losses, void* stream);

/* Patch crops of GridPatchSampler (models/sampler.py:262-296; utils/extract_glimpse.py:7-79 with mode='nearest',
 * padding_mode='zeros'): out[m, c, i, j] = img[rows[m, i], cols[m, j], c], zero outside the image.  img: device
 * [img_h, img_w, channels] fp32; rows [m, h], cols [m, w] device int64 index tables; out: device [m, channels, h, w]. */
int npp_gather_windows(const float* img, int32_t img_h, int32_t img_w, int32_t channels, const int64_t* rows,
                       const int64_t* cols, int64_t m, int32_t h, int32_t w, float* out, void* stream);

/* Candidate filter of GridPatchSampler.sample_patch_real (models/sampler.py:127-216) for an integer lattice: for each
 * of the n_samples fake centroids (centroids [n_samples, 2] int64, (row, col)) and each (i, j) in [-10, 10)^2, the
 * centroid c = cent + i * shift1 + j * shift2 (shifts4 = {s1_row, s1_col, s2_row, s2_col}, HOST int64) is kept iff it
 * lies strictly inside the image and its (2 half_h) x (2 half_w) window holds at most max_unknown unknown pixels
 * (sampler.py:156-190; pixels outside the image count as unknown).  sat: device int64 [img_h + 1, img_w + 1],
 * summed-area table of (mask < 0.5) with a zero first row / column.  keep: device uint8 [n_samples, 400], candidate
 * q = (i + 10) * 20 + (j + 10) in the order of the reference's meshgrid.  No plan needed. */
int npp_sampler_candidates(const int64_t* sat, int32_t img_h, int32_t img_w, const int64_t* centroids, int32_t n_samples,
                           const int64_t* shifts4, int32_t half_h, int32_t half_w, float max_unknown, uint8_t* keep,
                           void* stream);

/* Optional input pipelining for npp_train_step (the data-loader analogue of NPP_completion/train.py:164-181, where
 * the reference gathers the next batch's rows of the encoding table): encodes `coords` of a FUTURE step into the
 * plan's second encoding buffer on an internal stream.  It waits for everything already enqueued on `stream`
 * (including whatever produces `coords`) and for nothing enqueued afterwards, so call it BEFORE enqueueing the step
 * it should overlap with.  The step whose coords pointer and row count match picks the encoding up instead of
 * encoding again; `coords` must stay unchanged until then.  At most one prefetch per future step is kept. */
int npp_encode_prefetch(NppPlan* plan, const float* coords, int64_t n, void* stream);

/* By default npp_train_step goes straight from the split-K slabs to the Adam update and never writes the
 * gradient arena; turn this on to have it written as well (tests, gradient inspection). */
int npp_set_keep_grads(NppPlan* plan, int on);

/* Number of kernels the last npp_train_step / forward / backward call launched. */
int npp_last_launch_count(const NppPlan* plan);

/* Per-kernel-class device timing with CUDA events on the launching stream (bench.py's roofline).
 * Classes: 0 encode, 1 forward GEMMs (npp_gemm_kmajor), 2 head+loss, 3 dgrad GEMMs (npp_gemm_kmajor),
 * 4 wgrad GEMM (npp_gemm_wgrad), 5 gradient finalize, 6 Adam + shadow refresh.
 * npp_profile_read synchronises the device and returns accumulated milliseconds / launch counts
 * since npp_profile_enable(plan, 1). */
#define NPP_PROFILE_CLASSES 7
int npp_profile_enable(NppPlan* plan, int on);
int npp_profile_read(NppPlan* plan, int n_classes, double* ms, int64_t* launches);

/* Test hooks: copy an internal fp16 activation/gradient buffer ("h0".."h7","d0",...,"delta0",...,
 * "f1","hs","f2","hp","enc1","enc_aux") to a caller fp32 [n, width] device buffer. */
int npp_debug_width(NppPlan* plan, const char* name);
int npp_debug_copy(NppPlan* plan, const char* name, int64_t n, float* out, void* stream);
float npp_debug_grad_scale(NppPlan* plan, void* stream); /* synchronises */

/* Stand-alone tcgen05 GEMM checks (no plan): C[m,n] = A[m,k] . B[n,k]^T (fp16 in, fp32 out) and
 * C[m,n] = A[rows,m]^T . B[rows,n] summed over `splits` row ranges. Dimensions: m%128==0 not required
 * for the first (rows are masked), n%256==0, k%64==0. */
int npp_debug_gemm(const void* a, const void* b, float* c, int m, int n, int k, void* stream);
int npp_debug_gemm_bench(const void* a, const void* b, void* out0, void* out1, int m, int n, int k, int epi,
                         int iters, float* ms_out);
int npp_debug_wgrad(const void* a, const void* b, float* c_partials, int rows, int m, int n, int splits,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NPP_B200_H_ */
