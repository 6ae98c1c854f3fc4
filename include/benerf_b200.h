/*
 * benerf_b200 -- C ABI of the Blackwell-native BeNeRF render-and-image-formation engine.
 *
 * The reference (WU-CVGL/BeNeRF) has no FFI of its own: the drop-in boundary is the
 * Python call surface of model/nerf.py + model/optimize.py + spline.py.
 * benerf_b200/ (Python) keeps that surface and forwards every arithmetic step to the
 * entry points below; each one names the reference lines it replaces.
 *
 * Conventions
 *   - every function returns BNRF_OK (0) or a negative bnrf_status; nothing throws
 *     across the ABI; bnrf_last_error() gives the text of the last failure;
 *   - all pointers marked "device" are CUDA device pointers owned by the CALLER
 *     (PyTorch allocations in practice); the library never frees or retains them
 *     beyond the call, except its packed-weight cache inside the context;
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*),
 *     with no internal synchronisation and no host allocation on the hot path;
 *   - one context per (process, device); not thread-safe across concurrent calls on
 *     the same context (the reference is a single Python loop, train.py:153);
 *   - tensors are fp32 row-major, ray order is POSE-MAJOR: ray n = pose n / R,
 *     pixel n % R (model/nerf.py:242-243);
 *   - there is no CPU fallback: on a machine without an sm_100 device bnrf_create
 *     fails with BNRF_ERR_DEVICE.
 */
#ifndef BENERF_B200_H
#define BENERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNRF_ABI_VERSION 2   /* 2: bnrf_rng grew ray_base / offset_dev, bnrf_adam_step_sched takes advance_scratch, multi-segment render */

typedef enum {
    BNRF_OK = 0,
    BNRF_ERR_ARG = -1,        /* bad argument (null pointer, unsupported size) */
    BNRF_ERR_DEVICE = -2,     /* no usable sm_100 device / wrong device */
    BNRF_ERR_CUDA = -3,       /* a CUDA runtime call or kernel launch failed */
    BNRF_ERR_STATE = -4,      /* weights not set, workspace too small, ... */
    BNRF_ERR_NCCL = -5
} bnrf_status;

/* MLP arithmetic (cfg.mlp_mode).  Both run on the GPU; there is no host path. */
#define BNRF_MLP_TC_FP16X2 0  /* tcgen05.mma kind::f16 on CTA pairs, 2-term split operands (hi*hi + lo*hi + hi*lo), fp32 TMEM accumulate;
                               * hidden state kept in tensor memory, A operand read from TMEM (mlp_tc3.cu) */
#define BNRF_MLP_SIMT_FP32 1  /* plain fp32 FFMA; on-device cross-check of the tensor-core path */
#define BNRF_MLP_TC_PAIR_SS 3 /* round-1 CTA-pair kernel: activations as a shared-memory A operand (mlp_tc2.cu); cross-check of mode 0 */
#define BNRF_MLP_TC_1CTA 2    /* same arithmetic as mode 0 on single CTAs (cta_group::1); mode 0 runs CTA pairs (cta_group::2) */

/* Backward-pass GEMMs (cfg.gemm_mode). */
#define BNRF_GEMM_TC 0         /* tcgen05.mma kind::f16 on bf16 hi/lo split operands (3 MMAs per product), fp32 accumulate */
#define BNRF_GEMM_TC_PER_LINEAR 2   /* like BNRF_GEMM_TC, but the activation gradients take one launch per linear (tile_dgrad_kernel)
                                     * instead of the fused chain (dgrad_chain_kernel); kept as an on-device cross-check */
#define BNRF_GEMM_TC_1CTA 3         /* like BNRF_GEMM_TC with the dgrad chain on single CTAs (dgrad_chain_kernel) instead of CTA pairs
                                     * (dgrad_chain_pair_kernel, cta_group::2); kept as an on-device cross-check */
#define BNRF_GEMM_SIMT_FP32 1  /* plain fp32 FFMA (on-device cross-check; also used for shapes too small for a 128-row tile) */

/* Order of the 12 linears in bnrf_set_weights (the reference's state-dict order). */
enum {
    BNRF_L_PTS0 = 0, /* .. BNRF_L_PTS0 + 7 : pts_linears.0-7  (256,63) (256,256)x4 (256,319) (256,256)x2 */
    BNRF_L_VIEWS = 8,   /* views_linears.0 (128,283) */
    BNRF_L_FEATURE = 9, /* feature_linear  (256,256) */
    BNRF_L_ALPHA = 10,  /* alpha_linear    (1,256)   */
    BNRF_L_RGB = 11,    /* rgb_linear      (C,128)   */
    BNRF_NUM_LINEARS = 12
};

typedef struct bnrf_ctx bnrf_ctx;

/* Static configuration of the render path (args.* the reference reads on this path). */
typedef struct {
    int32_t n_samples;     /* args.N_samples   (64)                       model/nerf.py:297 */
    int32_t n_importance;  /* args.N_importance (64; 0 = coarse only)     model/nerf.py:319 */
    int32_t channels;      /* args.channels, 1 or 3                       model/nerf.py:64  */
    int32_t ndc;           /* args.ndc (always 1 in the reference, Q3)    model/nerf.py:278 */
    float near_, far_;     /* sampling range, Graph.render defaults 0, 1  model/nerf.py:239 */
    int32_t mlp_mode;      /* BNRF_MLP_*                                                    */
    int32_t gemm_mode;     /* BNRF_GEMM_*: arithmetic of the backward pass's dgrad / wgrad GEMMs                */
} bnrf_cfg;

/* The four random draws of one Graph.render.  Parity mode: device pointers to
 * pre-generated tensors.  Production mode: all four NULL and (seed, offset) key an in-kernel
 * Philox4x32-10 stream (statistically equivalent, not bit-equal to torch's generator). */
typedef struct {
    const float* t_rand;   /* device [N, S_c]       U[0,1)  model/nerf.py:305 */
    const float* noise_c;  /* device [N, S_c]       N(0,1)  model/nerf.py:135 */
    const float* u;        /* device [N, N_i]       U[0,1)  run_nerf_helpers.py:86 */
    const float* noise_f;  /* device [N, S_c + N_i] N(0,1)  model/nerf.py:135 */
    /* Parity-only override: device [N, S_c + N_i] sorted fine depths to use INSTEAD of resampling
     * (model/nerf.py:322-326).  The reference's inverse-CDF step divides fp32 rounding noise of its
     * cdf by the bin mass, so two valid fp32 evaluations of the coarse pass pick fine depths that
     * differ by up to 1e-4 in near-empty bins (DESIGN.md "conditioning"); injecting the reference's
     * depths lets the fine network be compared on identical samples.  NULL in production. */
    const float* z_fine;
    uint64_t seed, offset;
    uint64_t ray_base;          /* added to the ray index that keys a draw: a render whose rays are rows [ray_base, ...) of a larger
                                 * batch draws what bnrf_render_forward_multi draws for those rows */
    const uint64_t* offset_dev; /* NULL, or a device counter: the effective stream offset is offset + 64 * (*offset_dev).  A training
                                 * loop captured in a CUDA graph keeps its iteration count there (bnrf_step_advance), so that
                                 * replays draw fresh numbers although the launch arguments are frozen */
} bnrf_rng;

/* Outputs of Graph.render (model/nerf.py:336-343); any pointer may be NULL to skip it. */
typedef struct {
    float* rgb_map;   /* device [N, C]   */
    float* disp_map;  /* device [N]      */
    float* acc_map;   /* device [N]      */
    float* rgb0;      /* device [N, C]   coarse (n_importance > 0) */
    float* disp0;     /* device [N]      */
    float* acc0;      /* device [N]      */
    float* sigma;     /* device [N, S_f] relu(raw_sigma + noise) of the last network */
    float* depth_map; /* device [N]      (raw2output returns it; Graph.render drops it) */
    float* z_vals;    /* device [N, S_f] sorted sample depths of the last network (model/nerf.py:326; not returned upstream) */
} bnrf_outputs;

/* -------------------------------------------------------------------------------------- */
/* context                                                                                */

int bnrf_abi_version(void);
/* Replaces Graph.__init__ (model/nerf.py:152-158) as far as device state goes. */
int bnrf_create(bnrf_ctx** ctx, int device, const bnrf_cfg* cfg);
void bnrf_destroy(bnrf_ctx* ctx);
const char* bnrf_last_error(const bnrf_ctx* ctx);   /* ctx may be NULL: last create() error */

/* Borrow the 12 weight + 12 bias tensors of one network (0 coarse, 1 fine) in PyTorch
 * (out,in) row-major fp32 and repack them into the library's private tensor-core layout.
 * Call after every optimiser step.  Replaces the nn.Linear reads of model/nerf.py:93-112. */
int bnrf_set_weights(bnrf_ctx* ctx, int net, const float* const* weights /*[12] device*/,
                     const float* const* biases /*[12] device*/, void* stream);
/* bnrf_set_weights for both networks of a Graph at once (what a training loop does after every optimiser step): the repack runs
 * as 4 launches over both networks instead of 6 per network. */
int bnrf_set_weights_pair(bnrf_ctx* ctx, const float* const* w_coarse, const float* const* b_coarse, const float* const* w_fine,
                          const float* const* b_fine, void* stream);

/* Override the coarse sampling grid t_vals (host float[S], S == cfg.n_samples).  Default is
 * torch.linspace(0, 1, S) as the reference's CUDA device evaluates it (model/nerf.py:297); the
 * CPU build of torch.linspace differs from it by an ulp depending on the host's vector width,
 * so parity harnesses pass the oracle's own grid. */
int bnrf_set_sample_grid(bnrf_ctx* ctx, const float* t_vals_host, int S, void* stream);

/* -------------------------------------------------------------------------------------- */
/* a1/a2: pose interpolation -- spline.py:247-331 via model/optimize.py:58-111             */

/* knots device [4,6]; transform device [6] or NULL (added in se(3), optimize.py:86-89);
 * ts device [P] in [0,1] (0 and 1 are nudged by 1e-6, spline.py:249-252, without writing ts);
 * traj 0 = cubic B-spline, 1 = linear between knots 0 and 3; poses_out device [P,3,4]. */
int bnrf_spline_poses(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts,
                      int P, int traj, float* poses_out, void* stream);
/* get_pose_evt and get_pose_rgb of one iteration (model/nerf.py:208,211) in one launch: poses [0, n_plain) interpolate the knots as
 * they are (event camera, model/optimize.py:58-82), poses [n_plain, P) the knots + transform (RGB camera, :84-111).  The backward
 * call ADDS into d_knots [4,6] and d_transform [6] (the latter from the poses >= n_plain only). */
int bnrf_spline_poses_pair(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts, int P, int n_plain, int traj,
                           float* poses_out, void* stream);
int bnrf_spline_poses_pair_backward(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts, int P, int n_plain,
                                    int traj, const float* d_poses, float* d_knots, float* d_transform, void* stream);

/* -------------------------------------------------------------------------------------- */
/* a3-a10: Graph.render -- model/nerf.py:236-343                                            */

/* Bytes of caller-provided scratch needed by bnrf_render_forward for n_rays = P*R rays. */
size_t bnrf_workspace_bytes(const bnrf_ctx* ctx, int64_t n_rays);

/* poses device [P,3,4]; ray_idx device int64 [R] (flat pixel index j*W+i); K host float[9]
 * row-major; remap device [H,W,2] or NULL (TUM-VIE LUT, model/nerf.py:247-250).
 * Launches: ray setup (rays, view biases, stratified depths), then one per network pass -- with the default mlp_mode and
 * n_samples / n_samples + n_importance in {32, 64, 128}, raw2output (and, for the coarse pass, sample_pdf + sort) run inside the
 * MLP kernel and neither the per-sample MLP outputs nor the coarse weights are written to the workspace; other sample counts,
 * other mlp_modes and bnrf_render_forward_train use the stand-alone compositing / resampling kernels (same arithmetic, bit-identical
 * results; BNRF_NO_FUSE_COMPOSITE=1 in the environment of bnrf_create forces them). */
int bnrf_render_forward(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R,
                        int H, int W, const float* K, const float* remap, const bnrf_rng* rng,
                        const bnrf_outputs* out, void* workspace, size_t workspace_bytes, void* stream);

/* One Graph.render call as an argument block: several of them can be rendered as ONE ray batch (the rays of segment 0 first,
 * each segment pose-major).  A training iteration renders the event pose pair and the N blur poses (model/nerf.py:217,227);
 * as two segments of one batch every stage runs once over all their rays (half the launches, one partial tile instead of two). */
typedef struct {
    const float* poses;      /* device [P,3,4] */
    const int64_t* ray_idx;  /* device [R] flat pixel indices */
    int32_t P, R, H, W;
    float K[9];              /* host, row-major intrinsics */
    const float* remap;      /* device [H,W,2] or NULL */
} bnrf_render_seg;

/* bnrf_render_forward / bnrf_render_forward_train (saved != NULL) over the concatenated rays of n_segs <= 4 segments;
 * outputs, workspace and saved are sized for N = sum P_i R_i rays. */
int bnrf_render_forward_multi(bnrf_ctx* ctx, const bnrf_render_seg* segs, int n_segs, const bnrf_rng* rng, const bnrf_outputs* out,
                              void* workspace, size_t workspace_bytes, void* saved, size_t saved_bytes, void* stream);

/* BARF coarse-to-fine weighting (args.use_barf_c2f, model/nerf.py:16-26,75-88): the network input is cat([x, w (.) PE(x)]) with
 * per-channel weights w that depend on iter_step.  w_pts: host [63] (the first 3 = 1: the raw point), w_dir: host [27]; both NULL
 * switch the weighting off.  A linear layer sees w (.) e as columns of its weight matrix scaled by w, so the weights are folded
 * into the packed matrices at the next bnrf_set_weights (which the caller must issue: the pack depends on them) and the forward
 * kernels run unchanged; the backward pass scales the affected weight-gradient columns by the same w. */
int bnrf_set_encoding_weights(bnrf_ctx* ctx, const float* w_pts, const float* w_dir, void* stream);

/* -------------------------------------------------------------------------------------- */
/* a16: training -- the part of loss.backward() (train.py:340) that runs through Graph.render */

/* Gradient tables of one network: 12 weight and 12 bias tensors in PyTorch (out,in) layout, order as
 * bnrf_set_weights.  The backward pass ADDS into them (autograd .grad accumulation semantics). */
typedef struct {
    float* weights[BNRF_NUM_LINEARS];
    float* biases[BNRF_NUM_LINEARS];
} bnrf_param_grads;

/* Bytes of the caller-owned buffer in which bnrf_render_forward_train keeps what the backward pass needs
 * (rays, depths, raw outputs, densities and the activations of both networks as 16-bit tile matrices: ~10 KB per
 * sample).  The buffer and the backward workspace must be 256-byte aligned (cudaMalloc / torch allocations are): the
 * backward kernels read them with cp.async.bulk.  The backward call must see the weights the forward call used
 * (no bnrf_set_weights in between). */
size_t bnrf_saved_bytes(const bnrf_ctx* ctx, int64_t n_rays);
/* bnrf_render_forward that additionally fills `saved`.  Needs cfg.mlp_mode == BNRF_MLP_TC_FP16X2 or BNRF_MLP_TC_PAIR_SS. */
int bnrf_render_forward_train(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R,
                              int H, int W, const float* K, const float* remap, const bnrf_rng* rng,
                              const bnrf_outputs* out, void* workspace, size_t workspace_bytes,
                              void* saved, size_t saved_bytes, void* stream);
size_t bnrf_backward_workspace_bytes(const bnrf_ctx* ctx, int64_t n_rays);
/* d_rgb_map / d_rgb0: device [N,C] gradients of the loss w.r.t. the two colour outputs (either may be NULL;
 * the other outputs of Graph.render do not enter any loss of train.py:205-331).  Adds the parameter
 * gradients into grads_coarse / grads_fine and d L / d poses into d_poses (device [P,3,4]).  Same poses,
 * ray_idx, H, W, K, remap as the forward call.  z_vals carry no gradient (model/nerf.py:324 detaches them). */
int bnrf_render_backward(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R, int H, int W,
                         const float* K, const float* remap, const float* d_rgb_map, const float* d_rgb0,
                         const void* saved, size_t saved_bytes, const bnrf_param_grads* grads_coarse,
                         const bnrf_param_grads* grads_fine, float* d_poses, void* workspace,
                         size_t workspace_bytes, void* stream);
/* bnrf_render_backward for a batch rendered by bnrf_render_forward_multi: d_rgb_map / d_rgb0 device [N,C] over the concatenated
 * rays; d_poses[i]: device [P_i,3,4] of segment i (added into). */
int bnrf_render_backward_multi(bnrf_ctx* ctx, const bnrf_render_seg* segs, int n_segs, const float* d_rgb_map, const float* d_rgb0,
                               const void* saved, size_t saved_bytes, const bnrf_param_grads* grads_coarse,
                               const bnrf_param_grads* grads_fine, float* const* d_poses, void* workspace, size_t workspace_bytes,
                               void* stream);
/* Make `stream` wait until the FINE network's parameter gradients of the last bnrf_render_backward[_multi] enqueued on this context
 * are complete (the backward pass handles the fine network first): a data-parallel training loop starts the all-reduce of that
 * half of the gradient buffer on a second stream while the coarse network's backward pass still runs.  Capturable. */
int bnrf_wait_fine_gradients(bnrf_ctx* ctx, void* stream);
/* Backward of bnrf_spline_poses: adds d L / d knots (device [4,6]) and, when transform != NULL,
 * d L / d transform (device [6]) given d L / d poses (device [P,3,4]).  spline.py:247-331 under autograd. */
int bnrf_spline_poses_backward(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts,
                               int P, int traj, const float* d_poses, float* d_knots, float* d_transform,
                               void* stream);

/* Stage-level operators (the same kernels bnrf_render_forward chains; exported so each can be
 * checked 1:1 against the reference function it replaces). */

/* run_nerf_helpers.py:35-71 + model/nerf.py:241-279: rays_o/rays_d (NDC if cfg.ndc) and the
 * pre-NDC unit view directions, all device [N,3]. */
int bnrf_op_rays(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R, int H, int W,
                 const float* K, const float* remap, float* rays_o, float* rays_d, float* viewdirs, void* stream);
/* model/nerf.py:285-307: stratified depths z device [N,S] from t_rand device [N,S]. */
int bnrf_op_stratified(bnrf_ctx* ctx, const float* t_rand, int64_t n_rays, int S, float* z, void* stream);
/* model/embedder.py:9-34 + model/nerf.py:67-116: raw device [N,S,C+1] for pts = o + d*z. */
int bnrf_op_mlp(bnrf_ctx* ctx, int net, const float* rays_o, const float* rays_d, const float* viewdirs,
                const float* z, int64_t n_rays, int S, float* raw, void* stream);
/* model/nerf.py:118-148 (raw2output).  weights/depth/sigma may be NULL. */
int bnrf_op_composite(bnrf_ctx* ctx, const float* raw, const float* z, const float* rays_d, const float* noise,
                      int64_t n_rays, int S, float* rgb_map, float* disp_map, float* acc_map,
                      float* weights, float* depth_map, float* sigma, void* stream);
/* run_nerf_helpers.py:74-115 + model/nerf.py:322-326: z_fine device [N, S+K] sorted. */
int bnrf_op_resample(bnrf_ctx* ctx, const float* z_coarse, const float* weights, const float* u,
                     int64_t n_rays, int S, int K, float* z_fine, void* stream);

/* -------------------------------------------------------------------------------------- */
/* a11, a13, a14: image formation -- train.py:163-177,205-331, utils/event_utils.py:247-259 */

/* Blur model: out[r,c] = (sum_j rgb[j,r,c]) / P, summed j = 0..P-1 (train.py:307-318). */
int bnrf_blur_mean(const float* rgb /*device [P,R,C]*/, int P, int64_t R, int C, float* out /*device [R,C]*/, void* stream);
/* Event model: out[b,r] = L(rgb[b+1,r,:]) - L(rgb[b,r,:]) for b < B, L = log-brightness of the
 * gray value (C==3: 0.299/0.587/0.114, utils/img_utils.py:13-16); log_mode 0 = log(x+1e-9)
 * (BeNeRF_*), 1 = lin-log on 255x (E2NeRF_*), utils/math_utils.py:4-23.  B = 1 is the training
 * pair of train.py:166-173; B > 1 serves get_pose_evt(..., seg_num=B+1) renders. */
int bnrf_event_logdiff(const float* rgb /*device [B+1,R,C]*/, int B, int64_t R, int C, int log_mode,
                       float* out /*device [B,R]*/, void* stream);
/* Backward of the two operators above (train.py:340): g = d loss / d out; d_rgb has the shape of rgb and is OVERWRITTEN. */
int bnrf_blur_mean_backward(const float* g /*device [R,C]*/, int P, int64_t R, int C, float* d_rgb /*device [P,R,C]*/, void* stream);
int bnrf_event_logdiff_backward(const float* rgb /*device [B+1,R,C]*/, const float* g /*device [B,R]*/, int B, int64_t R, int C,
                                int log_mode, float* d_rgb /*device [B+1,R,C]*/, void* stream);
/* Scatter-add polarities into a float64 image (utils/event_utils.py:247-259; float64 per Q10).
 * x, y device int32 [E]; pol device float [E]; out device double [H,W], NOT cleared here. */
int bnrf_accumulate_events(const int32_t* x, const int32_t* y, const float* pol, int64_t E,
                           int H, int W, double* out, void* stream);
/* The same for `bins` consecutive windows of one time-sorted event array in one launch (the 256 event bins of
 * get_pose_evt(..., seg_num = 257), model/optimize.py:58-71, each accumulated as utils/event_utils.py:247-259 would):
 * window b = events [bounds[b], bounds[b+1]) is added to image b.  bounds device int64 [bins+1], non-decreasing, within
 * [0, E]; out device double [bins,H,W], NOT cleared here.  bins <= 4096. */
int bnrf_accumulate_events_binned(const int32_t* x, const int32_t* y, const float* pol, int64_t E, const int64_t* bounds,
                                  int bins, int H, int W, double* out, void* stream);

/* -------------------------------------------------------------------------------------- */
/* a14, a15: the loss block of one iteration, forward + gradients w.r.t. the rendered tensors -- train.py:163-177 (event pair,
 * target gather), 205-292 (event loss), 294-331 (blur mean + rgb loss), loss/imgloss.py:3-5 */

typedef struct {
    int32_t channels;         /* C: 1 | 3 */
    int32_t n_poses;          /* P: virtual poses of the blur render (args.num_interpolated_pose) */
    int32_t log_mode;         /* 0 = log(x + 1e-9) (BeNeRF_*), 1 = lin-log on 255 x (E2NeRF_*), utils/math_utils.py:4-23 */
    int32_t event_loss;       /* args.event_loss (train.py:205); 0: the event terms are 0 and get no gradient */
    int32_t rgb_loss;         /* args.rgb_loss (train.py:294) */
    float event_threshold;    /* > 0: mse(d, t * threshold) * event_coeff_syn (train.py:207-236); else both sides divided by their
                               * L2 norm over the ray batch, * event_coeff_real (train.py:238-292) */
    float event_coeff_syn, event_coeff_real, rgb_coeff;
} bnrf_loss_cfg;

/* Bytes of the device workspace of the two calls below (sums + the event differences of both levels). */
size_t bnrf_training_loss_workspace_bytes(int64_t R_e);
/* Stage 1.  evt_fine / evt_coarse: rgb_map / rgb0 of the event render, device [2 R_e, C] (start poses first); events_accu:
 * device float64 accumulated event image, gathered at idx_evt (device int64 [R_e]); blur_fine / blur_coarse: rgb_map / rgb0 of
 * the blur render, device [P R_b, C] pose-major; blur_target device [R_b, C].  Writes the gradients of the total loss w.r.t.
 * the four renders (same shapes; any may be NULL) and loss_out (device double [5]: total, event fine, event coarse, rgb fine,
 * rgb coarse) -- except, for the NORMALISED event loss, the event gradients and loss_out, which bnrf_training_loss_finish
 * writes.  workspace[0..4] then hold sum d^2 (fine), sum d t (fine), sum d^2 (coarse), sum d t (coarse), sum t^2 of THIS
 * rank's pixels as doubles: with pixel-sharded ranks the caller all-reduces these five numbers between the two calls, so that
 * every rank normalises by the norms of the whole batch exactly as a single process would. */
int bnrf_training_loss(const bnrf_loss_cfg* cfg, const float* evt_fine, const float* evt_coarse, const double* events_accu,
                       const int64_t* idx_evt, int64_t R_e, const float* blur_fine, const float* blur_coarse,
                       const float* blur_target, int64_t R_b, void* workspace, float* d_evt_fine, float* d_evt_coarse,
                       float* d_blur_fine, float* d_blur_coarse, double* loss_out, void* stream);
/* Stage 2 (a no-op unless cfg->event_loss and cfg->event_threshold <= 0): event gradients and loss_out of the normalised loss. */
int bnrf_training_loss_finish(const bnrf_loss_cfg* cfg, const float* evt_fine, const float* evt_coarse, const double* events_accu,
                              const int64_t* idx_evt, int64_t R_e, int64_t R_b, void* workspace, float* d_evt_fine,
                              float* d_evt_coarse, double* loss_out, void* stream);

/* -------------------------------------------------------------------------------------- */
/* f4: camera-response tone mappers -- model/component.py:38-149 (ColorToneMapper / LuminanceToneMapper, input_type "Gray"),
 * applied to the rendered tensors when args.optimize_rgb_crf / optimize_event_crf is set (train.py:176-192,
 * run_nerf_helpers.py:125-126,152-153) */

/* y[e] = sigmoid(L_{hidden+1}(relu(L_hidden(... relu(L_0(x[e])))))): L_0 = Linear(1, width), `hidden` x Linear(width, width),
 * L_{hidden+1} = Linear(width, 1).  weights / biases: hidden + 2 device pointers in layer order, PyTorch (out, in) layout
 * (the nn.Sequential's Linear modules).  width <= 256, hidden <= 4 (and hidden * width^2 floats must fit shared memory). */
int bnrf_crf_forward(int width, int hidden, const float* const* weights, const float* const* biases, const float* x, int64_t n,
                     float* y, void* stream);
/* g = d loss / d y -> dx (written) and the parameter gradients (ADDED into d_weights / d_biases, same shapes as the parameters). */
int bnrf_crf_backward(int width, int hidden, const float* const* weights, const float* const* biases, const float* x, const float* g,
                      int64_t n, float* dx, float* const* d_weights, float* const* d_biases, void* stream);

/* -------------------------------------------------------------------------------------- */
/* f3: fused optimiser tail -- train.py:343-394 (optimizer.step() x3, learning-rate decay, zero_grad),
 * model/optimize.py:36-55 (torch.optim.Adam, default betas / eps, no weight decay) */

typedef struct {
    int64_t begin, end;   /* element range [begin, end) of the flat buffers that forms one optimiser (param group) */
    float lr;             /* this step's learning rate of the group (the caller applies the exponential decay) */
    int32_t active;       /* 0: the reference would not call this optimiser's step() (e.g. optimize_trans = False): skipped */
} bnrf_adam_group;

/* One Adam step over flat fp32 device buffers of n elements: grads are first multiplied by grad_scale (1 / world size
 * after the gradient all-reduce), moments and parameters are updated in place with torch.optim.Adam's arithmetic
 * (step = 1-based count of this call, bias corrections from it), and grads are cleared when zero_grads != 0.
 * Elements outside every group are left untouched (their gradients are still cleared). */
int bnrf_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                   const bnrf_adam_group* groups, int n_groups, int64_t step, float beta1, float beta2, float eps,
                   float grad_scale, int zero_grads, void* stream);

typedef struct {
    int64_t begin, end;   /* element range of the group */
    float lr0;            /* initial learning rate (args.lrate / pose_lrate / transform_lrate) */
    float decay_rate;     /* args.decay_rate / decay_rate_pose / decay_rate_transform */
    int32_t active;
} bnrf_adam_sched_group;

/* bnrf_adam_step for a training loop captured in a CUDA graph: nothing that changes from iteration to iteration is a launch
 * argument.  step_dev: device counter holding train.py's global_step (0-based) of THIS iteration; the kernel derives the
 * 1-based Adam step count (bias corrections) and each group's learning rate of train.py:355-394 from it on the device:
 * lr = lr0 for the first iteration, lr0 * decay_rate ** ((global_step - 1) / decay_steps) afterwards (the reference updates the
 * rate AFTER optimizer.step(), with the not yet incremented global_step).
 * advance_scratch: NULL, or one device word, zero before the first call and otherwise left alone: the launch then also does
 * train.py's `global_step += 1` (*step_dev += 1, by the last thread block to finish), saving the bnrf_step_advance launch. */
int bnrf_adam_step_sched(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                         const bnrf_adam_sched_group* groups, int n_groups, uint64_t* step_dev, double decay_steps,
                         float beta1, float beta2, float eps, float grad_scale, int zero_grads, uint64_t* advance_scratch,
                         void* stream);
/* *step_dev += 1 (train.py: global_step += 1), enqueued as the last node of the captured iteration. */
int bnrf_step_advance(uint64_t* step_dev, void* stream);

/* -------------------------------------------------------------------------------------- */
/* measurement hooks (bench.py): CUDA-event timing of the dominant kernel on its own stream */

/* enable != 0: every MLP kernel launched through this context is bracketed by a cudaEvent pair
 * on the launching stream (up to 512 launches; later ones are counted but not timed).  Resets the
 * counters.  enable == 0 stops recording. */
int bnrf_profile(bnrf_ctx* ctx, int enable);
/* Synchronises the recorded events and returns: summed device time of the timed MLP launches (ms),
 * how many were timed, their summed ALGORITHMIC flops (2 * MACs of the reference's linears, SURVEY
 * 8-d: 593,408 MAC/sample for C=3), and the number of kernel launches of any kind issued through
 * the context since bnrf_profile(ctx, 1). */
int bnrf_profile_read(bnrf_ctx* ctx, double* mlp_ms, int64_t* mlp_timed, double* mlp_flops, int64_t* launches);

/* Bring-up probe (tests only): D[128,N] = A[128,64] * B[N,64]^T through the same shared-memory
 * swizzle, UMMA descriptors, tcgen05.mma and tcgen05.ld helpers as the MLP kernel.  A, B device
 * fp16 row-major; D device fp32 [128,N]; lbo_field = raw 14-bit leading-byte-offset field. */
int bnrf_debug_umma_probe(const void* A_half, const void* B_half, int N, int lbo_field, float* D, void* stream);
/* Host-only: the forward kernel's MMA issue schedule for one tile (tests; no device needed).  groups: max_groups >= 64 entries of
 * 4 ints (GEMM step, kind 0 = all columns / 1, 2 = lower, upper 128-column half, K-block or -1 for the encoded points, flags:
 * 1 first write of an accumulator half, 2 / 4 commit of half 0 / 1, 8 encoded points free, 16 wait for encoded points,
 * 32 wait for the previous layer's converted chunks).  Returns the number of groups, or a negative status. */
int bnrf_debug_tc3_schedule(int split, int32_t* groups, int max_groups, int64_t* stream_bytes_per_cta);
/* Same with the element formats of the instruction descriptor chosen per operand (a_bf16 / b_bf16: 0 = fp16, 1 = bf16):
 * only equal formats are executable on B200 (a mixed pair ends the launch with "illegal instruction" and poisons the CUDA
 * context), which is why the forward pass re-splits its activations as bf16 for the weight-gradient kernel. */
int bnrf_debug_umma_probe_fmt(const void* A16, const void* B16, int N, int a_bf16, int b_bf16, float* D, void* stream);

/* Bring-up probe (tests only) of the ".ts" MMA form on a CTA pair: D[256,N] = A[256,64] * B[N,64]^T with the A operand in
 * tensor memory (written with tcgen05.st at column a_col, lane = row, one 32-bit column = two consecutive K elements) and
 * each CTA holding N/2 rows of B in shared memory; N = 128 | 256, N <= a_col <= 480. */
int bnrf_debug_umma_ts_probe(const void* A_half, const void* B_half, int N, int a_col, float* D, void* stream);

/* Test hook: C[M,N] (op)= A_op * B_op through the backward pass's fp32 GEMM (sgemm.cu).  ta/tb: operand stored
 * transposed; epi 0 store, 1 accumulate, 2 atomic add (split contraction), 3 masked store with optional rank-1 term. */
int bnrf_debug_sgemm(bnrf_ctx* ctx, int ta, int tb, int64_t M, int N, int64_t K, const float* A, int64_t lda,
                     const float* B, int64_t ldb, float* C, int64_t ldc, int epi, const float* mask, int64_t ldm,
                     const float* r_row, int64_t r_stride, const float* r_col, void* stream);

/* Debug: when counters != NULL (device, [148][16] uint64) every tensor-core MLP launch records per-CTA clock64
 * stall accounting of its warp roles there (see mlp_tc.cu); NULL switches it off. */
int bnrf_debug_mlp_trace(bnrf_ctx* ctx, unsigned long long* counters);

/* Debug / tests: the two tensor-core kernels of the backward pass (csrc/bwd_tiles.cu) on caller-provided fp32 device
 * matrices; the conversion to the library's 16-bit tile matrices happens inside the call.
 *   dgrad: out[rows, N] = A[rows, K] . B[N, K]^T (+ r_row (x) r_col) (* (mask[rows, 256] > 0));  K in {128, 256};
 *          N = 256 (mask / rank-1 term optional) or N = 64 (fp32 store, or += when accumulate != 0)
 *   wgrad: dW[m, col0 + n] += sum_rows dz[rows, M]^T . h[rows, N] for n < n_valid; dB[M] += column sums of dz;
 *          with wrow: dWv[N] += sum_rows wrow[row] * h[row, :], dBv[0] += sum wrow.  M in {128, 256}, N in {64, 256}. */
int bnrf_debug_tile_dgrad(bnrf_ctx* ctx, int64_t rows, int K, int N, const float* A, const float* B, const float* mask,
                          const float* r_row, const float* r_col, int accumulate, float* out, void* stream);
int bnrf_debug_tile_wgrad(bnrf_ctx* ctx, int64_t rows, int M, int N, const float* dz, const float* h, const float* wrow,
                          float* dW, int ldw, int col0, int n_valid, float* dB, float* dWv, float* dBv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BENERF_B200_H */
