// K2/K3: pixels x poses -> NDC rays, view directions, stratified depths, per-ray view bias.
//
// Replaces run_nerf_helpers.py:35-71 (get_specific_rays / get_rays, ndc_rays) and
// model/nerf.py:241-308 (pose-major expansion, viewdirs before NDC, jittered depths).
// The reference materialises [N,3,4] repeated poses and ~25 temporaries; here a thread owns a
// ray and keeps everything in registers.  Built with -fmad=false and written with explicit
// round-to-nearest intrinsics in the reference's operation order: these values sit upstream
// of the 2^9 positional-encoding frequency, so a 1-ulp difference becomes 3e-5 in sin/cos.
#include "common.cuh"

namespace bnrf {

// One ray: NDC origin o, direction d and the unit view direction v of pixel ray_idx[n % R] seen from pose n / R.
__device__ __forceinline__ void ray_of(const float* __restrict__ poses, const int64_t* __restrict__ ray_idx, int64_t n, int R, int H, int W,
                                       float fx, float fy, float cx, float cy, const float* __restrict__ remap, int ndc,
                                       float (&o)[3], float (&d)[3], float (&v)[3]) {
    const int p = (int)(n / R);
    const int64_t pix = ray_idx[n % R];
    float fi = (float)(pix % W), fj = (float)(pix / W);
    if (remap) {                                   // TUM-VIE LUT (model/nerf.py:247-250)
        const float* e = remap + 2 * pix;
        fi = e[0]; fj = e[1];
    }
    const float* c2w = poses + (size_t)p * 12;
    // camera-frame direction (run_nerf_helpers.py:36-38)
    const float dx = __fdiv_rn(__fsub_rn(fi, cx), fx);
    const float dy = -__fdiv_rn(__fsub_rn(fj, cy), fy);
    const float dz = -1.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {                  // sum over the last dim of dirs * c2w[:3,:3]
        const float a = __fmul_rn(dx, c2w[r * 4 + 0]), b = __fmul_rn(dy, c2w[r * 4 + 1]), c = __fmul_rn(dz, c2w[r * 4 + 2]);
        d[r] = __fadd_rn(__fadd_rn(a, b), c);
        o[r] = c2w[r * 4 + 3];
    }
    // unit view direction from the PRE-ndc direction (model/nerf.py:272-275)
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
    for (int r = 0; r < 3; ++r) v[r] = __fdiv_rn(d[r], nrm);
    if (ndc) {                                     // run_nerf_helpers.py:46-71 with near = 1, focal = K[0][0]
        const float near = 1.0f;
        const float t = __fdiv_rn(-__fadd_rn(near, o[2]), d[2]);
#pragma unroll
        for (int r = 0; r < 3; ++r) o[r] = __fadd_rn(o[r], __fmul_rn(t, d[r]));
        const float sx = __fdiv_rn(-1.0f, __fdiv_rn((float)W, __fmul_rn(2.0f, fx)));
        const float sy = __fdiv_rn(-1.0f, __fdiv_rn((float)H, __fmul_rn(2.0f, fx)));
        const float o0 = __fdiv_rn(__fmul_rn(sx, o[0]), o[2]);
        const float o1 = __fdiv_rn(__fmul_rn(sy, o[1]), o[2]);
        const float o2 = __fadd_rn(1.0f, __fdiv_rn(2.0f * near, o[2]));
        const float d0 = __fmul_rn(sx, __fsub_rn(__fdiv_rn(d[0], d[2]), __fdiv_rn(o[0], o[2])));
        const float d1 = __fmul_rn(sy, __fsub_rn(__fdiv_rn(d[1], d[2]), __fdiv_rn(o[1], o[2])));
        const float d2 = __fdiv_rn(-2.0f * near, o[2]);
        o[0] = o0; o[1] = o1; o[2] = o2; d[0] = d0; d[1] = d1; d[2] = d2;
    }
}

__global__ void rays_kernel(const float* __restrict__ poses, const int64_t* __restrict__ ray_idx, int P, int R,
                            int H, int W, float fx, float fy, float cx, float cy, const float* __restrict__ remap,
                            int ndc, float* __restrict__ out_o, float* __restrict__ out_d, float* __restrict__ out_v) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (int64_t)P * R) return;
    float o[3], d[3], v[3];
    ray_of(poses, ray_idx, n, R, H, W, fx, fy, cx, cy, remap, ndc, o, d, v);
#pragma unroll
    for (int r = 0; r < 3; ++r) { out_o[n * 3 + r] = o[r]; out_d[n * 3 + r] = d[r]; out_v[n * 3 + r] = v[r]; }
}

// z = lower + (upper - lower) * t_rand over the S strata of [near, far]  (model/nerf.py:285-307)
__device__ __forceinline__ float stratified_depth(const float* __restrict__ t_vals, const float* __restrict__ t_rand, const bnrf_rng& rng,
                                                  int64_t e, int S, float near, float far) {
    const int s = (int)(e % S);
    auto grid = [&](int i) {   // near * (1 - t) + far * t
        const float t = t_vals[i];
        return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
    };
    const float zc = grid(s);
    const float lower = (s == 0) ? zc : __fmul_rn(0.5f, __fadd_rn(zc, grid(s - 1)));
    const float upper = (s == S - 1) ? zc : __fmul_rn(0.5f, __fadd_rn(grid(s + 1), zc));
    float r;
    if (t_rand) {
        r = t_rand[e];
    } else {
        uint32_t w[4];
        Philox::draw(rng.seed, rng_offset(rng), rng.ray_base + (uint64_t)(e / S), (uint32_t)s, kStreamTRand, w);
        r = Philox::uniform(w[0]);
    }
    return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), r));
}

__global__ void stratified_kernel(const float* __restrict__ t_vals, const float* __restrict__ t_rand, bnrf_rng rng,
                                  int64_t total, int S, float near, float far, float* __restrict__ z) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    z[e] = stratified_depth(t_vals, t_rand, rng, e, S, near, far);
}

// Per-ray constant part of views_linears.0: vb[n][j] = b[j] + sum_i W[j][256+i] * enc(viewdir)[i].
// All samples of a ray share the view direction, so the 27 direction features never enter the
// per-sample GEMM (model/nerf.py:80-88,103 concatenates them onto every sample instead).
__device__ __forceinline__ void encode_dir(const float (&v)[3], float (&enc)[kDirCh]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) enc[c] = v[c];
#pragma unroll
    for (int k = 0; k < kDirFreqs; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s, co;
            sincosf(v[c] * (float)(1 << k), &s, &co);
            enc[3 + 6 * k + c] = s;
            enc[3 + 6 * k + 3 + c] = co;
        }
}
__device__ __forceinline__ void viewbias_row(const float (&enc)[kDirCh], const float* __restrict__ w_dir, const float* __restrict__ bias,
                                             int lane, float* __restrict__ vb_row) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = lane + 32 * q;
        float acc = bias[j];
#pragma unroll
        for (int i = 0; i < kDirCh; ++i) acc = fmaf(w_dir[i * kHalf + j], enc[i], acc);
        vb_row[j] = acc;
    }
}

__global__ void viewbias_kernel(const float* __restrict__ view, const float* __restrict__ w_dir /*[27][128]*/,
                                const float* __restrict__ bias /*[128]*/, int64_t n_rays, float* __restrict__ vb) {
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (ray >= n_rays) return;
    float enc[kDirCh];
    const float v[3] = {view[ray * 3], view[ray * 3 + 1], view[ray * 3 + 2]};
    encode_dir(v, enc);
    viewbias_row(enc, w_dir, bias, lane, vb + ray * kHalf);
}

// Everything of a render that depends only on the poses, the pixels and the weights, in ONE launch (the training step is a chain
// of small launches around three big ones): blocks [0, ray_blocks) own four rays each, one per warp -- every lane computes the ray
// (a few dozen flops), lane 0 stores o / d / view, the warp computes the per-ray view bias of the coarse and of the fine network and,
// in training mode, keeps the encoded view direction for the weight gradient of views_linears.0; the remaining blocks draw the
// stratified depths, one thread per sample.
struct RaySetup {
    RaySetupSeg seg[4];
    int n_segs, ndc;
    int64_t n;
    float *o, *d, *view;
    const float* w_dir[2]; const float* vbias[2]; float* vb[2];      // vb[1] == NULL: no fine network
    float* pe_dir;                                                   // [n, 32] or NULL
    const float* dir_scale;                                          // BARF c2f weights of the 27 direction channels, or NULL
    const float* t_vals; const float* t_rand; bnrf_rng rng; int S; float near, far; float* z;
    unsigned ray_blocks;
};

__global__ void __launch_bounds__(128) ray_setup_kernel(const __grid_constant__ RaySetup a) {
    if (blockIdx.x >= a.ray_blocks) {
        const int64_t e = (int64_t)(blockIdx.x - a.ray_blocks) * 128 + threadIdx.x;
        if (e < a.n * a.S) a.z[e] = stratified_depth(a.t_vals, a.t_rand, a.rng, e, a.S, a.near, a.far);
        return;
    }
    const int lane = threadIdx.x % 32;
    const int64_t n = (int64_t)blockIdx.x * 4 + threadIdx.x / 32;
    if (n >= a.n) return;
    int si = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i) if (i < a.n_segs && n >= a.seg[i].off) si = i;
    const RaySetupSeg& sg = a.seg[si];
    float o[3], d[3], v[3];
    ray_of(sg.poses, sg.ray_idx, n - sg.off, sg.R, sg.H, sg.W, sg.fx, sg.fy, sg.cx, sg.cy, sg.remap, a.ndc, o, d, v);
    if (lane < 3) {
        float ov = o[0], dv = d[0], vv = v[0];
        if (lane == 1) { ov = o[1]; dv = d[1]; vv = v[1]; }
        if (lane == 2) { ov = o[2]; dv = d[2]; vv = v[2]; }
        a.o[n * 3 + lane] = ov; a.d[n * 3 + lane] = dv; a.view[n * 3 + lane] = vv;
    }
    float enc[kDirCh];
    encode_dir(v, enc);
    viewbias_row(enc, a.w_dir[0], a.vbias[0], lane, a.vb[0] + n * kHalf);
    if (a.vb[1]) viewbias_row(enc, a.w_dir[1], a.vbias[1], lane, a.vb[1] + n * kHalf);
    if (a.pe_dir) {
        float e = 0.f;
#pragma unroll
        for (int i = 0; i < kDirCh; ++i) if (lane == i) e = enc[i];
        if (a.dir_scale && lane < kDirCh) e *= a.dir_scale[lane];    // BARF c2f: the weight gradient sees the weighted encoding
        a.pe_dir[n * 32 + lane] = e;                                 // lanes 27..31 write the zero padding
    }
}

int launch_ray_setup(bnrf_ctx* ctx, const bnrf_render_seg* segs, int n_segs, const bnrf_rng* rng, int S, float* o, float* d, float* view,
                     float* vb_c, float* vb_f, float* pe_dir, float* z, cudaStream_t st) {
    if (!segs || n_segs <= 0 || n_segs > 4 || !o || !d || !view || !vb_c || !z || S != ctx->cfg.n_samples)
        return fail(ctx, BNRF_ERR_ARG, "ray_setup: bad argument (S must equal cfg.n_samples)");
    RaySetup a{};
    int64_t off = 0;
    for (int i = 0; i < n_segs; ++i) {
        const bnrf_render_seg& sg = segs[i];
        if (!sg.poses || !sg.ray_idx || sg.P <= 0 || sg.R <= 0) return fail(ctx, BNRF_ERR_ARG, "ray_setup: bad segment %d", i);
        a.seg[i] = RaySetupSeg{sg.poses, sg.ray_idx, sg.remap, sg.R, sg.H, sg.W, sg.K[0], sg.K[4], sg.K[2], sg.K[5], off};
        off += (int64_t)sg.P * sg.R;
    }
    a.n_segs = n_segs; a.ndc = ctx->cfg.ndc; a.n = off;
    a.o = o; a.d = d; a.view = view;
    const bool pair = mlp_mode_is_pair(ctx->cfg.mlp_mode);       // feature_linear merged into the view layer: the bias carries W_views . b_feature
    for (int net = 0; net < 2; ++net) {
        const NetParams& np = ctx->net[net];
        a.w_dir[net] = np.w_dir; a.vbias[net] = pair ? np.bias9m : np.bias[9];
    }
    a.vb[0] = vb_c; a.vb[1] = vb_f;
    a.pe_dir = pe_dir; a.dir_scale = ctx->enc_scaled ? ctx->enc_scale + 64 : nullptr;
    bnrf_rng r = rng ? *rng : bnrf_rng{};
    a.t_vals = ctx->t_vals; a.t_rand = r.t_rand; a.rng = r; a.S = S; a.near = ctx->cfg.near_; a.far = ctx->cfg.far_; a.z = z;
    a.ray_blocks = (unsigned)ceil_div(a.n, 4);
    ray_setup_kernel<<<a.ray_blocks + (unsigned)ceil_div(a.n * S, 128), 128, 0, st>>>(a);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int launch_rays(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R, int H, int W, const float* K,
                const float* remap, float* o, float* d, float* view, cudaStream_t st) {
    if (!poses || !ray_idx || !K || !o || !d || !view || P <= 0 || R <= 0) return fail(ctx, BNRF_ERR_ARG, "rays: bad argument");
    const int64_t n = (int64_t)P * R;
    rays_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, st>>>(poses, ray_idx, P, R, H, W, K[0], K[4], K[2], K[5], remap,
                                                           ctx->cfg.ndc, o, d, view);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int launch_stratified(bnrf_ctx* ctx, const float* t_rand, const bnrf_rng* rng, int64_t n, int S, float* z, cudaStream_t st) {
    if (!z || n <= 0 || S != ctx->cfg.n_samples) return fail(ctx, BNRF_ERR_ARG, "stratified: bad argument (S must equal cfg.n_samples)");
    bnrf_rng r = rng ? *rng : bnrf_rng{};
    stratified_kernel<<<(unsigned)ceil_div(n * S, 256), 256, 0, st>>>(ctx->t_vals, t_rand, r, n * S, S, ctx->cfg.near_,
                                                                     ctx->cfg.far_, z);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int launch_viewbias(bnrf_ctx* ctx, int net, const float* view, int64_t n, float* vb, cudaStream_t st) {
    const NetParams& np = ctx->net[net];
    // the CTA-pair kernel runs feature_linear merged into the view layer: its per-ray bias carries W_views . b_feature too
    const float* bias = mlp_mode_is_pair(ctx->cfg.mlp_mode) ? np.bias9m : np.bias[9];
    viewbias_kernel<<<(unsigned)ceil_div(n, 4), 128, 0, st>>>(view, np.w_dir, bias, n, vb);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bnrf
