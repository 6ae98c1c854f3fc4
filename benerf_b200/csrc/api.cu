// C-ABI entry points (include/benerf_b200.h): context, weight cache, render orchestration.
#include <stdarg.h>
#include <math.h>
#include "common.cuh"
#include "bwd_tiles.cuh"

namespace bnrf {

char g_create_error[512] = "";

int fail(bnrf_ctx* ctx, int code, const char* fmt, ...) {
    char* dst = ctx ? ctx->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

// One launch repacks every tensor of a network (bnrf_set_weights runs after every optimiser step): a table of segments,
// blockIdx.y = segment.  Transposes (in_f > 0): dst[(k_dst + k) * N + n] = src[n * in_f + k_src + k] for k < k_count,
// n < out_f (PyTorch (out,in) -> k-major); plain copies (in_f == 0): dst[i] = src[i] for i < k_count.
// amax (nullable): max |value written| is folded into *amax (as uint bits) -- the per-GEMM-step maxima from which the fp16
// pre-scale of the tensor-core weight stream is derived, gathered while the matrix is transposed anyway.
struct PackSeg { const float* src; float* dst; int out_f, in_f, k_src, k_dst, k_count, N; const float* kscale; /* [k_count] per input channel, or NULL */
                 unsigned int* amax; };
struct PackTable { PackSeg seg[56]; int n; };
__global__ void pack_all_kernel(const __grid_constant__ PackTable t) {
    const PackSeg& s = t.seg[blockIdx.y];
    const int total = s.in_f ? s.k_count * s.out_f : s.k_count;
    float m = 0.0f;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        if (s.in_f) {
            const int k = idx / s.out_f, n = idx % s.out_f;
            float v = s.src[(size_t)n * s.in_f + s.k_src + k];
            if (s.kscale) v *= s.kscale[k];
            s.dst[(size_t)(s.k_dst + k) * s.N + n] = v;
            m = fmaxf(m, fabsf(v));
        } else {
            s.dst[idx] = s.src[idx];
        }
    }
    if (s.amax) {
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x % 32 == 0 && m > 0.0f) atomicMax(s.amax, __float_as_uint(m));
    }
}

// wt9m[k][j] = sum_f wt8[k][f] * wt9[f][j];  bias9m[j] = b_views[j] + sum_f b_feature[f] * wt9[f][j]   (block k, thread j;
// block 256 computes the bias)
struct MergeViews { const float *wt8, *wt9, *b_feature, *b_views; float *wt9m, *bias9m; unsigned int* amax; /* slot 10, nullable */ };
struct MergeViewsPair { MergeViews net[2]; };
__global__ void merge_views_kernel(const __grid_constant__ MergeViewsPair p) {      // blockIdx.y = network
    const MergeViews& m = p.net[blockIdx.y];
    const int k = blockIdx.x, j = threadIdx.x;
    const float* a = (k < kWidth) ? m.wt8 + (size_t)k * kWidth : m.b_feature;
    float acc = 0.f;
#pragma unroll 8
    for (int f = 0; f < kWidth; ++f) acc = fmaf(a[f], m.wt9[(size_t)f * kHalf + j], acc);
    if (k < kWidth) {
        m.wt9m[(size_t)k * kHalf + j] = acc;
        if (m.amax) {
            float v = fabsf(acc);
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (j % 32 == 0 && v > 0.0f) atomicMax(m.amax, __float_as_uint(v));
        }
    } else {
        m.bias9m[j] = m.b_views[j] + acc;
    }
}

int pack_tc_stream(bnrf_ctx*, int net, cudaStream_t);   // mlp_tc.cu

int alloc_net(bnrf_ctx* ctx, int n) {
    NetParams& np = ctx->net[n];
    memset(&np, 0, sizeof(np));
    for (int s = 0; s < 10; ++s) {
        BNRF_CUDA(ctx, cudaMalloc(&np.wt[s], (size_t)gemm_k(s) * gemm_n(s) * sizeof(float)));
        BNRF_CUDA(ctx, cudaMemset(np.wt[s], 0, (size_t)gemm_k(s) * gemm_n(s) * sizeof(float)));   // padding rows stay zero for good
        BNRF_CUDA(ctx, cudaMalloc(&np.bias[s], gemm_n(s) * sizeof(float)));
    }
    BNRF_CUDA(ctx, cudaMalloc(&np.wt9m, (size_t)kWidth * kHalf * sizeof(float)));
    BNRF_CUDA(ctx, cudaMalloc(&np.bias9m, kHalf * sizeof(float)));
    {
        const float* table[11];
        for (int s = 0; s < 10; ++s) table[s] = np.wt[s];
        table[10] = np.wt9m;
        BNRF_CUDA(ctx, cudaMalloc(&np.wt_table, 11 * sizeof(float*)));
        BNRF_CUDA(ctx, cudaMemcpy(np.wt_table, table, 11 * sizeof(float*), cudaMemcpyHostToDevice));
    }
    BNRF_CUDA(ctx, cudaMalloc(&np.absmax, 16 * sizeof(unsigned int)));
    BNRF_CUDA(ctx, cudaMemset(np.absmax, 0, 16 * sizeof(unsigned int)));
    BNRF_CUDA(ctx, cudaMalloc(&np.scale, 16 * sizeof(float)));
    BNRF_CUDA(ctx, cudaMalloc(&np.w_alpha, kWidth * sizeof(float)));
    BNRF_CUDA(ctx, cudaMalloc(&np.b_alpha, sizeof(float)));
    BNRF_CUDA(ctx, cudaMalloc(&np.w_rgb, 3 * kHalf * sizeof(float)));
    BNRF_CUDA(ctx, cudaMalloc(&np.b_rgb, 3 * sizeof(float)));
    BNRF_CUDA(ctx, cudaMemset(np.w_rgb, 0, 3 * kHalf * sizeof(float)));                               // rows >= C stay zero
    BNRF_CUDA(ctx, cudaMemset(np.b_rgb, 0, 3 * sizeof(float)));
    BNRF_CUDA(ctx, cudaMalloc(&np.w_dir, kDirCh * kHalf * sizeof(float)));
    BNRF_CUDA(ctx, cudaMalloc(&np.tc_stream, tc_stream_halfs() * sizeof(__half)));
    BNRF_CUDA(ctx, cudaMalloc(&np.tc2_stream, tc2_stream_halfs() * sizeof(__half)));
    BNRF_CUDA(ctx, cudaMalloc(&np.tc3_stream, tc3_stream_halfs() * sizeof(__half)));
    BNRF_CUDA(ctx, cudaMalloc(&np.tc_scale, 16 * sizeof(float)));
    BNRF_CUDA(ctx, cudaMalloc(&np.dg_img, dgrad_images_bytes()));
    BNRF_CUDA(ctx, cudaMalloc(&np.dgc_stream, dgrad_chain_stream_bytes()));
    BNRF_CUDA(ctx, cudaMalloc(&np.dgp_stream, dgrad_chain_pair_stream_bytes()));
    return BNRF_OK;
}

void free_net(bnrf_ctx* ctx, int n) {
    NetParams& np = ctx->net[n];
    for (int s = 0; s < 10; ++s) { cudaFree(np.wt[s]); cudaFree(np.bias[s]); }
    cudaFree(np.w_alpha); cudaFree(np.b_alpha); cudaFree(np.w_rgb); cudaFree(np.b_rgb); cudaFree(np.w_dir);
    cudaFree(np.tc_stream); cudaFree(np.tc2_stream); cudaFree(np.tc3_stream); cudaFree(np.tc_scale); cudaFree(np.dg_img); cudaFree(np.dgc_stream); cudaFree(np.dgp_stream);
    cudaFree(np.wt_table); cudaFree(np.absmax); cudaFree(np.scale); cudaFree(np.wt9m); cudaFree(np.bias9m);
    memset(&np, 0, sizeof(np));
}

// Segments of one network.  fused: the per-step maxima are gathered here (absmax slots 0..9) instead of by absmax_kernel.
static void pack_segments(bnrf_ctx* ctx, int n, const float* const* w, const float* const* b, PackTable& t, bool fused) {
    NetParams& np = ctx->net[n];
    const int C = ctx->cfg.channels;
    auto tr = [&](const float* W, int out_f, int in_f, int k_src, int k_dst, int k_count, int N, float* dst, int slot, const float* kscale = nullptr) {
        t.seg[t.n++] = PackSeg{W, dst, out_f, in_f, k_src, k_dst, k_count, N, kscale, (fused && slot >= 0) ? np.absmax + slot : nullptr};
    };
    auto cp = [&](const float* src, int cnt, float* dst) { t.seg[t.n++] = PackSeg{src, dst, 0, 0, 0, 0, cnt, 0, nullptr, nullptr}; };
    // BARF c2f (bnrf_set_encoding_weights): the encoding channels' weights scale the weight-matrix columns that read them
    const float* sp = ctx->enc_scaled ? ctx->enc_scale : nullptr;
    const float* sd = ctx->enc_scaled ? ctx->enc_scale + 64 : nullptr;
    // GEMM step s <- reference linear: 0-7 pts_linears, 8 feature_linear, 9 views_linears.0 (feature block)
    tr(w[BNRF_L_PTS0], kWidth, kPtsCh, 0, 0, kPtsCh, kWidth, np.wt[0], 0, sp);
    for (int l = 1; l < 8; ++l) {
        if (l == 5) {   // cat([input_pts, h]) -> [pe64 | h256]  (model/nerf.py:98)
            tr(w[l], kWidth, kPtsCh + kWidth, 0, 0, kPtsCh, kWidth, np.wt[5], 5, sp);
            tr(w[l], kWidth, kPtsCh + kWidth, kPtsCh, kPtsChPad, kWidth, kWidth, np.wt[5], 5);
        } else {
            tr(w[l], kWidth, kWidth, 0, 0, kWidth, kWidth, np.wt[l], l);
        }
    }
    tr(w[BNRF_L_FEATURE], kWidth, kWidth, 0, 0, kWidth, kWidth, np.wt[8], 8);
    tr(w[BNRF_L_VIEWS], kHalf, kWidth + kDirCh, 0, 0, kWidth, kHalf, np.wt[9], 9);       // cat([feature, dirs]) (model/nerf.py:103)
    tr(w[BNRF_L_VIEWS], kHalf, kWidth + kDirCh, kWidth, 0, kDirCh, kHalf, np.w_dir, -1, sd);
    for (int l = 0; l < 8; ++l) cp(b[l], kWidth, np.bias[l]);
    cp(b[BNRF_L_FEATURE], kWidth, np.bias[8]);
    cp(b[BNRF_L_VIEWS], kHalf, np.bias[9]);
    cp(w[BNRF_L_ALPHA], kWidth, np.w_alpha);
    cp(b[BNRF_L_ALPHA], 1, np.b_alpha);
    cp(w[BNRF_L_RGB], C * kHalf, np.w_rgb);
    cp(b[BNRF_L_RGB], C, np.b_rgb);
}

static MergeViews merge_args(NetParams& np, bool fused) {
    return MergeViews{np.wt[8], np.wt[9], np.bias[8], np.bias[9], np.wt9m, np.bias9m, fused ? np.absmax + 10 : nullptr};
}

int pack_weights(bnrf_ctx* ctx, int n, const float* const* w, const float* const* b, cudaStream_t st) {
    NetParams& np = ctx->net[n];
    PackTable t{};
    pack_segments(ctx, n, w, b, t, false);
    pack_all_kernel<<<dim3(16, t.n), 256, 0, st>>>(t);
    BNRF_LAUNCH_CHECK(ctx);
    MergeViewsPair mv{};
    mv.net[0] = merge_args(np, false);
    merge_views_kernel<<<dim3(kWidth + 1, 1), kHalf, 0, st>>>(mv);
    BNRF_LAUNCH_CHECK(ctx);
    int rc = pack_tc_stream(ctx, n, st);
    if (rc != BNRF_OK) return rc;
    np.ready = true;
    np.dg_dirty = true;
    return BNRF_OK;
}

// Both networks of a Graph repacked together (the training loop does this after every optimiser step): 4 launches instead of 12 --
// transpose + per-step maxima, merged view step + its maximum, tensor-core stream (the pre-scale is derived from the maxima inside
// the packing kernel), and, for a training context, the dgrad chain's transposed stream.
int pack_weights_pair(bnrf_ctx* ctx, const float* const* w0, const float* const* b0, const float* const* w1, const float* const* b1,
                      cudaStream_t st) {
    if (ctx->cfg.mlp_mode != BNRF_MLP_TC_FP16X2) {           // the other kernels keep the per-network pipeline
        int rc = pack_weights(ctx, 0, w0, b0, st);
        return rc ? rc : pack_weights(ctx, 1, w1, b1, st);
    }
    PackTable t{};
    for (int n = 0; n < 2; ++n) {
        BNRF_CUDA(ctx, cudaMemsetAsync(ctx->net[n].absmax, 0, 16 * sizeof(unsigned int), st));
        pack_segments(ctx, n, n ? w1 : w0, n ? b1 : b0, t, true);
    }
    pack_all_kernel<<<dim3(16, t.n), 256, 0, st>>>(t);
    BNRF_LAUNCH_CHECK(ctx);
    MergeViewsPair mv{};
    for (int n = 0; n < 2; ++n) mv.net[n] = merge_args(ctx->net[n], true);
    merge_views_kernel<<<dim3(kWidth + 1, 2), kHalf, 0, st>>>(mv);
    BNRF_LAUNCH_CHECK(ctx);
    int rc = pack_tc3_stream_pair(ctx, st);
    if (rc != BNRF_OK) return rc;
    for (int n = 0; n < 2; ++n) { ctx->net[n].ready = true; ctx->net[n].dg_dirty = true; }
    if (ctx->cfg.gemm_mode == BNRF_GEMM_TC) {                // training contexts: the dgrad chain's operand stream, both networks
        if ((rc = pack_dgrad_chain_pair_stream_both(ctx, st))) return rc;
        for (int n = 0; n < 2; ++n) ctx->net[n].dg_dirty = false;
    }
    return BNRF_OK;
}

// Workspace carve-up for bnrf_render_forward (all regions 256-byte aligned).
struct Workspace {
    float *o, *d, *view, *vb, *vb_f, *z_c, *z_f, *raw, *w_c;
    size_t bytes;
};
static Workspace carve(const bnrf_cfg& c, int64_t n, void* base) {
    Workspace w;
    size_t off = 0;
    auto take = [&](size_t floats) {
        float* p = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
        off += (floats * sizeof(float) + 255) / 256 * 256;
        return p;
    };
    const int Sc = c.n_samples, Sf = c.n_samples + c.n_importance, Smax = Sf;
    w.o = take(n * 3); w.d = take(n * 3); w.view = take(n * 3);
    w.vb = take(n * kHalf);
    w.vb_f = take(c.n_importance > 0 ? n * kHalf : 0);
    w.z_c = take(n * Sc);
    w.z_f = take(c.n_importance > 0 ? n * Sf : 0);
    w.raw = take(n * Smax * (c.channels + 1));
    w.w_c = take(n * Sc);
    w.bytes = off;
    return w;
}

static int run_mlp(bnrf_ctx* ctx, int net, const float* o, const float* d, const float* vb, const float* z, int64_t n,
                   int S, float* raw, const ActPtrs* acts, cudaStream_t st, const FuseComposite* fuse = nullptr) {
    if (!ctx->net[net].ready) return fail(ctx, BNRF_ERR_STATE, "weights of network %d not set (bnrf_set_weights)", net);
    const double macs = 63.0 * 256 + 4 * 65536.0 + 319.0 * 256 + 2 * 65536.0 + 256 + 65536.0 + 283.0 * 128 + 128.0 * ctx->cfg.channels;
    MlpTimer timer(ctx, st, 2.0 * macs * (double)n * (double)S);
    if (acts && !mlp_mode_is_pair(ctx->cfg.mlp_mode))
        return fail(ctx, BNRF_ERR_STATE, "training mode (saved activations) needs mlp_mode BNRF_MLP_TC_FP16X2 or BNRF_MLP_TC_PAIR_SS");
    if (ctx->cfg.mlp_mode == BNRF_MLP_SIMT_FP32) return launch_mlp_simt(ctx, net, o, d, vb, z, n, S, raw, st);
    if (ctx->cfg.mlp_mode == BNRF_MLP_TC_1CTA) return launch_mlp_tc(ctx, net, o, d, vb, z, n, S, raw, st);
    if (ctx->cfg.mlp_mode == BNRF_MLP_TC_PAIR_SS) return launch_mlp_tc2(ctx, net, o, d, vb, z, n, S, raw, acts, st);
    return launch_mlp_tc3(ctx, net, o, d, vb, z, n, S, raw, acts, st, fuse);
}

}  // namespace bnrf

using namespace bnrf;

extern "C" {

int bnrf_abi_version(void) { return BNRF_ABI_VERSION; }

const char* bnrf_last_error(const bnrf_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int bnrf_create(bnrf_ctx** out, int device, const bnrf_cfg* cfg) {
    if (!out || !cfg) return fail(nullptr, BNRF_ERR_ARG, "bnrf_create: null argument");
    *out = nullptr;
    if (cfg->n_samples < 3 || cfg->n_importance < 0 || cfg->n_samples + cfg->n_importance > kMaxSamples)
        return fail(nullptr, BNRF_ERR_ARG, "bnrf_create: need 3 <= n_samples and n_samples + n_importance <= %d", kMaxSamples);
    if (cfg->gemm_mode != BNRF_GEMM_TC && cfg->gemm_mode != BNRF_GEMM_SIMT_FP32 && cfg->gemm_mode != BNRF_GEMM_TC_PER_LINEAR && cfg->gemm_mode != BNRF_GEMM_TC_1CTA) return fail(nullptr, BNRF_ERR_ARG, "bnrf_create: unknown gemm_mode %d", cfg->gemm_mode);
    if (cfg->channels != 1 && cfg->channels != 3) return fail(nullptr, BNRF_ERR_ARG, "bnrf_create: channels must be 1 or 3");
    if (cfg->mlp_mode != BNRF_MLP_TC_FP16X2 && cfg->mlp_mode != BNRF_MLP_SIMT_FP32 && cfg->mlp_mode != BNRF_MLP_TC_1CTA && cfg->mlp_mode != BNRF_MLP_TC_PAIR_SS)
        return fail(nullptr, BNRF_ERR_ARG, "bnrf_create: unknown mlp_mode %d", cfg->mlp_mode);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return fail(nullptr, BNRF_ERR_DEVICE, "bnrf_create: CUDA device %d not available (%d visible); there is no CPU path", device, count);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, BNRF_ERR_DEVICE, "bnrf_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    bnrf_ctx* ctx = new bnrf_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->cfg = *cfg;
    ctx->sm_count = prop.multiProcessorCount;
    { const char* e = getenv("BNRF_NO_FUSE_COMPOSITE"); ctx->no_fuse = e && e[0] == '1'; }
    auto bail = [&](int rc) { strncpy(g_create_error, ctx->err, 511); bnrf_destroy(ctx); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(fail(ctx, BNRF_ERR_CUDA, "cudaSetDevice failed"));
    int rc;
    for (int n = 0; n < 2; ++n)
        if ((rc = alloc_net(ctx, n)) != BNRF_OK) return bail(rc);
    if (cudaMalloc(&ctx->enc_scale, 96 * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&ctx->t_vals, kMaxSamples * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&ctx->tile_counter, 64 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&ctx->err_flag, sizeof(unsigned int)) != cudaSuccess)
        return bail(fail(ctx, BNRF_ERR_CUDA, "cudaMalloc failed"));
    cudaMemset(ctx->tile_counter, 0, 64 * sizeof(int));
    cudaMemset(ctx->err_flag, 0, sizeof(unsigned int));
    if (cudaEventCreateWithFlags(&ctx->fine_grads_done, cudaEventDisableTiming) != cudaSuccess) return bail(fail(ctx, BNRF_ERR_CUDA, "cudaEventCreate failed"));
    // default sampling grid = torch.linspace(0, 1, S) as the CUDA kernel of the reference's device computes it:
    // start + step*i below the midpoint, end - step*(S-1-i) above, single rounding (fma).
    const int S = cfg->n_samples;
    float host[kMaxSamples];
    const float step = 1.0f / (float)(S - 1);
    for (int i = 0; i < S; ++i) host[i] = (i < S / 2) ? fmaf(step, (float)i, 0.0f) : fmaf(-step, (float)(S - i - 1), 1.0f);
    if (cudaMemcpy(ctx->t_vals, host, S * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(fail(ctx, BNRF_ERR_CUDA, "cudaMemcpy failed"));
    *out = ctx;
    return BNRF_OK;
}

void bnrf_destroy(bnrf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int n = 0; n < 2; ++n) free_net(ctx, n);
    cudaFree(ctx->t_vals); cudaFree(ctx->tile_counter); cudaFree(ctx->err_flag); cudaFree(ctx->enc_scale);
    if (ctx->prof_ev[0]) for (int i = 0; i < 2 * 512; ++i) cudaEventDestroy(ctx->prof_ev[i]);
    if (ctx->fine_grads_done) cudaEventDestroy(ctx->fine_grads_done);
    delete ctx;
}

int bnrf_profile(bnrf_ctx* ctx, int enable) {
    if (!ctx) return BNRF_ERR_ARG;
    if (enable && !ctx->prof_ev[0])
        for (int i = 0; i < 2 * 512; ++i) BNRF_CUDA(ctx, cudaEventCreate(&ctx->prof_ev[i]));
    ctx->prof_enabled = enable ? 1 : 0;
    if (enable) { ctx->prof_used = 0; ctx->prof_flops = 0.0; ctx->launches = 0; }
    return BNRF_OK;
}

int bnrf_profile_read(bnrf_ctx* ctx, double* mlp_ms, int64_t* mlp_timed, double* mlp_flops, int64_t* launches) {
    if (!ctx) return BNRF_ERR_ARG;
    double total = 0.0;
    for (int i = 0; i < ctx->prof_used; ++i) {
        BNRF_CUDA(ctx, cudaEventSynchronize(ctx->prof_ev[2 * i + 1]));
        float ms = 0.f;
        BNRF_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        total += ms;
    }
    if (mlp_ms) *mlp_ms = total;
    if (mlp_timed) *mlp_timed = ctx->prof_used;
    if (mlp_flops) *mlp_flops = ctx->prof_flops;
    if (launches) *launches = ctx->launches;
    return BNRF_OK;
}

int bnrf_debug_mlp_trace(bnrf_ctx* ctx, unsigned long long* counters) {
    if (!ctx) return BNRF_ERR_ARG;
    ctx->trace = counters;
    return BNRF_OK;
}

int bnrf_set_sample_grid(bnrf_ctx* ctx, const float* t_vals_host, int S, void* stream) {
    if (!ctx || !t_vals_host || S != ctx->cfg.n_samples) return fail(ctx, BNRF_ERR_ARG, "set_sample_grid: S must equal cfg.n_samples");
    BNRF_CUDA(ctx, cudaMemcpyAsync(ctx->t_vals, t_vals_host, S * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    BNRF_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));   // the host buffer is borrowed only for the call
    return BNRF_OK;
}

int bnrf_set_encoding_weights(bnrf_ctx* ctx, const float* w_pts, const float* w_dir, void* stream) {
    if (!ctx || (w_pts == nullptr) != (w_dir == nullptr)) return fail(ctx, BNRF_ERR_ARG, "set_encoding_weights: pass both weight vectors or neither");
    float host[96];
    for (int i = 0; i < 96; ++i) host[i] = 1.0f;
    if (w_pts) {
        memcpy(host, w_pts, kPtsCh * sizeof(float));
        memcpy(host + 64, w_dir, kDirCh * sizeof(float));
    }
    BNRF_CUDA(ctx, cudaMemcpyAsync(ctx->enc_scale, host, sizeof(host), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    BNRF_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));   // the host buffer is borrowed only for the call
    ctx->enc_scaled = w_pts != nullptr;
    for (int n = 0; n < 2; ++n) ctx->net[n].ready = false;         // the packed matrices no longer match: bnrf_set_weights must follow
    return BNRF_OK;
}

int bnrf_set_weights(bnrf_ctx* ctx, int net, const float* const* weights, const float* const* biases, void* stream) {
    if (!ctx || !weights || !biases || net < 0 || net > 1) return fail(ctx, BNRF_ERR_ARG, "set_weights: bad argument");
    for (int i = 0; i < BNRF_NUM_LINEARS; ++i)
        if (!weights[i] || !biases[i]) return fail(ctx, BNRF_ERR_ARG, "set_weights: linear %d is null", i);
    return pack_weights(ctx, net, weights, biases, (cudaStream_t)stream);
}

int bnrf_set_weights_pair(bnrf_ctx* ctx, const float* const* w_coarse, const float* const* b_coarse, const float* const* w_fine,
                          const float* const* b_fine, void* stream) {
    if (!ctx || !w_coarse || !b_coarse || !w_fine || !b_fine) return fail(ctx, BNRF_ERR_ARG, "set_weights_pair: bad argument");
    for (int i = 0; i < BNRF_NUM_LINEARS; ++i)
        if (!w_coarse[i] || !b_coarse[i] || !w_fine[i] || !b_fine[i]) return fail(ctx, BNRF_ERR_ARG, "set_weights_pair: linear %d is null", i);
    return pack_weights_pair(ctx, w_coarse, b_coarse, w_fine, b_fine, (cudaStream_t)stream);
}

int bnrf_spline_poses(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts, int P, int traj,
                      float* poses_out, void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    return launch_spline(ctx, knots, transform, ts, P, 0, traj, poses_out, (cudaStream_t)stream);
}

int bnrf_spline_poses_pair(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts, int P, int n_plain, int traj,
                           float* poses_out, void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    return launch_spline(ctx, knots, transform, ts, P, n_plain, traj, poses_out, (cudaStream_t)stream);
}

int bnrf_spline_poses_pair_backward(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts, int P, int n_plain,
                                    int traj, const float* d_poses, float* d_knots, float* d_transform, void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    return launch_spline_backward(ctx, knots, transform, ts, P, n_plain, traj, d_poses, d_knots, d_transform, (cudaStream_t)stream);
}

int bnrf_spline_poses_backward(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts, int P, int traj,
                               const float* d_poses, float* d_knots, float* d_transform, void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    return launch_spline_backward(ctx, knots, transform, ts, P, 0, traj, d_poses, d_knots, d_transform, (cudaStream_t)stream);
}

size_t bnrf_workspace_bytes(const bnrf_ctx* ctx, int64_t n_rays) {
    if (!ctx || n_rays <= 0) return 0;
    return carve(ctx->cfg, n_rays, nullptr).bytes;
}

// Shared body of bnrf_render_forward / bnrf_render_forward_train.  With `saved` the per-ray tensors the backward pass
// needs (rays, depths, raw outputs, densities) are placed in the caller's saved buffer instead of the scratch
// workspace and the MLP kernel also writes its activations there.
static int render_impl(bnrf_ctx* ctx, const bnrf_render_seg* segs, int n_segs, const bnrf_rng* rng, const bnrf_outputs* out,
                       void* workspace, size_t workspace_bytes, void* saved, size_t saved_bytes, void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    if (!segs || n_segs <= 0 || n_segs > 4 || !out || !workspace) return fail(ctx, BNRF_ERR_ARG, "render_forward: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const bnrf_cfg& c = ctx->cfg;
    int64_t n = 0;
    for (int i = 0; i < n_segs; ++i) {
        if (!segs[i].poses || !segs[i].ray_idx || segs[i].P <= 0 || segs[i].R <= 0) return fail(ctx, BNRF_ERR_ARG, "render_forward: bad segment %d", i);
        n += (int64_t)segs[i].P * segs[i].R;
    }
    Workspace w = carve(c, n, workspace);
    if (workspace_bytes < w.bytes) return fail(ctx, BNRF_ERR_STATE, "render_forward: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    SavedLayout s{};
    float* raw_c = w.raw; float* raw_f = w.raw;
    float* sig_c = nullptr; float* sig_f = nullptr;
    if (saved) {
        s = carve_saved(c, n, saved);
        if (saved_bytes < s.bytes) return fail(ctx, BNRF_ERR_STATE, "render_forward_train: saved buffer %zu < %zu bytes", saved_bytes, s.bytes);
        w.o = s.o; w.d = s.d; w.view = s.view; w.z_c = s.z_c; w.z_f = s.z_f;
        raw_c = s.raw_c; raw_f = s.raw_f; sig_c = s.sig_c; sig_f = s.sig_f;
    }
    bnrf_rng r = rng ? *rng : bnrf_rng{};
    const bool fine = c.n_importance > 0;
    const int Sc = c.n_samples, Sf = c.n_samples + c.n_importance;
    int rc;
    // rays of all segments (laid out one after the other, each pose-major), view bias of both networks, stratified depths
    if ((rc = launch_ray_setup(ctx, segs, n_segs, &r, Sc, w.o, w.d, w.view, w.vb, fine ? w.vb_f : nullptr, saved ? s.pe_dir : nullptr, w.z_c, st))) return rc;
    // A forward-only render on the default MLP kernel composites in that kernel's epilogue when a ray's samples fit a tile
    // (S = 32, 64, 128): `raw` never reaches HBM and the pass is one launch.  Training mode keeps `raw` for the backward pass.
    const bool can_fuse = !saved && c.mlp_mode == BNRF_MLP_TC_FP16X2 && !ctx->no_fuse;
    // coarse composite: outputs go to rgb0/disp0/acc0 when a fine pass follows (model/nerf.py:319-343)
    float* sigma_c_out = saved ? sig_c : (fine ? nullptr : out->sigma);
    bool resampled = false;                  // the coarse launch has already produced the fine depths
    if (can_fuse && mlp_tc3_can_fuse_composite(Sc)) {
        FuseComposite fz{w.d, r.noise_c, r, kStreamNoiseC, fine ? out->rgb0 : out->rgb_map, fine ? out->disp0 : out->disp_map,
                         fine ? out->acc0 : out->acc_map, w.w_c, fine ? nullptr : out->depth_map, sigma_c_out, nullptr, nullptr, 0, 0};
        if (fine && !r.z_fine) {
            int sort_n = 1;
            while (sort_n < Sf) sort_n <<= 1;
            if ((128 / Sc) * (2 * Sc + sort_n) * 4 <= 5120) {            // the resampler's scratch fits beside the tile: fuse it too
                fz.z_f = w.z_f; fz.u = r.u; fz.K = c.n_importance; fz.sort_n = sort_n; fz.weights = nullptr;
                resampled = true;
            }
        }
        if ((rc = run_mlp(ctx, 0, w.o, w.d, w.vb, w.z_c, n, Sc, raw_c, nullptr, st, &fz))) return rc;
    } else {
        if ((rc = run_mlp(ctx, 0, w.o, w.d, w.vb, w.z_c, n, Sc, raw_c, saved ? &s.acts_c : nullptr, st))) return rc;
        if ((rc = launch_composite(ctx, raw_c, w.z_c, w.d, r.noise_c, &r, kStreamNoiseC, n, Sc,
                                   fine ? out->rgb0 : out->rgb_map, fine ? out->disp0 : out->disp_map,
                                   fine ? out->acc0 : out->acc_map, w.w_c, fine ? nullptr : out->depth_map, sigma_c_out, st))) return rc;
    }
    if (!fine) {
        if (saved && out->sigma) BNRF_CUDA(ctx, cudaMemcpyAsync(out->sigma, sig_c, (size_t)n * Sc * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (out->z_vals) BNRF_CUDA(ctx, cudaMemcpyAsync(out->z_vals, w.z_c, (size_t)n * Sc * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return BNRF_OK;
    }
    if (r.z_fine) {
        BNRF_CUDA(ctx, cudaMemcpyAsync(w.z_f, r.z_fine, (size_t)n * Sf * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if (!resampled && (rc = launch_resample(ctx, w.z_c, w.w_c, r.u, &r, n, Sc, c.n_importance, w.z_f, st))) {
        return rc;
    }
    if (out->z_vals) BNRF_CUDA(ctx, cudaMemcpyAsync(out->z_vals, w.z_f, (size_t)n * Sf * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (can_fuse && mlp_tc3_can_fuse_composite(Sf)) {
        const FuseComposite fz{w.d, r.noise_f, r, kStreamNoiseF, out->rgb_map, out->disp_map, out->acc_map, nullptr, out->depth_map, out->sigma};
        if ((rc = run_mlp(ctx, 1, w.o, w.d, w.vb_f, w.z_f, n, Sf, raw_f, nullptr, st, &fz))) return rc;
    } else {
        if ((rc = run_mlp(ctx, 1, w.o, w.d, w.vb_f, w.z_f, n, Sf, raw_f, saved ? &s.acts_f : nullptr, st))) return rc;
        if ((rc = launch_composite(ctx, raw_f, w.z_f, w.d, r.noise_f, &r, kStreamNoiseF, n, Sf, out->rgb_map, out->disp_map,
                                   out->acc_map, nullptr, out->depth_map, saved ? sig_f : out->sigma, st))) return rc;
    }
    if (saved && out->sigma) BNRF_CUDA(ctx, cudaMemcpyAsync(out->sigma, sig_f, (size_t)n * Sf * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return BNRF_OK;
}

static bnrf_render_seg one_seg(const float* poses, const int64_t* ray_idx, int P, int R, int H, int W, const float* K, const float* remap) {
    bnrf_render_seg s{};
    s.poses = poses; s.ray_idx = ray_idx; s.P = P; s.R = R; s.H = H; s.W = W; s.remap = remap;
    if (K) memcpy(s.K, K, sizeof(s.K));
    return s;
}

int bnrf_render_forward(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R, int H, int W,
                        const float* K, const float* remap, const bnrf_rng* rng, const bnrf_outputs* out,
                        void* workspace, size_t workspace_bytes, void* stream) {
    if (!K) return fail(ctx, BNRF_ERR_ARG, "render_forward: K is null");
    const bnrf_render_seg s = one_seg(poses, ray_idx, P, R, H, W, K, remap);
    return render_impl(ctx, &s, 1, rng, out, workspace, workspace_bytes, nullptr, 0, stream);
}

int bnrf_render_forward_multi(bnrf_ctx* ctx, const bnrf_render_seg* segs, int n_segs, const bnrf_rng* rng, const bnrf_outputs* out,
                              void* workspace, size_t workspace_bytes, void* saved, size_t saved_bytes, void* stream) {
    return render_impl(ctx, segs, n_segs, rng, out, workspace, workspace_bytes, saved, saved_bytes, stream);
}

int bnrf_render_forward_train(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R, int H, int W,
                              const float* K, const float* remap, const bnrf_rng* rng, const bnrf_outputs* out,
                              void* workspace, size_t workspace_bytes, void* saved, size_t saved_bytes, void* stream) {
    if (!saved || !K) return fail(ctx, BNRF_ERR_ARG, "render_forward_train: saved buffer / K is null");
    const bnrf_render_seg s = one_seg(poses, ray_idx, P, R, H, W, K, remap);
    return render_impl(ctx, &s, 1, rng, out, workspace, workspace_bytes, saved, saved_bytes, stream);
}

int bnrf_op_rays(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R, int H, int W, const float* K,
                 const float* remap, float* rays_o, float* rays_d, float* viewdirs, void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    return launch_rays(ctx, poses, ray_idx, P, R, H, W, K, remap, rays_o, rays_d, viewdirs, (cudaStream_t)stream);
}

int bnrf_op_stratified(bnrf_ctx* ctx, const float* t_rand, int64_t n_rays, int S, float* z, void* stream) {
    if (!ctx || !t_rand) return fail(ctx, BNRF_ERR_ARG, "op_stratified: bad argument");
    return launch_stratified(ctx, t_rand, nullptr, n_rays, S, z, (cudaStream_t)stream);
}

int bnrf_op_mlp(bnrf_ctx* ctx, int net, const float* rays_o, const float* rays_d, const float* viewdirs, const float* z,
                int64_t n_rays, int S, float* raw, void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    if (!rays_o || !rays_d || !viewdirs || !z || !raw || n_rays <= 0 || S <= 0 || net < 0 || net > 1)
        return fail(ctx, BNRF_ERR_ARG, "op_mlp: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    float* vb = nullptr;                      // operator-level call: scratch for the per-ray view bias
    BNRF_CUDA(ctx, cudaMallocAsync(&vb, (size_t)n_rays * kHalf * sizeof(float), st));
    int rc = ctx->net[net].ready ? launch_viewbias(ctx, net, viewdirs, n_rays, vb, st)
                                 : fail(ctx, BNRF_ERR_STATE, "weights of network %d not set", net);
    if (rc == BNRF_OK) rc = run_mlp(ctx, net, rays_o, rays_d, vb, z, n_rays, S, raw, nullptr, st);
    cudaFreeAsync(vb, st);
    return rc;
}

int bnrf_op_composite(bnrf_ctx* ctx, const float* raw, const float* z, const float* rays_d, const float* noise,
                      int64_t n_rays, int S, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                      float* depth_map, float* sigma, void* stream) {
    if (!ctx || !noise) return fail(ctx, BNRF_ERR_ARG, "op_composite: bad argument");
    return launch_composite(ctx, raw, z, rays_d, noise, nullptr, 0, n_rays, S, rgb_map, disp_map, acc_map, weights,
                            depth_map, sigma, (cudaStream_t)stream);
}

int bnrf_op_resample(bnrf_ctx* ctx, const float* z_coarse, const float* weights, const float* u, int64_t n_rays, int S,
                     int K, float* z_fine, void* stream) {
    if (!ctx || !u) return fail(ctx, BNRF_ERR_ARG, "op_resample: bad argument");
    return launch_resample(ctx, z_coarse, weights, u, nullptr, n_rays, S, K, z_fine, (cudaStream_t)stream);
}

}  // extern "C"
