// Shared declarations of the benerf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/benerf_b200.h"

namespace bnrf {

constexpr int kWidth = 256;        // hidden width W          (model/optimize.py:9)
constexpr int kHalf = 128;         // view-branch width W/2   (model/nerf.py:58)
constexpr int kPtsFreqs = 10;      // multires                (config.py:89)
constexpr int kDirFreqs = 4;       // multires_views          (config.py:91)
constexpr int kPtsCh = 63;         // 3 + 3*2*10
constexpr int kPtsChPad = 64;
constexpr int kDirCh = 27;         // 3 + 3*2*4
constexpr int kMaxSamples = 512;   // S_c + N_i limit of the per-ray warp kernels
constexpr int kNumGemmSteps = 10;  // L0..L7, feature, views

// ---- packed parameters of one network (private cache inside the context) -------------
struct NetParams {
    // fp32 "k-major" copies used by the SIMT path and by every epilogue:
    float* wt[10];        // W^T [K_pad][N] for the 10 GEMM steps (L5 reordered to [pe64 | h256])
    float* bias[10];      // [N]
    float* w_alpha;       // [256]
    float* b_alpha;       // [1]
    float* w_rgb;         // [3][128] (rows >= C are zero)
    float* b_rgb;         // [3]
    float* w_dir;         // [27][128]  view-direction block of views_linears.0, transposed
    // feature_linear has no activation, so it composes with the feature block of views_linears.0 into ONE linear
    // h7 -> view layer (model/nerf.py:102-105): wt9m[k][j] = sum_f W_feature[f][k] * W_views[j][f], bias9m = b_views +
    // W_views[:, :256] . b_feature.  The CTA-pair kernel and the backward pass run this merged step (one GEMM less).
    float* wt9m;          // [256][128] k-major
    float* bias9m;        // [128]
    // tensor-core stream: fp16 hi/lo tiles in the SW128 K-major shared-memory image,
    // in the exact order the TMA producer consumes them (mlp_tc.cu).
    __half* tc_stream;    // single-CTA kernel (mlp_tc.cu)
    __half* tc2_stream;   // CTA-pair kernel (mlp_tc2.cu): [rank][stage], each CTA's half of the output columns
    __half* tc3_stream;   // CTA-pair kernel with A in tensor memory (mlp_tc3.cu): [rank][stage], 8 KB stages in N-half issue order
    float* tc_scale;      // [10] 2^-s per GEMM step undoing the fp16 weight pre-scale
    const float** wt_table;   // device copy of {wt[0..9], wt9m} (the packing kernels index it by GEMM step; 10 = merged)
    unsigned int* absmax;     // device [16] scratch of the per-step max |W| (self-clearing)
    float* scale;             // device [16] 2^s per GEMM step
    // backward pass (bwd_tiles.cu): the 11 dgrad B operands as bf16 hi/lo K-major blocks, packed lazily by the first
    // backward call after bnrf_set_weights
    unsigned char* dg_img;
    unsigned char* dgc_stream;   // dgrad_chain.cu: the transposed weights as [N x 32] bf16 SW64 stages in consumption order
    unsigned char* dgp_stream;   // dgrad_chain2.cu (CTA pairs): [rank][stage], each CTA's half of the output columns, SW128
    bool dg_dirty;
    bool ready;
};

// Activations the training-mode forward pass keeps for one network (written by mlp_tc3.cu / mlp_tc2.cu, read by backward.cu).
struct ActPtrs {
    float* pe_f32;             // [rows, 64] encoded points, fp32 (encoding backward)
    float* h9_f32;             // [rows, 128] view-layer activations, fp32 (rgb head)
    unsigned char* pe_tiles;   // bf16 hi/lo tile matrix of width 64 (bwd_tiles.cuh)
    unsigned char* h_tiles;    // 8 bf16 hi/lo tile matrices of width 256: h0..h7; t_alloc tiles each
    unsigned char* mask_bits;  // ReLU masks of h0..h7, 1 bit per activation, laid out like the tiles: [8][t_alloc][4096] bytes,
                               // byte c = the 8 elements of 16-byte chunk c of the tile's hi part (bit e = element e)
    int64_t t_alloc;
};

inline int gemm_k(int step) { return step == 0 ? 64 : (step == 5 ? 320 : 256); }
inline int gemm_n(int step) { return step == 9 ? 128 : 256; }

}  // namespace bnrf

struct bnrf_ctx {
    int device;
    int sm_count;
    bnrf_cfg cfg;
    bnrf::NetParams net[2];
    float* t_vals;            // device [n_samples] sampling grid (linspace(0,1,S) by default)
    float* enc_scale;         // device [64 + 32]: BARF c2f weights of the 63 point / 27 direction encoding channels (1 when off)
    bool enc_scaled;          // bnrf_set_encoding_weights is in effect
    cudaEvent_t fine_grads_done;  // recorded by the render backward pass once the fine network's parameter gradients are complete
    bool no_fuse;             // BNRF_NO_FUSE_COMPOSITE=1: keep compositing as its own launch (A/B measurements)
    int* tile_counter;        // device scratch for the persistent tile scheduler
    unsigned int* err_flag;   // device: set by kernels on watchdog timeout
    unsigned long long* trace; // device [sm_count][16] stall counters of the last MLP launch, or NULL (bnrf_debug_mlp_trace)
    // measurement hooks (bnrf_profile)
    int prof_enabled;
    int prof_used;
    cudaEvent_t prof_ev[2 * 512];
    double prof_flops;
    long long launches;
    char err[512];
};

namespace bnrf {

extern char g_create_error[512];

int fail(bnrf_ctx* ctx, int code, const char* fmt, ...);

#define BNRF_CUDA(ctx, expr)                                                              \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return bnrf::fail((ctx), BNRF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,       \
                              cudaGetErrorString(e__), __FILE__, __LINE__);               \
    } while (0)

#define BNRF_LAUNCH_CHECK(ctx) do { (ctx)->launches++; BNRF_CUDA(ctx, cudaGetLastError()); } while (0)

struct MlpTimer {      // brackets one MLP launch with events when profiling is on
    bnrf_ctx* ctx; cudaStream_t st; int slot;
    MlpTimer(bnrf_ctx* c, cudaStream_t s, double flops) : ctx(c), st(s), slot(-1) {
        if (c->prof_enabled && c->prof_used < 512) {
            slot = c->prof_used++;
            c->prof_flops += flops;
            cudaEventRecord(c->prof_ev[2 * slot], st);
        }
    }
    ~MlpTimer() { if (slot >= 0) cudaEventRecord(ctx->prof_ev[2 * slot + 1], st); }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- Philox4x32-10 counter RNG (production mode; parity mode reads tensors) ------------
struct Philox {
    __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    // counter = (ray lo, ray hi, sample, stream id); key = seed
    __device__ static inline void draw(uint64_t seed, uint64_t offset, uint64_t ray, uint32_t sample,
                                       uint32_t stream, uint32_t (&out)[4]) {
        uint32_t c[4] = {(uint32_t)ray, (uint32_t)(ray >> 32), sample, stream ^ (uint32_t)(offset * 0x9E3779B9u)};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    __device__ static inline float uniform(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1)
    __device__ static inline float normal(uint32_t a, uint32_t b) {
        const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);                                   // (0,1]
        const float u2 = uniform(b);
        return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
};
// effective Philox stream offset of a render call: host-side call counter + 64 x the device-resident iteration counter
__device__ inline uint64_t rng_offset(const bnrf_rng& r) { return r.offset + (r.offset_dev ? 64ull * __ldg(r.offset_dev) : 0ull); }
enum : uint32_t { kStreamTRand = 1, kStreamNoiseC = 2, kStreamU = 3, kStreamNoiseF = 4 };

// ---- warp scans in double (torch's CPU cumsum / cumprod accumulate fp32 inputs in double, composite.cu) ----
__device__ inline double warp_excl_scan_mul(double v, int lane, double& total) {
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc *= t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    double ex = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? 1.0 : ex;
}
__device__ inline double warp_excl_scan_add(double v, int lane, double& total) {
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    double ex = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? 0.0 : ex;
}
__device__ inline double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Alpha compositing of a render pass as the epilogue of the forward MLP kernel (mlp_tc3.cu; inference mode, S in {32, 64, 128} so that
// a ray's samples never straddle a 128-row tile): the arguments of launch_composite except raw and z, which the kernel has.
struct FuseComposite {
    const float* rays_d; const float* noise; bnrf_rng rng; uint32_t stream_id;
    float *rgb_map, *disp_map, *acc_map, *weights, *depth_map, *sigma;
    // coarse pass with a fine pass behind it: also the inverse-CDF resampling (sample_pdf + sort, composite.cuh) on the weights just
    // produced -- z_f [n, S + K] receives the fine depths (NULL: no resampling); u [n, K] injected draws or NULL (Philox)
    float* z_f; const float* u; int K, sort_n;
};

// ---- launchers implemented in the individual .cu files ---------------------------------
int launch_spline(bnrf_ctx*, const float* knots, const float* transform, const float* ts, int P, int n_plain, int traj,
                  float* poses, cudaStream_t);
int launch_spline_backward(bnrf_ctx*, const float* knots, const float* transform, const float* ts, int P, int n_plain, int traj,
                           const float* d_poses, float* d_knots, float* d_transform, cudaStream_t);
int launch_rays(bnrf_ctx*, const float* poses, const int64_t* ray_idx, int P, int R, int H, int W, const float* K,
                const float* remap, float* o, float* d, float* view, cudaStream_t);
int launch_stratified(bnrf_ctx*, const float* t_rand, const bnrf_rng* rng, int64_t n, int S, float* z, cudaStream_t);
int launch_viewbias(bnrf_ctx*, int net, const float* view, int64_t n, float* vb, cudaStream_t);
// rays + view directions of all segments, per-ray view bias of both networks, encoded view directions (pe_dir, training) and the
// stratified depths in one launch (rays.cu)
struct RaySetupSeg { const float* poses; const int64_t* ray_idx; const float* remap; int R, H, W; float fx, fy, cx, cy; int64_t off; };
int launch_ray_setup(bnrf_ctx*, const bnrf_render_seg* segs, int n_segs, const bnrf_rng* rng, int S, float* o, float* d, float* view,
                     float* vb_c, float* vb_f, float* pe_dir, float* z, cudaStream_t);
int launch_mlp_simt(bnrf_ctx*, int net, const float* o, const float* d, const float* vb, const float* z,
                    int64_t n, int S, float* raw, cudaStream_t);
int launch_mlp_tc(bnrf_ctx*, int net, const float* o, const float* d, const float* vb, const float* z,
                  int64_t n, int S, float* raw, cudaStream_t);
int launch_composite(bnrf_ctx*, const float* raw, const float* z, const float* d, const float* noise,
                     const bnrf_rng* rng, uint32_t stream_id, int64_t n, int S, float* rgb, float* disp, float* acc,
                     float* weights, float* depth, float* sigma, cudaStream_t);
int launch_resample(bnrf_ctx*, const float* zc, const float* w, const float* u, const bnrf_rng* rng, int64_t n,
                    int S, int K, float* zf, cudaStream_t);
int pack_weights(bnrf_ctx*, int net, const float* const* w, const float* const* b, cudaStream_t);
int alloc_net(bnrf_ctx*, int net);
void free_net(bnrf_ctx*, int net);
size_t tc_stream_halfs();
size_t tc2_stream_halfs();
int pack_tc2_stream(bnrf_ctx*, int net, const float* const* table_dev, const float* scale_dev, cudaStream_t);
int launch_mlp_tc2(bnrf_ctx*, int net, const float* o, const float* d, const float* vb, const float* z,
                   int64_t n, int S, float* raw, const ActPtrs* acts /*NULL unless training*/, cudaStream_t);
size_t tc3_stream_halfs();
int pack_tc3_stream_pair(bnrf_ctx*, cudaStream_t);                  // both networks, pre-scale derived from NetParams::absmax in the kernel
int pack_dgrad_chain_pair_stream_both(bnrf_ctx*, cudaStream_t);    // both networks in one launch
int pack_tc3_stream(bnrf_ctx*, int net, const float* const* table_dev, const float* scale_dev, cudaStream_t);
bool mlp_tc3_can_fuse_composite(int S);
int launch_mlp_tc3(bnrf_ctx*, int net, const float* o, const float* d, const float* vb, const float* z,
                   int64_t n, int S, float* raw, const ActPtrs* acts /*NULL unless training*/, cudaStream_t,
                   const FuseComposite* fuse = nullptr /*inference only: composite in the epilogue, raw is not written*/);
// the two CTA-pair kernels run nine GEMM steps (feature_linear merged into the view layer) and take the merged view bias
inline bool mlp_mode_is_pair(int mode) { return mode == BNRF_MLP_TC_FP16X2 || mode == BNRF_MLP_TC_PAIR_SS; }

// Tensors the forward pass keeps for the backward pass (bnrf_render_forward_train), carved from the caller's buffer.
struct SavedLayout {
    float *o, *d, *view;            // [N,3] NDC origin / direction, pre-NDC unit view direction
    float* pe_dir;                  // [N,32] encoded view direction (27 channels, BARF-weighted, zero padded)
    float *z_c, *raw_c, *sig_c;     // [N,S_c], [N,S_c,C+1], [N,S_c] relu(raw_sigma + noise)
    float *z_f, *raw_f, *sig_f;     // fine network
    ActPtrs acts_c, acts_f;
    size_t bytes;
};
SavedLayout carve_saved(const bnrf_cfg& c, int64_t n, void* base);
size_t dgrad_images_bytes();   // backward.cu: the 11 dgrad weight images of one network
size_t dgrad_chain_stream_bytes();
size_t dgrad_chain_pair_stream_bytes();
int pack_dgrad_chain_pair_stream(bnrf_ctx*, int net, cudaStream_t);
int launch_dgrad_chain_pair(bnrf_ctx*, int net, const unsigned char* dz9_tiles, const unsigned char* mask_bits, int64_t t_alloc,
                            const float* d_sigma, int64_t d_sigma_stride, int64_t rows, int64_t dz_tile_count, unsigned char* dz_tiles,
                            float* d_pe, cudaStream_t);
int pack_dgrad_chain_stream(bnrf_ctx*, int net, cudaStream_t);
int launch_dgrad_chain(bnrf_ctx*, int net, const unsigned char* dz9_tiles, const unsigned char* mask_bits, int64_t t_alloc,
                       const float* d_sigma, int64_t d_sigma_stride, int64_t rows, int64_t dz_tile_count, unsigned char* dz_tiles,
                       float* d_pe, cudaStream_t);

}  // namespace bnrf
