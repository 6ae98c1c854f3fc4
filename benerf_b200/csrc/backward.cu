// Backward pass of Graph.render (the part of train.py:340 loss.backward() that runs through
// model/nerf.py:236-343): d(rgb_map, rgb0) -> d(NeRF parameters of both networks), d(poses).
//
// The forward pass in training mode (bnrf_render_forward_train / _multi) keeps, per network, the encoded
// points and the hidden activations of every sample as bf16 hi/lo tile matrices (written by
// the tensor-core kernel's epilogue, mlp_tc3.cu; format in bwd_tiles.cuh) plus raw / z / sigma.
// Per network, six launches:
//   composite_backward   raw2output (model/nerf.py:118-148): sigmoid, relu(sigma+noise), alpha,
//                        exclusive cumprod, sum(w*rgb) -- one warp per ray, suffix scan
//   heads_fused          rgb_linear and the view layer's ReLU: dZ9 tiles, d view bias, sum dZ9, dW / dB of rgb_linear
//   dgrad chain          d activations of the whole network in ONE launch (dgrad_chain2.cu), dZ_l tile matrices out
//   wgrad (pair, 64-wide) d weights / d biases of the wide linears (wgrad_pair.cu) and of the two encoded-point blocks
//                        (bwd_tiles.cu), operands straight from the tile matrices
//   net_tail             by block range: sin/cos encoding (model/embedder.py:9-34) and pts = o + d*z summed over a ray's
//                        samples; direction encoding + per-ray view bias; direction block of views_linears.0;
//                        feature_linear / views_linears.0 from the merged contraction
// and once: rays_backward  ndc_rays + get_specific_rays + viewdirs (run_nerf_helpers.py:35-71) -> d poses, all segments.
// z_vals carry no gradient (stratified depths have no parameters; z_samples are detached,
// model/nerf.py:324), exactly as in the reference.  d poses -> d knots is pose.cu (dual numbers).
#include "common.cuh"
#include "bwd_tiles.cuh"
#include "tc_ptx.cuh"

namespace bnrf {

namespace {

constexpr int kWarps = 4;

__device__ inline float wsumf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------ raw2output backward
// g_rgb [N,C] = dL/d rgb_map.  Writes d_raw [N,S,C+1]; adds dL/d|rays_d| to d_dnorm [N].
template <int C>
__global__ void composite_backward_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                          const float* __restrict__ sigma, const float* __restrict__ rays_d,
                                          const float* __restrict__ g_rgb, int64_t n_rays, int S,
                                          float* __restrict__ d_raw, float* __restrict__ d_dnorm) {
    const int lane = threadIdx.x % 32;
    const int64_t ray = (int64_t)blockIdx.x * kWarps + threadIdx.x / 32;
    if (ray >= n_rays) return;
    const int K = (S + 31) / 32;
    constexpr int MAXK = kMaxSamples / 32;
    const float* rd = rays_d + ray * 3;
    const float dn = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
    const float* zr = z + ray * S;
    const float* rr = raw + ray * S * (C + 1);
    const float* sg = sigma + ray * S;
    float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C; ++c) g[c] = g_rgb[ray * C + c];
    float em[MAXK], gap[MAXK];                  // exp(-sigma*delta) = 1 - alpha; z gap (delta / |d|)
    double prod = 1.0;
    for (int k = 0; k < K; ++k) {
        const int s = lane * K + k;
        em[k] = 1.0f; gap[k] = 0.0f;
        if (s < S) {
            gap[k] = (s + 1 < S) ? zr[s + 1] - zr[s] : 1e10f;
            em[k] = expf(-sg[s] * (gap[k] * dn));
            prod *= (double)((1.0f - (1.0f - em[k])) + 1e-10f);
        }
    }
    // exclusive prefix product of (1 - alpha + 1e-10) across lanes
    double inc = prod;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc *= t;
    }
    double t_run = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) t_run = 1.0;
    float Tk[MAXK], Ak[MAXK];
    double local = 0.0;                          // sum of A_k * w_k over this lane's samples
    for (int k = 0; k < K; ++k) {
        const int s = lane * K + k;
        Tk[k] = 0.f; Ak[k] = 0.f;
        if (s < S) {
            const float alpha = 1.0f - em[k];
            const float T = (float)t_run;
            float A = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) A = fmaf(g[c], 1.0f / (1.0f + expf(-rr[s * (C + 1) + c])), A);
            Tk[k] = T; Ak[k] = A;
            local += (double)A * (double)(alpha * T);
            t_run *= (double)((1.0f - alpha) + 1e-10f);
        }
    }
    // suffix sum over later samples: lanes after this one, then this lane's own later samples
    double incs = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incs, o);
        if (lane >= o) incs += t;
    }
    const double total = __shfl_sync(0xffffffffu, incs, 31);
    double after = total - incs;                 // sum over lanes > lane
    float dn_acc = 0.f;
    for (int k = K - 1; k >= 0; --k) {
        const int s = lane * K + k;
        if (s < S) {
            const float alpha = 1.0f - em[k];
            const float f = (1.0f - alpha) + 1e-10f;
            const float T = Tk[k], A = Ak[k];
            const float w = alpha * T;
            const float dalpha = A * T - (float)(after / (double)f);
            const float delta = gap[k] * dn;
            const float sig = sg[s];
            const float dsig = dalpha * delta * em[k];
            d_raw[(ray * S + s) * (C + 1) + C] = (sig > 0.0f) ? dsig : 0.0f;
            const float ddelta = dalpha * sig * em[k];
            dn_acc = fmaf(ddelta, gap[k], dn_acc);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float col = 1.0f / (1.0f + expf(-rr[s * (C + 1) + c]));
                d_raw[(ray * S + s) * (C + 1) + c] = g[c] * w * col * (1.0f - col);
            }
            after += (double)A * (double)w;
        }
    }
    dn_acc = wsumf(dn_acc);
    if (lane == 0) d_dnorm[ray] += dn_acc;
}

// ------------------------------------------------------------------ rgb head backward, fused
// One pass over d_raw [rows, C+1] and the saved view-layer activations H9 [rows, 128] (528 B per row) does what
// heads_backward_kernel + rgb_head_wgrad_kernel + sum_samples_kernel do in three (2.6 KB per row, the fp32 dZ9 matrix written
// and re-read): dZ9 as a bf16 hi/lo tile matrix (for the dgrad chain and the weight-gradient contraction), the per-ray sum
// dvb [n, 128] (gradient of the view bias), its total s [128] (bias of the merged view step), and rgb_linear's weight / bias
// gradient.  A block of 256 threads = 16 row slots x 16 column chunks of 8 walks whole rays (their S rows are consecutive).
template <int C>
__global__ void __launch_bounds__(256) heads_fused_kernel(const float* __restrict__ d_raw, const float* __restrict__ h9,
                                                          const float* __restrict__ w_rgb, int64_t n_rays, int S, int64_t rows,
                                                          int64_t rows_pad, unsigned char* __restrict__ dz9_tiles, float* __restrict__ dvb,
                                                          float* __restrict__ s_views, float* __restrict__ dW, float* __restrict__ dB) {
    __shared__ float red[16][kHalf];
    const int slot = threadIdx.x >> 4, c8 = threadIdx.x & 15;
    float wr[C][8], accW[C][8], accB[C], tot[8];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        accB[c] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { wr[c][j] = w_rgb[c * kHalf + c8 * 8 + j]; accW[c][j] = 0.f; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tot[j] = 0.f;
    auto store_tile = [&](int64_t row, const float* v) {
        uint4 hi, lo;
        bwt::split8_bf16_pub(v, hi, lo);
        const int64_t tile = row / bwt::kTileRows;
        const int r = (int)(row % bwt::kTileRows);
        unsigned char* t = dz9_tiles + (size_t)tile * bwt::tile_bytes(kHalf) + (size_t)(c8 / 8) * bwt::kKbBytes + (size_t)r * 128 + ((uint32_t)((c8 & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(t) = hi;
        *reinterpret_cast<uint4*>(t + bwt::tile_part_bytes(kHalf)) = lo;
    };
    for (int64_t ray = blockIdx.x; ray < n_rays; ray += gridDim.x) {
        float sum[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] = 0.f;
        for (int sidx = slot; sidx < S; sidx += 16) {
            const int64_t row = ray * S + sidx;
            float g[C];
            if (C == 3) {
                const float4 q = *reinterpret_cast<const float4*>(d_raw + row * 4);
                g[0] = q.x; g[1 % C] = q.y; g[2 % C] = q.z;
            } else {
                g[0] = d_raw[row * (C + 1)];
            }
            const float4 a = *reinterpret_cast<const float4*>(h9 + row * kHalf + c8 * 8), b = *reinterpret_cast<const float4*>(h9 + row * kHalf + c8 * 8 + 4);
            const float h[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float x = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) { x = fmaf(g[c], wr[c][j], x); accW[c][j] = fmaf(g[c], h[j], accW[c][j]); }
                v[j] = (h[j] > 0.0f) ? x : 0.0f;
                sum[j] += v[j];
            }
            if (c8 == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) accB[c] += g[c];
            }
            store_tile(row, v);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { red[slot][c8 * 8 + j] = sum[j]; tot[j] += sum[j]; }
        __syncthreads();
        if (threadIdx.x < kHalf) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) t += red[q][threadIdx.x];
            dvb[ray * kHalf + threadIdx.x] = t;
        }
        __syncthreads();
    }
    // block totals: the merged view step's bias gradient s, then rgb_linear's weight and bias gradient
    auto block_reduce_add = [&](const float* vals, float* dst) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[slot][c8 * 8 + j] = vals[j];
        __syncthreads();
        if (threadIdx.x < kHalf) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) t += red[q][threadIdx.x];
            if (t != 0.0f) atomicAdd(dst + threadIdx.x, t);
        }
        __syncthreads();
    };
    if (s_views) block_reduce_add(tot, s_views);
#pragma unroll
    for (int c = 0; c < C; ++c) block_reduce_add(accW[c], dW + c * kHalf);
    if (c8 == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) red[slot][c] = accB[c];
    }
    __syncthreads();
    if (threadIdx.x < C) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) t += red[q][threadIdx.x];
        atomicAdd(dB + threadIdx.x, t);
    }
    // rows of the last tile (pair) beyond `rows` take part in the weight-gradient contraction: zeros
    if (blockIdx.x == 0) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int64_t row = rows + slot; row < rows_pad; row += 16) store_tile(row, z);
    }
}

// ------------------------------------------------------------------ views_linears.0: direction block and bias
// dW[j][256 + i] += sum_rays dvb[ray][j] * enc[ray][i] (i < 27),  dB[j] += sum_rays dvb[ray][j]; blockIdx.x = i (27 = bias)
// (role of net_tail_kernel: 128 threads = the 128 outputs j)
__device__ __forceinline__ void viewdir_wgrad_block(unsigned vblock, const float* __restrict__ dvb, const float* __restrict__ pe_dir /*[N,32]*/,
                                                    int64_t n_rays, float* __restrict__ dW /*[128, 283]*/, float* __restrict__ dB) {
    constexpr int64_t kRaysPerBlock = 128;
    const int j = threadIdx.x, i = (int)(vblock % (kDirCh + 1));
    const int64_t r0 = (int64_t)(vblock / (kDirCh + 1)) * kRaysPerBlock;
    const int64_t r1 = (r0 + kRaysPerBlock < n_rays) ? r0 + kRaysPerBlock : n_rays;
    float acc = 0.f;
#pragma unroll 4
    for (int64_t ray = r0; ray < r1; ++ray) acc = fmaf(dvb[ray * kHalf + j], i < kDirCh ? pe_dir[ray * 32 + i] : 1.0f, acc);
    if (i < kDirCh) atomicAdd(dW + (size_t)j * (kWidth + kDirCh) + kWidth + i, acc);
    else atomicAdd(dB + j, acc);
}

// ------------------------------------------------------------------ feature_linear + feature block of views_linears.0
// The forward pass runs them as one merged linear (common.cuh: wt9m), the backward pass contracts once, G = sum dZ9 (x) h7
// [128, 256] and s = sum dZ9 [128]; with feature = W_f h7 + b_f:
//   dW_views[j][f] += sum_k G[j][k] W_f[f][k] + s[j] b_f[f]      (blocks 0..127: j, thread f)
//   dW_f[f][k]     += sum_j W_v[j][f] G[j][k]                     (blocks 128..383: f, thread k)
//   dB_f[f]        += sum_j W_v[j][f] s[j]
// wt8[k][f] = W_f[f][k], wt9[f][j] = W_v[j][f] (the k-major copies of the weight cache).
__device__ __forceinline__ void views_feature_wgrad_block(unsigned vblock, const float* __restrict__ G, const float* __restrict__ s,
                                                          const float* __restrict__ wt8, const float* __restrict__ wt9,
                                                          const float* __restrict__ b_f, float* __restrict__ dW_views /*[128, 283]*/,
                                                          float* __restrict__ dW_f /*[256, 256]*/, float* __restrict__ dB_f) {
    const int t = (int)(vblock & 1u) * 128 + threadIdx.x;        // two 128-thread blocks per output row
    vblock >>= 1;
    if (vblock < kHalf) {
        const int j = (int)vblock, f = t;
        float acc = s[j] * b_f[f];
#pragma unroll 8
        for (int k = 0; k < kWidth; ++k) acc = fmaf(G[j * kWidth + k], wt8[(size_t)k * kWidth + f], acc);
        dW_views[(size_t)j * (kWidth + kDirCh) + f] += acc;
    } else {
        const int f = (int)vblock - kHalf, k = t;
        float acc = 0.f, bs = 0.f;
#pragma unroll 8
        for (int j = 0; j < kHalf; ++j) {
            const float wv = wt9[(size_t)f * kHalf + j];
            acc = fmaf(wv, G[j * kWidth + k], acc);
            bs = fmaf(wv, s[j], bs);
        }
        dW_f[(size_t)f * kWidth + k] += acc;
        if (k == 0) dB_f[f] += bs;
    }
}

// ------------------------------------------------------------------ view-direction branch
// per ray: encoding of the unit view direction (27 values, padded to 32), d enc = dvb * W_dir^T, d view
__device__ __forceinline__ void viewdir_backward_ray(int64_t ray, int lane, const float* __restrict__ view, const float* __restrict__ dvb,
                                                     const float* __restrict__ w_dir /*[27][128]*/, float* __restrict__ d_view /*[N,3] +=*/) {
    const float v[3] = {view[ray * 3], view[ray * 3 + 1], view[ray * 3 + 2]};
    float enc[kDirCh];
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
    float g4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) g4[q] = dvb[ray * kHalf + lane + 32 * q];
    float denc[kDirCh];
#pragma unroll
    for (int i = 0; i < kDirCh; ++i) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) a = fmaf(g4[q], w_dir[i * kHalf + lane + 32 * q], a);
        denc[i] = wsumf(a);
    }
    if (lane < 3) {
        float dv = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (lane != c) continue;
            dv = denc[c];
#pragma unroll
            for (int k = 0; k < kDirFreqs; ++k) {
                const float f = (float)(1 << k);
                dv += f * (enc[3 + 6 * k + 3 + c] * denc[3 + 6 * k + c] - enc[3 + 6 * k + c] * denc[3 + 6 * k + 3 + c]);
            }
        }
        d_view[ray * 3 + lane] += dv;
    }
}

// ------------------------------------------------------------------ point encoding + pts = o + d*z
// per ray: d_o += sum_s d_pts, d_d += sum_s z * d_pts with d_pts from d_pe through the saved sin/cos
__device__ __forceinline__ void pe_ray_backward_ray(int64_t ray, int lane, const float* __restrict__ pe, const float* __restrict__ d_pe,
                                                    const float* __restrict__ z, int S, float* __restrict__ d_o, float* __restrict__ d_d) {
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    for (int s = lane; s < S; s += 32) {
        const int64_t row = ray * S + s;
        const float4* p4 = reinterpret_cast<const float4*>(pe + row * kPtsChPad);
        const float4* g4 = reinterpret_cast<const float4*>(d_pe + row * kPtsChPad);
        float p[64], g[64];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 a = p4[i], b = g4[i];
            p[4 * i] = a.x; p[4 * i + 1] = a.y; p[4 * i + 2] = a.z; p[4 * i + 3] = a.w;
            g[4 * i] = b.x; g[4 * i + 1] = b.y; g[4 * i + 2] = b.z; g[4 * i + 3] = b.w;
        }
        const float zz = z[row];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float dx = g[c];
#pragma unroll
            for (int k = 0; k < kPtsFreqs; ++k)
                dx += (float)(1 << k) * (p[3 + 6 * k + 3 + c] * g[3 + 6 * k + c] - p[3 + 6 * k + c] * g[3 + 6 * k + 3 + c]);
            go[c] += dx;
            gd[c] = fmaf(zz, dx, gd[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { go[c] = wsumf(go[c]); gd[c] = wsumf(gd[c]); }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { d_o[ray * 3 + c] += go[c]; d_d[ray * 3 + c] += gd[c]; }
    }
}

// ------------------------------------------------------------------ what follows the tensor-core kernels of one network, in one launch
// Four independent pieces of work, told apart by block index (128 threads each; the first needs ~140 registers, three blocks per SM):
// the encoded-point gradients of a ray's samples summed into d o / d d (4 rays per block), the view-direction branch (4 rays per
// block), the direction block of views_linears.0 with its bias, and feature_linear + the feature block of views_linears.0 from
// the contraction G.
struct NetTail {
    const float* pe; const float* d_pe; const float* z; int S; float* g_o; float* g_d;          // A
    const float* view; const float* dvb; const float* w_dir; float* g_v;                       // B
    const float* pe_dir; float* dW_views; float* dB_views;                                     // C
    const float* G; const float* s; const float* wt8; const float* wt9; const float* b_f; float* dW_f; float* dB_f;   // D
    int64_t n;
    unsigned nA, nB, nC;
};
__global__ void __launch_bounds__(128) net_tail_kernel(const __grid_constant__ NetTail a) {
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    unsigned b = blockIdx.x;
    if (b < a.nA) {
        const int64_t ray = (int64_t)b * 4 + warp;
        if (ray < a.n) pe_ray_backward_ray(ray, lane, a.pe, a.d_pe, a.z, a.S, a.g_o, a.g_d);
        return;
    }
    b -= a.nA;
    if (b < a.nB) {
        const int64_t ray = (int64_t)b * 4 + warp;
        if (ray < a.n) viewdir_backward_ray(ray, lane, a.view, a.dvb, a.w_dir, a.g_v);
        return;
    }
    b -= a.nB;
    if (b < a.nC) { viewdir_wgrad_block(b, a.dvb, a.pe_dir, a.n, a.dW_views, a.dB_views); return; }
    views_feature_wgrad_block(b - a.nC, a.G, a.s, a.wt8, a.wt9, a.b_f, a.dW_views, a.dW_f, a.dB_f);
}

// ------------------------------------------------------------------ rays: ndc, viewdirs, R*dir, origin -> d poses
struct RaysBwdSeg { const float* poses; const int64_t* ray_idx; const float* remap; float* d_poses; int R, H, W; float fx, fy, cx, cy; int64_t off; };
struct RaysBwd { RaysBwdSeg seg[4]; int n_segs, ndc; int64_t n; const float *g_o, *g_d, *g_v, *g_dn; };
__global__ void rays_backward_kernel(const __grid_constant__ RaysBwd a) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = n < a.n;
    float* dst = nullptr;                       // this ray's pose gradient, [12] +=
    float gp[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) gp[i] = 0.f;
    if (live) {
        int si = 0;
#pragma unroll
        for (int i = 1; i < 4; ++i) if (i < a.n_segs && n >= a.seg[i].off) si = i;
        const RaysBwdSeg& sg = a.seg[si];
        const float* poses = sg.poses; const float* remap = sg.remap;
        const int R = sg.R, H = sg.H, W = sg.W, ndc = a.ndc;
        const float fx = sg.fx, fy = sg.fy, cx = sg.cx, cy = sg.cy;
        const float *g_o_in = a.g_o, *g_d_in = a.g_d, *g_v_in = a.g_v, *g_dn_in = a.g_dn;
        const int64_t nl = n - sg.off;
        const int p = (int)(nl / R);
        dst = sg.d_poses + (size_t)p * 12;
        const int64_t pix = sg.ray_idx[nl % R];
        float fi = (float)(pix % W), fj = (float)(pix / W);
        if (remap) { fi = remap[2 * pix]; fj = remap[2 * pix + 1]; }
        const float* c2w = poses + (size_t)p * 12;
        const float dir[3] = {(fi - cx) / fx, -(fj - cy) / fy, -1.0f};
        float d[3], o[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            d[r] = dir[0] * c2w[r * 4] + dir[1] * c2w[r * 4 + 1] + dir[2] * c2w[r * 4 + 2];
            o[r] = c2w[r * 4 + 3];
        }
        float gO[3] = {g_o_in[n * 3], g_o_in[n * 3 + 1], g_o_in[n * 3 + 2]};
        float gD[3] = {g_d_in[n * 3], g_d_in[n * 3 + 1], g_d_in[n * 3 + 2]};
        const float gV[3] = {g_v_in[n * 3], g_v_in[n * 3 + 1], g_v_in[n * 3 + 2]};
        const float g_dn = g_dn_in[n];
        float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};   // w.r.t. the pre-ndc origin / direction
        if (ndc) {
            const float near = 1.0f;
            const float t = -(near + o[2]) / d[2];
            float op[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) op[r] = o[r] + t * d[r];
            const float sx = -1.0f / ((float)W / (2.0f * fx)), sy = -1.0f / ((float)H / (2.0f * fx));
            // post-ndc direction (for the |rays_d| term of raw2output, model/nerf.py:124)
            const float D[3] = {sx * (d[0] / d[2] - op[0] / op[2]), sy * (d[1] / d[2] - op[1] / op[2]), -2.0f * near / op[2]};
            const float Dn = sqrtf(D[0] * D[0] + D[1] * D[1] + D[2] * D[2]);
#pragma unroll
            for (int r = 0; r < 3; ++r) gD[r] += g_dn * D[r] / Dn;
            const float i2 = 1.0f / op[2], i22 = i2 * i2;
            float gop[3];
            gop[0] = sx * i2 * (gO[0] - gD[0]);
            gop[1] = sy * i2 * (gO[1] - gD[1]);
            gop[2] = i22 * (-sx * op[0] * gO[0] - sy * op[1] * gO[1] - 2.0f * near * gO[2]
                            + sx * op[0] * gD[0] + sy * op[1] * gD[1] + 2.0f * near * gD[2]);
            const float id2 = 1.0f / d[2];
            gd[0] = sx * gD[0] * id2;
            gd[1] = sy * gD[1] * id2;
            gd[2] = -(sx * d[0] * gD[0] + sy * d[1] * gD[1]) * id2 * id2;
            float gt = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) { go[r] = gop[r]; gd[r] += t * gop[r]; gt += gop[r] * d[r]; }
            go[2] += -gt * id2;
            gd[2] += gt * (near + o[2]) * id2 * id2;
        } else {
            const float Dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
            for (int r = 0; r < 3; ++r) { go[r] = gO[r]; gd[r] = gD[r] + g_dn * d[r] / Dn; }
        }
        // viewdirs = d / |d| (pre-ndc, model/nerf.py:272-275)
        const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const float v[3] = {d[0] / nrm, d[1] / nrm, d[2] / nrm};
        const float vg = v[0] * gV[0] + v[1] * gV[1] + v[2] * gV[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) gd[r] += (gV[r] - v[r] * vg) / nrm;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            gp[r * 4 + 0] = gd[r] * dir[0]; gp[r * 4 + 1] = gd[r] * dir[1]; gp[r * 4 + 2] = gd[r] * dir[2];
            gp[r * 4 + 3] = go[r];
        }
    }
    // warp-level reduction when the whole warp works on one pose (the common case in pose-major order); lane 0 is live whenever
    // any lane is
    float* dst0 = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), 0));
    const bool uniform = __all_sync(0xffffffffu, !live || dst == dst0);
    const bool any_live = __any_sync(0xffffffffu, live);
    if (uniform) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const float s = wsumf(gp[i]);
            if ((threadIdx.x % 32) == 0 && any_live) atomicAdd(dst0 + i, s);
        }
    } else if (live) {
#pragma unroll
        for (int i = 0; i < 12; ++i) atomicAdd(dst + i, gp[i]);
    }
}

}  // namespace

// ------------------------------------------------------------------ per-network MLP backward
// d_raw [rows, C+1] -> parameter gradients (PyTorch layouts, accumulated) and d_pe [rows,64], dvb [n,128].
struct BwdBuffers {
    float *d_raw, *d_pe, *dz9, *dvb, *pe_dir;
    unsigned char* dz_tiles;     // 8 bf16 tile matrices of width 256: dZ0..dZ7
    float* g_views;              // [128, 256] fp32: sum_rows dZ9 (x) h7, from which both merged linears get their gradients
    float* s_views;              // [128] fp32: sum_rows dZ9
    unsigned char* dz9_tiles;    // bf16 tile matrix of width 128
    int64_t tiles;
};

// dgrad B operands of one network, in the order the backward pass uses them
struct DgImage { int step, k0, N, K; };
static const DgImage kDgImages[10] = {
    {10, 0, 256, 128},                                 // views_linears.0 (feature block) . feature_linear, merged: dZ9 -> d h7
    {7, 0, 256, 256}, {6, 0, 256, 256},
    {5, kPtsChPad, 256, 256}, {5, 0, 64, 256},         // pts_linears.5: h4 block, encoded-points block
    {4, 0, 256, 256}, {3, 0, 256, 256}, {2, 0, 256, 256}, {1, 0, 256, 256},
    {0, 0, 64, 256},                                   // pts_linears.0: encoded points
};
static size_t dg_image_offset(int i) {
    size_t off = 0;
    for (int j = 0; j < i; ++j) off += bwt::dgrad_image_bytes(kDgImages[j].N, kDgImages[j].K);
    return off;
}
size_t dgrad_images_bytes() { return dg_image_offset(10); }

static int pack_dgrad_images(bnrf_ctx* ctx, int net, cudaStream_t st) {
    // only the operand image of the configured dgrad path is rebuilt (this runs after every optimiser step)
    NetParams& np = ctx->net[net];
    int rc = BNRF_OK;
    if (ctx->cfg.gemm_mode == BNRF_GEMM_TC) {
        rc = pack_dgrad_chain_pair_stream(ctx, net, st);
    } else if (ctx->cfg.gemm_mode == BNRF_GEMM_TC_1CTA) {
        rc = pack_dgrad_chain_stream(ctx, net, st);
    } else {
        bwt::DgImageTable t{};
        for (int i = 0; i < 10; ++i) {
            const DgImage& d = kDgImages[i];
            const float* w = d.step == 10 ? np.wt9m : np.wt[d.step];
            t.seg[t.n++] = bwt::DgImageSeg{w + (size_t)d.k0 * d.K, np.dg_img + dg_image_offset(i), d.N, d.K};
        }
        rc = bwt::pack_dgrad_images(ctx, t, st);
    }
    if (rc) return rc;
    np.dg_dirty = false;
    return BNRF_OK;
}

static int mlp_backward(bnrf_ctx* ctx, int net, int64_t n, int S, const ActPtrs& acts, const BwdBuffers& w,
                        float* const* dW, float* const* dB, cudaStream_t st) {
    NetParams& np = ctx->net[net];
    const int C = ctx->cfg.channels;
    const int64_t rows = n * S;
    const int tiles = (int)bwt::tile_count(rows);
    const size_t mat = (size_t)w.tiles * bwt::tile_bytes(kWidth);                       // one gradient tile matrix
    const size_t hmat = (size_t)acts.t_alloc * bwt::tile_bytes(kWidth);                 // one activation tile matrix
    auto DZ = [&](int l) { return w.dz_tiles + (size_t)l * mat; };                      // l = 0..7
    auto H = [&](int l) { return acts.h_tiles + (size_t)l * hmat; };                    // l = 0..7
    int rc;
    if (np.dg_dirty && (rc = pack_dgrad_images(ctx, net, st))) return rc;
    auto dgrad = [&](const unsigned char* A, int K, int img, int epi, const unsigned char* mask, unsigned char* out_tiles, float* out_f32,
                     const float* r_row, int64_t r_stride, const float* r_col) {
        bwt::DgradArgs a{};
        a.a_tiles = A; a.K = K; a.b_img = np.dg_img + dg_image_offset(img); a.N = kDgImages[img].N; a.epi = epi;
        a.mask_tiles = mask; a.r_row = r_row; a.r_stride = r_stride; a.r_col = r_col;
        a.out_tiles = out_tiles; a.out_f32 = out_f32; a.ld_out = kPtsChPad; a.rows = rows; a.tiles = tiles;
        return bwt::launch_tile_dgrad(ctx, a, st);
    };

    // ---- heads: rgb_linear (128 -> C) backward, the ReLU of the view layer, the per-ray view-bias gradient: ONE pass ----
    BNRF_CUDA(ctx, cudaMemsetAsync(w.g_views, 0, (size_t)(kHalf * kWidth + kHalf) * sizeof(float), st));
    {
        const int64_t rows_pad = bwt::tile_alloc(rows) * bwt::kTileRows;  // incl. the pad tile of an odd tile count (CTA pairs walk tiles in pairs)
        const int64_t cap = 4 * (int64_t)ctx->sm_count;
        const unsigned grid = (unsigned)(n < cap ? n : cap);
        if (C == 3) heads_fused_kernel<3><<<grid, 256, 0, st>>>(w.d_raw, acts.h9_f32, np.w_rgb, n, S, rows, rows_pad, w.dz9_tiles, w.dvb, w.s_views,
                                                                dW[BNRF_L_RGB], dB[BNRF_L_RGB]);
        else heads_fused_kernel<1><<<grid, 256, 0, st>>>(w.d_raw, acts.h9_f32, np.w_rgb, n, S, rows, rows_pad, w.dz9_tiles, w.dvb, w.s_views,
                                                         dW[BNRF_L_RGB], dB[BNRF_L_RGB]);
        BNRF_LAUNCH_CHECK(ctx);
    }
    // ---- dgrad chain: dZ9 -> d feature -> dZ7 -> ... -> dZ0 -> d encoding ----
    if (ctx->cfg.gemm_mode == BNRF_GEMM_TC) {
        // one launch per network: the gradient of a tile stays on the SM across all linears; CTA pairs, each CTA streams half
        // of every weight tile (dgrad_chain2.cu)
        if ((rc = launch_dgrad_chain_pair(ctx, net, w.dz9_tiles, acts.mask_bits, acts.t_alloc, w.d_raw + C, C + 1, rows, w.tiles,
                                          w.dz_tiles, w.d_pe, st))) return rc;
    } else if (ctx->cfg.gemm_mode != BNRF_GEMM_TC_PER_LINEAR) {
        // the same chain on single CTAs (dgrad_chain.cu; BNRF_GEMM_TC_1CTA)
        if ((rc = launch_dgrad_chain(ctx, net, w.dz9_tiles, acts.mask_bits, acts.t_alloc, w.d_raw + C, C + 1, rows, w.tiles, w.dz_tiles,
                                     w.d_pe, st))) return rc;
    } else {
        if ((rc = dgrad(w.dz9_tiles, kHalf, 0, bwt::DG_TILE_MASKED, H(7), DZ(7), nullptr, w.d_raw + C, C + 1, np.w_alpha))) return rc;   // merged step + alpha_linear
        if ((rc = dgrad(DZ(7), kWidth, 1, bwt::DG_TILE_MASKED, H(6), DZ(6), nullptr, nullptr, 0, nullptr))) return rc;
        if ((rc = dgrad(DZ(6), kWidth, 2, bwt::DG_TILE_MASKED, H(5), DZ(5), nullptr, nullptr, 0, nullptr))) return rc;
        if ((rc = dgrad(DZ(5), kWidth, 3, bwt::DG_TILE_MASKED, H(4), DZ(4), nullptr, nullptr, 0, nullptr))) return rc;              // cat([pe, h4]) (model/nerf.py:98)
        if ((rc = dgrad(DZ(5), kWidth, 4, bwt::DG_F32_STORE, nullptr, nullptr, w.d_pe, nullptr, 0, nullptr))) return rc;
        for (int l = 4; l >= 1; --l)
            if ((rc = dgrad(DZ(l), kWidth, 9 - l, bwt::DG_TILE_MASKED, H(l - 1), DZ(l - 1), nullptr, nullptr, 0, nullptr))) return rc;
        if ((rc = dgrad(DZ(0), kWidth, 9, bwt::DG_F32_ACCUM, nullptr, nullptr, w.d_pe, nullptr, 0, nullptr))) return rc;
    }
    // ---- weight / bias gradients: the eight wide contractions on CTA pairs (wgrad_pair.cu), the two encoded-points blocks on
    //      single CTAs (bwd_tiles.cu); gemm_mode tc_linear / tc_chain1 keep everything on the single-CTA kernel (cross-check) ----
    const bool pair = ctx->cfg.gemm_mode == BNRF_GEMM_TC;
    bwt::WgradParams p{};
    p.tiles = tiles; p.rows = rows;
    bwt::WgradPairParams pp{};
    pp.tiles = tiles; pp.rows = rows;
    auto job = [&](const unsigned char* dz, int M, const unsigned char* h, int N, float* dWl, int ldw, int col0, int n_valid, float* dBl) {
        bwt::WgradJob& j = p.job[p.n_jobs++];
        j.dz_tiles = dz; j.M = M; j.h_tiles = h; j.N = N; j.dW = dWl; j.ldw = ldw; j.col0 = col0; j.n_valid = n_valid; j.dB = dBl;
        return &j;
    };
    auto wide = [&](const unsigned char* dz, const unsigned char* h, float* dWl, int ldw, int col0, float* dBl) {
        if (!pair) { job(dz, kWidth, h, kWidth, dWl, ldw, col0, kWidth, dBl); return; }
        bwt::WgradPairJob& j = pp.job[pp.n_jobs++];
        j.a_tiles = dz; j.b_tiles = h; j.NB = kWidth; j.dW = dWl; j.ldw = ldw; j.col0 = col0; j.n_valid = kWidth; j.dB = dBl;
    };
    // the bias of layer 5 rides on its wide block when that runs on the pair kernel (bias off the tensor core), else on the narrow one
    // (BARF c2f: the saved encoding is unweighted; the weight-gradient columns of its channels take the channel weights)
    job(DZ(0), kWidth, acts.pe_tiles, kPtsChPad, dW[0], kPtsCh, 0, kPtsCh, dB[0])->col_scale = ctx->enc_scaled ? ctx->enc_scale : nullptr;
    job(DZ(5), kWidth, acts.pe_tiles, kPtsChPad, dW[5], kPtsCh + kWidth, 0, kPtsCh, pair ? nullptr : dB[5])->col_scale = ctx->enc_scaled ? ctx->enc_scale : nullptr;
    for (int l = 1; l < 8; ++l) {
        if (l == 5) wide(DZ(5), H(4), dW[5], kPtsCh + kWidth, kPtsCh, pair ? dB[5] : nullptr);
        else wide(DZ(l), H(l - 1), dW[l], kWidth, 0, dB[l]);
    }
    // feature_linear and the feature block of views_linears.0 (merged in the forward pass): ONE contraction G = sum_rows dZ9 (x) h7
    // (s = sum_rows dZ9 comes from heads_fused_kernel); views_feature_wgrad_kernel below turns them into both gradients.
    // alpha_linear reads the same h7 slices.
    if (pair) {
        bwt::WgradPairJob& j = pp.job[pp.n_jobs++];
        j.a_tiles = H(7); j.b_tiles = w.dz9_tiles; j.NB = kHalf; j.dW = w.g_views; j.ldw = kWidth; j.transposed = 1;
        j.wrow = w.d_raw + C; j.wrow_base = w.d_raw; j.wrow_col = C; j.wrow_stride = C + 1; j.dWv = dW[BNRF_L_ALPHA]; j.dBv = dB[BNRF_L_ALPHA];
        if ((rc = bwt::launch_tile_wgrad_pair(ctx, pp, st))) return rc;
    } else {
        bwt::WgradJob* j = job(w.dz9_tiles, kHalf, H(7), kWidth, w.g_views, kWidth, 0, kWidth, nullptr);
        j->wrow = w.d_raw + C; j->wrow_stride = C + 1; j->dWv = dW[BNRF_L_ALPHA]; j->dBv = dB[BNRF_L_ALPHA];
    }
    if ((rc = bwt::launch_tile_wgrad(ctx, p, st))) return rc;
    return BNRF_OK;        // net_tail_kernel (render_backward_impl) turns g_views / s_views into the feature_linear / views_linears.0 gradients
}

// ------------------------------------------------------------------ saved-tensor and workspace carve-ups
SavedLayout carve_saved(const bnrf_cfg& c, int64_t n, void* base) {
    SavedLayout s;
    size_t off = 0;
    auto take_bytes = [&](size_t bytes) {
        char* p = base ? static_cast<char*>(base) + off : nullptr;
        off += (bytes + 1023) / 1024 * 1024;            // tile matrices are read by cp.async.bulk: keep everything 1 KB aligned
        return p;
    };
    auto take = [&](size_t floats) { return reinterpret_cast<float*>(take_bytes(floats * sizeof(float))); };
    const int Sc = c.n_samples, Sf = c.n_samples + c.n_importance;
    const bool fine = c.n_importance > 0;
    auto acts = [&](int64_t rows) {
        ActPtrs a{};
        a.t_alloc = rows > 0 ? bwt::tile_alloc(rows) : 0;
        a.pe_f32 = take(rows * kPtsChPad);
        a.h9_f32 = take(rows * kHalf);
        a.pe_tiles = reinterpret_cast<unsigned char*>(take_bytes((size_t)a.t_alloc * bwt::tile_bytes(kPtsChPad)));
        a.h_tiles = reinterpret_cast<unsigned char*>(take_bytes(8 * (size_t)a.t_alloc * bwt::tile_bytes(kWidth)));
        a.mask_bits = reinterpret_cast<unsigned char*>(take_bytes(8 * (size_t)a.t_alloc * 4096));
        return a;
    };
    s.o = take(n * 3); s.d = take(n * 3); s.view = take(n * 3); s.pe_dir = take(n * 32);
    s.z_c = take(n * Sc); s.raw_c = take(n * Sc * (c.channels + 1)); s.sig_c = take(n * Sc);
    s.acts_c = acts(n * Sc);
    s.z_f = take(fine ? n * Sf : 0); s.raw_f = take(fine ? n * Sf * (c.channels + 1) : 0); s.sig_f = take(fine ? n * Sf : 0);
    s.acts_f = acts(fine ? n * Sf : 0);
    s.bytes = off;
    return s;
}

struct BwdWorkspace { BwdBuffers b; float *g_o, *g_d, *g_v, *g_dn; size_t bytes; };
static BwdWorkspace carve_bwd(const bnrf_cfg& c, int64_t n, void* base) {
    BwdWorkspace w;
    size_t off = 0;
    auto take_bytes = [&](size_t bytes) {
        char* p = base ? static_cast<char*>(base) + off : nullptr;
        off += (bytes + 1023) / 1024 * 1024;
        return p;
    };
    auto take = [&](size_t floats) { return reinterpret_cast<float*>(take_bytes(floats * sizeof(float))); };
    const int64_t rows = n * (c.n_samples + c.n_importance);
    w.b.tiles = bwt::tile_alloc(rows);
    w.b.d_raw = take(rows * (c.channels + 1));
    w.b.dz_tiles = reinterpret_cast<unsigned char*>(take_bytes(8 * (size_t)w.b.tiles * bwt::tile_bytes(kWidth)));
    w.b.g_views = take(kHalf * kWidth + kHalf); w.b.s_views = w.b.g_views + kHalf * kWidth;
    w.b.dz9_tiles = reinterpret_cast<unsigned char*>(take_bytes((size_t)w.b.tiles * bwt::tile_bytes(kHalf)));
    w.b.d_pe = take(rows * kPtsChPad); w.b.dz9 = nullptr;
    w.b.dvb = take(n * kHalf); w.b.pe_dir = nullptr;          // the encoded view directions come with the saved tensors
    w.g_o = take(n * 3); w.g_d = take(n * 3); w.g_v = take(n * 3); w.g_dn = take(n);
    w.bytes = off;
    return w;
}

}  // namespace bnrf

using namespace bnrf;

extern "C" {

size_t bnrf_saved_bytes(const bnrf_ctx* ctx, int64_t n_rays) {
    if (!ctx || n_rays <= 0) return 0;
    return carve_saved(ctx->cfg, n_rays, nullptr).bytes;
}

size_t bnrf_backward_workspace_bytes(const bnrf_ctx* ctx, int64_t n_rays) {
    if (!ctx || n_rays <= 0) return 0;
    return carve_bwd(ctx->cfg, n_rays, nullptr).bytes;
}

static int render_backward_impl(bnrf_ctx* ctx, const bnrf_render_seg* segs, int n_segs, const float* d_rgb_map, const float* d_rgb0,
                                const void* saved, size_t saved_bytes, const bnrf_param_grads* grads_coarse,
                                const bnrf_param_grads* grads_fine, float* const* d_poses, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    if (!segs || n_segs <= 0 || n_segs > 4 || !saved || !workspace || !d_poses) return fail(ctx, BNRF_ERR_ARG, "render_backward: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const bnrf_cfg& c = ctx->cfg;
    int64_t n = 0;
    for (int i = 0; i < n_segs; ++i) {
        if (!segs[i].poses || !segs[i].ray_idx || !d_poses[i] || segs[i].P <= 0 || segs[i].R <= 0) return fail(ctx, BNRF_ERR_ARG, "render_backward: bad segment %d", i);
        n += (int64_t)segs[i].P * segs[i].R;
    }
    const bool fine = c.n_importance > 0;
    const int Sc = c.n_samples, Sf = c.n_samples + c.n_importance;
    const SavedLayout s = carve_saved(c, n, const_cast<void*>(saved));
    if (saved_bytes < s.bytes) return fail(ctx, BNRF_ERR_STATE, "render_backward: saved buffer %zu < %zu bytes", saved_bytes, s.bytes);
    BwdWorkspace w = carve_bwd(c, n, workspace);
    if (workspace_bytes < w.bytes) return fail(ctx, BNRF_ERR_STATE, "render_backward: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    // with a fine pass, rgb_map comes from the fine network and rgb0 from the coarse one (model/nerf.py:319-343)
    const float* g_fine = fine ? d_rgb_map : nullptr;
    const float* g_coarse = fine ? d_rgb0 : d_rgb_map;
    if ((g_fine && !grads_fine) || (g_coarse && !grads_coarse)) return fail(ctx, BNRF_ERR_ARG, "render_backward: gradient tables missing");
    // the four per-ray gradient accumulators are carved one after the other (carve_bwd): one memset node
    BNRF_CUDA(ctx, cudaMemsetAsync(w.g_o, 0, (size_t)(reinterpret_cast<char*>(w.g_dn + n) - reinterpret_cast<char*>(w.g_o)), st));
    int rc;
    for (int net = 1; net >= 0; --net) {
        const float* g = net ? g_fine : g_coarse;
        if (!g) continue;
        const int S = net ? Sf : Sc;
        const float *raw = net ? s.raw_f : s.raw_c, *z = net ? s.z_f : s.z_c, *sig = net ? s.sig_f : s.sig_c;
        const ActPtrs& acts = net ? s.acts_f : s.acts_c;
        const bnrf_param_grads* pg = net ? grads_fine : grads_coarse;
        const unsigned grid = (unsigned)ceil_div(n, kWarps);
        if (c.channels == 3) composite_backward_kernel<3><<<grid, 32 * kWarps, 0, st>>>(raw, z, sig, s.d, g, n, S, w.b.d_raw, w.g_dn);
        else composite_backward_kernel<1><<<grid, 32 * kWarps, 0, st>>>(raw, z, sig, s.d, g, n, S, w.b.d_raw, w.g_dn);
        BNRF_LAUNCH_CHECK(ctx);
        if ((rc = mlp_backward(ctx, net, n, S, acts, w.b, pg->weights, pg->biases, st))) return rc;
        {   // everything after the tensor-core kernels of this network, one launch
            const NetParams& np = ctx->net[net];
            NetTail t{};
            t.pe = acts.pe_f32; t.d_pe = w.b.d_pe; t.z = z; t.S = S; t.g_o = w.g_o; t.g_d = w.g_d;
            t.view = s.view; t.dvb = w.b.dvb; t.w_dir = np.w_dir; t.g_v = w.g_v;
            t.pe_dir = s.pe_dir; t.dW_views = pg->weights[BNRF_L_VIEWS]; t.dB_views = pg->biases[BNRF_L_VIEWS];
            t.G = w.b.g_views; t.s = w.b.s_views; t.wt8 = np.wt[8]; t.wt9 = np.wt[9]; t.b_f = np.bias[8];
            t.dW_f = pg->weights[BNRF_L_FEATURE]; t.dB_f = pg->biases[BNRF_L_FEATURE];
            t.n = n; t.nA = t.nB = (unsigned)ceil_div(n, 4); t.nC = (unsigned)((kDirCh + 1) * ceil_div(n, 128));
            net_tail_kernel<<<t.nA + t.nB + t.nC + 2 * (kHalf + kWidth), 128, 0, st>>>(t);
            BNRF_LAUNCH_CHECK(ctx);
        }
        // the fine network is done first: a data-parallel caller may start reducing its gradients while the coarse network's
        // backward pass runs (bnrf_wait_fine_gradients)
        if (net == 1) BNRF_CUDA(ctx, cudaEventRecord(ctx->fine_grads_done, st));
    }
    {
        RaysBwd a{};
        int64_t off = 0;
        for (int i = 0; i < n_segs; ++i) {
            const bnrf_render_seg& sg = segs[i];
            a.seg[i] = RaysBwdSeg{sg.poses, sg.ray_idx, sg.remap, d_poses[i], sg.R, sg.H, sg.W, sg.K[0], sg.K[4], sg.K[2], sg.K[5], off};
            off += (int64_t)sg.P * sg.R;
        }
        a.n_segs = n_segs; a.ndc = c.ndc; a.n = off; a.g_o = w.g_o; a.g_d = w.g_d; a.g_v = w.g_v; a.g_dn = w.g_dn;
        rays_backward_kernel<<<(unsigned)ceil_div(off, 128), 128, 0, st>>>(a);
        BNRF_LAUNCH_CHECK(ctx);
    }
    return BNRF_OK;
}

int bnrf_wait_fine_gradients(bnrf_ctx* ctx, void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    BNRF_CUDA(ctx, cudaStreamWaitEvent((cudaStream_t)stream, ctx->fine_grads_done, 0));
    return BNRF_OK;
}

int bnrf_render_backward(bnrf_ctx* ctx, const float* poses, const int64_t* ray_idx, int P, int R, int H, int W,
                         const float* K, const float* remap, const float* d_rgb_map, const float* d_rgb0,
                         const void* saved, size_t saved_bytes, const bnrf_param_grads* grads_coarse,
                         const bnrf_param_grads* grads_fine, float* d_poses, void* workspace, size_t workspace_bytes,
                         void* stream) {
    if (!ctx) return BNRF_ERR_ARG;
    if (!K || !d_poses) return fail(ctx, BNRF_ERR_ARG, "render_backward: bad argument");
    bnrf_render_seg sg{};
    sg.poses = poses; sg.ray_idx = ray_idx; sg.P = P; sg.R = R; sg.H = H; sg.W = W; sg.remap = remap;
    memcpy(sg.K, K, sizeof(sg.K));
    float* dp[1] = {d_poses};
    return render_backward_impl(ctx, &sg, 1, d_rgb_map, d_rgb0, saved, saved_bytes, grads_coarse, grads_fine, dp, workspace, workspace_bytes, stream);
}

int bnrf_render_backward_multi(bnrf_ctx* ctx, const bnrf_render_seg* segs, int n_segs, const float* d_rgb_map, const float* d_rgb0,
                               const void* saved, size_t saved_bytes, const bnrf_param_grads* grads_coarse,
                               const bnrf_param_grads* grads_fine, float* const* d_poses, void* workspace, size_t workspace_bytes,
                               void* stream) {
    return render_backward_impl(ctx, segs, n_segs, d_rgb_map, d_rgb0, saved, saved_bytes, grads_coarse, grads_fine, d_poses, workspace,
                                workspace_bytes, stream);
}

}  // extern "C"
