// Camera-response tone mappers (f4): ColorToneMapper / LuminanceToneMapper of model/component.py:38-149 with input_type "Gray",
// the only variant model/optimize.py:15-20 constructs -- y = sigmoid(L_{h+1}(relu(L_h(... relu(L_0(x)))))) with L_0: 1 -> width,
// `hidden` layers width -> width, L_{h+1}: width -> 1, applied element-wise to a rendered [N, 1] tensor (train.py:176-192,
// run_nerf_helpers.py:125-126,152-153).  Forward and backward as one launch each: a warp owns an element, lanes own hidden units.
//
// The work is tiny (N <= a few 10^4 elements x width 128): the kernels are written for latency (everything of an element stays in
// the warp's registers / shared memory), not for a roofline.
#include "common.cuh"

namespace bnrf {
namespace {

constexpr int kMaxWidth = 256, kMaxHidden = 4, kWarps = 8;

struct CrfParams {
    const float* w0; const float* b0;                     // [width], [width]        Linear(1, width)
    const float* wh[kMaxHidden]; const float* bh[kMaxHidden];   // [width, width] (out, in), [width]
    const float* w1; const float* b1;                     // [width], [1]            Linear(width, 1)
    int width, hidden;
};
struct CrfGrads {
    float* w0; float* b0; float* wh[kMaxHidden]; float* bh[kMaxHidden]; float* w1; float* b1;
};

__device__ inline float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// activations of one element: act[l][j], l = 0 .. hidden (post-ReLU outputs of L_0 .. L_hidden), in the warp's shared memory
__device__ inline float crf_forward_one(const CrfParams& p, float x, float* act, int lane) {
    const int W = p.width;
    for (int j = lane; j < W; j += 32) act[j] = fmaxf(fmaf(p.w0[j], x, p.b0[j]), 0.0f);
    __syncwarp();
    for (int l = 0; l < p.hidden; ++l) {
        const float* in = act + l * W;
        float* out = act + (l + 1) * W;
        for (int k = lane; k < W; k += 32) {
            const float* row = p.wh[l] + (size_t)k * W;
            float a = p.bh[l][k];
            for (int j = 0; j < W; ++j) a = fmaf(row[j], in[j], a);
            out[k] = fmaxf(a, 0.0f);
        }
        __syncwarp();
    }
    const float* last = act + p.hidden * W;
    float s = 0.f;
    for (int j = lane; j < W; j += 32) s = fmaf(p.w1[j], last[j], s);
    return wsum(s) + p.b1[0];
}

__global__ void __launch_bounds__(32 * kWarps) crf_forward_kernel(const CrfParams p, const float* __restrict__ x, int64_t n, float* __restrict__ y) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    float* act = sm + (size_t)warp * (p.hidden + 1) * p.width;
    for (int64_t e = (int64_t)blockIdx.x * kWarps + warp; e < n; e += (int64_t)gridDim.x * kWarps) {
        const float raw = crf_forward_one(p, x[e], act, lane);
        if (lane == 0) y[e] = 1.0f / (1.0f + expf(-raw));
        __syncwarp();
    }
}

// g = d loss / d y.  dx[e] is written; parameter gradients are ADDED (per-warp partial sums, flushed with atomics at the end;
// the hidden layers' [width, width] gradients go through a block-shared accumulator).
__global__ void __launch_bounds__(32 * kWarps) crf_backward_kernel(const CrfParams p, const CrfGrads gr, const float* __restrict__ x,
                                                                  const float* __restrict__ g, int64_t n, float* __restrict__ dx) {
    extern __shared__ float sm[];
    const int W = p.width, H = p.hidden;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    float* act = sm + (size_t)warp * (2 * H + 2) * W;          // [H+1][W] activations, then [H+1][W] deltas
    float* del = act + (size_t)(H + 1) * W;
    float* acc_h = sm + (size_t)kWarps * (2 * H + 2) * W;        // [H][W][W] block-shared accumulator of the hidden weight gradients
    for (int i = threadIdx.x; i < H * W * W; i += blockDim.x) acc_h[i] = 0.f;
    __syncthreads();
    constexpr int Q = kMaxWidth / 32;
    float gw0[Q], gb0[Q], gw1[Q], gb1 = 0.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) { gw0[q] = 0.f; gb0[q] = 0.f; gw1[q] = 0.f; }
    for (int64_t e = (int64_t)blockIdx.x * kWarps + warp; e < n; e += (int64_t)gridDim.x * kWarps) {
        const float xe = x[e];
        const float raw = crf_forward_one(p, xe, act, lane);
        const float ye = 1.0f / (1.0f + expf(-raw));
        const float d_raw = g[e] * ye * (1.0f - ye);
        gb1 += d_raw;
        // output layer: d last[j] = d_raw * w1[j] (through the ReLU of the layer that produced last)
        const float* last = act + (size_t)H * W;
        float* dl = del + (size_t)H * W;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int j = lane + 32 * q;
            if (j < W) {
                gw1[q] = fmaf(d_raw, last[j], gw1[q]);
                dl[j] = last[j] > 0.0f ? d_raw * p.w1[j] : 0.0f;          // delta of the pre-activation of layer H
            }
        }
        __syncwarp();
        for (int l = H - 1; l >= 0; --l) {                               // hidden layer l: out = relu(Wh[l] in + bh[l]), in = act[l], out = act[l+1]
            const float* in = act + (size_t)l * W;
            const float* dout = del + (size_t)(l + 1) * W;
            float* din = del + (size_t)l * W;
            float* accw = acc_h + (size_t)l * W * W;
            for (int k = 0; k < W; ++k) {                                // weight / bias gradient: dW[k][j] += dout[k] in[j]
                const float dk = dout[k];
                if (dk != 0.0f)
                    for (int j = lane; j < W; j += 32) atomicAdd(accw + (size_t)k * W + j, dk * in[j]);
            }
            for (int j = lane; j < W; j += 32) {                         // d in[j] = sum_k dout[k] Wh[k][j], through in's own ReLU
                float a = 0.f;
                for (int k = 0; k < W; ++k) a = fmaf(dout[k], p.wh[l][(size_t)k * W + j], a);
                din[j] = in[j] > 0.0f ? a : 0.0f;
            }
            __syncwarp();
            if (lane == 0)
                for (int k = 0; k < W; ++k) if (dout[k] != 0.0f) atomicAdd(gr.bh[l] + k, dout[k]);
        }
        // input layer: pre_0[j] = w0[j] x + b0[j]
        const float* d0 = del;
        float dxe = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int j = lane + 32 * q;
            if (j < W) {
                gw0[q] = fmaf(d0[j], xe, gw0[q]);
                gb0[q] += d0[j];
                dxe = fmaf(d0[j], p.w0[j], dxe);
            }
        }
        dxe = wsum(dxe);
        if (lane == 0) dx[e] = dxe;
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int j = lane + 32 * q;
        if (j < W) {
            if (gw0[q] != 0.0f) atomicAdd(gr.w0 + j, gw0[q]);
            if (gb0[q] != 0.0f) atomicAdd(gr.b0 + j, gb0[q]);
            if (gw1[q] != 0.0f) atomicAdd(gr.w1 + j, gw1[q]);
        }
    }
    if (lane == 0 && gb1 != 0.0f) atomicAdd(gr.b1, gb1);
    __syncthreads();
    for (int i = threadIdx.x; i < H * W * W; i += blockDim.x) {
        const float v = acc_h[i];
        if (v != 0.0f) atomicAdd(gr.wh[i / (W * W)] + i % (W * W), v);
    }
}

int check(int width, int hidden, const float* const* weights, const float* const* biases) {
    if (width <= 0 || width > kMaxWidth || hidden < 0 || hidden > kMaxHidden || !weights || !biases) return BNRF_ERR_ARG;
    // the backward pass keeps [hidden][width][width] partial weight gradients in shared memory
    if (((size_t)kWarps * (2 * hidden + 2) * width + (size_t)hidden * width * width) * sizeof(float) > 232448) return BNRF_ERR_ARG;
    for (int l = 0; l < hidden + 2; ++l)
        if (!weights[l] || !biases[l]) return BNRF_ERR_ARG;
    return BNRF_OK;
}

CrfParams make_params(int width, int hidden, const float* const* w, const float* const* b) {
    CrfParams p{};
    p.width = width; p.hidden = hidden;
    p.w0 = w[0]; p.b0 = b[0];
    for (int l = 0; l < hidden; ++l) { p.wh[l] = w[1 + l]; p.bh[l] = b[1 + l]; }
    p.w1 = w[hidden + 1]; p.b1 = b[hidden + 1];
    return p;
}

}  // namespace
}  // namespace bnrf

using namespace bnrf;

extern "C" {

int bnrf_crf_forward(int width, int hidden, const float* const* weights, const float* const* biases, const float* x, int64_t n,
                     float* y, void* stream) {
    int rc = check(width, hidden, weights, biases);
    if (rc) return rc;
    if (n == 0) return BNRF_OK;
    if (!x || !y || n < 0) return BNRF_ERR_ARG;
    const CrfParams p = make_params(width, hidden, weights, biases);
    const size_t smem = (size_t)kWarps * (hidden + 1) * width * sizeof(float);
    const int64_t want = (n + kWarps - 1) / kWarps;
    crf_forward_kernel<<<(unsigned)(want < 1184 ? want : 1184), 32 * kWarps, smem, (cudaStream_t)stream>>>(p, x, n, y);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

int bnrf_crf_backward(int width, int hidden, const float* const* weights, const float* const* biases, const float* x, const float* g,
                      int64_t n, float* dx, float* const* d_weights, float* const* d_biases, void* stream) {
    int rc = check(width, hidden, weights, biases);
    if (rc) return rc;
    if (n == 0) return BNRF_OK;
    if (!x || !g || !dx || n < 0 || !d_weights || !d_biases) return BNRF_ERR_ARG;
    for (int l = 0; l < hidden + 2; ++l)
        if (!d_weights[l] || !d_biases[l]) return BNRF_ERR_ARG;
    const CrfParams p = make_params(width, hidden, weights, biases);
    CrfGrads gr{};
    gr.w0 = d_weights[0]; gr.b0 = d_biases[0];
    for (int l = 0; l < hidden; ++l) { gr.wh[l] = d_weights[1 + l]; gr.bh[l] = d_biases[1 + l]; }
    gr.w1 = d_weights[hidden + 1]; gr.b1 = d_biases[hidden + 1];
    const size_t smem = ((size_t)kWarps * (2 * hidden + 2) * width + (size_t)hidden * width * width) * sizeof(float);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(crf_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return BNRF_ERR_CUDA;
    const int64_t want = (n + kWarps - 1) / kWarps;
    crf_backward_kernel<<<(unsigned)(want < 296 ? want : 296), 32 * kWarps, smem, (cudaStream_t)stream>>>(p, gr, x, g, n, dx);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

}  // extern "C"
