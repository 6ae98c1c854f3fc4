// K4/K5 (fp32 SIMT variant): positional encoding + the whole 8x256 MLP for a tile of 64
// samples per CTA, activations resident in shared memory, plain FFMA.
//
// This is the on-device fp32 cross-check of the tensor-core kernel (mlp_tc.cu), selected with
// cfg.mlp_mode = BNRF_MLP_SIMT_FP32.  Same inputs, same outputs, same weight cache; it never
// touches HBM between layers either, but it is bound by the fp32 pipe (~70 TFLOP/s), not by
// tcgen05.  Replaces model/embedder.py:9-34 + model/nerf.py:67-116.
#include "common.cuh"

namespace bnrf {

namespace simt {
constexpr int TM = 64;       // samples (rows) per CTA
constexpr int KC = 16;       // weight k-chunk staged in shared memory
constexpr int NT = 256;      // threads
constexpr int kSmemFloats = (kPtsChPad + 2 * kWidth) * TM + KC * kWidth + 4 * TM;
}  // namespace simt

struct SimtParams {
    const float* wt[10];
    const float* bias[10];
    const float* w_alpha; const float* b_alpha;
    const float* w_rgb; const float* b_rgb;
};

// acc[r][c] += sum_k in[k][row r] * W^T[k][col c] for one input segment of K rows.
template <int N>
__device__ inline void gemm_segment(const float* __restrict__ in, int K, const float* __restrict__ wt, float* wbuf,
                                    float (&acc)[8][N / 32], int tx, int ty, int tid) {
    using namespace simt;
    constexpr int CPT = N / 32;                 // columns per thread
    constexpr int F4 = KC * N / 4 / NT;         // float4 loads per thread per chunk
    float4 stage[F4];
    const float4* src = reinterpret_cast<const float4*>(wt);
#pragma unroll
    for (int i = 0; i < F4; ++i) stage[i] = __ldg(src + tid + i * NT);
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();                        // previous chunk fully consumed
#pragma unroll
        for (int i = 0; i < F4; ++i) reinterpret_cast<float4*>(wbuf)[tid + i * NT] = stage[i];
        __syncthreads();
        if (k0 + KC < K) {
#pragma unroll
            for (int i = 0; i < F4; ++i) stage[i] = __ldg(src + (size_t)(k0 + KC) * N / 4 + tid + i * NT);
        }
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float a[8], b[CPT];
            const float4 a0 = *reinterpret_cast<const float4*>(in + (k0 + kk) * TM + tx * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(in + (k0 + kk) * TM + tx * 8 + 4);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int c = 0; c < CPT; c += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(wbuf + kk * N + ty * CPT + c);
                b[c] = bv.x; b[c + 1] = bv.y; b[c + 2] = bv.z; b[c + 3] = bv.w;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
    }
}

template <int N, bool RELU>
__device__ inline void store_layer(float (&acc)[8][N / 32], const float* __restrict__ bias, const float* __restrict__ row_bias,
                                   float* out, int tx, int ty) {
    using namespace simt;
    constexpr int CPT = N / 32;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int col = ty * CPT + c;
        const float bcol = bias ? bias[col] : 0.0f;
        float v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float x = acc[r][c] + bcol;
            if (row_bias) x += row_bias[(tx * 8 + r) * kHalf + col];   // per-ray view bias staged per row
            v[r] = RELU ? fmaxf(x, 0.0f) : x;
        }
        *reinterpret_cast<float4*>(out + col * TM + tx * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(out + col * TM + tx * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

template <int C>
__global__ void __launch_bounds__(simt::NT, 1)
mlp_simt_kernel(SimtParams p, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                const float* __restrict__ viewbias, const float* __restrict__ z, int64_t rows, int S,
                float* __restrict__ raw) {
    using namespace simt;
    extern __shared__ __align__(16) float sm[];
    float* pe = sm;                         // [64][TM]   encoded point (k-major)
    float* hA = pe + kPtsChPad * TM;        // [256][TM]
    float* hB = hA + kWidth * TM;           // [256][TM]
    float* wbuf = hB + kWidth * TM;         // [KC][256]
    float* part = wbuf + KC * kWidth;       // [4][TM] partial sums of the sigma head
    const int tid = threadIdx.x, tx = tid % 8, ty = tid / 8;
    const int64_t row0 = (int64_t)blockIdx.x * TM;

    // ---- front end: pts = o + d*z, positional encoding (model/embedder.py:13-28 layout) ----
    {
        const int r = tid % TM, q = tid / TM;               // 4 threads per row, frequencies k = q, q+4, q+8
        const int64_t row = row0 + r;
        float x[3] = {0.f, 0.f, 0.f};
        if (row < rows) {
            const int64_t ray = row / S;
#pragma unroll
            for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(rays_d[ray * 3 + c], z[row]));
        }
        if (q == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) pe[c * TM + r] = x[c];
            pe[63 * TM + r] = 0.0f;                         // K padding
        }
        for (int k = q; k < kPtsFreqs; k += 4) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float s, co;
                sincosf(x[c] * (float)(1 << k), &s, &co);
                pe[(3 + 6 * k + c) * TM + r] = s;
                pe[(3 + 6 * k + 3 + c) * TM + r] = co;
            }
        }
    }
    __syncthreads();

    float acc[8][8];
    auto zero = [&]() {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
    };
    float* cur = hA; float* nxt = hB;
    // L0: 63(+1) -> 256
    zero();
    gemm_segment<256>(pe, kPtsChPad, p.wt[0], wbuf, acc, tx, ty, tid);
    store_layer<256, true>(acc, p.bias[0], nullptr, cur, tx, ty);
    __syncthreads();
    // L1..L7 (L5 takes [pe | h], model/nerf.py:98)
    for (int l = 1; l < 8; ++l) {
        zero();
        const float* w = p.wt[l];
        if (l == 5) {
            gemm_segment<256>(pe, kPtsChPad, w, wbuf, acc, tx, ty, tid);
            w += (size_t)kPtsChPad * kWidth;
        }
        gemm_segment<256>(cur, kWidth, w, wbuf, acc, tx, ty, tid);
        store_layer<256, true>(acc, p.bias[l], nullptr, nxt, tx, ty);
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }
    // sigma head on h7 (alpha_linear, model/nerf.py:101): 4 partial sums per row
    {
        const int r = tid % TM, q = tid / TM;
        float s = 0.0f;
        for (int k = q * 64; k < q * 64 + 64; ++k) s = fmaf(cur[k * TM + r], p.w_alpha[k], s);
        part[q * TM + r] = s;
    }
    // feature head, no activation (model/nerf.py:102)
    zero();
    gemm_segment<256>(cur, kWidth, p.wt[8], wbuf, acc, tx, ty, tid);
    store_layer<256, false>(acc, p.bias[8], nullptr, nxt, tx, ty);
    __syncthreads();
    // view layer: feature part by GEMM, direction part + bias pre-reduced per ray (rays.cu viewbias_kernel)
    {
        float* vb_rows = cur;                                // reuse: [TM][128] per-row copy of the ray's view bias
        for (int i = tid; i < TM * kHalf; i += NT) {
            const int r = i / kHalf, j = i % kHalf;
            const int64_t row = row0 + r;
            vb_rows[i] = (row < rows) ? viewbias[(row / S) * kHalf + j] : 0.0f;
        }
        __syncthreads();
        float acc4[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc4[r][c] = 0.0f;
        gemm_segment<128>(nxt, kWidth, p.wt[9], wbuf, acc4, tx, ty, tid);
        __syncthreads();                                     // all reads of nxt (feature) done before overwrite
        store_layer<128, true>(acc4, nullptr, vb_rows, nxt, tx, ty);   // nxt[0..127][TM] = relu(view layer)
        __syncthreads();
    }
    // rgb head (model/nerf.py:109) + output cat([rgb, sigma]) (model/nerf.py:110)
    {
        const int r = tid % TM, q = tid / TM;
        const int64_t row = row0 + r;
        if (row < rows) {
            if (q < C) {
                float s = p.b_rgb[q];
                for (int k = 0; k < kHalf; ++k) s = fmaf(nxt[k * TM + r], p.w_rgb[q * kHalf + k], s);
                raw[row * (C + 1) + q] = s;
            } else if (q == 3) {
                raw[row * (C + 1) + C] = ((part[r] + part[TM + r]) + (part[2 * TM + r] + part[3 * TM + r])) + p.b_alpha[0];
            }
        }
    }
}

int launch_mlp_simt(bnrf_ctx* ctx, int net, const float* o, const float* d, const float* vb, const float* z,
                    int64_t n, int S, float* raw, cudaStream_t st) {
    using namespace simt;
    const NetParams& np = ctx->net[net];
    SimtParams p;
    for (int i = 0; i < 10; ++i) { p.wt[i] = np.wt[i]; p.bias[i] = np.bias[i]; }
    p.w_alpha = np.w_alpha; p.b_alpha = np.b_alpha; p.w_rgb = np.w_rgb; p.b_rgb = np.b_rgb;
    const int64_t rows = n * S;
    const size_t smem = kSmemFloats * sizeof(float);
    const unsigned grid = (unsigned)ceil_div(rows, TM);
    if (ctx->cfg.channels == 3) {
        BNRF_CUDA(ctx, cudaFuncSetAttribute(mlp_simt_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_simt_kernel<3><<<grid, NT, smem, st>>>(p, o, d, vb, z, rows, S, raw);
    } else {
        BNRF_CUDA(ctx, cudaFuncSetAttribute(mlp_simt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_simt_kernel<1><<<grid, NT, smem, st>>>(p, o, d, vb, z, rows, S, raw);
    }
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bnrf
