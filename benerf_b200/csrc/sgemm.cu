// fp32 GEMM of the backward pass (dgrad / wgrad of the NeRF linears, train.py:340 through autograd).
//
// C[M,N] (op)= sum_k A_op(m,k) * B_op(k,n), row-major operands with leading dimensions, either
// operand optionally transposed, three epilogues:
//   store / accumulate            dgrad:  dA = dZ * W      (optionally C += acc)
//   masked store + rank-1 term    dgrad fused with the ReLU backward of the previous layer:
//                                 C = (acc + r_row[m] * r_col[n]) * (mask[m,n] > 0)
//   atomic add                    wgrad:  dW += dZ^T * A, the contraction over the rows split across blockIdx.z
// 128 x 128 x 16 tiles, 256 threads, 8 x 8 register tile, next tile prefetched into registers.
// Plain FFMA in the reference's own precision (fp32): gradients match torch autograd to rounding.
// The tensor-core (tcgen05) version of these per-layer GEMMs is the next step.
#include "common.cuh"
#include "sgemm.cuh"

namespace bnrf {

namespace {
constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(NT, 2) sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int64_t k_begin = (int64_t)blockIdx.z * g.k_chunk;
    const int64_t k_end = (k_begin + g.k_chunk < g.K) ? k_begin + g.k_chunk : g.K;
    if (k_begin >= k_end && g.epi == GEMM_ATOMIC) return;

    float4 ra[2], rb[2];
    // ---- global -> registers (zero-filled outside the matrix) ----
    auto load_tile = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * NT;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (!TA) {                                   // A[m*lda + k]: 4 consecutive k of one row
                const int64_t m = m0 + f / 4, k = k0 + (f % 4) * 4;
                if (m < g.M) {
                    const float* p = g.A + m * g.lda + k;
                    if (g.vec_a && k + 3 < k_end) { const float4 q = __ldg(reinterpret_cast<const float4*>(p)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (k + j < k_end) v[j] = __ldg(p + j);
                    }
                }
            } else {                                     // A[k*lda + m]: 4 consecutive m of one k
                const int64_t k = k0 + f / 32, m = m0 + (f % 32) * 4;
                if (k < k_end) {
                    const float* p = g.A + k * g.lda + m;
                    if (g.vec_a && m + 3 < g.M) { const float4 q = __ldg(reinterpret_cast<const float4*>(p)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (m + j < g.M) v[j] = __ldg(p + j);
                    }
                }
            }
            ra[i] = make_float4(v[0], v[1], v[2], v[3]);
            float w[4] = {0.f, 0.f, 0.f, 0.f};
            if (!TB) {                                   // B[k*ldb + n]: 4 consecutive n of one k
                const int64_t k = k0 + f / 32;
                const int n = n0 + (f % 32) * 4;
                if (k < k_end) {
                    const float* p = g.B + k * g.ldb + n;
                    if (g.vec_b && n + 3 < g.N) { const float4 q = __ldg(reinterpret_cast<const float4*>(p)); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (n + j < g.N) w[j] = __ldg(p + j);
                    }
                }
            } else {                                     // B[n*ldb + k]: 4 consecutive k of one n
                const int n = n0 + f / 4;
                const int64_t k = k0 + (f % 4) * 4;
                if (n < g.N) {
                    const float* p = g.B + (int64_t)n * g.ldb + k;
                    if (g.vec_b && k + 3 < k_end) { const float4 q = __ldg(reinterpret_cast<const float4*>(p)); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (k + j < k_end) w[j] = __ldg(p + j);
                    }
                }
            }
            rb[i] = make_float4(w[0], w[1], w[2], w[3]);
        }
    };
    auto store_tile = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * NT;
            if (!TA) {
                const int m = f / 4, k = (f % 4) * 4;
                As[k][m] = ra[i].x; As[k + 1][m] = ra[i].y; As[k + 2][m] = ra[i].z; As[k + 3][m] = ra[i].w;
            } else {
                *reinterpret_cast<float4*>(&As[f / 32][(f % 32) * 4]) = ra[i];
            }
            if (!TB) {
                *reinterpret_cast<float4*>(&Bs[f / 32][(f % 32) * 4]) = rb[i];
            } else {
                const int n = f / 4, k = (f % 4) * 4;
                Bs[k][n] = rb[i].x; Bs[k + 1][n] = rb[i].y; Bs[k + 2][n] = rb[i].z; Bs[k + 3][n] = rb[i].w;
            }
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    if (k_begin < k_end) load_tile(k_begin);
    for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
        __syncthreads();                                 // previous tile fully consumed
        store_tile();
        __syncthreads();
        if (k0 + BK < k_end) load_tile(k0 + BK);         // in flight while this tile is multiplied
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]), b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
        const float rr = (g.epi == GEMM_MASKED && g.r_row) ? g.r_row[m * g.r_stride] : 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n >= g.N) continue;
            float* c = g.C + m * g.ldc + n;
            float v = acc[i][j];
            if (g.epi == GEMM_STORE) *c = v;
            else if (g.epi == GEMM_ACCUM) *c += v;
            else if (g.epi == GEMM_ATOMIC) atomicAdd(c, v);
            else {
                if (g.r_row) v = fmaf(rr, g.r_col[n], v);
                *c = (g.mask[m * g.ldm + n] > 0.0f) ? v : 0.0f;
            }
        }
    }
}
}  // namespace

int launch_sgemm(bnrf_ctx* ctx, bool ta, bool tb, GemmArgs g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.K < 0) return fail(ctx, BNRF_ERR_ARG, "sgemm: bad shape");
    if (ctx->cfg.gemm_mode != BNRF_GEMM_SIMT_FP32 && gemm_tc_eligible(g)) return launch_gemm_tc(ctx, ta, tb, g, st);
    g.vec_a = ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && g.lda % 4 == 0) ? 1 : 0;
    g.vec_b = ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0 && g.ldb % 4 == 0) ? 1 : 0;
    int splits = 1;
    if (g.epi == GEMM_ATOMIC) {
        // wgrad: few output tiles, long contraction -> split the rows so that ~4 waves of CTAs exist
        const int64_t tiles = ceil_div(g.M, BM) * ceil_div(g.N, BN);
        const int64_t want = ceil_div((int64_t)4 * 2 * ctx->sm_count, tiles);
        const int64_t max_splits = ceil_div(g.K, 4 * BK);
        splits = (int)(want < max_splits ? want : max_splits);
        if (splits < 1) splits = 1;
        g.k_chunk = ceil_div(ceil_div(g.K, splits), BK) * BK;
        splits = (int)ceil_div(g.K, g.k_chunk);
    } else {
        g.k_chunk = g.K > 0 ? g.K : 1;
    }
    const dim3 grid((unsigned)ceil_div(g.M, BM), (unsigned)ceil_div(g.N, BN), (unsigned)splits);
    if (!ta && !tb) sgemm_kernel<false, false><<<grid, NT, 0, st>>>(g);
    else if (!ta && tb) sgemm_kernel<false, true><<<grid, NT, 0, st>>>(g);
    else if (ta && !tb) sgemm_kernel<true, false><<<grid, NT, 0, st>>>(g);
    else sgemm_kernel<true, true><<<grid, NT, 0, st>>>(g);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bnrf

// Test hook: C = A_op * B_op through the kernel above (tests/test_gpu_backward.py checks it against torch.matmul).
extern "C" int bnrf_debug_sgemm(bnrf_ctx* ctx, int ta, int tb, int64_t M, int N, int64_t K, const float* A, int64_t lda,
                                const float* B, int64_t ldb, float* C, int64_t ldc, int epi, const float* mask, int64_t ldm,
                                const float* r_row, int64_t r_stride, const float* r_col, void* stream) {
    using namespace bnrf;
    if (!ctx || !A || !B || !C) return BNRF_ERR_ARG;
    if (epi == GEMM_MASKED && !mask) return fail(ctx, BNRF_ERR_ARG, "sgemm: masked epilogue needs a mask");
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.epi = epi;
    g.mask = mask; g.ldm = ldm; g.r_row = r_row; g.r_stride = r_stride; g.r_col = r_col;
    return launch_sgemm(ctx, ta != 0, tb != 0, g, (cudaStream_t)stream);
}
