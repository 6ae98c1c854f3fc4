// Interface of the backward pass's fp32 GEMM (sgemm.cu).
#pragma once
#include "common.cuh"

namespace bnrf {

enum { GEMM_STORE = 0, GEMM_ACCUM = 1, GEMM_ATOMIC = 2, GEMM_MASKED = 3 };

struct GemmArgs {
    int64_t M; int N; int64_t K;      // C[M,N] (op)= sum_k A_op(m,k) * B_op(k,n)
    const float* A; int64_t lda;      // not transposed: A[m*lda + k]; transposed: A[k*lda + m]
    const float* B; int64_t ldb;      // not transposed: B[k*ldb + n]; transposed: B[n*ldb + k]
    float* C; int64_t ldc;
    int epi;                          // GEMM_*
    const float* mask; int64_t ldm;   // GEMM_MASKED: C = (acc + r_row[m*r_stride] * r_col[n]) * (mask[m*ldm + n] > 0)
    const float* r_row; int64_t r_stride; const float* r_col;   // optional rank-1 term (r_row == NULL: none)
    int64_t k_chunk;                  // set by the launcher (contraction range per blockIdx.z)
    int vec_a, vec_b;                 // set by the launcher (16-byte aligned, ld % 4 == 0)
};

// Dispatches to the tensor-core kernel (gemm_tc.cu) when cfg.gemm_mode == BNRF_GEMM_TC and the shape qualifies,
// else to the fp32 FFMA kernel.
int launch_sgemm(bnrf_ctx* ctx, bool ta, bool tb, GemmArgs g, cudaStream_t st);
bool gemm_tc_eligible(const GemmArgs& g);
int launch_gemm_tc(bnrf_ctx* ctx, bool ta, bool tb, GemmArgs g, cudaStream_t st);

}  // namespace bnrf
