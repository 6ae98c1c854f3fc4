// Bring-up probe for tcgen05.mma with the A operand in tensor memory (".ts" form) on a CTA pair: the layout conventions
// the forward / dgrad-chain kernels rely on, validated in isolation (tests/test_gpu_probe.py).
//
//   D[256 x N] = A[256 x 64] * B[N x 64]^T,  fp16 inputs in plain row-major global memory, fp32 out.
//   CTA r of the pair owns rows [128 r, 128 r + 128) of A -- written by its own threads into its own TMEM with
//   tcgen05.st (lane = row, one 32-bit column = two consecutive K elements) at column a_col -- and rows
//   [r N/2, (r+1) N/2) of B in its shared memory (SWIZZLE_128B K-major).  One thread of CTA 0 issues the four K = 16
//   MMAs with cta_group::2, M = 256.
#include "tc_ptx.cuh"

namespace bnrf {
using namespace tcp;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_ts_probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, int N, int a_col, float* __restrict__ D) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t done_bar;
    const int warp = threadIdx.x / 32;
    const uint32_t rank = cluster_ctarank();
    const int nh = N / 2;
    for (int e = threadIdx.x; e < nh * 64; e += 128)
        *reinterpret_cast<__half*>(sm + sw128_offset(e / 64, e % 64)) = B[(size_t)(rank * nh + e / 64) * 64 + e % 64];
    if (threadIdx.x == 0) { mbar_init(smem_u32(&done_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    {   // this thread's row of A: 64 fp16 = 32 packed columns
        const int r = threadIdx.x;
        uint32_t v[32];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)(rank * 128 + r) * 64);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = src[j];
        tc_st32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)a_col, v);
        tc_st_wait();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = make_idesc(256, N);
        for (int kk = 0; kk < 4; ++kk)
            tc_mma_pair_ts_f16(tmem, tmem + (uint32_t)a_col + (uint32_t)kk * 8u, make_desc(base + kk * 32, 0), idesc, kk > 0);
        tc_commit_pair(smem_u32(&done_bar));
    }
    mbar_wait(smem_u32(&done_bar), 0, nullptr, 0);
    tc_fence_after();
    const int r = threadIdx.x;
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + r) * N + c0 + j] = v[j];
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

}  // namespace bnrf

extern "C" int bnrf_debug_umma_ts_probe(const void* A_half, const void* B_half, int N, int a_col, float* D, void* stream) {
    using namespace bnrf;
    if (!A_half || !B_half || !D || (N != 128 && N != 256) || a_col < N || a_col + 32 > 512) return BNRF_ERR_ARG;
    const size_t smem = 16384 + 1024;
    if (cudaFuncSetAttribute(umma_ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return BNRF_ERR_CUDA;
    umma_ts_probe_kernel<<<2, 128, smem, (cudaStream_t)stream>>>((const __half*)A_half, (const __half*)B_half, N, a_col, D);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}
