// Tensor-core GEMM of the backward pass: C[M,N] (op)= A_op[M,K] * B_op[K,N], fp32 in HBM, tcgen05 inside.
//
// Same contract as sgemm.cu (GemmArgs, the four epilogues) for the shapes the dgrad / wgrad of the NeRF linears
// produce: N <= 256, any M and K.  Operands are read as fp32, split on the fly into bf16 hi + lo (bf16 keeps the fp32
// exponent range, so gradients of any magnitude need no scaling) and multiplied with three MMAs per K=16 slice
// (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM): ~2^-17 relative per product, fp32-equivalent for gradients.
//
// Persistent CTAs, 21 warps:
//   warps 0-15  loaders: fp32 global -> bf16 hi/lo -> shared memory, canonical SWIZZLE_128B K-major tiles.  Either
//               operand may be stored K-contiguous (dgrad: dZ [rows,N_l], W^T [K_l,N_l]) or MN-contiguous (wgrad: the
//               contraction runs over the ROWS of dZ [rows,N_l] and A [rows,K_l]); in the second case a thread owns one
//               output index and gathers 8 consecutive rows, which is a register-level transpose with coalesced loads.
//   warps 16-19 epilogue: tcgen05.ld -> store / accumulate / ReLU-masked store (+ rank-1 term) / red.global.add
//   warp 20     one thread issues tcgen05.mma (cta_group::1, M = 128, N = round_up(N, 32))
// Two shared-memory stages of one 64-deep K-block each, two TMEM accumulators (epilogue of tile i overlaps tile i+1).
#include "tc_ptx.cuh"
#include "sgemm.cuh"

namespace bnrf {
namespace gtc {
using namespace tcp;

constexpr int TM = 128, BK = 64, NT_MAX = 256;
constexpr int LOADER_WARPS = 16, LOADER_THREADS = LOADER_WARPS * 32;
constexpr int EPI_WARP0 = LOADER_WARPS, MMA_WARP = LOADER_WARPS + 4;
constexpr int MAXC = (TM + NT_MAX) * 8 / LOADER_THREADS;   // 16-byte chunks per loader thread per stage (6)
constexpr int THREADS = (LOADER_WARPS + 4 + 1) * 32;
constexpr uint32_t A_BYTES = TM * BK * 2;           // 16 KB
constexpr uint32_t B_BYTES = NT_MAX * BK * 2;       // 32 KB
constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = A_BYTES, OFF_B_HI = 2 * A_BYTES, OFF_B_LO = 2 * A_BYTES + B_BYTES;
constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 96 KB
constexpr int NSTAGE = 2;
constexpr uint32_t OFF_BAR = NSTAGE * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128 + 1024;
enum { BAR_FULL = 0, BAR_EMPTY = NSTAGE, BAR_ACC_FULL = 2 * NSTAGE, BAR_ACC_EMPTY = 2 * NSTAGE + 2, BAR_COUNT = 2 * NSTAGE + 4 };

// kind::f16 instruction descriptor with bf16 operands, fp32 accumulate, both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split_store8_bf16(const float* v, unsigned char* hi_dst, unsigned char* lo_dst) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        const float b0 = __uint_as_float(h[i] << 16), b1 = __uint_as_float(h[i] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(v[2 * i + 1] - b1), "f"(v[2 * i] - b0));
    }
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

// One stage = one K-block of both operands: A [128 x 64] and B [NT x 64] values -> bf16 hi/lo tiles.  A thread owns up to
// MAXC 16-byte chunks (8 consecutive k of one row of a tile); ALL its global loads are issued before the first conversion
// so that ~100 KB per SM are in flight (the loaders are latency-bound otherwise).  Out-of-range elements are read from a
// clamped address and zeroed by predicate, which keeps the loads unconditional.
// k_contig operand: P[mn*ld + k]; otherwise P[k*ld + mn] (a thread then gathers 8 consecutive k = rows of the source).
__device__ __forceinline__ void load_stage(const GemmArgs& g, bool a_kc, bool b_kc, int64_t m0, int NT, int64_t k0, int64_t ke,
                                           unsigned char* st, int tid) {
    const int chunks_a = TM * 8, total = chunks_a + NT * 8;
    const bool fullk = k0 + BK <= ke;
    float v[MAXC][8];
    uint32_t off_[MAXC];
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
        const int f = tid + u * LOADER_THREADS;
        const bool is_b = f >= chunks_a;
        const int fl = is_b ? f - chunks_a : f;
        const int rows = is_b ? NT : TM;
        const bool kc = is_b ? b_kc : a_kc;
        const float* P = is_b ? g.B : g.A;
        const int64_t ld = is_b ? g.ldb : g.lda;
        const int64_t MN = is_b ? (int64_t)g.N : g.M;
        const int64_t mn0 = is_b ? 0 : m0;
        const bool vec = (is_b ? g.vec_b : g.vec_a) != 0;
        int mn, c8;
        if (kc) { mn = fl >> 3; c8 = fl & 7; } else { mn = fl % rows; c8 = fl / rows; }
        off_[u] = (is_b ? OFF_B_HI : OFF_A_HI) + sw128_offset(mn, c8 * 8);
        const bool live = f < total && mn0 + mn < MN;
        const int64_t k = k0 + c8 * 8;
        if (kc) {
            const float* p = live ? P + (mn0 + mn) * ld + k : P;
            if (vec && fullk) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
                v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u][j] = live ? v[u][j] : 0.0f;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool ok = live && k + j < ke;
                    const float x = __ldg(ok ? p + j : P);
                    v[u][j] = ok ? x : 0.0f;
                }
            }
        } else {
            const float* p = live ? P + k * ld + mn0 + mn : P;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = live && (fullk || k + j < ke);
                const float x = __ldg(ok ? p + (int64_t)j * ld : P);
                v[u][j] = ok ? x : 0.0f;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
        if (tid + u * LOADER_THREADS < total) {
            // the lo tile of either operand sits one tile size after its hi tile
            const uint32_t lo_delta = (off_[u] >= OFF_B_HI) ? B_BYTES : A_BYTES;
            split_store8_bf16(v[u], st + off_[u], st + off_[u] + lo_delta);
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(GemmArgs g, int ta, int tb, int NT, int m_tiles, int splits,
                                                             unsigned int* err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + OFF_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * BAR_COUNT);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(BAR_FULL + i), LOADER_WARPS); mbar_init(bar(BAR_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(BAR_ACC_FULL + i), 1); mbar_init(bar(BAR_ACC_EMPTY + i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int total_tiles = m_tiles * splits;
    const int my_tiles = (total_tiles > (int)blockIdx.x) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_k = [&](int tile, int64_t& kb, int64_t& ke) {
        const int sp = tile / m_tiles;
        kb = (int64_t)sp * g.k_chunk;
        ke = (kb + g.k_chunk < g.K) ? kb + g.k_chunk : g.K;
    };

    if (warp < LOADER_WARPS) {
        // ================= loaders =================
        const int tid = threadIdx.x;
        uint32_t cnt = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int64_t m0 = (int64_t)(tile % m_tiles) * TM;
            int64_t kb, ke;
            tile_k(tile, kb, ke);
            for (int64_t k0 = kb; k0 < ke; k0 += BK, ++cnt) {
                const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                mbar_wait(bar(BAR_EMPTY + s), ph ^ 1u, err_flag, 21);
                unsigned char* st = sm + s * STAGE_BYTES;
                load_stage(g, !ta, tb != 0, m0, NT, k0, ke, st, tid);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(BAR_FULL + s));
            }
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(TM, NT);
            uint32_t cnt = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                int64_t kb, ke;
                tile_k(tile, kb, ke);
                const uint32_t acc = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
                mbar_wait(bar(BAR_ACC_EMPTY + acc), (use & 1u) ^ 1u, err_flag, 22);
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * 256u;
                uint32_t accumulate = 0;
                for (int64_t k0 = kb; k0 < ke; k0 += BK, ++cnt) {
                    const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                    mbar_wait(bar(BAR_FULL + s), ph, err_flag, 23);
                    tc_fence_after();
                    const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16) {
                        const uint64_t ah = make_desc(st + OFF_A_HI + k16 * 32, 0), al = make_desc(st + OFF_A_LO + k16 * 32, 0);
                        const uint64_t bh = make_desc(st + OFF_B_HI + k16 * 32, 0), bl = make_desc(st + OFF_B_LO + k16 * 32, 0);
                        tc_mma_f16(d_tmem, ah, bh, idesc, accumulate);
                        accumulate = 1;
                        tc_mma_f16(d_tmem, al, bh, idesc, 1);
                        tc_mma_f16(d_tmem, ah, bl, idesc, 1);
                    }
                    tc_commit(bar(BAR_EMPTY + s));
                }
                tc_commit(bar(BAR_ACC_FULL + acc));
            }
        }
    } else {
        // ================= epilogue (4 warps: TMEM lane quarter q = warp % 4) =================
        const int q = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const bool vec_c = (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && g.ldc % 4 == 0;
        const bool vec_m = g.epi == GEMM_MASKED && (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0 && g.ldm % 4 == 0 &&
                           (!g.r_row || (reinterpret_cast<uintptr_t>(g.r_col) & 15) == 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int64_t m = (int64_t)(tile % m_tiles) * TM + q * 32 + lane;
            int64_t kb, ke;
            tile_k(tile, kb, ke);
            const uint32_t acc = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
            mbar_wait(bar(BAR_ACC_FULL + acc), use & 1u, err_flag, 24);
            tc_fence_after();
            const bool empty = !(kb < ke);                  // no K-block was multiplied: the accumulator is stale
            const float rr = (g.epi == GEMM_MASKED && g.r_row && m < g.M) ? g.r_row[m * g.r_stride] : 0.0f;
            for (int c0 = 0; c0 < NT; c0 += 32) {
                float v[32];
                tc_ld32(lane_addr + acc * 256u + (uint32_t)c0, v);
                if (m < g.M && !(empty && g.epi != GEMM_STORE && g.epi != GEMM_MASKED)) {
                    float* crow = g.C + m * g.ldc;
                    if (empty) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.0f;
                    }
                    const bool full = c0 + 32 <= g.N;
                    if (g.epi == GEMM_MASKED) {
                        const float* mrow = g.mask + m * g.ldm;
                        if (full && vec_c && vec_m) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 mk = __ldg(reinterpret_cast<const float4*>(mrow + c0 + j));
                                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                                if (g.r_row) {
                                    const float4 rc = __ldg(reinterpret_cast<const float4*>(g.r_col + c0 + j));
                                    o.x = fmaf(rr, rc.x, o.x); o.y = fmaf(rr, rc.y, o.y); o.z = fmaf(rr, rc.z, o.z); o.w = fmaf(rr, rc.w, o.w);
                                }
                                o.x = mk.x > 0.0f ? o.x : 0.0f; o.y = mk.y > 0.0f ? o.y : 0.0f;
                                o.z = mk.z > 0.0f ? o.z : 0.0f; o.w = mk.w > 0.0f ? o.w : 0.0f;
                                *reinterpret_cast<float4*>(crow + c0 + j) = o;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int n = c0 + j;
                                if (n >= g.N) break;
                                float x = v[j];
                                if (g.r_row) x = fmaf(rr, g.r_col[n], x);
                                crow[n] = (mrow[n] > 0.0f) ? x : 0.0f;
                            }
                        }
                    } else if (g.epi == GEMM_STORE && full && vec_c) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = c0 + j;
                            if (n >= g.N) break;
                            if (g.epi == GEMM_STORE) crow[n] = v[j];
                            else if (g.epi == GEMM_ACCUM) crow[n] += v[j];
                            else atomicAdd(crow + n, v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_ACC_EMPTY + acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

}  // namespace gtc

bool gemm_tc_eligible(const GemmArgs& g) { return g.N >= 32 && g.N <= gtc::NT_MAX && g.M >= 64 && g.K >= 64; }

int launch_gemm_tc(bnrf_ctx* ctx, bool ta, bool tb, GemmArgs g, cudaStream_t st) {
    using namespace gtc;
    g.vec_a = ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && g.lda % 4 == 0) ? 1 : 0;
    g.vec_b = ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0 && g.ldb % 4 == 0) ? 1 : 0;
    const int NT = (g.N + 31) / 32 * 32;
    const int64_t m_tiles64 = ceil_div(g.M, TM);
    if (m_tiles64 > 0x3fffffff) return fail(ctx, BNRF_ERR_ARG, "gemm: too many rows");
    const int m_tiles = (int)m_tiles64;
    int splits = 1;
    if (g.epi == GEMM_ATOMIC) {                       // wgrad: few output tiles, long contraction -> one wave of CTAs
        const int64_t want = ceil_div((int64_t)ctx->sm_count, m_tiles);
        const int64_t max_splits = ceil_div(g.K, 4 * BK);
        splits = (int)(want < max_splits ? want : max_splits);
        if (splits < 1) splits = 1;
        g.k_chunk = ceil_div(ceil_div(g.K, splits), BK) * BK;
        splits = (int)ceil_div(g.K, g.k_chunk);
    } else {
        g.k_chunk = g.K;
    }
    const int64_t total = (int64_t)m_tiles * splits;
    const int grid = (int)(total < ctx->sm_count ? total : ctx->sm_count);
    BNRF_CUDA(ctx, cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    gemm_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(g, ta ? 1 : 0, tb ? 1 : 0, NT, m_tiles, splits, ctx->err_flag);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bnrf
