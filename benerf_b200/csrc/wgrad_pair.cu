// Weight gradients of the 256-wide linears on CTA pairs: dW_l += dZ_l^T . H_{l-1}, dB_l += colsum(dZ_l) for every wide linear of a
// network in ONE launch (autograd of model/nerf.py:93-112, the d-weights half of train.py:340).
//
// Same contraction as tile_wgrad_kernel (bwd_tiles.cu): both operands are bf16 hi/lo tile matrices read as MN-major UMMA operands
// (the contraction runs over their ROWS), 3 MMAs per K = 16 slice, the whole [M x N] fp32 partial of dW stays in TMEM while a CTA
// streams row tiles, one red.global per element at the end.  What changes is who reads what:
//   * tcgen05.mma cta_group::2, M = 256: CTA r of a pair loads only ITS 128 columns of dZ (its half of dW's rows) and ITS half of
//     H's columns; the hardware exchanges the B halves.  A single CTA needs 4 KB (A) + 8 KB (B) of shared-memory operand reads
//     per 128-cycle MMA = 96 B/cycle, plus 26 B/cycle of bulk-copy fills, plus 21 B/cycle of LSU reads for the bias sums: 143 of
//     the 128 B/cycle the port has (ncu, round 1: 0.72 of the HBM peak, 66 % of its LSU wavefronts "bank conflicted").  The pair
//     reads 4 + 4 KB per MMA = 64 B/cycle and stages 32 KB instead of 64 KB per 32-row slice.
//   * the bias gradient is summed out of the staged slices by two otherwise idle warps per CTA (16 KB of LSU reads per 32 KB stage;
//     N = 16 MMAs against a tile of ones were tried instead: each re-reads the 4 KB A slice, and the kernel ran 1.4x slower).
//   * work is balanced exactly: the (job, row-tile) line, weighted by bytes per tile, is cut into equal spans, one per cluster;
//     a span that crosses a job boundary flushes its accumulator and continues with the next job.
// The merged view step (G = sum dZ9 (x) h7, backward.cu) runs with the roles swapped -- A = h7 (M = 256), B = dZ9 (N = 128) --
// and is flushed transposed; alpha_linear's weight gradient rides on its staged h7 slices as before.
// The two narrow jobs (encoded points, N = 64) stay on tile_wgrad_kernel.
#include "tc_ptx.cuh"
#include "bwd_tiles.cuh"

namespace bnrf {
namespace bwt {
using namespace tcp;

namespace wgp {
constexpr int THREADS = 384;                 // warps 0, 3 producers, 1 MMA issuer (leader), 2 relay (peer), 4-7 flush, 8-11 column sums
constexpr uint32_t SLICE = 4096;             // 32 rows of one 64-column block
constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = 2 * SLICE, OFF_B_HI = 4 * SLICE, OFF_B_LO = 6 * SLICE, OFF_WROW = 8 * SLICE, STAGE = 8 * SLICE + 1024;
                                             // OFF_WROW: the 32 rows of the matrix wrow is a column of (view job only; <= 32 x 16 B)
constexpr int NSTAGE = 6;
constexpr uint32_t OFF_BAR = NSTAGE * STAGE;
enum { FULL = 0, EMPTY = NSTAGE, ACC_FULL = 2 * NSTAGE, ACC_EMPTY, NBAR };
constexpr uint32_t SMEM = OFF_BAR + 8 * NBAR + 16 + 1024;
static_assert(SMEM <= 232448, "shared memory budget");

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {   // kind::f16, bf16 x bf16 -> fp32, both operands MN-major
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
}  // namespace wgp

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(wgp::THREADS, 1)
tile_wgrad_pair_kernel(const __grid_constant__ WgradPairParams p, unsigned int* err_flag, unsigned long long* __restrict__ trace) {
    using namespace wgp;
    const long long k_t0 = clock64();
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + OFF_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * NBAR);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const int cluster = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
    const uint32_t lbar0 = mapa_u32(bar0, 0);
    auto lbar = [&](int i) { return lbar0 + 8u * (uint32_t)i; };

    // this cluster's span of the work line and the (job, tile range) segments it covers: identical in every thread
    const int64_t lo = p.units * cluster / n_clusters, hi = p.units * (cluster + 1) / n_clusters;
    auto seg_range = [&](int j, int& t0, int& t1) {
        const WgradPairJob& jb = p.job[j];
        auto cut = [&](int64_t u) {
            int64_t t = (u - jb.unit0 + jb.weight - 1) / jb.weight;       // first tile starting at or after u
            return (int)(t < 0 ? 0 : (t > p.tiles ? p.tiles : t));
        };
        t0 = cut(lo); t1 = cut(hi);
        return t1 > t0;
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(FULL + i), rank == 0 ? 3 : 2); mbar_init(bar(EMPTY + i), 5); }   // FULL: two producers (+ the relay); EMPTY: MMA commit + 4 column-sum warps
        mbar_init(bar(ACC_FULL), 1);
        mbar_init(bar(ACC_EMPTY), 8);                          // 4 flush warps x 2 CTAs, on the leader
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 || warp == 3) {
        // ================= producers: warp 0 streams this CTA's A slices, warp 3 its B slices; ONE elected thread each issues the
        //                   4 KB bulk copies of a stage back to back.  (Issued from different lanes of a warp, every cp.async.bulk
        //                   sits in a uniformisation loop of ~150 cycles: 16 copies per 64 KB stage paced the single-CTA kernel at
        //                   0.72 of the HBM peak.) =================
        if (elect_one()) {
            const bool is_a = warp == 0;
            uint32_t cnt = 0;
            for (int j = 0; j < p.n_jobs; ++j) {
                int t0, t1;
                if (!seg_range(j, t0, t1)) continue;
                const WgradPairJob& jb = p.job[j];
                const int W = is_a ? 256 : jb.NB;
                const int nblk = W / 128;                             // 64-column blocks of the operand this CTA loads (2, or 1 for NB = 128)
                const unsigned char* src0 = (is_a ? jb.a_tiles : jb.b_tiles) + (size_t)(nblk * rank) * kKbBytes;
                const size_t part = tile_part_bytes(W), tstride = tile_bytes(W);
                const uint32_t off_hi = is_a ? OFF_A_HI : OFF_B_HI, off_lo = is_a ? OFF_A_LO : OFF_B_LO;
                const uint32_t iters = (uint32_t)(t1 - t0) * 4u;
                for (uint32_t i = 0; i < iters; ++i, ++cnt) {
                    const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                    mbar_wait(bar(EMPTY + s), ph ^ 1u, err_flag, 91);
                    // the view job's row weights travel with the B slices: 32 rows x wrow_stride floats, contiguous (rows of the last
                    // tile beyond p.rows are read from the caller's workspace behind the matrix and masked by the consumer)
                    const uint32_t wbytes = (!is_a && jb.wrow) ? 32u * (uint32_t)jb.wrow_stride * 4u : 0u;
                    mbar_expect_tx(bar(FULL + s), (uint32_t)(2 * nblk) * SLICE + wbytes);
                    const unsigned char* src = src0 + ((size_t)t0 + (i >> 2)) * tstride + (size_t)(i & 3u) * SLICE;
                    const uint32_t st = base + s * STAGE;
                    if (wbytes) {
                        const int64_t row0 = ((int64_t)t0 + (i >> 2)) * kTileRows + (int64_t)(i & 3u) * 32;
                        tma_bulk_load(st + OFF_WROW, jb.wrow_base + row0 * jb.wrow_stride, wbytes, bar(FULL + s));
                    }
                    for (int blk = 0; blk < nblk; ++blk) {
                        tma_bulk_load(st + off_hi + (uint32_t)blk * SLICE, src + (size_t)blk * kKbBytes, SLICE, bar(FULL + s));
                        tma_bulk_load(st + off_lo + (uint32_t)blk * SLICE, src + part + (size_t)blk * kKbBytes, SLICE, bar(FULL + s));
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ================= peer only: forward "my half of the stage has landed" to the leader's FULL =================
        if (lane == 0 && rank == 1) {
            uint32_t cnt = 0;
            for (int j = 0; j < p.n_jobs; ++j) {
                int t0, t1;
                if (!seg_range(j, t0, t1)) continue;
                const uint32_t iters = (uint32_t)(t1 - t0) * 4u;
                for (uint32_t i = 0; i < iters; ++i, ++cnt) {
                    const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                    mbar_wait(bar(FULL + s), ph, err_flag, 92);
                    mbar_arrive_cluster(lbar(FULL + s));
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA, one thread) =================
        if (rank == 0 && elect_one()) {
            uint32_t cnt = 0, seg = 0;
            for (int j = 0; j < p.n_jobs; ++j) {
                int t0, t1;
                if (!seg_range(j, t0, t1)) continue;
                const WgradPairJob& jb = p.job[j];
                const uint32_t idesc = idesc_mn(256, jb.NB);
                if (seg > 0) { mbar_wait_cluster(bar(ACC_EMPTY), (seg - 1) & 1u, err_flag, 93); tc_fence_after(); }   // previous partial flushed
                const uint32_t iters = (uint32_t)(t1 - t0) * 4u;
                for (uint32_t i = 0; i < iters; ++i, ++cnt) {
                    const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                    mbar_wait_cluster(bar(FULL + s), ph, err_flag, 94);
                    tc_fence_after();
                    const uint32_t st = base + s * STAGE;
#pragma unroll
                    for (int k16 = 0; k16 < 2; ++k16) {
                        const uint64_t ah = desc_mn(st + OFF_A_HI + k16 * 2048, SLICE, 1024), al = desc_mn(st + OFF_A_LO + k16 * 2048, SLICE, 1024);
                        const uint64_t bh = desc_mn(st + OFF_B_HI + k16 * 2048, SLICE, 1024), bl = desc_mn(st + OFF_B_LO + k16 * 2048, SLICE, 1024);
                        const uint32_t acc = (i | (uint32_t)k16) ? 1u : 0u;
                        tc_mma_pair_f16(tmem, ah, bh, idesc, acc);
                        tc_mma_pair_f16(tmem, al, bh, idesc, 1);
                        tc_mma_pair_f16(tmem, ah, bl, idesc, 1);
                    }
                    tc_commit_pair(bar(EMPTY + s));
                }
                tc_commit_pair(bar(ACC_FULL));
                ++seg;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= flush: this CTA's 128 accumulator rows (lane = row) -> dW (+ bias column) =================
        const int q = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t seg = 0;
        for (int j = 0; j < p.n_jobs; ++j) {
            int t0, t1;
            if (!seg_range(j, t0, t1)) {
                // not a consumer of this job's stages, but EMPTY expects this warp's arrival only for aux jobs: nothing to do
                continue;
            }
            const WgradPairJob& jb = p.job[j];
            // free the stages: flush warps do not read them (the EMPTY count covers the MMA commit and, for the alpha job, warps 8-9)
            mbar_wait(bar(ACC_FULL), seg & 1u, err_flag, 95);
            tc_fence_after();
            const int m = (int)rank * 128 + q * 32 + lane;
            if (!jb.transposed) {
                float* drow = jb.dW + (size_t)m * jb.ldw + jb.col0;
                const bool vec = (reinterpret_cast<uintptr_t>(drow) & 15) == 0;       // per-thread: depends on m * ldw
                for (int c0 = 0; c0 < jb.NB; c0 += 32) {
                    float v[32];
                    tc_ld32(lane_addr + (uint32_t)c0, v);
                    if (vec && c0 + 32 <= jb.n_valid) {
#pragma unroll
                        for (int k = 0; k < 32; k += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + k), "f"(v[k]), "f"(v[k + 1]), "f"(v[k + 2]), "f"(v[k + 3]) : "memory");
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            if (c0 + k < jb.n_valid) atomicAdd(drow + c0 + k, v[k]);
                    }
                }
            } else {
                for (int c0 = 0; c0 < jb.NB; c0 += 32) {
                    float v[32];
                    tc_ld32(lane_addr + (uint32_t)c0, v);
#pragma unroll
                    for (int k = 0; k < 32; ++k) atomicAdd(jb.dW + (size_t)(c0 + k) * jb.ldw + m, v[k]);     // lanes = consecutive m: coalesced
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lbar(ACC_EMPTY));
            ++seg;
        }
    } else if (warp >= 8) {
        // ================= column sums out of the staged A slices: bias gradient dB[m] += sum_rows A[row][m]; on the merged view job
        //                   (A = h7) alpha_linear's weight gradient dWv[m] += sum_rows wrow[row] * A[row][m].  Four warps: warp (blk, part)
        //                   reads the hi or lo part of 64-column block blk; a lane owns 16-byte chunk (lane & 7) -- 8 columns -- of the rows
        //                   8 (lane >> 3) .. + 8 with LDS.128 (a 4 KB slice = 8 loads per lane), the four row groups are combined once per
        //                   job.  (Two warps walking 32 rows with 4-byte loads and a shuffle per row took ~2x the stage time of the other
        //                   jobs on the view job: its clusters ran 2.5x longer than the rest, ncu sm__cycles_active 40 % of elapsed.) ======
        const int blk = (warp - 8) & 1, part = (warp - 8) >> 1;
        const int chunk = lane & 7, rg = lane >> 3;
        uint32_t cnt = 0;
        for (int j = 0; j < p.n_jobs; ++j) {
            int t0, t1;
            if (!seg_range(j, t0, t1)) continue;
            const WgradPairJob& jb = p.job[j];
            const uint32_t iters = (uint32_t)(t1 - t0) * 4u;
            const bool aux = jb.wrow != nullptr, bias = jb.dB != nullptr;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float bs = 0.f;
            for (uint32_t i = 0; i < iters; ++i, ++cnt) {
                const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                // always wait for the stage: an arrival must not run ahead into a later phase of EMPTY[s]
                mbar_wait(bar(FULL + s), ph, err_flag, 96);
                float wv = 1.0f;
                if (aux) {              // lane = row of the slice; the weight is column wrow_col of the staged rows
                    const int64_t row = ((int64_t)t0 + (i >> 2)) * kTileRows + (int64_t)(i & 3u) * 32 + lane;
                    wv = row < p.rows ? reinterpret_cast<const float*>(sm + s * STAGE + OFF_WROW)[lane * jb.wrow_stride + jb.wrow_col] : 0.0f;
                    bs += wv;
                }
                if (aux || bias) {
                    const unsigned char* src = sm + s * STAGE + (part ? OFF_A_LO : OFF_A_HI) + blk * SLICE;
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int r = rg * 8 + rr;
                        const float w = aux ? __shfl_sync(0xffffffffu, wv, r) : 1.0f;
                        const uint4 q = *reinterpret_cast<const uint4*>(src + (uint32_t)r * 128u + ((uint32_t)(chunk ^ (r & 7)) << 4));
                        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            acc[2 * e] = fmaf(w, __uint_as_float(u[e] << 16), acc[2 * e]);
                            acc[2 * e + 1] = fmaf(w, __uint_as_float(u[e] & 0xffff0000u), acc[2 * e + 1]);
                        }
                    }
                    __syncwarp();
                }
                if (lane == 0) mbar_arrive(bar(EMPTY + s));          // every stage expects these four warps (uniform barrier count)
            }
            if (aux || bias) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
                    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
                }
                if (rg == 0) {
                    float* dst = (aux ? jb.dWv : jb.dB) + (int)rank * 128 + blk * 64 + chunk * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) atomicAdd(dst + e, acc[e]);
                }
                if (aux && rank == 0 && warp == 8 && jb.dBv) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) bs += __shfl_xor_sync(0xffffffffu, bs, o);
                    if (lane == 0) atomicAdd(jb.dBv, bs);
                }
            }
        }
    }

    if (trace && threadIdx.x == 0) {       // debug (bnrf_debug_mlp_trace): this CTA's busy time, stages and segments
        unsigned long long stages = 0, segs = 0, first = 99;
        for (int j = 0; j < p.n_jobs; ++j) {
            int t0, t1;
            if (seg_range(j, t0, t1)) { stages += 4ull * (unsigned long long)(t1 - t0); ++segs; if (first == 99) first = j; }
        }
        trace[blockIdx.x * 4 + 0] = (unsigned long long)(clock64() - k_t0);
        trace[blockIdx.x * 4 + 1] = stages; trace[blockIdx.x * 4 + 2] = segs; trace[blockIdx.x * 4 + 3] = first;
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

int launch_tile_wgrad_pair(bnrf_ctx* ctx, WgradPairParams& p, cudaStream_t st) {
    using namespace wgp;
    if (p.n_jobs <= 0 || p.n_jobs > kMaxPairJobs || p.tiles <= 0) return fail(ctx, BNRF_ERR_ARG, "tile_wgrad_pair: bad job list");
    int64_t units = 0;
    for (int j = 0; j < p.n_jobs; ++j) {
        WgradPairJob& jb = p.job[j];
        if (jb.NB != 256 && jb.NB != 128) return fail(ctx, BNRF_ERR_ARG, "tile_wgrad_pair: bad shape");
        // time per row tile: measured equal for the 64 KB (NB = 256) and the 48 KB (NB = 128) stages -- a CTA's stream is paced by
        // the latency of its NSTAGE slices in flight, and all of them together by the HBM read peak -- so every job weighs the same
        jb.weight = 4;
        jb.unit0 = units;
        units += (int64_t)jb.weight * p.tiles;
    }
    p.units = units;
    int64_t clusters = ctx->sm_count / 2;
    const int64_t max_useful = (int64_t)p.n_jobs * p.tiles;       // at least one tile per cluster
    if (clusters > max_useful) clusters = max_useful;
    static bool configured = false;
    if (!configured) {
        BNRF_CUDA(ctx, cudaFuncSetAttribute(tile_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        configured = true;
    }
    tile_wgrad_pair_kernel<<<(unsigned)(2 * clusters), THREADS, SMEM, st>>>(p, ctx->err_flag, ctx->trace);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bwt
}  // namespace bnrf
