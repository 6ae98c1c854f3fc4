// The dgrad chain (dgrad_chain.cu) on CTA pairs: tcgen05.mma cta_group::2 with M = 256, each CTA of a TPC owns the gradient
// of its own 128-sample tile (A operand in its shared memory, accumulators in its TMEM) and HALF of every transposed
// weight tile, exactly as the forward kernel does (mlp_tc2.cu) and for the same reason: the single-CTA chain streams the
// whole 2 MB weight image per tile through every SM (43 B/cycle/SM, the whole-chip L2 ceiling); the pair halves that.
//
// Cross-CTA protocol (all mbarriers live at the same shared-memory offset in both CTAs):
//   W_FULL[slot]      leader: local expect_tx arrival + relayed arrival from the peer (peer warp 2); peer: local only
//   A0_FULL           leader: its own dZ9 loader's expect_tx arrival + the peer's relayed arrival (peer warp 5 forwards
//                     its local A0_LOCAL): both first A operands have landed
//   W_EMPTY[slot], ACC_FULL[2], A_FREE   tcgen05.commit multicast to both CTAs
//   A_READY[8]        on the leader only: one elected-lane arrival per epilogue warp of BOTH CTAs (16 per phase)
//   A_LOCAL[4], SPILLED   per CTA: the epilogue warps tell their own spill thread that a 64-column block of the new dZ is
//                     complete; the spill thread tells them when the bulk-copy engine has read it
// Steps, masks, rank-1 term and the fp32 encoded-points blocks are those of dgrad_chain.cu.
#include "tc_ptx.cuh"
#include "bwd_tiles.cuh"

namespace bnrf {
namespace dgp {
using namespace tcp;

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 768;                   // 8 role warps (6 used) + 16 epilogue warps
constexpr int NS = 6;                              // weight ring depth
constexpr uint32_t STAGE_BYTES = 16384;            // this CTA's [128 x 64] bf16 SW128 half of one K-block of W_hi or W_lo
constexpr uint32_t KBLOCK_BYTES = 16384;
constexpr uint32_t OFF_A_HI = 0;
constexpr uint32_t OFF_A_LO = 4 * KBLOCK_BYTES;
constexpr uint32_t OFF_W = 8 * KBLOCK_BYTES;
constexpr uint32_t OFF_BAR = OFF_W + NS * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 512 + 1024;   // barrier block + alignment slack
constexpr uint32_t TMEM_COLS = 512;
constexpr int NUM_STEPS = 9, NUM_PASSES = 10;

enum { BAR_W_FULL = 0, BAR_W_EMPTY = BAR_W_FULL + NS, BAR_A0_FULL = BAR_W_EMPTY + NS, BAR_A0_LOCAL, BAR_A_FREE, BAR_SPILLED,
       BAR_A_LOCAL = BAR_SPILLED + 4, BAR_A_READY = BAR_A_LOCAL + 4, BAR_ACC_FULL = BAR_A_READY + 8 /* [buffer][N-half] */,
       BAR_COUNT = BAR_ACC_FULL + 4 };
static_assert(8 * BAR_COUNT + 4 <= 512 && SMEM_BYTES <= 232448, "barrier block / shared memory budget");

// pass p: 0 = s0 (merged view step, K = 128), 1..3 = s1..s3, 4 = s3' (encoded-points block of layer 5), 5..8 = s4..s7, 9 = s8
struct PassInfo { int slot /*weight table slot: forward GEMM step, 10 = merged step*/, k0, N, K; };
__host__ __device__ inline PassInfo pass_info(int p) {
    switch (p) {
        case 0: return {10, 0, 256, 128};
        case 1: return {7, 0, 256, 256};
        case 2: return {6, 0, 256, 256};
        case 3: return {5, kPtsChPad, 256, 256};
        case 4: return {5, 0, 64, 256};
        case 5: return {4, 0, 256, 256};
        case 6: return {3, 0, 256, 256};
        case 7: return {2, 0, 256, 256};
        case 8: return {1, 0, 256, 256};
        default: return {0, 0, 64, 256};
    }
}
// The LAST SPLIT_KB K-blocks of every 256-deep, 256-wide pass are issued in two N-halves (columns 0..127, then 128..255): the
// epilogue of the first half -- which rewrites A K-blocks 0 and 1 in place -- then runs under the MMAs of the second half, which
// read K-blocks 2 / 3 only, instead of after the whole step has drained.  Stage order of such a pass: (kb 0..FULL_KB-1: W_hi 16 KB,
// W_lo 16 KB), then (half 0: kb FULL_KB..3: W_hi 8 KB, W_lo 8 KB), (half 1: ...).  Pass 0 (K = 128: its second K-block is A block 1, which the first half-epilogue
// would overwrite) and the two N = 64 passes are not split.
constexpr int SPLIT_KB = 1;                        // trailing K-blocks of a 256 x 256 pass issued in N-halves
constexpr int FULL_KB = 4 - SPLIT_KB;
__host__ __device__ inline bool pass_split(int p) { return p != 0 && p != 4 && p != 9; }
__host__ __device__ inline int pass_stages(int p) { return p == 0 ? 4 : (pass_split(p) ? 2 * FULL_KB + 4 * SPLIT_KB : 8); }
struct StageDesc { int kb, lo, half /* -1: all N */; uint32_t bytes; };
__host__ __device__ inline StageDesc stage_desc(int p, int i) {
    if (!pass_split(p)) return {i >> 1, i & 1, -1, (p == 4 || p == 9) ? STAGE_BYTES / 4 : STAGE_BYTES};
    if (i < 2 * FULL_KB) return {i >> 1, i & 1, -1, STAGE_BYTES};
    const int j = i - 2 * FULL_KB;                 // half 0: kb FULL_KB .. 3 (hi, lo each), then half 1
    return {FULL_KB + ((j >> 1) % SPLIT_KB), j & 1, j / (2 * SPLIT_KB), STAGE_BYTES / 2};
}
__host__ __device__ inline size_t pass_bytes(int p) { return p == 0 ? 4 * (size_t)STAGE_BYTES : ((p == 4 || p == 9) ? 2 * (size_t)STAGE_BYTES : 8 * (size_t)STAGE_BYTES); }
__host__ __device__ inline size_t pass_offset_bytes(int p) {
    size_t off = 0;
    for (int q = 0; q < p; ++q) off += pass_bytes(q);
    return off;
}
__host__ __device__ inline size_t stage_offset_in_pass(int p, int i) {
    size_t off = 0;
    for (int q = 0; q < i; ++q) off += stage_desc(p, q).bytes;
    return off;
}
__host__ __device__ inline size_t rank_stream_bytes() { return pass_offset_bytes(NUM_PASSES); }
constexpr int STAGES_PER_TILE = 4 + 7 * (2 * FULL_KB + 4 * SPLIT_KB) + 2 * 8;

__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
dgrad_chain_pair_kernel(const unsigned char* __restrict__ stream, const unsigned char* __restrict__ dz9_tiles,
                        const unsigned char* __restrict__ mask_bits, int64_t t_alloc, const float* __restrict__ d_sigma,
                        int64_t d_sigma_stride, const float* __restrict__ w_alpha, int64_t rows, int num_pairs, int64_t dz_tile_count,
                        unsigned char* __restrict__ dz_tiles, float* __restrict__ d_pe, unsigned int* err_flag,
                        unsigned long long* __restrict__ trace) {
    // trace (debug, normally NULL): timeline of CTA 0, tile pair 2: [s * 16 + 0] epilogue sees ACC half 0, [1 + kh] chunk kh handed over,
    // [9] ACC half 1 seen; [s * 16 + 10] MMA issuer reaches the step, [11] commits half 0, [12] commits half 1 (tools/chain_trace.py)
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + OFF_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * BAR_COUNT);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const int cluster = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
    const uint32_t lbar0 = mapa_u32(bar0, 0);
    auto lbar = [&](int i) { return lbar0 + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(bar(BAR_W_FULL + i), rank == 0 ? 2 : 1); mbar_init(bar(BAR_W_EMPTY + i), 1); }
        mbar_init(bar(BAR_A0_FULL), 2);
        mbar_init(bar(BAR_A0_LOCAL), 1);
        mbar_init(bar(BAR_A_FREE), 1);
        for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_SPILLED + i), 1);
        for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_A_LOCAL + i), 32);      // 16 epilogue warps x 2 K-halves
        for (int i = 0; i < 8; ++i) mbar_init(bar(BAR_A_READY + i), 32);      // 16 epilogue warps x 2 CTAs
        for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_ACC_FULL + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int my_iters = (num_pairs > cluster) ? (num_pairs - 1 - cluster) / n_clusters + 1 : 0;
    auto tile_of = [&](int it) { return (int64_t)2 * ((int64_t)cluster + (int64_t)it * n_clusters) + rank; };

    if (warp == 0) {
        // ================= weight producer (both CTAs: own half of every weight tile) =================
        if (elect_one()) {
            uint32_t cnt = 0;
            auto load = [&](const unsigned char* src, uint32_t bytes) {
                const uint32_t slot = cnt % NS, ph = (cnt / NS) & 1u;
                mbar_wait(bar(BAR_W_EMPTY + slot), ph ^ 1u, err_flag, 71);
                mbar_expect_tx(bar(BAR_W_FULL + slot), bytes);
                tma_bulk_load(base + OFF_W + slot * STAGE_BYTES, src, bytes, bar(BAR_W_FULL + slot));
                ++cnt;
            };
            for (int it = 0; it < my_iters; ++it) {
                const unsigned char* rs = stream + (size_t)rank * rank_stream_bytes();
                for (int p = 0; p < NUM_PASSES; ++p) {
                    if (p == 4) continue;                        // consumed inside pass 3, below
                    const unsigned char* src = rs + pass_offset_bytes(p);
                    const int n = pass_stages(p);
                    for (int i = 0; i < n; ++i) {
                        if (p == 3 && i == 2 * FULL_KB) {        // the encoded-points pass of layer 5 runs before the two N-halves of pass 3
                            const unsigned char* s4 = rs + pass_offset_bytes(4);
                            for (int j = 0; j < pass_stages(4); ++j, s4 += STAGE_BYTES / 4) load(s4, STAGE_BYTES / 4);
                        }
                        const uint32_t bytes = stage_desc(p, i).bytes;
                        load(src, bytes);
                        src += bytes;
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ================= peer only: forward "my half of the stage has landed" to the leader's W_FULL =================
        if (lane == 0 && rank == 1) {
            const uint32_t total = (uint32_t)my_iters * STAGES_PER_TILE;
            for (uint32_t cnt = 0; cnt < total; ++cnt) {
                const uint32_t slot = cnt % NS, ph = (cnt / NS) & 1u;
                mbar_wait(bar(BAR_W_FULL + slot), ph, err_flag, 72);
                mbar_arrive_cluster(lbar(BAR_W_FULL + slot));
            }
        }
    } else if (warp == 4) {
        // ================= dZ9 loader: this CTA's first A operand (K = 128: two K-blocks, hi and lo parts) =================
        if (elect_one()) {
            const int full_bar = rank == 0 ? BAR_A0_FULL : BAR_A0_LOCAL;
            for (int it = 0; it < my_iters; ++it) {
                const int64_t tile = tile_of(it);
                mbar_wait(bar(BAR_A_FREE), ((uint32_t)it & 1u) ^ 1u, err_flag, 73);       // last MMAs of the previous tile pair have read A
                if (it > 0) mbar_wait(bar(BAR_SPILLED + 3), ((uint32_t)(it - 1) * 8u + 7u) & 1u, err_flag, 74);   // ... and dZ0 has been copied out
                const unsigned char* src = dz9_tiles + (size_t)tile * bwt::tile_bytes(kHalf);
                mbar_expect_tx(bar(full_bar), 4u * KBLOCK_BYTES);
                tma_bulk_load(base + OFF_A_HI, src, 2u * KBLOCK_BYTES, bar(full_bar));
                tma_bulk_load(base + OFF_A_LO, src + bwt::tile_part_bytes(kHalf), 2u * KBLOCK_BYTES, bar(full_bar));
            }
        }
    } else if (warp == 5) {
        // ================= peer only: forward "my dZ9 has landed" to the leader's A0_FULL =================
        if (lane == 0 && rank == 1) {
            for (int it = 0; it < my_iters; ++it) {
                mbar_wait(bar(BAR_A0_LOCAL), (uint32_t)it & 1u, err_flag, 75);
                mbar_arrive_cluster(lbar(BAR_A0_FULL));
            }
        }
    } else if (warp == 3) {
        // ================= spill: this CTA's dZ_l tiles -> global memory through the bulk-copy engine =================
        if (elect_one()) {
            const size_t dz_mat = (size_t)dz_tile_count * bwt::tile_bytes(kWidth);
            uint32_t sgen = 0;
            for (int it = 0; it < my_iters; ++it) {
                const int64_t tile = tile_of(it);
                for (int s = 0; s < NUM_STEPS - 1; ++s, ++sgen) {
                    unsigned char* gt = dz_tiles + (size_t)(7 - s) * dz_mat + (size_t)tile * bwt::tile_bytes(kWidth);
                    for (int kb = 0; kb < 4; ++kb) {
                        mbar_wait(bar(BAR_A_LOCAL + kb), sgen & 1u, err_flag, 76);
                        bulk_store(gt + (size_t)kb * KBLOCK_BYTES, base + OFF_A_HI + kb * KBLOCK_BYTES, KBLOCK_BYTES);
                        bulk_store(gt + (size_t)(4 + kb) * KBLOCK_BYTES, base + OFF_A_LO + kb * KBLOCK_BYTES, KBLOCK_BYTES);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");      // one group per 64-column block ...
                        if (kb > 0) {                                                   // ... so that block kb-1 is released as soon as it has been read
                            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            mbar_arrive(bar(BAR_SPILLED + kb - 1));
                        }
                    }
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    mbar_arrive(bar(BAR_SPILLED + 3));
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA, one thread) =================
        if (rank == 0 && elect_one()) {
            uint32_t wcnt = 0, agen = 0;
            const uint32_t idesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)((2 * TILE_M) >> 4) << 24);
            const uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)((2 * TILE_M) >> 4) << 24);
            // one 64-deep K-block: the W_hi stage feeds A_hi * W_hi and A_lo * W_hi, the W_lo stage A_hi * W_lo
            auto kblock = [&](uint32_t d_tmem, uint32_t idesc, int kb, bool wait_a, uint32_t& accumulate) {
                const uint32_t a_hi = base + OFF_A_HI + kb * KBLOCK_BYTES, a_lo = base + OFF_A_LO + kb * KBLOCK_BYTES;
                {
                    const uint32_t slot = wcnt % NS, ph = (wcnt / NS) & 1u;
                    mbar_wait_cluster(bar(BAR_W_FULL + slot), ph, err_flag, 77);
                    tc_fence_after();
                    const uint32_t w = base + OFF_W + slot * STAGE_BYTES;
#pragma unroll
                    for (int hk = 0; hk < 2; ++hk) {
                        if (wait_a) { mbar_wait_cluster(bar(BAR_A_READY + kb * 2 + hk), agen & 1u, err_flag, 78); tc_fence_after(); }
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint32_t ko = (uint32_t)(hk * 2 + kk) * 32u;
                            const uint64_t bd = make_desc(w + ko, 0);
                            tc_mma_pair_f16(d_tmem, make_desc(a_hi + ko, 0), bd, idesc, accumulate);
                            accumulate = 1;
                            tc_mma_pair_f16(d_tmem, make_desc(a_lo + ko, 0), bd, idesc, 1);
                        }
                    }
                    tc_commit_pair(bar(BAR_W_EMPTY + slot));
                    ++wcnt;
                }
                {
                    const uint32_t slot = wcnt % NS, ph = (wcnt / NS) & 1u;
                    mbar_wait_cluster(bar(BAR_W_FULL + slot), ph, err_flag, 79);
                    tc_fence_after();
                    const uint32_t w = base + OFF_W + slot * STAGE_BYTES;
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16)
                        tc_mma_pair_f16(d_tmem, make_desc(a_hi + k16 * 32, 0), make_desc(w + k16 * 32, 0), idesc, 1);
                    tc_commit_pair(bar(BAR_W_EMPTY + slot));
                    ++wcnt;
                }
            };
            const uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)((2 * TILE_M) >> 4) << 24);
            for (int it = 0; it < my_iters; ++it) {
                for (int s = 0; s < NUM_STEPS; ++s) {
                    const uint32_t accb = ((uint32_t)it * NUM_STEPS + (uint32_t)s) & 1u;
                    const uint32_t d_tmem = tmem + accb * 256u;
                    const bool split = s >= 1 && s <= NUM_STEPS - 2;              // the 256-deep, 256-wide steps
                    uint32_t accumulate = 0;
                    const bool tl = trace && blockIdx.x == 0 && it == 2;
                    if (tl) trace[s * 16 + 10] = (unsigned long long)clock64();
                    if (s == 0) { mbar_wait_cluster(bar(BAR_A0_FULL), (uint32_t)it & 1u, err_flag, 80); tc_fence_after(); }
                    if (!split) {
                        const int n_kb = (s == 0) ? 2 : 4;
                        for (int kb = 0; kb < n_kb; ++kb) kblock(d_tmem, (s == NUM_STEPS - 1) ? idesc64 : idesc256, kb, s > 0, accumulate);
                        tc_commit_pair(bar(BAR_ACC_FULL + accb * 2 + 0));
                        tc_commit_pair(bar(BAR_ACC_FULL + accb * 2 + 1));
                    } else {
                        for (int kb = 0; kb < FULL_KB; ++kb) kblock(d_tmem, idesc256, kb, true, accumulate);
                        if (s == 3) {   // encoded-points block of layer 5 from the same dZ5 (all four K-blocks), into the first 64 columns of the
                                        // OTHER buffer (its epilogue has consumed them: A_READY 0..5 were waited above); before the halves, so that
                                        // only the second half still reads A when the first half's epilogue starts rewriting K-blocks 0 and 1
                            uint32_t acc2 = 0;
                            for (int kb = 0; kb < 4; ++kb) kblock(tmem + (accb ^ 1u) * 256u, idesc64, kb, kb >= FULL_KB, acc2);
                        }
                        uint32_t a0 = 1, a1 = 1;
                        for (int kb = FULL_KB; kb < 4; ++kb) kblock(d_tmem, idesc128, kb, true, a0);            // columns 0..127 complete
                        tc_commit_pair(bar(BAR_ACC_FULL + accb * 2 + 0));
                        if (tl) trace[s * 16 + 11] = (unsigned long long)clock64();
                        for (int kb = FULL_KB; kb < 4; ++kb) kblock(d_tmem + 128u, idesc128, kb, false, a1);    // columns 128..255
                        tc_commit_pair(bar(BAR_ACC_FULL + accb * 2 + 1));
                        if (tl) trace[s * 16 + 12] = (unsigned long long)clock64();
                    }
                    if (s == NUM_STEPS - 1) tc_commit_pair(bar(BAR_A_FREE));
                    if (s >= 1) ++agen;
                }
            }
        }
    } else if (warp >= 8) {
        // ================= epilogue: 16 warps; the four warps of a TMEM lane quarter q split every 32-column K-half into 8-column pieces.
        // (The step is a serial chain -- last hand-off of step s -> K-block 3 of step s+1 -> first hand-off of step s+1 -- so its
        // period is the epilogue's time for its 8 hand-offs plus ~2.5 k cycles: 8 warps at 2 per scheduler were latency-bound at
        // ~600 cycles per hand-off, tools/chain_trace.py; 16 warps do half the work each with twice the latency hiding.) =================
        const int q = warp & 3;
        const int c4 = (warp - 8) >> 2;                     // 8-column piece of every K-half
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        for (int it = 0; it < my_iters; ++it) {
            const int64_t tile = tile_of(it);
            const int64_t row = tile * TILE_M + r;
            const bool live = row < rows;
            float* pe_row = d_pe + row * kPtsChPad + c4 * 16;
            for (int s = 0; s < NUM_STEPS; ++s) {
                const int b = (int)(((uint32_t)it * NUM_STEPS + (uint32_t)s) & 1u);
                unsigned long long mrow[4] = {~0ull, ~0ull, ~0ull, ~0ull};
                if (s < NUM_STEPS - 1) {
                    const unsigned char* mk = mask_bits + ((size_t)(7 - s) * (size_t)t_alloc + (size_t)tile) * 4096 + (size_t)r * 8;
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) mrow[kb] = __ldg(reinterpret_cast<const unsigned long long*>(mk + kb * 1024));
                }
                const float rr = (s == 0 && live) ? __ldg(d_sigma + row * d_sigma_stride) : 0.0f;
                float4 wq[2];
                if (s == 0) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) wq[j] = __ldg(reinterpret_cast<const float4*>(w_alpha + c4 * 8) + j);
                }
                // both N-half barriers of a buffer complete once per step: parity of the buffer's use count
                const uint32_t acc_par = ((((uint32_t)it * NUM_STEPS + (uint32_t)s) >> 1) & 1u);
                mbar_wait(bar(BAR_ACC_FULL + b * 2 + 0), acc_par, err_flag, 81);
                tc_fence_after();
                const bool tl = trace && blockIdx.x == 0 && it == 2 && threadIdx.x == 256;
                if (tl) trace[s * 16 + 0] = (unsigned long long)clock64();
                if (s == 3) {           // d pe from layer 5, parked in the other buffer by MMAs issued before the first half's commit; read
                                        // NOW: the next step's MMAs overwrite that buffer as soon as the first chunks below are handed over
                    uint32_t pv[16];
                    tc_ld16_issue(lane_addr + (uint32_t)(b ^ 1) * 256u + (uint32_t)c4 * 16u, pv);
                    tc_ld16_wait(pv);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            reinterpret_cast<float4*>(pe_row)[j] = make_float4(__uint_as_float(pv[4 * j]), __uint_as_float(pv[4 * j + 1]),
                                                                               __uint_as_float(pv[4 * j + 2]), __uint_as_float(pv[4 * j + 3]));
                    }
                }
                if (s < NUM_STEPS - 1) {
                    const uint32_t acc_addr = lane_addr + (uint32_t)b * 256u + (uint32_t)c4 * 8u;
                    uint32_t va[8], vb[8];
                    tc_ld8_issue(acc_addr, va);
#pragma unroll
                    for (int kh = 0; kh < 8; ++kh) {
                        uint32_t (&cur)[8] = (kh & 1) ? vb : va;
                        uint32_t (&nxt)[8] = (kh & 1) ? va : vb;
                        if (kh == 4) {          // columns 128..255: the second N-half
                            mbar_wait(bar(BAR_ACC_FULL + b * 2 + 1), acc_par, err_flag, 82);
                            tc_fence_after();
                            if (tl) trace[s * 16 + 9] = (unsigned long long)clock64();
                            tc_ld8_issue(acc_addr + 4 * 32, cur);
                        }
                        tc_ld8_wait(cur);
                        // the 64-column block this chunk overwrites (dZ of the previous step) has been copied out
                        if (s >= 1 && (kh & 1) == 0) mbar_wait(bar(BAR_SPILLED + (kh >> 1)), ((uint32_t)it * 8u + (uint32_t)(s - 1)) & 1u, err_flag, 61);
                        float4 wn[2];
                        if (kh < 7) {
                            if (kh != 3) tc_ld8_issue(acc_addr + (kh + 1) * 32, nxt);       // (not across the N-half boundary)
                            if (s == 0) {
#pragma unroll
                                for (int j = 0; j < 2; ++j) wn[j] = __ldg(reinterpret_cast<const float4*>(w_alpha + c4 * 8 + (kh + 1) * 32) + j);
                            }
                        }
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(cur[j]);
                        if (s == 0) {
#pragma unroll
                            for (int j = 0; j < 8; j += 4) {
                                const float4 wv = wq[j >> 2];
                                v[j] = fmaf(rr, wv.x, v[j]); v[j + 1] = fmaf(rr, wv.y, v[j + 1]);
                                v[j + 2] = fmaf(rr, wv.z, v[j + 2]); v[j + 3] = fmaf(rr, wv.w, v[j + 3]);
                            }
                        }
                        {
                            const int lc = (kh & 1) * 4 + c4;                   // 16-byte chunk of the 128-byte tile row = 8 columns = one mask byte
                            const uint32_t bits = (uint32_t)(mrow[kh >> 1] >> (8 * (lc ^ (r & 7)))) & 0xffu;
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : 0.0f;
                        }
                        {
                            const uint32_t off = (kh >> 1) * KBLOCK_BYTES + sw128_offset(r, (kh & 1) * 32 + c4 * 8);
                            uint4 hi, lo;
                            bwt::split8_bf16_pub(v, hi, lo);
                            *reinterpret_cast<uint4*>(sm + OFF_A_HI + off) = hi;
                            *reinterpret_cast<uint4*>(sm + OFF_A_LO + off) = lo;
                        }
                        tc_fence_before();
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive_cluster(lbar(BAR_A_READY + kh));       // MMA issuer (leader)
                            mbar_arrive(bar(BAR_A_LOCAL + (kh >> 1)));          // this CTA's spill thread
                        }
                        if (tl) trace[s * 16 + 1 + kh] = (unsigned long long)clock64();
                        if (s == 0 && kh < 7) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) wq[j] = wn[j];
                        }
                    }
                } else {
                    mbar_wait(bar(BAR_ACC_FULL + b * 2 + 1), acc_par, err_flag, 82);        // keep both barriers' phases in step
                    uint32_t pv[16];
                    tc_ld16_issue(lane_addr + (uint32_t)b * 256u + (uint32_t)c4 * 16u, pv);
                    tc_ld16_wait(pv);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float4 o = reinterpret_cast<float4*>(pe_row)[j];
                            o.x += __uint_as_float(pv[4 * j]); o.y += __uint_as_float(pv[4 * j + 1]);
                            o.z += __uint_as_float(pv[4 * j + 2]); o.w += __uint_as_float(pv[4 * j + 3]);
                            reinterpret_cast<float4*>(pe_row)[j] = o;
                        }
                    }
                    tc_fence_before();
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

// rank r's stage (pass p, K-block kb, hi/lo): element (n, k) of its [N/2 x 64] SW128 tile <- wt[slot][(k0 + r*N/2 + n) * K + kb*64 + k]
struct ChainPackNets { const float* const* wt[2]; unsigned char* stream[2]; };
__global__ void pack_chain_pair_stream_kernel(const __grid_constant__ ChainPackNets nets) {       // blockIdx.z = 2 * network + rank
    const float* const* __restrict__ wt = nets.wt[blockIdx.z >> 1];
    unsigned char* __restrict__ stream = nets.stream[blockIdx.z >> 1];
    const int p = blockIdx.y, rank = blockIdx.z & 1;
    const PassInfo pi = pass_info(p);
    const int i = blockIdx.x;
    if (i >= pass_stages(p)) return;
    const StageDesc sd = stage_desc(p, i);
    // B rows = output features of the pass: rank r holds [r N/2, (r+1) N/2) of an all-N stage, [128 h + 64 r, + 64) of N-half h
    const int nrows = (int)(sd.bytes / 128u);
    const int f0 = sd.half < 0 ? rank * (pi.N / 2) : 128 * sd.half + 64 * rank;
    const float* w = wt[pi.slot] + (size_t)(pi.k0 + f0) * pi.K;
    unsigned char* dst = stream + (size_t)rank * rank_stream_bytes() + pass_offset_bytes(p) + stage_offset_in_pass(p, i);
    for (int e = threadIdx.x; e < nrows * 64; e += blockDim.x) {
        const int n = e >> 6, k = e & 63;
        const float v = w[(size_t)n * pi.K + sd.kb * 64 + k];
        uint32_t hb, lb;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hb) : "f"(0.0f), "f"(v));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lb) : "f"(0.0f), "f"(v - __uint_as_float(hb << 16)));
        *reinterpret_cast<unsigned short*>(dst + sw128_offset(n, k)) = (unsigned short)((sd.lo ? lb : hb) & 0xffffu);
    }
}

}  // namespace dgp

size_t dgrad_chain_pair_stream_bytes() { return 2 * dgp::rank_stream_bytes(); }

int pack_dgrad_chain_pair_stream(bnrf_ctx* ctx, int net, cudaStream_t st) {
    NetParams& np = ctx->net[net];
    dgp::ChainPackNets nets{};
    nets.wt[0] = np.wt_table; nets.stream[0] = np.dgp_stream;
    dgp::pack_chain_pair_stream_kernel<<<dim3(2 * dgp::FULL_KB + 4 * dgp::SPLIT_KB, dgp::NUM_PASSES, 2), 256, 0, st>>>(nets);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int pack_dgrad_chain_pair_stream_both(bnrf_ctx* ctx, cudaStream_t st) {
    dgp::ChainPackNets nets{};
    for (int n = 0; n < 2; ++n) { nets.wt[n] = ctx->net[n].wt_table; nets.stream[n] = ctx->net[n].dgp_stream; }
    dgp::pack_chain_pair_stream_kernel<<<dim3(2 * dgp::FULL_KB + 4 * dgp::SPLIT_KB, dgp::NUM_PASSES, 4), 256, 0, st>>>(nets);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int launch_dgrad_chain_pair(bnrf_ctx* ctx, int net, const unsigned char* dz9_tiles, const unsigned char* mask_bits, int64_t t_alloc,
                            const float* d_sigma, int64_t d_sigma_stride, int64_t rows, int64_t dz_tile_count, unsigned char* dz_tiles,
                            float* d_pe, cudaStream_t st) {
    using namespace dgp;
    const NetParams& np = ctx->net[net];
    const int64_t pairs64 = ceil_div(rows, 2 * TILE_M);
    if (pairs64 > 0x3fffffff) return fail(ctx, BNRF_ERR_ARG, "dgrad chain: too many rows");
    if (dz_tile_count < 2 * pairs64 || t_alloc < 2 * pairs64) return fail(ctx, BNRF_ERR_STATE, "dgrad chain: tile matrices too small for CTA pairs");
    const int pairs = (int)pairs64;
    const int max_clusters = ctx->sm_count / 2;
    const int clusters = pairs < max_clusters ? pairs : max_clusters;
    BNRF_CUDA(ctx, cudaFuncSetAttribute(dgrad_chain_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    dgrad_chain_pair_kernel<<<2 * clusters, NUM_THREADS, SMEM_BYTES, st>>>(np.dgp_stream, dz9_tiles, mask_bits, t_alloc, d_sigma, d_sigma_stride,
                                                                            np.w_alpha, rows, pairs, dz_tile_count, dz_tiles, d_pe, ctx->err_flag,
                                                                            ctx->trace ? ctx->trace + 3072 : nullptr);   // (its own region of the debug buffer)
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bnrf
