// K4/K5 on CTA pairs: positional encoding + the 8x256 NeRF MLP with tcgen05.mma cta_group::2 (sm_100a).
//
// Same dataflow as mlp_tc.cu (hidden state resident in shared memory / TMEM, fp16 hi/lo split
// operands, 3 MMAs per K=16 slice, fp32 accumulation) but two CTAs of one TPC work as a pair on
// 256 rows: each CTA owns 128 rows (its A operand and its accumulators) and HALF of every
// weight tile (128 of the 256 output columns).  One thread of the leader CTA issues
// tcgen05.mma.cta_group::2 with M = 256; the hardware reads A from both CTAs' shared memory, the
// two B halves from one CTA each, and writes each CTA's 128 x N accumulator into its own TMEM.
//
// Why: measured on the single-CTA kernel (profiles/r01_mlp_tc_stall_trace.txt) every SM re-streamed
// the whole 2.3 MB weight image per 128-row tile -- 43 B/cycle/SM, the whole-chip L2 ceiling -- and
// SS-mode operand reads (96 B/cycle) + TMA fills + epilogue stores exceeded the 128 B/cycle shared
// memory port.  The pair halves both: 21 B/cycle/SM from L2, 64 B/cycle of operand reads.
//
// Cross-CTA protocol (all mbarriers live at the same shared-memory offset in both CTAs):
//   W_FULL[slot]   leader: 1 local expect_tx arrival + 1 relayed arrival from the peer (warp 2 of the
//                  peer waits on its own W_FULL and forwards); peer: local only
//   W_EMPTY[slot], ACC_FULL[2], PE_EMPTY   tcgen05.commit multicast to both CTAs
//   A_READY[8], PE_FULL   on the leader only: one elected-lane arrival per producing warp of BOTH
//                  CTAs (mapa + mbarrier.arrive.release.cluster), 16 resp. 8 per phase
//
// Nine GEMM steps per tile, not ten: feature_linear has no activation, so it is composed with the feature block of
// views_linears.0 into one 256 -> 128 linear when the weights are packed (common.cuh: wt9m, bias9m); step 8 reads h7.
//
// Replaces model/embedder.py:9-34 + model/nerf.py:67-116.
#include "tc_ptx.cuh"

namespace bnrf {
namespace tc2 {
using namespace tcp;

constexpr int TILE_M = 128;                        // rows per CTA (256 per pair)
constexpr int NUM_THREADS = 512;
constexpr int NS = 4;                              // weight ring depth
constexpr uint32_t STAGE_BYTES = 16384;            // this CTA's [128 x 64] fp16 SW128 half of one K-block of W_hi or W_lo
constexpr uint32_t KBLOCK_BYTES = 16384;           // one [128 x 64] fp16 SW128 A tile
constexpr uint32_t OFF_A_HI = 0;
constexpr uint32_t OFF_A_LO = 4 * KBLOCK_BYTES;
constexpr uint32_t OFF_PE_HI = 8 * KBLOCK_BYTES;
constexpr uint32_t OFF_PE_LO = 9 * KBLOCK_BYTES;
constexpr uint32_t OFF_W = 10 * KBLOCK_BYTES;
constexpr uint32_t OFF_BAR = OFF_W + NS * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;   // + alignment slack
constexpr int NUM_STEPS = 9;                       // L0..L7, then feature_linear and the view layer merged into one (common.cuh: wt9m)
constexpr int STAGES_PER_TILE = 68;                // 34 K-blocks x (hi, lo)
constexpr int N256_STAGES = 60;                    // stages of the eight 256-wide steps; the view layer's 8 stages are half size
constexpr uint32_t TMEM_COLS = 512;

enum { BAR_W_FULL = 0, BAR_W_EMPTY = BAR_W_FULL + NS, BAR_PE_FULL = BAR_W_EMPTY + NS, BAR_PE_EMPTY,
       BAR_A_READY, BAR_ACC_FULL = BAR_A_READY + 8, BAR_COUNT = BAR_ACC_FULL + 2 };

__host__ __device__ inline size_t stage_offset_bytes(int i) {
    return i < N256_STAGES ? (size_t)i * STAGE_BYTES : (size_t)N256_STAGES * STAGE_BYTES + (size_t)(i - N256_STAGES) * (STAGE_BYTES / 2);
}
__host__ __device__ inline size_t rank_stream_bytes() { return stage_offset_bytes(STAGES_PER_TILE); }

// TRAIN: the epilogue / front end also write what the backward pass reads (ActPtrs, common.cuh): every A operand
// they build (encoded points, h0..h7, feature) also goes to a bf16 hi/lo tile matrix in global memory (bwd_tiles.cuh: the backward GEMMs consume them without conversion), plus fp32 copies of the encoding and of
// the view layer for the two SIMT backward kernels.
template <int C, bool TRAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
mlp_tc2_kernel(TcParams p, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
               const float* __restrict__ viewbias, const float* __restrict__ z, int64_t rows, int S, int num_pairs,
               float* __restrict__ raw, const ActPtrs acts, unsigned int* err_flag, unsigned long long* __restrict__ trace) {
    // trace (debug, normally NULL): per-CTA stall accounting, 16 counters of clock64 cycles --
    //   0 kernel total   1 mma: wait PE_FULL   2 mma: wait A_READY   3 mma: wait W_FULL   4 mma: loop total
    //   5 tma: wait W_EMPTY   6 epilogue(warp 8): wait ACC_FULL   7 epilogue: loop total
    //   8 front end(warp 4): wait PE_EMPTY   9 front end: loop total
    const long long k_t0 = clock64();
    auto timed_wait = [&](uint32_t b, uint32_t parity, unsigned int code, unsigned long long& acc) {
        if (trace) {
            const long long t = clock64();
            mbar_wait(b, parity, err_flag, code);
            acc += (unsigned long long)(clock64() - t);
        } else {
            mbar_wait(b, parity, err_flag, code);
        }
    };
    auto timed_wait_cluster = [&](uint32_t b, uint32_t parity, unsigned int code, unsigned long long& acc) {
        if (trace) {
            const long long t = clock64();
            mbar_wait_cluster(b, parity, err_flag, code);
            acc += (unsigned long long)(clock64() - t);
        } else {
            mbar_wait_cluster(b, parity, err_flag, code);
        }
    };
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + OFF_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * BAR_COUNT);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();                  // 0 = leader (issues the MMAs), 1 = peer
    const int cluster = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
    const uint32_t lbar0 = mapa_u32(bar0, 0);                 // the leader's barrier block, as a shared::cluster address
    auto lbar = [&](int i) { return lbar0 + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(bar(BAR_W_FULL + i), rank == 0 ? 2 : 1); mbar_init(bar(BAR_W_EMPTY + i), 1); }
        mbar_init(bar(BAR_PE_FULL), 8);                       // 4 front-end warps x 2 CTAs
        mbar_init(bar(BAR_PE_EMPTY), 1);
        for (int i = 0; i < 8; ++i) mbar_init(bar(BAR_A_READY + i), 16);   // 8 epilogue warps x 2 CTAs
        for (int i = 0; i < 2; ++i) mbar_init(bar(BAR_ACC_FULL + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();                                       // barrier inits + TMEM of both CTAs visible before any remote arrive / MMA
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int my_iters = (num_pairs > cluster) ? (num_pairs - 1 - cluster) / n_clusters + 1 : 0;
    auto tile_of = [&](int it) { return (int64_t)2 * ((int64_t)cluster + (int64_t)it * n_clusters) + rank; };

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own half of every weight tile) =================
        if (elect_one()) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.stream) + (size_t)rank * rank_stream_bytes();
            uint32_t cnt = 0;
            unsigned long long w_empty = 0;
            for (int it = 0; it < my_iters; ++it) {
                for (int i = 0; i < STAGES_PER_TILE; ++i, ++cnt) {
                    const uint32_t slot = cnt % NS, ph = (cnt / NS) & 1u;
                    timed_wait(bar(BAR_W_EMPTY + slot), ph ^ 1u, 1, w_empty);
                    const uint32_t bytes = i < N256_STAGES ? STAGE_BYTES : STAGE_BYTES / 2;
                    mbar_expect_tx(bar(BAR_W_FULL + slot), bytes);
                    tma_bulk_load(base + OFF_W + slot * STAGE_BYTES, src + stage_offset_bytes(i), bytes, bar(BAR_W_FULL + slot));
                }
            }
            if (trace) trace[blockIdx.x * 16 + 5] = w_empty;
        }
    } else if (warp == 2) {
        // ================= peer only: forward "my half of the stage has landed" to the leader's W_FULL =================
        if (lane == 0 && rank == 1) {
            const uint32_t total = (uint32_t)my_iters * STAGES_PER_TILE;
            unsigned long long dummy = 0;
            for (uint32_t cnt = 0; cnt < total; ++cnt) {
                const uint32_t slot = cnt % NS, ph = (cnt / NS) & 1u;
                timed_wait(bar(BAR_W_FULL + slot), ph, 8, dummy);
                mbar_arrive_cluster(lbar(BAR_W_FULL + slot));
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA, one thread) =================
        if (rank == 0 && elect_one()) {
            uint32_t wcnt = 0;           // weight stages consumed
            uint32_t agen = 0;           // generation of the a_ready barriers (one per producing epilogue)
            unsigned long long w_pe = 0, w_a = 0, w_w = 0;
            const long long m_t0 = clock64();
            for (int it = 0; it < my_iters; ++it) {
                for (int t = 0; t < NUM_STEPS; ++t) {
                    const int N = (t == NUM_STEPS - 1) ? 128 : 256;
                    const uint32_t accb = ((uint32_t)it * NUM_STEPS + (uint32_t)t) & 1u;   // accumulators alternate over ALL steps (9 per tile: odd)
                    const uint32_t idesc = make_idesc(2 * TILE_M, N);
                    const uint32_t d_tmem = tmem + accb * 256u;
                    const bool has_pe = (t == 0 || t == 5);
                    const int n_act = (t == 0) ? 0 : 4;
                    uint32_t accumulate = 0;
                    for (int kb = has_pe ? -1 : 0; kb < n_act; ++kb) {
                        uint32_t a_hi, a_lo;
                        if (kb < 0) {
                            if (t == 0) timed_wait_cluster(bar(BAR_PE_FULL), (uint32_t)it & 1u, 2, w_pe);
                            a_hi = base + OFF_PE_HI; a_lo = base + OFF_PE_LO;
                        } else {
                            a_hi = base + OFF_A_HI + kb * KBLOCK_BYTES; a_lo = base + OFF_A_LO + kb * KBLOCK_BYTES;
                        }
                        {   // W_hi stage of this K-block: A_hi * W_hi and A_lo * W_hi
                            const uint32_t slot = wcnt % NS, ph = (wcnt / NS) & 1u;
                            timed_wait_cluster(bar(BAR_W_FULL + slot), ph, 4, w_w);
                            tc_fence_after();
                            const uint32_t w = base + OFF_W + slot * STAGE_BYTES;
#pragma unroll
                            for (int hk = 0; hk < 2; ++hk) {
                                if (kb >= 0) {
                                    timed_wait_cluster(bar(BAR_A_READY + kb * 2 + hk), agen & 1u, 3, w_a);
                                    tc_fence_after();
                                    if (trace && blockIdx.x == 0 && it == 5 && hk == 0) trace[148 * 16 + t * 8 + 1 + kb] = (unsigned long long)clock64();
                                }
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
                        {   // W_lo stage: A_hi * W_lo
                            const uint32_t slot = wcnt % NS, ph = (wcnt / NS) & 1u;
                            timed_wait_cluster(bar(BAR_W_FULL + slot), ph, 5, w_w);
                            tc_fence_after();
                            const uint32_t w = base + OFF_W + slot * STAGE_BYTES;
#pragma unroll
                            for (int k16 = 0; k16 < 4; ++k16)
                                tc_mma_pair_f16(d_tmem, make_desc(a_hi + k16 * 32, 0), make_desc(w + k16 * 32, 0), idesc, 1);
                            tc_commit_pair(bar(BAR_W_EMPTY + slot));
                            ++wcnt;
                        }
                        if (kb < 0 && t == 5) tc_commit_pair(bar(BAR_PE_EMPTY));   // encoded tiles no longer needed
                    }
                    tc_commit_pair(bar(BAR_ACC_FULL + accb));
                    if (trace && blockIdx.x == 0 && it == 5) trace[148 * 16 + t * 8 + 5] = (unsigned long long)clock64();
                    if (t >= 1) ++agen;                                        // steps 1..8 each consumed one generation
                }
            }
            if (trace) {
                trace[blockIdx.x * 16 + 1] = w_pe; trace[blockIdx.x * 16 + 2] = w_a; trace[blockIdx.x * 16 + 3] = w_w;
                trace[blockIdx.x * 16 + 4] = (unsigned long long)(clock64() - m_t0);
            }
        }
    } else if (warp >= 8) {
        // ================= epilogue: 8 warps, warp pair (w, w+4) shares TMEM lane quarter q and splits the columns =================
        const int q = warp & 3;                             // TMEM lane quarter this warp may access
        const int ch = (warp - 8) >> 2;                     // column half: 16 of every 32-column K-half
        const int r = q * 32 + lane;                        // row in tile == TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        // partial heads of the row, parked in a 16-byte chunk of A kb0 that the ch = 1 warp itself owns (columns 16..23 of row r):
        // the training-mode spill above re-reads A, and only a chunk's owner may overwrite it
        float* xchg = reinterpret_cast<float*>(sm + OFF_A_HI + sw128_offset(r, 16));
        uint32_t acc_uses[2] = {0, 0};
        unsigned long long w_acc = 0;
        const long long e_t0 = clock64();
        for (int it = 0; it < my_iters; ++it) {
            const int64_t tile = tile_of(it);
            const int64_t row = tile * TILE_M + r;
            float sigma_acc = 0.0f;
            for (int t = 0; t < NUM_STEPS; ++t) {
                const int b = (int)(((uint32_t)it * NUM_STEPS + (uint32_t)t) & 1u);
                // the smem carve-out leaves no L1: every __ldg is an L2 round trip, so the per-layer constants of the first
                // chunk are fetched BEFORE blocking on the accumulator and later chunks prefetch one chunk ahead
                const float inv_scale = __ldg(p.inv_scale + (t == NUM_STEPS - 1 ? 10 : t));      // slot 10 = the merged step
                const float* bias = p.bias[t < 8 ? t : 0] + ch * 16;
                float4 bq[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(bias) + j);
                timed_wait(bar(BAR_ACC_FULL + b), acc_uses[b] & 1u, 6, w_acc);
                ++acc_uses[b];
                tc_fence_after();
                const bool tl = trace && blockIdx.x == 0 && it == 5 && threadIdx.x == 256;
                if (tl) trace[148 * 16 + 128 + t * 8 + 0] = (unsigned long long)clock64();
                const uint32_t acc_addr = lane_addr + (uint32_t)b * 256u + (uint32_t)ch * (t < 8 ? 16u : 32u);
                if (t < 8) {
                    // hand-off granularity = one K-half (32 columns): every warp converts 16 columns of each K-half, so the
                    // MMA of the next layer can start after 1/8 of the epilogue.
                    uint32_t va[16], vb[16];
                    tc_ld16_issue(acc_addr, va);
#pragma unroll
                    for (int kh = 0; kh < 8; ++kh) {
                        uint32_t (&cur)[16] = (kh & 1) ? vb : va;
                        uint32_t (&nxt)[16] = (kh & 1) ? va : vb;
                        tc_ld16_wait(cur);
                        if (tl && kh == 0) trace[148 * 16 + 128 + t * 8 + 5] = (unsigned long long)clock64();
                        float4 bn[4];
                        if (kh < 7) {
                            tc_ld16_issue(acc_addr + (kh + 1) * 32, nxt);   // next K-half's accumulators in flight while this one is converted
#pragma unroll
                            for (int j = 0; j < 4; ++j) bn[j] = __ldg(reinterpret_cast<const float4*>(bias + (kh + 1) * 32) + j);
                        }
                        const int col0 = kh * 32;
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 bv = bq[j >> 2];
                            v[j] = fmaf(__uint_as_float(cur[j]), inv_scale, bv.x);
                            v[j + 1] = fmaf(__uint_as_float(cur[j + 1]), inv_scale, bv.y);
                            v[j + 2] = fmaf(__uint_as_float(cur[j + 2]), inv_scale, bv.z);
                            v[j + 3] = fmaf(__uint_as_float(cur[j + 3]), inv_scale, bv.w);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);                                 // ReLU
                        if (t == 7) {                                          // sigma head on h7 (model/nerf.py:101)
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 wa = __ldg(reinterpret_cast<const float4*>(p.w_alpha + ch * 16 + col0 + j));
                                sigma_acc = fmaf(v[j], wa.x, sigma_acc); sigma_acc = fmaf(v[j + 1], wa.y, sigma_acc);
                                sigma_acc = fmaf(v[j + 2], wa.z, sigma_acc); sigma_acc = fmaf(v[j + 3], wa.w, sigma_acc);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t off = (kh >> 1) * KBLOCK_BYTES + sw128_offset(r, (kh & 1) * 32 + ch * 16 + j * 8);
                            split_store8(v + 8 * j, sm + OFF_A_HI + off, sm + OFF_A_LO + off);
                        }
                        if (tl && kh == 0) trace[148 * 16 + 128 + t * 8 + 6] = (unsigned long long)clock64();
                        tc_fence_before();
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(lbar(BAR_A_READY + kh));   // one arrival per warp, on the leader's barrier
                        if (tl && (kh & 1) == 0) trace[148 * 16 + 128 + t * 8 + 1 + (kh >> 1)] = (unsigned long long)clock64();
                        if (kh < 7) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) bq[j] = bn[j];
                        }
                    }
                    if (TRAIN) {
                        // Spill h_t for the backward pass AFTER the whole layer has been handed
                        // over, while the tensor core is busy with the next one.  The A tile in shared memory and the bf16 tile in
                        // global memory have the same layout, so ANY thread can convert ANY 16-byte chunk: the 8 epilogue warps
                        // walk the 64 KB hi / lo parts in 512-byte spans (fp16 hi + lo = the value to 2^-22 -> bf16 hi / lo), which
                        // makes both the shared-memory reads and the global stores fully coalesced.  Stored from the hand-off
                        // loop instead (each thread its own 16-byte pieces, 128 B apart) the stores cost 32 LSU passes per
                        // instruction and 40 % of the kernel.
                        named_bar_sync(9, 256);                    // every epilogue warp of this CTA has written its share of A
                        const int ew = warp - 8;
                        unsigned char* gt = acts.h_tiles + ((size_t)t * (size_t)acts.t_alloc + (size_t)tile) * (8 * KBLOCK_BYTES);
                        unsigned char* mk = acts.mask_bits + ((size_t)t * (size_t)acts.t_alloc + (size_t)tile) * 4096;
#pragma unroll 4
                        for (int i = 0; i < 16; ++i) {
                            const uint32_t c = (uint32_t)((i * 8 + ew) * 32 + lane);      // 16-byte chunk of the 64 KB hi (and lo) part
                            const uint32_t off = c * 16u;
                            const uint4 hw = *reinterpret_cast<const uint4*>(sm + OFF_A_HI + off);
                            const uint4 lw = *reinterpret_cast<const uint4*>(sm + OFF_A_LO + off);
                            const uint32_t hh[4] = {hw.x, hw.y, hw.z, hw.w}, ll[4] = {lw.x, lw.y, lw.z, lw.w};
                            float v[8];
                            uint32_t bits = 0;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[e]));
                                const float2 cc = __half22float2(*reinterpret_cast<const __half2*>(&ll[e]));
                                v[2 * e] = a.x + cc.x; v[2 * e + 1] = a.y + cc.y;
                                bits |= ((hh[e] & 0x0000ffffu) ? 1u : 0u) << (2 * e);                      // post-ReLU: h > 0 <=> hi half != 0
                                bits |= ((hh[e] & 0xffff0000u) ? 1u : 0u) << (2 * e + 1);
                            }
                            split_store8_bf16_global(v, gt + off, gt + 4 * KBLOCK_BYTES + off);
                            mk[c] = (unsigned char)bits;                                                  // ReLU mask, 1 bit per activation
                        }
                        named_bar_sync(10, 256);                   // ... and nobody overwrites A before all of it has been read
                    }
                } else {
                    // view layer output (128 cols, 64 per warp of the pair) -> ReLU -> rgb head; write cat([rgb, sigma]) (model/nerf.py:103-110)
                    const int64_t ray = (row < rows) ? row / S : 0;
                    const float* vbp = viewbias + ray * kHalf + ch * 32;
                    float rgb[3] = {0.f, 0.f, 0.f};
                    uint32_t va[32], vb[32];
                    tc_ld32_issue(acc_addr, va);
                    tc_ld32_issue(acc_addr + 64, vb);
                    tc_ld_wait(va);
                    tc_ld_wait(vb);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t (&cur)[32] = h ? vb : va;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int col = h * 64 + j;                          // relative to this warp's first column
                            const float4 bv = __ldg(reinterpret_cast<const float4*>(vbp + col));
                            const float x0 = fmaxf(fmaf(__uint_as_float(cur[j]), inv_scale, bv.x), 0.0f);
                            const float x1 = fmaxf(fmaf(__uint_as_float(cur[j + 1]), inv_scale, bv.y), 0.0f);
                            const float x2 = fmaxf(fmaf(__uint_as_float(cur[j + 2]), inv_scale, bv.z), 0.0f);
                            const float x3 = fmaxf(fmaf(__uint_as_float(cur[j + 3]), inv_scale, bv.w), 0.0f);
                            if (TRAIN && row < rows)
                                *reinterpret_cast<float4*>(acts.h9_f32 + row * kHalf + ch * 32 + col) = make_float4(x0, x1, x2, x3);
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                const float4 wr = __ldg(reinterpret_cast<const float4*>(p.w_rgb + c * kHalf + ch * 32 + col));
                                rgb[c] = fmaf(x0, wr.x, rgb[c]); rgb[c] = fmaf(x1, wr.y, rgb[c]);
                                rgb[c] = fmaf(x2, wr.z, rgb[c]); rgb[c] = fmaf(x3, wr.w, rgb[c]);
                            }
                        }
                    }
                    tc_fence_before();
                    // combine the two column halves of each row: the upper-half warp parks its partial sums in the pair's own
                    // rows of the (idle between t = 9 and the next tile's first epilogue) A tile, the lower-half warp adds and writes.
                    if (ch == 1) {
                        *reinterpret_cast<float4*>(xchg) = make_float4(rgb[0], rgb[1], rgb[2], sigma_acc);
                        named_bar_arrive(1 + q, 64);
                        named_bar_sync(5 + q, 64);            // partner has read: the A rows may be overwritten again
                    } else {
                        named_bar_sync(1 + q, 64);
                        const float4 o = *reinterpret_cast<const float4*>(xchg);
                        named_bar_arrive(5 + q, 64);
                        if (row < rows) {
                            const float sg = sigma_acc + o.w + __ldg(p.b_alpha);
                            if (C == 3) {
                                *reinterpret_cast<float4*>(raw + row * 4) =
                                    make_float4(rgb[0] + o.x + __ldg(p.b_rgb), rgb[1] + o.y + __ldg(p.b_rgb + 1), rgb[2] + o.z + __ldg(p.b_rgb + 2), sg);
                            } else {
                                *reinterpret_cast<float2*>(raw + row * 2) = make_float2(rgb[0] + o.x + __ldg(p.b_rgb), sg);
                            }
                        }
                    }
                }
            }
        }
        if (trace && threadIdx.x == 256) {
            trace[blockIdx.x * 16 + 6] = w_acc; trace[blockIdx.x * 16 + 7] = (unsigned long long)(clock64() - e_t0);
        }
    } else if (warp >= 4) {
        // ================= front end: encode the next tile =================
        const int r = threadIdx.x - 128;
        unsigned long long w_pee = 0;
        const long long f_t0 = clock64();
        for (int it = 0; it < my_iters; ++it) {
            const int64_t tile = tile_of(it);
            const int64_t row = tile * TILE_M + r;
            float enc[64];
            float x[3] = {0.f, 0.f, 0.f};
            const bool live = row < rows;
            if (live) {
                const int64_t ray = row / S;
                const float zz = __ldg(z + row);
#pragma unroll
                for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(__ldg(rays_o + ray * 3 + c), __fmul_rn(__ldg(rays_d + ray * 3 + c), zz));
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) enc[c] = x[c];
#pragma unroll
            for (int k = 0; k < kPtsFreqs; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float s, co;
                    sincosf(x[c] * (float)(1 << k), &s, &co);
                    enc[3 + 6 * k + c] = live ? s : 0.0f;
                    enc[3 + 6 * k + 3 + c] = live ? co : 0.0f;
                }
            enc[63] = 0.0f;
            if (TRAIN && live) {
                float4* dst = reinterpret_cast<float4*>(acts.pe_f32 + row * kPtsChPad);
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[j] = make_float4(enc[4 * j], enc[4 * j + 1], enc[4 * j + 2], enc[4 * j + 3]);
            }
            timed_wait(bar(BAR_PE_EMPTY), ((uint32_t)it & 1u) ^ 1u, 7, w_pee);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                const uint32_t off = sw128_offset(r, c8 * 8);
                split_store8(enc + 8 * c8, sm + OFF_PE_HI + off, sm + OFF_PE_LO + off);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lbar(BAR_PE_FULL));
            if (TRAIN) {
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    unsigned char* gt = acts.pe_tiles + (size_t)tile * (2 * KBLOCK_BYTES) + sw128_offset(r, c8 * 8);
                    split_store8_bf16_global(enc + 8 * c8, gt, gt + KBLOCK_BYTES);
                }
            }
        }
        if (trace && threadIdx.x == 128) {
            trace[blockIdx.x * 16 + 8] = w_pee; trace[blockIdx.x * 16 + 9] = (unsigned long long)(clock64() - f_t0);
        }
    }

    tc_fence_before();
    cluster_sync_all();              // the leader's MMAs read the peer's shared memory and write its TMEM: leave together
    if (trace && threadIdx.x == 0) trace[blockIdx.x * 16 + 0] = (unsigned long long)(clock64() - k_t0);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------- weight stream (two ranks)
// Stage i of a tile = (GEMM step t, K-block, hi/lo) in MMA consumption order; the encoded-point K-block comes
// first for steps 0 and 5.  Rank r holds output columns [r * N/2, (r+1) * N/2) of every stage.
struct StageInfo { int step, k0, nh, lo; };
__host__ __device__ inline StageInfo stage_info(int i) {
    int kbi = i >> 1, lo = i & 1, t = 0;
    const int kbs[NUM_STEPS] = {1, 4, 4, 4, 4, 5, 4, 4, 4};
    while (kbi >= kbs[t]) { kbi -= kbs[t]; ++t; }
    return {t, kbi * 64, t == NUM_STEPS - 1 ? 64 : 128, lo};   // wt[t] rows are already ordered [pe64 | h256]; step 8 = merged
}

__global__ void pack_stream_kernel(const float* const* __restrict__ wt, const float* __restrict__ scale, __half* __restrict__ stream) {
    const int i = blockIdx.x, rank = blockIdx.y;
    const StageInfo si = stage_info(i);
    const int ti = si.step == NUM_STEPS - 1 ? 10 : si.step;      // table slot 10 = the merged feature + view step (common.cuh)
    const float* w = wt[ti];
    const float sc = scale[ti];
    const int N = 2 * si.nh;
    unsigned char* dst = reinterpret_cast<unsigned char*>(stream) + (size_t)rank * rank_stream_bytes() + stage_offset_bytes(i);
    for (int e = threadIdx.x; e < si.nh * 64; e += blockDim.x) {
        const int k = e / si.nh, n = e % si.nh;             // coalesced over n in the k-major source
        const float v = w[(size_t)(si.k0 + k) * N + rank * si.nh + n] * sc;
        const __half hi = __float2half_rn(v);
        const __half out = si.lo ? __float2half_rn(v - __half2float(hi)) : hi;
        *reinterpret_cast<__half*>(dst + sw128_offset(n, k)) = out;
    }
}

}  // namespace tc2

size_t tc2_stream_halfs() { return 2 * tc2::rank_stream_bytes() / sizeof(__half); }

// scale (device [10], computed by pack_tc_stream) -> the pair kernel's stream
int pack_tc2_stream(bnrf_ctx* ctx, int net, const float* const* table_dev, const float* scale_dev, cudaStream_t st) {
    using namespace tc2;
    NetParams& np = ctx->net[net];
    pack_stream_kernel<<<dim3(STAGES_PER_TILE, 2), 256, 0, st>>>(table_dev, scale_dev, np.tc2_stream);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

template <int C, bool TRAIN>
static int launch_one(bnrf_ctx* ctx, const tcp::TcParams& p, int clusters, const float* o, const float* d, const float* vb,
                      const float* z, int64_t rows, int S, int pairs, float* raw, const ActPtrs& acts, cudaStream_t st) {
    using namespace tc2;
    BNRF_CUDA(ctx, cudaFuncSetAttribute(mlp_tc2_kernel<C, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    mlp_tc2_kernel<C, TRAIN><<<2 * clusters, NUM_THREADS, SMEM_BYTES, st>>>(p, o, d, vb, z, rows, S, pairs, raw, acts, ctx->err_flag, ctx->trace);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int launch_mlp_tc2(bnrf_ctx* ctx, int net, const float* o, const float* d, const float* vb, const float* z,
                   int64_t n, int S, float* raw, const ActPtrs* acts, cudaStream_t st) {
    using namespace tc2;
    const NetParams& np = ctx->net[net];
    TcParams p;
    p.stream = np.tc2_stream; p.inv_scale = np.tc_scale;
    for (int i = 0; i < 10; ++i) p.bias[i] = np.bias[i];
    p.w_alpha = np.w_alpha; p.b_alpha = np.b_alpha; p.w_rgb = np.w_rgb; p.b_rgb = np.b_rgb;
    const int64_t rows = n * S;
    const int64_t pairs64 = ceil_div(rows, 2 * TILE_M);
    if (pairs64 > 0x3fffffff) return fail(ctx, BNRF_ERR_ARG, "mlp: too many rows");
    const int pairs = (int)pairs64;
    const int max_clusters = ctx->sm_count / 2;
    const int clusters = pairs < max_clusters ? pairs : max_clusters;
    const bool c3 = ctx->cfg.channels == 3;
    if (acts) {
        if (acts->t_alloc < 2 * (int64_t)pairs) return fail(ctx, BNRF_ERR_STATE, "mlp: activation tile matrices too small");
        return c3 ? launch_one<3, true>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, *acts, st)
                  : launch_one<1, true>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, *acts, st);
    }
    const ActPtrs none{};
    return c3 ? launch_one<3, false>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, none, st)
              : launch_one<1, false>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, none, st);
}

}  // namespace bnrf
