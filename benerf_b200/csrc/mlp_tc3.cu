// K4/K5 on CTA pairs with the hidden state in TENSOR MEMORY: positional encoding + the 8x256 NeRF MLP, tcgen05.mma
// cta_group::2 in the ".ts" form (A operand read from TMEM), sm_100a.
//
// Same arithmetic as mlp_tc2.cu (fp16 hi/lo split operands, 3 MMAs per K = 16 slice, fp32 accumulation, nine GEMM steps per
// tile with feature_linear merged into the view layer), different dataflow.  mlp_tc2 keeps the activations of a tile as a
// 128 KB shared-memory A operand that the epilogue rewrites IN PLACE, so a layer's epilogue cannot start before the layer's
// last MMA has read A, and every layer boundary drains the tensor pipe (~1.7 k cycles x 8, profiles/r01_mlp_tc2_stall_trace.txt);
// shared memory is full, so neither a second tile nor a deeper weight ring fits.  Here:
//
//   * The epilogue converts a 32-column chunk of the fp32 accumulator to fp16 hi/lo and writes it back with tcgen05.st over
//     the very columns it came from: 32 fp32 columns = 16 packed hi + 16 packed lo columns (lane = row, one 32-bit column =
//     two consecutive K elements).  The accumulator buffer of layer t thereby BECOMES the A operand of layer t + 1, and the
//     other 256-column buffer (layer t's A, dead once its MMAs retire) receives layer t + 1's accumulator: two buffers
//     ping-pong, TMEM (512 columns) is exactly enough and the 128 KB of shared memory are free.
//   * Because A and D never alias, the tail of a layer can be issued in N-halves (make_schedule: three full K-blocks, then the
//     last K-block against columns 0..127 and against 128..255, each half with its own ACC_FULL barrier): the epilogue of
//     columns 0..127 runs while the tensor core still works on columns 128..255, and its output is exactly the first two
//     K-blocks of the next layer, which therefore starts without draining the pipe.  (Splitting EVERY K-block made the kernel
//     bound by the tensor-memory read port: an A operand from TMEM costs 4 KB per MMA whatever N is.)
//   * Shared memory holds the encoded points (SS-form MMAs for those K-blocks), a deep ring of 16 KB weight stages, the
//     epilogue constants, and -- in training mode -- a 3-slot staging ring from which ONE thread bulk-stores the bf16 hi/lo
//     activation tiles and ReLU mask bits the backward pass reads (no re-read / convert pass over A as in mlp_tc2).
//
// Cross-CTA protocol (all mbarriers live at the same shared-memory offset in both CTAs):
//   W_FULL[slot]   leader: 1 local expect_tx arrival + 1 relayed arrival from the peer; peer: local only
//   W_EMPTY[slot], ACC_FULL[buffer][N-half], PE_EMPTY   tcgen05.commit multicast to both CTAs
//   A_READY[8], PE_FULL   on the leader only: one elected-lane arrival per producing warp of BOTH CTAs (8 per phase)
//   ST_FULL[3], ST_EMPTY[3]   per CTA (training mode): staging slot filled by the 8 epilogue warps / read by the copy engine
//
// FUSE (forward-only renders with S = 32 / 64 / 128): the epilogue parks (rgb, sigma) of every row in shared memory instead of storing
// `raw`, and the front-end warps run NeRF.raw2output (model/nerf.py:118-148) on them between two encodings (RAW_FULL / RAW_EMPTY).
//
// Replaces model/embedder.py:9-34 + model/nerf.py:67-116.
#include <stdlib.h>
#include "tc_ptx.cuh"
#include "bwd_tiles.cuh"
#include "composite.cuh"

namespace bnrf {
namespace tc3 {
using namespace tcp;

constexpr int TILE_M = 128;                        // rows per CTA (256 per pair)
constexpr int NUM_THREADS = 512;
constexpr uint32_t SLOT_W_BYTES = 16384;           // weight ring slot: this CTA's [128 n x 64 k] fp16 SW128 half of one K-block of W_hi or W_lo
                                                   // (N-half groups and the view layer fill 64 rows = 8 KB of it)
constexpr uint32_t KBLOCK_BYTES = 16384;           // one [128 x 64] 16-bit SW128 tile
constexpr int NUM_STEPS = 9;
constexpr int MAX_GROUPS = 64;
constexpr uint32_t TMEM_COLS = 512;
constexpr int NSLOT = 3;                           // training mode: staging slots of one 64-column block (hi, lo, mask bits)
constexpr uint32_t SLOT_BYTES = 2 * KBLOCK_BYTES + 1024;
constexpr uint32_t CONST_FLOATS = 8 * 256 + 256 + 3 * 128;   // biases of L0..L7, w_alpha, w_rgb
constexpr uint32_t XCHG_BYTES = TILE_M * 16;

template <bool TRAIN> struct Cfg {
    static constexpr int NS = TRAIN ? 5 : 10;      // weight ring depth
    static constexpr uint32_t OFF_PE_HI = 0;
    static constexpr uint32_t OFF_PE_LO = KBLOCK_BYTES;
    static constexpr uint32_t OFF_W = 2 * KBLOCK_BYTES;
    static constexpr uint32_t OFF_ST = OFF_W + NS * SLOT_W_BYTES;
    static constexpr uint32_t OFF_CONST = OFF_ST + (TRAIN ? NSLOT * SLOT_BYTES : 0);
    static constexpr uint32_t OFF_XCHG = OFF_CONST + CONST_FLOATS * 4;
    static constexpr uint32_t OFF_FUSE = OFF_XCHG + XCHG_BYTES;            // fused compositing: [128] float4 (rgb, sigma) of the tile, then 4 warp
    static constexpr uint32_t FUSE_BYTES = 2048 + 256 + 512 + 5120;        // products + 4 x 5 partial sums (double); the tile's 128 weights; resampling
    static constexpr uint32_t OFF_BAR = OFF_FUSE + (TRAIN ? 0 : FUSE_BYTES);   // scratch of up to 4 rays (cdf, bins, sort buffer: <= 5 KB)
    static constexpr int BAR_W_FULL = 0, BAR_W_EMPTY = NS, BAR_PE_FULL = 2 * NS, BAR_PE_EMPTY = 2 * NS + 1, BAR_A_READY = 2 * NS + 2,
                         BAR_ACC_FULL = BAR_A_READY + 8, BAR_ST_FULL = BAR_ACC_FULL + 4, BAR_ST_EMPTY = BAR_ST_FULL + NSLOT,
                         BAR_RAW_FULL = BAR_ST_EMPTY + NSLOT, BAR_RAW_EMPTY = BAR_RAW_FULL + 1, BAR_COUNT = BAR_RAW_EMPTY + 1;
    static constexpr uint32_t SMEM_BYTES = OFF_BAR + 8 * BAR_COUNT + 16 + 1024;   // + tmem slot + alignment slack
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(OFF_ST % 1024 == 0 && OFF_CONST % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
};

// One tile = a list of MMA groups in issue order, built on the host (make_schedule) and passed by value.  A group is one
// 64-deep K-block of one GEMM step against either all output columns (kind FULL: N = 256, or 128 for the view layer) or one
// 128-column N-half; it consumes two weight stages (W_hi: A_hi W_hi + A_lo W_hi, then W_lo: A_hi W_lo).
//   A layer is issued as [FULL kb 0 .. 3-S] [N0: kb 4-S .. 3] [N1: kb 4-S .. 3] with S = `split` trailing K-blocks in halves: the
// epilogue of columns 0..127 then runs under the N1 MMAs.  Splitting costs a second read of the A operand (4 KB per MMA from
// TMEM, whatever N is; measured: with every K-block split the kernel is bound by the ~64 B/cycle TMEM read port, A fetches +
// the epilogue's tcgen05.ld, at ~100 cycles per 64-cycle MMA), so only the tail of a layer is split.
enum : uint8_t { G_FULL = 0, G_H0 = 1, G_H1 = 2 };
enum : uint8_t { F_FIRST = 1, F_ACC0 = 2, F_ACC1 = 4, F_PE_EMPTY = 8, F_PE_FULL = 16, F_WAIT_A = 32 };
struct GroupDesc { uint8_t t, kind; int8_t kb /* -1: encoded points */; uint8_t flags; };
struct Schedule { GroupDesc g[MAX_GROUPS]; int n; };
__host__ __device__ inline uint32_t group_stage_bytes(const GroupDesc& g) { return (g.kind == G_FULL && g.t < NUM_STEPS - 1) ? 16384u : 8192u; }

static Schedule make_schedule(int split) {
    Schedule sc{};
    auto push = [&](int t, int kind, int kb, int flags) { sc.g[sc.n++] = GroupDesc{(uint8_t)t, (uint8_t)kind, (int8_t)kb, (uint8_t)flags}; };
    for (int t = 0; t < NUM_STEPS; ++t) {
        const int first = sc.n;
        if (t == 0) {
            if (split > 0) { push(0, G_H0, -1, F_PE_FULL | F_ACC0); push(0, G_H1, -1, F_ACC1); }
            else push(0, G_FULL, -1, F_PE_FULL | F_ACC0 | F_ACC1);
        } else if (t == NUM_STEPS - 1) {
            for (int kb = 0; kb < 4; ++kb) push(t, G_FULL, kb, F_WAIT_A | (kb == 3 ? F_ACC0 | F_ACC1 : 0));   // one N-half: completes both barriers
        } else {
            if (t == 5) push(t, G_FULL, -1, F_PE_EMPTY);
            const int nfull = 4 - split;
            for (int kb = 0; kb < nfull; ++kb) push(t, G_FULL, kb, F_WAIT_A | ((split == 0 && kb == 3) ? F_ACC0 | F_ACC1 : 0));
            for (int kb = nfull; kb < 4; ++kb) push(t, G_H0, kb, F_WAIT_A | (kb == 3 ? F_ACC0 : 0));
            for (int kb = nfull; kb < 4; ++kb) push(t, G_H1, kb, kb == 3 ? F_ACC1 : 0);
        }
        // the first group that touches each accumulator half overwrites it
        bool seen[2] = {false, false};
        for (int i = first; i < sc.n; ++i) {
            GroupDesc& g = sc.g[i];
            const bool t0 = g.kind != G_H1, t1 = g.kind != G_H0;
            if ((t0 && !seen[0]) || (t1 && !seen[1])) g.flags |= F_FIRST;
            seen[0] |= t0; seen[1] |= t1;
        }
    }
    return sc;
}
static size_t schedule_stream_bytes(const Schedule& sc) {
    size_t b = 0;
    for (int i = 0; i < sc.n; ++i) b += 2 * group_stage_bytes(sc.g[i]);
    return b;
}

__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

template <int C, bool TRAIN, bool FUSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
mlp_tc3_kernel(const __grid_constant__ Schedule sched, const __grid_constant__ FuseComposite fz, TcParams p, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
               const float* __restrict__ viewbias, const float* __restrict__ z, int64_t rows, int S, int num_pairs, size_t stream_bytes,
               float* __restrict__ raw, const ActPtrs acts, unsigned int* err_flag, unsigned long long* __restrict__ trace) {
    // trace (debug, normally NULL): per-CTA stall accounting in clock64 cycles, [blockIdx.x * 16 + i] --
    //   0 kernel total   1 mma: wait PE_FULL   2 mma: wait A_READY   3 mma: wait W_FULL   4 mma: loop total
    //   5 tma: wait W_EMPTY   6 epilogue(warp 8): wait ACC_FULL   7 epilogue: loop total   8 front end: wait PE_EMPTY
    //   9 front end: loop total   10 epilogue: wait ST_EMPTY
    // and, from [148 * 16], a timeline of CTA 0's tile 3: mma [t * 8 + j], epilogue [128 + t * 8 + j] (tools/mlp_trace.py)
    using L = Cfg<TRAIN>;
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
    constexpr int NS = L::NS;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + L::OFF_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L::OFF_BAR + 8 * L::BAR_COUNT);
    float* c_bias = reinterpret_cast<float*>(sm + L::OFF_CONST);       // [8][256]
    float* c_walpha = c_bias + 8 * 256;                                // [256]
    float* c_wrgb = c_walpha + 256;                                    // [3][128]
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();                  // 0 = leader (issues the MMAs), 1 = peer
    const int cluster = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
    const uint32_t lbar0 = mapa_u32(bar0, 0);                 // the leader's barrier block, as a shared::cluster address
    auto lbar = [&](int i) { return lbar0 + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(bar(L::BAR_W_FULL + i), rank == 0 ? 2 : 1); mbar_init(bar(L::BAR_W_EMPTY + i), 1); }
        mbar_init(bar(L::BAR_PE_FULL), 8);                    // 4 front-end warps x 2 CTAs
        mbar_init(bar(L::BAR_PE_EMPTY), 1);
        for (int i = 0; i < 8; ++i) mbar_init(bar(L::BAR_A_READY + i), 8);     // 4 epilogue warps (one column half) x 2 CTAs
        for (int i = 0; i < 4; ++i) mbar_init(bar(L::BAR_ACC_FULL + i), 1);
        for (int i = 0; i < NSLOT; ++i) { mbar_init(bar(L::BAR_ST_FULL + i), 8); mbar_init(bar(L::BAR_ST_EMPTY + i), 1); }
        mbar_init(bar(L::BAR_RAW_FULL), 4); mbar_init(bar(L::BAR_RAW_EMPTY), 4);   // fused compositing: 4 epilogue warps -> 4 front-end warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    // epilogue constants: every epilogue thread reads the same words, so they live in shared memory (broadcast LDS) instead of
    // costing an L2 round trip per chunk
    for (int i = threadIdx.x; i < 8 * 256; i += NUM_THREADS) c_bias[i] = __ldg(p.bias[i >> 8] + (i & 255));
    for (int i = threadIdx.x; i < 256; i += NUM_THREADS) c_walpha[i] = __ldg(p.w_alpha + i);
    for (int i = threadIdx.x; i < 3 * 128; i += NUM_THREADS) c_wrgb[i] = __ldg(p.w_rgb + i);
    tc_fence_before();
    cluster_sync_all();                                       // barrier inits + TMEM of both CTAs visible before any remote arrive / MMA
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int my_iters = (num_pairs > cluster) ? (num_pairs - 1 - cluster) / n_clusters + 1 : 0;
    auto tile_of = [&](int it) { return (int64_t)2 * ((int64_t)cluster + (int64_t)it * n_clusters) + rank; };

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own quarter of every weight tile) =================
        if (elect_one()) {
            const unsigned char* src0 = reinterpret_cast<const unsigned char*>(p.stream) + (size_t)rank * stream_bytes;
            uint32_t cnt = 0;
            unsigned long long w_empty = 0;
            for (int it = 0; it < my_iters; ++it) {
                const unsigned char* src = src0;
                for (int g = 0; g < sched.n; ++g) {
                    const uint32_t bytes = group_stage_bytes(sched.g[g]);
                    for (int i = 0; i < 2; ++i, ++cnt, src += bytes) {
                        const uint32_t slot = cnt % NS, ph = (cnt / NS) & 1u;
                        timed_wait(bar(L::BAR_W_EMPTY + slot), ph ^ 1u, 1, w_empty);
                        mbar_expect_tx(bar(L::BAR_W_FULL + slot), bytes);
                        tma_bulk_load(base + L::OFF_W + slot * SLOT_W_BYTES, src, bytes, bar(L::BAR_W_FULL + slot));
                    }
                }
            }
            if (trace) trace[blockIdx.x * 16 + 5] = w_empty;
        }
    } else if (warp == 2) {
        // ================= peer only: forward "my part of the stage has landed" to the leader's W_FULL =================
        if (lane == 0 && rank == 1) {
            const uint32_t total = (uint32_t)my_iters * 2u * (uint32_t)sched.n;
            for (uint32_t cnt = 0; cnt < total; ++cnt) {
                const uint32_t slot = cnt % NS, ph = (cnt / NS) & 1u;
                mbar_wait(bar(L::BAR_W_FULL + slot), ph, err_flag, 8);
                mbar_arrive_cluster(lbar(L::BAR_W_FULL + slot));
            }
        }
    } else if (warp == 3) {
        // ================= training mode: staging slots -> activation tile matrices, through the bulk-copy engine =================
        if (TRAIN && elect_one()) {
            uint32_t g = 0;
            for (int it = 0; it < my_iters; ++it) {
                const int64_t tile = tile_of(it);
                for (int t = 0; t < 8; ++t) {
                    unsigned char* gt = acts.h_tiles + ((size_t)t * (size_t)acts.t_alloc + (size_t)tile) * (8 * KBLOCK_BYTES);
                    unsigned char* mk = acts.mask_bits + ((size_t)t * (size_t)acts.t_alloc + (size_t)tile) * 4096;
                    for (int kb = 0; kb < 4; ++kb, ++g) {
                        const uint32_t slot = g % NSLOT;
                        mbar_wait(bar(L::BAR_ST_FULL + slot), (g / NSLOT) & 1u, err_flag, 9);
                        const uint32_t s0 = base + L::OFF_ST + slot * SLOT_BYTES;
                        bulk_store(gt + (size_t)kb * KBLOCK_BYTES, s0, KBLOCK_BYTES);
                        bulk_store(gt + (size_t)(4 + kb) * KBLOCK_BYTES, s0 + KBLOCK_BYTES, KBLOCK_BYTES);
                        bulk_store(mk + (size_t)kb * 1024, s0 + 2 * KBLOCK_BYTES, 1024);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        if (g > 0) {                                   // the previous block has been read: its slot is free again
                            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            mbar_arrive(bar(L::BAR_ST_EMPTY + (g - 1) % NSLOT));
                        }
                    }
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA, one thread) =================
        if (rank == 0 && elect_one()) {
            uint32_t wcnt = 0;           // weight stages consumed
            const uint32_t idesc128 = make_idesc(2 * TILE_M, 128), idesc256 = make_idesc(2 * TILE_M, 256);
            unsigned long long w_pe = 0, w_a = 0, w_w = 0;
            const long long m_t0 = clock64();
            for (int it = 0; it < my_iters; ++it) {
                const bool tl = trace && blockIdx.x == 0 && it == 3;
                for (int g = 0; g < sched.n; ++g) {
                    const GroupDesc gd = sched.g[g];
                    const int t = gd.t, kb = gd.kb;
                    const uint32_t accb = ((uint32_t)it * NUM_STEPS + (uint32_t)t) & 1u;   // accumulators alternate over ALL steps (9 per tile: odd)
                    const uint32_t d_tmem = tmem + accb * 256u + (gd.kind == G_H1 ? 128u : 0u);
                    const uint32_t a_buf = tmem + (accb ^ 1u) * 256u;                      // layer t's input = the converted accumulator of layer t - 1
                    const uint32_t idesc = (gd.kind == G_FULL && t < NUM_STEPS - 1) ? idesc256 : idesc128;
                    uint32_t accumulate = (gd.flags & F_FIRST) ? 0u : 1u;
                    const uint32_t a_par = ((uint32_t)it * 8u + (uint32_t)(t - 1)) & 1u;   // generation of the A_READY barriers (steps 1..8)
                    const bool wait_a = (gd.flags & F_WAIT_A) != 0;                        // N1 groups re-read chunks the N0 group already waited for
                    if (gd.flags & F_PE_FULL) timed_wait_cluster(bar(L::BAR_PE_FULL), (uint32_t)it & 1u, 2, w_pe);
                    if (t == 1 && kb == 0 && wait_a) {
                        // first write into the buffer the PREVIOUS tile's view-layer epilogue read: chunk 0 and 1 are produced by the two
                        // column-half warp groups, so both arrivals mean every epilogue warp has left that epilogue
                        timed_wait_cluster(bar(L::BAR_A_READY + 0), a_par, 3, w_a);
                        timed_wait_cluster(bar(L::BAR_A_READY + 1), a_par, 3, w_a);
                    }
                    {   // W_hi stage of this K-block: A_hi * W_hi and A_lo * W_hi
                        const uint32_t slot = wcnt % NS, ph = (wcnt / NS) & 1u;
                        timed_wait_cluster(bar(L::BAR_W_FULL + slot), ph, 4, w_w);
                        tc_fence_after();
                        if (tl && (gd.flags & F_FIRST) && gd.kind != G_H1) trace[148 * 16 + t * 8 + 0] = (unsigned long long)clock64();
                        const uint32_t w = base + L::OFF_W + slot * SLOT_W_BYTES;
#pragma unroll
                        for (int hk = 0; hk < 2; ++hk) {
                            if (wait_a) {
                                timed_wait_cluster(bar(L::BAR_A_READY + kb * 2 + hk), a_par, 3, w_a);
                                tc_fence_after();
                                if (tl && hk == 0) trace[148 * 16 + t * 8 + 1 + kb] = (unsigned long long)clock64();
                            }
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                const uint64_t bd = make_desc(w + (uint32_t)(hk * 2 + kk) * 32u, 0);
                                if (kb < 0) {
                                    const uint32_t ko = (uint32_t)(hk * 2 + kk) * 32u;
                                    tc_mma_pair_f16(d_tmem, make_desc(base + L::OFF_PE_HI + ko, 0), bd, idesc, accumulate);
                                    tc_mma_pair_f16(d_tmem, make_desc(base + L::OFF_PE_LO + ko, 0), bd, idesc, 1);
                                } else {
                                    const uint32_t a_hi = a_buf + (uint32_t)(kb * 2 + hk) * 32u + (uint32_t)kk * 8u;
                                    tc_mma_pair_ts_f16(d_tmem, a_hi, bd, idesc, accumulate);
                                    tc_mma_pair_ts_f16(d_tmem, a_hi + 16u, bd, idesc, 1);
                                }
                                accumulate = 1;
                            }
                        }
                        tc_commit_pair(bar(L::BAR_W_EMPTY + slot));
                        ++wcnt;
                    }
                    {   // W_lo stage: A_hi * W_lo
                        const uint32_t slot = wcnt % NS, ph = (wcnt / NS) & 1u;
                        timed_wait_cluster(bar(L::BAR_W_FULL + slot), ph, 5, w_w);
                        tc_fence_after();
                        const uint32_t w = base + L::OFF_W + slot * SLOT_W_BYTES;
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) {
                            const uint64_t bd = make_desc(w + (uint32_t)k16 * 32u, 0);
                            if (kb < 0) tc_mma_pair_f16(d_tmem, make_desc(base + L::OFF_PE_HI + (uint32_t)k16 * 32u, 0), bd, idesc, 1);
                            else tc_mma_pair_ts_f16(d_tmem, a_buf + (uint32_t)(kb * 2 + (k16 >> 1)) * 32u + (uint32_t)(k16 & 1) * 8u, bd, idesc, 1);
                        }
                        tc_commit_pair(bar(L::BAR_W_EMPTY + slot));
                        ++wcnt;
                    }
                    if (gd.flags & F_PE_EMPTY) tc_commit_pair(bar(L::BAR_PE_EMPTY));       // encoded tiles no longer needed
                    // both N-half barriers of a buffer complete exactly once per step (the view layer's single pass completes both),
                    // so they share one parity, (global step >> 1) & 1
                    if (gd.flags & F_ACC0) { tc_commit_pair(bar(L::BAR_ACC_FULL + accb * 2 + 0)); if (tl) trace[148 * 16 + t * 8 + 5] = (unsigned long long)clock64(); }
                    if (gd.flags & F_ACC1) { tc_commit_pair(bar(L::BAR_ACC_FULL + accb * 2 + 1)); if (tl) trace[148 * 16 + t * 8 + 6] = (unsigned long long)clock64(); }
                }
            }
            if (trace) {
                trace[blockIdx.x * 16 + 1] = w_pe; trace[blockIdx.x * 16 + 2] = w_a; trace[blockIdx.x * 16 + 3] = w_w;
                trace[blockIdx.x * 16 + 4] = (unsigned long long)(clock64() - m_t0);
            }
        }
    } else if (warp >= 8) {
        // ================= epilogue: 8 warps; the two warps of a TMEM lane quarter take alternate 32-column chunks =================
        const int q = warp & 3;                             // TMEM lane quarter this warp may access
        const int ch = (warp - 8) >> 2;                     // chunk parity (steps 0..7) / column half (view layer)
        const int r = q * 32 + lane;                        // row in tile == TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        float* xchg = reinterpret_cast<float*>(sm + L::OFF_XCHG) + r * 4;   // partial heads of the row, handed from the ch = 1 warp to its partner
        uint32_t sg = 0;                                    // staged 64-column blocks so far (training mode)
        unsigned long long w_acc = 0, w_st = 0;
        const long long e_t0 = clock64();
        for (int it = 0; it < my_iters; ++it) {
            const bool tl = trace && blockIdx.x == 0 && it == 3 && threadIdx.x == 256;
            const int64_t tile = tile_of(it);
            const int64_t row = tile * TILE_M + r;
            float sigma_acc = 0.0f;
            for (int t = 0; t < NUM_STEPS; ++t) {
                const uint32_t gstep = (uint32_t)it * NUM_STEPS + (uint32_t)t;
                const int b = (int)(gstep & 1u);
                const uint32_t acc_par = (gstep >> 1) & 1u;         // both N-half barriers of a buffer complete once per use of the buffer
                const float inv_scale = __ldg(p.inv_scale + (t == NUM_STEPS - 1 ? 10 : t));      // slot 10 = the merged step
                if (t < 8) {
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        timed_wait(bar(L::BAR_ACC_FULL + b * 2 + h), acc_par, 6, w_acc);
                        tc_fence_after();
                        if (tl) trace[148 * 16 + 128 + t * 8 + 3 * h] = (unsigned long long)clock64();
#pragma unroll 1
                        for (int i = 0; i < 2; ++i) {
                            const int kh = 4 * h + 2 * i + ch;                      // this warp's chunk: columns [32 kh, 32 kh + 32)
                            const uint32_t caddr = lane_addr + (uint32_t)b * 256u + (uint32_t)kh * 32u;
                            uint32_t va[16], vb[16];
                            tc_ld16_issue(caddr, va);
                            tc_ld16_issue(caddr + 16, vb);
                            tc_ld16_wait(va);
                            tc_ld16_wait(vb);
                            const float4* bias4 = reinterpret_cast<const float4*>(c_bias + t * 256 + kh * 32);
                            float v[32];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 bv = bias4[j];
                                const uint32_t* src = (j < 4) ? va + 4 * j : vb + 4 * (j - 4);
                                v[4 * j] = fmaxf(fmaf(__uint_as_float(src[0]), inv_scale, bv.x), 0.0f);
                                v[4 * j + 1] = fmaxf(fmaf(__uint_as_float(src[1]), inv_scale, bv.y), 0.0f);
                                v[4 * j + 2] = fmaxf(fmaf(__uint_as_float(src[2]), inv_scale, bv.z), 0.0f);
                                v[4 * j + 3] = fmaxf(fmaf(__uint_as_float(src[3]), inv_scale, bv.w), 0.0f);
                            }
                            if (t == 7) {                                          // sigma head on h7 (model/nerf.py:101)
                                const float4* wa4 = reinterpret_cast<const float4*>(c_walpha + kh * 32);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 wa = wa4[j];
                                    sigma_acc = fmaf(v[4 * j], wa.x, sigma_acc); sigma_acc = fmaf(v[4 * j + 1], wa.y, sigma_acc);
                                    sigma_acc = fmaf(v[4 * j + 2], wa.z, sigma_acc); sigma_acc = fmaf(v[4 * j + 3], wa.w, sigma_acc);
                                }
                            }
                            // fp16 hi / lo, packed two K elements per 32-bit column, written back over the accumulator chunk:
                            // columns [0, 16) of the chunk = hi, [16, 32) = lo
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                uint32_t hw, lw;
                                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
                                const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hw));
                                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lw) : "f"(v[2 * j + 1] - back.y), "f"(v[2 * j] - back.x));
                                va[j] = hw; vb[j] = lw;
                            }
                            tc_st16(caddr, va);
                            tc_st16(caddr + 16, vb);
                            tc_st_wait();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(lbar(L::BAR_A_READY + kh));   // one arrival per warp, on the leader's barrier
                            if (TRAIN) {
                                // AFTER the hand-off (the next layer's MMAs may already run on this chunk): the same 32 activations as
                                // bf16 hi / lo + ReLU mask bits into the staging slot of their 64-column block, in the tile-matrix layout
                                // (bwd_tiles.cuh); slot kb of layer t is block number sg + kb
                                // (the empty asm keeps the compiler from starting the conversions below before the hand-off and carrying
                                // their results across it in spilled registers)
#pragma unroll
                                for (int j = 0; j < 32; ++j) { asm volatile("" : "+f"(v[j])); }
                                const int kb = kh >> 1;
                                const uint32_t blk = sg + (uint32_t)kb, slot = blk % NSLOT;
                                timed_wait(bar(L::BAR_ST_EMPTY + slot), ((blk / NSLOT) & 1u) ^ 1u, 10, w_st);
                                unsigned char* s0 = sm + L::OFF_ST + slot * SLOT_BYTES;
#pragma unroll
                                for (int c8 = 0; c8 < 4; ++c8) {
                                    uint32_t bits = 0;
#pragma unroll
                                    for (int e = 0; e < 8; ++e) bits |= (v[8 * c8 + e] > 0.0f ? 1u : 0u) << e;
                                    const uint32_t off = sw128_offset(r, (kh & 1) * 32 + c8 * 8);
                                    uint4 hi, lo;
                                    bwt::split8_bf16_pub(v + 8 * c8, hi, lo);
                                    *reinterpret_cast<uint4*>(s0 + off) = hi;
                                    *reinterpret_cast<uint4*>(s0 + KBLOCK_BYTES + off) = lo;
                                    s0[2 * KBLOCK_BYTES + (off >> 4)] = (unsigned char)bits;
                                }
                                fence_proxy_async();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(bar(L::BAR_ST_FULL + slot));
                            }
                            if (tl) trace[148 * 16 + 128 + t * 8 + 3 * h + 1 + i] = (unsigned long long)clock64();
                        }
                    }
                    sg += 4;
                } else {
                    // view layer output (128 cols, 64 per warp of the pair) -> ReLU -> rgb head; write cat([rgb, sigma]) (model/nerf.py:103-110)
                    timed_wait(bar(L::BAR_ACC_FULL + b * 2), acc_par, 6, w_acc);
                    tc_fence_after();
                    if (tl) trace[148 * 16 + 128 + t * 8] = (unsigned long long)clock64();
                    const uint32_t acc_addr = lane_addr + (uint32_t)b * 256u + (uint32_t)ch * 32u;
                    const int64_t ray = (row < rows) ? row / S : 0;
                    const float* vbp = viewbias + ray * kHalf + ch * 32;
                    float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t va[16], vb[16];
                        tc_ld16_issue(acc_addr + hh * 64, va);
                        tc_ld16_issue(acc_addr + hh * 64 + 16, vb);
                        tc_ld16_wait(va);
                        tc_ld16_wait(vb);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int col = hh * 64 + j;                         // relative to this warp's first column
                            const uint32_t* src = (j < 16) ? va + j : vb + (j - 16);
                            const float4 bv = __ldg(reinterpret_cast<const float4*>(vbp + col));
                            const float x0 = fmaxf(fmaf(__uint_as_float(src[0]), inv_scale, bv.x), 0.0f);
                            const float x1 = fmaxf(fmaf(__uint_as_float(src[1]), inv_scale, bv.y), 0.0f);
                            const float x2 = fmaxf(fmaf(__uint_as_float(src[2]), inv_scale, bv.z), 0.0f);
                            const float x3 = fmaxf(fmaf(__uint_as_float(src[3]), inv_scale, bv.w), 0.0f);
                            if (TRAIN && row < rows)
                                *reinterpret_cast<float4*>(acts.h9_f32 + row * kHalf + ch * 32 + col) = make_float4(x0, x1, x2, x3);
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                const float4 wr = *reinterpret_cast<const float4*>(c_wrgb + c * kHalf + ch * 32 + col);
                                rgb[c] = fmaf(x0, wr.x, rgb[c]); rgb[c] = fmaf(x1, wr.y, rgb[c]);
                                rgb[c] = fmaf(x2, wr.z, rgb[c]); rgb[c] = fmaf(x3, wr.w, rgb[c]);
                            }
                        }
                    }
                    tc_fence_before();
                    // combine the two column halves of each row: the upper-half warp parks its partial sums, the lower-half warp adds and writes
                    if (ch == 1) {
                        *reinterpret_cast<float4*>(xchg) = make_float4(rgb[0], rgb[1], rgb[2], sigma_acc);
                        named_bar_arrive(1 + q, 64);
                        named_bar_sync(5 + q, 64);            // partner has read: the slot may be overwritten again
                    } else {
                        named_bar_sync(1 + q, 64);
                        const float4 o = *reinterpret_cast<const float4*>(xchg);
                        named_bar_arrive(5 + q, 64);
                        const float sgm = sigma_acc + o.w + __ldg(p.b_alpha);
                        float out_c[3] = {rgb[0] + o.x + __ldg(p.b_rgb), 0.f, 0.f};
                        if (C == 3) { out_c[1] = rgb[1] + o.y + __ldg(p.b_rgb + 1); out_c[2] = rgb[2] + o.z + __ldg(p.b_rgb + 2); }
                        if (!FUSE) {
                            if (row < rows) {
                                if (C == 3) *reinterpret_cast<float4*>(raw + row * 4) = make_float4(out_c[0], out_c[1], out_c[2], sgm);
                                else *reinterpret_cast<float2*>(raw + row * 2) = make_float2(out_c[0], sgm);
                            }
                        } else {
                            // fused compositing: hand the four values to the front-end warps (idle most of a tile), which run
                            // NeRF.raw2output on them off the critical path; `raw` never reaches HBM
                            timed_wait(bar(L::BAR_RAW_EMPTY), ((uint32_t)it & 1u) ^ 1u, 13, w_acc);
                            reinterpret_cast<float4*>(sm + L::OFF_FUSE)[r] = make_float4(out_c[0], out_c[1], out_c[2], sgm);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar(L::BAR_RAW_FULL));
                        }
                    }
                }
            }
        }
        if (trace && threadIdx.x == 256) {
            trace[blockIdx.x * 16 + 6] = w_acc; trace[blockIdx.x * 16 + 7] = (unsigned long long)(clock64() - e_t0);
            trace[blockIdx.x * 16 + 10] = w_st;
        }
    } else if (warp >= 4) {
        // ================= front end: encode the next tile =================
        const int r = threadIdx.x - 128;
        unsigned long long w_pee = 0;
        const long long f_t0 = clock64();
        // fused compositing (FUSE): NeRF.raw2output (model/nerf.py:118-148) of tile `itp` from the (rgb, sigma) rows the epilogue warps
        // parked in shared memory.  The samples of a ray are S consecutive rows of the tile (S = 32, 64 or 128), one per lane of
        // S / 32 of these four warps.  Same arithmetic as composite_kernel (composite.cu): fp32 steps rounded separately, scans in
        // double.  Runs between two encodings, off the critical path of the tensor pipe.
        const int q = warp - 4;
        auto composite_tile = [&](int itp) {
            const int64_t row = tile_of(itp) * TILE_M + r;
            unsigned long long w_raw = 0;
            timed_wait(bar(L::BAR_RAW_FULL), (uint32_t)itp & 1u, 14, w_raw);
            const float4 rv = reinterpret_cast<const float4*>(sm + L::OFF_FUSE)[r];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(L::BAR_RAW_EMPTY));
            const float out_c[3] = {rv.x, rv.y, rv.z};
            const float sgm = rv.w;
            const bool valid = row < rows;                 // rows = n S: a ray is inside the batch or not at all
            const int s_idx = r & (S - 1), wpr = S >> 5;   // sample index; warps per ray
            const int64_t ray_c = valid ? row / S : 0;
            float a_ = 0.0f, zi = 0.0f;
            double fct = 1.0;
            if (valid) {
                zi = z[row];
                float dist = (s_idx + 1 < S) ? __fsub_rn(z[row + 1], zi) : 1e10f;
                const float* rd = fz.rays_d + ray_c * 3;
                const float dn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])), __fmul_rn(rd[2], rd[2])));
                dist = __fmul_rn(dist, dn);
                float nz;
                if (fz.noise) {
                    nz = fz.noise[row];
                } else {
                    uint32_t w4[4];
                    Philox::draw(fz.rng.seed, rng_offset(fz.rng), fz.rng.ray_base + (uint64_t)ray_c, (uint32_t)s_idx, fz.stream_id, w4);
                    nz = Philox::normal(w4[0], w4[1]);
                }
                const float dens = fmaxf(__fadd_rn(sgm, nz), 0.0f);      // relu(sigma_raw + noise)
                if (fz.sigma) fz.sigma[row] = dens;
                a_ = __fsub_rn(1.0f, expf(__fmul_rn(-dens, dist)));
                fct = (double)__fadd_rn(__fsub_rn(1.0f, a_), 1e-10f);
            }
            double* scratch = reinterpret_cast<double*>(sm + L::OFF_FUSE + 2048);   // [4] warp products, [4][5] warp sums
            double w_total;
            const double ex = warp_excl_scan_mul(fct, lane, w_total);
            if (lane == 0) scratch[q] = w_total;
            named_bar_sync(9, 128);
            const int q0 = q & ~(wpr - 1);                   // first warp of this ray
            double pre = 1.0;
            for (int j = q0; j < q; ++j) pre *= scratch[j];
            const float T = (float)(pre * ex);               // cumprod rounds every prefix to fp32
            const float wgt = __fmul_rn(a_, T);
            if (valid && fz.weights) fz.weights[row] = wgt;
            float* wbuf = reinterpret_cast<float*>(sm + L::OFF_FUSE + 2048 + 256);           // the tile's weights, for the resampler
            if (fz.z_f) wbuf[r] = wgt;
            double part[5];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                part[c] = (c < C && valid) ? (double)__fmul_rn(wgt, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-out_c[c])))) : 0.0;   // sigmoid
            part[3] = valid ? (double)__fmul_rn(wgt, zi) : 0.0;
            part[4] = valid ? (double)wgt : 0.0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                part[k] = warp_sum(part[k]);
                if (lane == 0) scratch[4 + q * 5 + k] = part[k];
            }
            named_bar_sync(10, 128);
            if (q == q0 && lane == 0 && valid) {
                for (int j = q0 + 1; j < q0 + wpr; ++j)
#pragma unroll
                    for (int k = 0; k < 5; ++k) part[k] += scratch[4 + j * 5 + k];
                const float depth = (float)part[3], acc = (float)part[4];
                if (fz.rgb_map)
                    for (int c = 0; c < C; ++c) fz.rgb_map[ray_c * C + c] = (float)part[c];
                if (fz.depth_map) fz.depth_map[ray_c] = depth;
                if (fz.acc_map) fz.acc_map[ray_c] = acc;
                if (fz.disp_map) {                   // 1 / max(1e-10, depth / acc); NaN when acc == 0 (Q14)
                    const float rr = __fdiv_rn(depth, acc);
                    const float m = (rr != rr) ? rr : fmaxf(1e-10f, rr);
                    fz.disp_map[ray_c] = __fdiv_rn(1.0f, m);
                }
            }
            if (fz.z_f && q == q0 && valid) {
                // sample_pdf + sort (run_nerf_helpers.py:74-115, model/nerf.py:322-326) by the first warp of the ray, on the weights in
                // shared memory: the coarse weights never reach HBM.  (wbuf is rewritten only after the next call's first barrier.)
                float* sc = reinterpret_cast<float*>(sm + L::OFF_FUSE + 2048 + 256 + 512) + (q0 / wpr) * (2 * S + fz.sort_n);
                resample_ray(z + ray_c * S, wbuf + q0 * 32, fz.u ? fz.u + ray_c * fz.K : nullptr, fz.rng, ray_c, S, fz.K, fz.sort_n,
                             sc, sc + S, sc + 2 * S, fz.z_f + ray_c * (S + fz.K), lane);
            }
        };
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
            timed_wait(bar(L::BAR_PE_EMPTY), ((uint32_t)it & 1u) ^ 1u, 7, w_pee);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                const uint32_t off = sw128_offset(r, c8 * 8);
                split_store8(enc + 8 * c8, sm + L::OFF_PE_HI + off, sm + L::OFF_PE_LO + off);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lbar(L::BAR_PE_FULL));
            if (FUSE && it > 0) composite_tile(it - 1);
            if (TRAIN) {
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    unsigned char* gt = acts.pe_tiles + (size_t)tile * (2 * KBLOCK_BYTES) + sw128_offset(r, c8 * 8);
                    split_store8_bf16_global(enc + 8 * c8, gt, gt + KBLOCK_BYTES);
                }
            }
        }
        if (FUSE && my_iters > 0) composite_tile(my_iters - 1);
        if (trace && threadIdx.x == 128) {
            trace[blockIdx.x * 16 + 8] = w_pee; trace[blockIdx.x * 16 + 9] = (unsigned long long)(clock64() - f_t0);
        }
    }

    tc_fence_before();
    cluster_sync_all();              // the leader's MMAs read the peer's shared / tensor memory: leave together
    if (trace && threadIdx.x == 0) trace[blockIdx.x * 16 + 0] = (unsigned long long)(clock64() - k_t0);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------- weight stream (two ranks)
// Stages in MMA issue order: group g -> (W_hi, W_lo).  Accumulator column c is output feature c.  The hardware concatenates the
// two CTAs' B rows along N, so rank r holds features [128 r, 128 r + 128) of a FULL group, [128 h + 64 r, + 64) of N-half h,
// and [64 r, 64 r + 64) of the 128-wide view layer.
struct PackTable3 { uint32_t off[2 * MAX_GROUPS]; };

// 2^s with max|W| * 2^s in [1024, 2048) (the fp16 pre-scale of a GEMM step's weights, cf. tc::scale_kernel)
__device__ inline float prescale_from_absmax(unsigned int bits, float* inv) {
    const float m = __uint_as_float(bits);
    int s = 0;
    if (m > 0.0f && isfinite(m)) s = 10 - ilogbf(m);
    s = max(-24, min(24, s));
    *inv = exp2f((float)-s);
    return exp2f((float)s);
}

struct PackNet3 { const float* const* wt; const unsigned int* absmax; float* scale; float* inv_scale; __half* stream; };
struct PackNets3 { PackNet3 net[2]; };
// pack_stream3_kernel for both networks (blockIdx.z), deriving each step's pre-scale from the maxima pack_all_kernel /
// merge_views_kernel gathered; block (0, 0, net) also publishes scale / inv_scale for the MLP kernel's epilogue.
__global__ void pack_stream3_pair_kernel(const __grid_constant__ Schedule sched, const __grid_constant__ PackTable3 tab, size_t rank_bytes,
                                         const __grid_constant__ PackNets3 nets) {
    const PackNet3& pn = nets.net[blockIdx.z];
    const int i = blockIdx.x, rank = blockIdx.y;
    if (i == 0 && rank == 0 && threadIdx.x < 11) {
        float inv;
        pn.scale[threadIdx.x] = prescale_from_absmax(pn.absmax[threadIdx.x], &inv);
        pn.inv_scale[threadIdx.x] = inv;
    }
    const GroupDesc gd = sched.g[i >> 1];
    const int lo = i & 1;
    const bool view = gd.t == NUM_STEPS - 1;
    const int ti = view ? 10 : gd.t;
    const float* w = pn.wt[ti];
    float inv;
    const float sc = prescale_from_absmax(pn.absmax[ti], &inv);
    const int N = view ? 128 : 256;
    const int k0 = gd.kb < 0 ? 0 : (gd.t == 5 ? kPtsChPad : 0) + 64 * gd.kb;
    const int nrows = (int)(group_stage_bytes(gd) / 128u);
    const int n0 = view ? 64 * rank : (gd.kind == G_FULL ? 128 * rank : 128 * (gd.kind - G_H0) + 64 * rank);
    unsigned char* dst = reinterpret_cast<unsigned char*>(pn.stream) + (size_t)rank * rank_bytes + tab.off[i];
    for (int e = threadIdx.x; e < nrows * 64; e += blockDim.x) {
        const int k = e / nrows, n = e % nrows;
        const float v = w[(size_t)(k0 + k) * N + n0 + n] * sc;
        const __half hi = __float2half_rn(v);
        const __half out = lo ? __float2half_rn(v - __half2float(hi)) : hi;
        *reinterpret_cast<__half*>(dst + sw128_offset(n, k)) = out;
    }
}
__global__ void pack_stream3_kernel(const __grid_constant__ Schedule sched, const __grid_constant__ PackTable3 tab, size_t rank_bytes,
                                    const float* const* __restrict__ wt, const float* __restrict__ scale, __half* __restrict__ stream) {
    const int i = blockIdx.x, rank = blockIdx.y;
    const GroupDesc gd = sched.g[i >> 1];
    const int lo = i & 1;
    const bool view = gd.t == NUM_STEPS - 1;
    const int ti = view ? 10 : gd.t;                            // table slot 10 = the merged feature + view step (common.cuh)
    const float* w = wt[ti];
    const float sc = scale[ti];
    const int N = view ? 128 : 256;
    const int k0 = gd.kb < 0 ? 0 : (gd.t == 5 ? kPtsChPad : 0) + 64 * gd.kb;   // wt[5] rows are ordered [pe64 | h256]
    const int nrows = (int)(group_stage_bytes(gd) / 128u);
    const int n0 = view ? 64 * rank : (gd.kind == G_FULL ? 128 * rank : 128 * (gd.kind - G_H0) + 64 * rank);
    unsigned char* dst = reinterpret_cast<unsigned char*>(stream) + (size_t)rank * rank_bytes + tab.off[i];
    for (int e = threadIdx.x; e < nrows * 64; e += blockDim.x) {
        const int k = e / nrows, n = e % nrows;                 // coalesced over n in the k-major source
        const float v = w[(size_t)(k0 + k) * N + n0 + n] * sc;
        const __half hi = __float2half_rn(v);
        const __half out = lo ? __float2half_rn(v - __half2float(hi)) : hi;
        *reinterpret_cast<__half*>(dst + sw128_offset(n, k)) = out;
    }
}

// trailing K-blocks of a layer issued in N-halves (see Schedule); BNRF_TC3_SPLIT overrides for experiments
static int tc3_split() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("BNRF_TC3_SPLIT");
        v = e ? atoi(e) : 1;
        if (v < 0 || v > 2) v = 1;
    }
    return v;
}
static const Schedule& tc3_schedule() {
    static Schedule sc = make_schedule(tc3_split());
    return sc;
}

}  // namespace tc3

size_t tc3_stream_halfs() { return 2 * (size_t)tc3::MAX_GROUPS * 2 * 16384 / sizeof(__half); }   // upper bound over all schedules

int pack_tc3_stream(bnrf_ctx* ctx, int net, const float* const* table_dev, const float* scale_dev, cudaStream_t st) {
    using namespace tc3;
    NetParams& np = ctx->net[net];
    const Schedule& sc = tc3_schedule();
    PackTable3 tab{};
    uint32_t off = 0;
    for (int g = 0; g < sc.n; ++g)
        for (int i = 0; i < 2; ++i) { tab.off[2 * g + i] = off; off += group_stage_bytes(sc.g[g]); }
    pack_stream3_kernel<<<dim3(2 * sc.n, 2), 256, 0, st>>>(sc, tab, schedule_stream_bytes(sc), table_dev, scale_dev, np.tc3_stream);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int pack_tc3_stream_pair(bnrf_ctx* ctx, cudaStream_t st) {
    using namespace tc3;
    const Schedule& sc = tc3_schedule();
    PackTable3 tab{};
    uint32_t off = 0;
    for (int g = 0; g < sc.n; ++g)
        for (int i = 0; i < 2; ++i) { tab.off[2 * g + i] = off; off += group_stage_bytes(sc.g[g]); }
    PackNets3 nets{};
    for (int n = 0; n < 2; ++n) {
        NetParams& np = ctx->net[n];
        nets.net[n] = PackNet3{np.wt_table, np.absmax, np.scale, np.tc_scale, np.tc3_stream};
    }
    pack_stream3_pair_kernel<<<dim3(2 * sc.n, 2, 2), 256, 0, st>>>(sc, tab, schedule_stream_bytes(sc), nets);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

template <int C, bool TRAIN, bool FUSE = false>
static int launch_one3(bnrf_ctx* ctx, const tcp::TcParams& p, int clusters, const float* o, const float* d, const float* vb,
                       const float* z, int64_t rows, int S, int pairs, float* raw, const ActPtrs& acts, cudaStream_t st,
                       const FuseComposite& fz = FuseComposite{}) {
    using namespace tc3;
    static bool configured = false;                 // per instantiation: the attribute is a property of the function
    if (!configured) {
        BNRF_CUDA(ctx, cudaFuncSetAttribute(mlp_tc3_kernel<C, TRAIN, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<TRAIN>::SMEM_BYTES));
        configured = true;
    }
    const Schedule& sc = tc3_schedule();
    mlp_tc3_kernel<C, TRAIN, FUSE><<<2 * clusters, NUM_THREADS, Cfg<TRAIN>::SMEM_BYTES, st>>>(sc, fz, p, o, d, vb, z, rows, S, pairs, schedule_stream_bytes(sc), raw,
                                                                                            acts, ctx->err_flag, ctx->trace);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

// a ray's samples must be S consecutive rows of ONE 128-row tile, a whole number of warps
bool mlp_tc3_can_fuse_composite(int S) { return S == 32 || S == 64 || S == 128; }

int launch_mlp_tc3(bnrf_ctx* ctx, int net, const float* o, const float* d, const float* vb, const float* z,
                   int64_t n, int S, float* raw, const ActPtrs* acts, cudaStream_t st, const FuseComposite* fuse) {
    using namespace tc3;
    const NetParams& np = ctx->net[net];
    TcParams p;
    p.stream = np.tc3_stream; p.inv_scale = np.tc_scale;
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
        return c3 ? launch_one3<3, true>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, *acts, st)
                  : launch_one3<1, true>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, *acts, st);
    }
    const ActPtrs none{};
    if (fuse) {
        if (!mlp_tc3_can_fuse_composite(S) || !fuse->rays_d) return fail(ctx, BNRF_ERR_ARG, "mlp: fused compositing needs S in {32, 64, 128}");
        if (fuse->z_f && (128 / S) * (2 * S + fuse->sort_n) * 4 > 5120) return fail(ctx, BNRF_ERR_ARG, "mlp: fused resampling scratch too small");
        return c3 ? launch_one3<3, false, true>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, none, st, *fuse)
                  : launch_one3<1, false, true>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, none, st, *fuse);
    }
    return c3 ? launch_one3<3, false>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, none, st)
              : launch_one3<1, false>(ctx, p, clusters, o, d, vb, z, rows, S, pairs, raw, none, st);
}

}  // namespace bnrf

// Host-only (no device needed): the MMA issue schedule of one tile as the forward kernel receives it, 4 ints per group --
// GEMM step t, kind (0 all columns, 1 / 2 the lower / upper 128-column half), K-block (-1 = encoded points), flags -- and the bytes
// of weight stream one tile consumes per CTA.  tests/test_library.py checks its invariants on the CPU.
extern "C" int bnrf_debug_tc3_schedule(int split, int32_t* groups, int max_groups, int64_t* stream_bytes_per_cta) {
    using namespace bnrf::tc3;
    if (split < 0 || split > 4 || !groups || max_groups < MAX_GROUPS) return BNRF_ERR_ARG;
    const Schedule sc = make_schedule(split);
    for (int i = 0; i < sc.n; ++i) {
        groups[4 * i] = sc.g[i].t; groups[4 * i + 1] = sc.g[i].kind; groups[4 * i + 2] = sc.g[i].kb; groups[4 * i + 3] = sc.g[i].flags;
    }
    if (stream_bytes_per_cta) *stream_bytes_per_cta = (int64_t)schedule_stream_bytes(sc);
    return sc.n;
}
