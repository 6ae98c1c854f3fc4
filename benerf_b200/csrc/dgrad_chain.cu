// Backward pass, activation gradients of a whole network in ONE launch: the dgrad chain of NeRF.forward
// (autograd of model/nerf.py:93-112) with the gradient of a 128-sample tile resident on the SM from dZ9 down to the
// encoded points -- the mirror image of the forward kernel (mlp_tc.cu):
//
//   s0  dZ7 = relu'(h7) * (dZ9 . (W_views[:, :256] . W_feature) + d sigma (x) w_alpha)   (K = 128; feature_linear has no
//                                                               activation, so it is merged into the view layer: common.cuh wt9m)
//   s1..s7  dZ_{l-1} = relu'(h_{l-1}) * (dZ_l . W_l)              l = 7..1; at l = 5 only the h4 block of cat([pe, h4]) ...
//   s3' d pe  = dZ5 . W_5[:, :63]                                 ... and the encoded-points block goes out as fp32
//   s8  d pe += dZ0 . W_0
//
// Same machinery as the forward kernel: one persistent CTA per SM, A operand (the current dZ, bf16 hi/lo, 3 MMAs per
// K=16 slice) in shared memory in the SWIZZLE_128B K-major layout and rewritten in place by the epilogue warps, fp32
// accumulators ping-ponging between two 256-column TMEM buffers, transposed weights streaming through a 6-stage ring of
// [N x 32] SWIZZLE_64B tiles filled by cp.async.bulk from an L2-resident image, layer hand-off per 32-column K-half.
// Differences: the first A operand (dZ9, written by heads_fused_kernel as a tile matrix) arrives by cp.async.bulk; the
// epilogue applies the ReLU mask from 1 bit per activation emitted by the training-mode forward pass (32 B per row and
// layer instead of re-reading activations); and every dZ_l, which in shared memory already IS a tile of its tile matrix,
// is copied to global memory by the bulk-copy engine (one thread, cp.async.bulk.global.shared) for the weight-gradient
// kernel (bwd_tiles.cu) -- the epilogue warps issue no global stores, whose completion their release-arrive would
// await.  HBM traffic per sample: 8 x 1 KB of dZ out + 0.8 KB in (dZ9 tile, mask bits), against 9 x 2.5 KB for the
// per-linear kernels it replaces (tile_dgrad_kernel, kept as a cross-check: gemm_mode BNRF_GEMM_TC_PER_LINEAR).
#include "tc_ptx.cuh"
#include "bwd_tiles.cuh"

namespace bnrf {
namespace dgc {
using namespace tcp;

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 512;
constexpr int NS = 6;                              // weight ring depth
constexpr uint32_t STAGE_BYTES = 16384;            // one [256 x 32] bf16 SW64 tile
constexpr uint32_t KBLOCK_BYTES = 16384;           // one [128 x 64] bf16 SW128 A block
constexpr uint32_t OFF_A_HI = 0;
constexpr uint32_t OFF_A_LO = 4 * KBLOCK_BYTES;
constexpr uint32_t OFF_W = 8 * KBLOCK_BYTES;
constexpr uint32_t OFF_BAR = OFF_W + NS * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr uint32_t TMEM_COLS = 512;

enum { BAR_W_FULL = 0, BAR_W_EMPTY = BAR_W_FULL + NS, BAR_A0_FULL = BAR_W_EMPTY + NS, BAR_A_FREE, BAR_SPILLED, BAR_A_READY = BAR_SPILLED + 4,
       BAR_ACC_FULL = BAR_A_READY + 8, BAR_COUNT = BAR_ACC_FULL + 2 };

__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

// ---- weight stream: stage i = (pass, K-block, K-half, hi/lo) in MMA consumption order ----
// pass p: 0 = s0, 1..3 = s1..s3, 4 = s3' (encoded-points block of layer 5), 5..8 = s4..s7, 9 = s8
struct PassInfo { int step /*slot of the weight table: forward GEMM step, 10 = merged feature + view step*/, k0, N, K; };
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
constexpr int NUM_PASSES = 10;
constexpr int NUM_STEPS = 9;
__host__ __device__ inline int pass_stages(int p) { return p == 0 ? 8 : 16; }            // K/64 blocks x 2 halves x (hi, lo)
__host__ __device__ inline uint32_t pass_stage_bytes(int p) { return (p == 4 || p == 9) ? STAGE_BYTES / 4 : STAGE_BYTES; }
__host__ __device__ inline size_t pass_offset_bytes(int p) {
    size_t off = 0;
    for (int q = 0; q < p; ++q) off += (size_t)pass_stages(q) * pass_stage_bytes(q);
    return off;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
dgrad_chain_kernel(const unsigned char* __restrict__ stream, const unsigned char* __restrict__ dz9_tiles,
                   const unsigned char* __restrict__ mask_bits, int64_t t_alloc, const float* __restrict__ d_sigma, int64_t d_sigma_stride,
                   const float* __restrict__ w_alpha, int64_t rows, int num_tiles, int64_t dz_tile_count,
                   unsigned char* __restrict__ dz_tiles, float* __restrict__ d_pe, unsigned int* err_flag) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + OFF_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * BAR_COUNT);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(bar(BAR_W_FULL + i), 1); mbar_init(bar(BAR_W_EMPTY + i), 1); }
        mbar_init(bar(BAR_A0_FULL), 1);
        mbar_init(bar(BAR_A_FREE), 1);
        for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_SPILLED + i), 1);
        for (int i = 0; i < 8; ++i) mbar_init(bar(BAR_A_READY + i), 256);
        for (int i = 0; i < 2; ++i) mbar_init(bar(BAR_ACC_FULL + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int my_tiles = (num_tiles > (int)blockIdx.x) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const unsigned char* src = stream;
                for (int p = 0; p < NUM_PASSES; ++p) {
                    const uint32_t bytes = pass_stage_bytes(p);
                    const int n = pass_stages(p);
                    for (int i = 0; i < n; ++i, ++cnt, src += bytes) {
                        const uint32_t slot = cnt % NS, ph = (cnt / NS) & 1u;
                        mbar_wait(bar(BAR_W_EMPTY + slot), ph ^ 1u, err_flag, 51);
                        mbar_expect_tx(bar(BAR_W_FULL + slot), bytes);
                        tma_bulk_load(base + OFF_W + slot * STAGE_BYTES, src, bytes, bar(BAR_W_FULL + slot));
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ================= dZ9 loader: the tile's first A operand (K = 128: two K-blocks, hi and lo parts) =================
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                mbar_wait(bar(BAR_A_FREE), ((uint32_t)it & 1u) ^ 1u, err_flag, 52);       // last MMAs of the previous tile have read A
                if (it > 0) mbar_wait(bar(BAR_SPILLED + 3), ((uint32_t)(it - 1) * 8u + 7u) & 1u, err_flag, 58);   // ... and its dZ0 has been copied out
                const unsigned char* src = dz9_tiles + (size_t)tile * bwt::tile_bytes(kHalf);
                mbar_expect_tx(bar(BAR_A0_FULL), 4u * KBLOCK_BYTES);
                tma_bulk_load(base + OFF_A_HI, src, 2u * KBLOCK_BYTES, bar(BAR_A0_FULL));
                tma_bulk_load(base + OFF_A_LO, src + bwt::tile_part_bytes(kHalf), 2u * KBLOCK_BYTES, bar(BAR_A0_FULL));
            }
        }
    } else if (warp == 3) {
        // ================= spill: every dZ_l the epilogue builds in shared memory IS a tile of its tile matrix; the bulk-copy engine
        // writes it to global memory for the weight-gradient kernel, so the epilogue warps issue no global stores =================
        if (lane == 0) {
            const size_t dz_mat = (size_t)dz_tile_count * bwt::tile_bytes(kWidth);
            uint32_t sgen = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                for (int s = 0; s < NUM_STEPS - 1; ++s, ++sgen) {
                    unsigned char* gt = dz_tiles + (size_t)(7 - s) * dz_mat + (size_t)tile * bwt::tile_bytes(kWidth);
                    for (int kb = 0; kb < 4; ++kb) {
                        mbar_wait(bar(BAR_A_READY + 2 * kb), sgen & 1u, err_flag, 59);
                        mbar_wait(bar(BAR_A_READY + 2 * kb + 1), sgen & 1u, err_flag, 60);
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
        // ================= MMA issuer =================
        if (elect_one()) {
            uint32_t wcnt = 0, agen = 0;
            auto kblock = [&](uint32_t d_tmem, uint32_t idesc, int kb, int hk, uint32_t& accumulate) {
                const uint32_t a_hi = base + OFF_A_HI + kb * KBLOCK_BYTES, a_lo = base + OFF_A_LO + kb * KBLOCK_BYTES;
                {   // hi weight stage: A_hi * W_hi and A_lo * W_hi
                    const uint32_t slot = wcnt % NS, ph = (wcnt / NS) & 1u;
                    mbar_wait(bar(BAR_W_FULL + slot), ph, err_flag, 53);
                    tc_fence_after();
                    const uint32_t w = base + OFF_W + slot * STAGE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint64_t bd = make_desc_sw64(w + kk * 32);
                        tc_mma_f16(d_tmem, make_desc(a_hi + (hk * 2 + kk) * 32, 0), bd, idesc, accumulate);
                        accumulate = 1;
                        tc_mma_f16(d_tmem, make_desc(a_lo + (hk * 2 + kk) * 32, 0), bd, idesc, 1);
                    }
                    tc_commit(bar(BAR_W_EMPTY + slot));
                    ++wcnt;
                }
                {   // lo weight stage: A_hi * W_lo
                    const uint32_t slot = wcnt % NS, ph = (wcnt / NS) & 1u;
                    mbar_wait(bar(BAR_W_FULL + slot), ph, err_flag, 54);
                    tc_fence_after();
                    const uint32_t w = base + OFF_W + slot * STAGE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
                        tc_mma_f16(d_tmem, make_desc(a_hi + (hk * 2 + kk) * 32, 0), make_desc_sw64(w + kk * 32), idesc, 1);
                    tc_commit(bar(BAR_W_EMPTY + slot));
                    ++wcnt;
                }
            };
            const uint32_t idesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            const uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            for (int it = 0; it < my_tiles; ++it) {
                for (int s = 0; s < NUM_STEPS; ++s) {
                    const uint32_t accb = ((uint32_t)it * NUM_STEPS + (uint32_t)s) & 1u;      // accumulators alternate over ALL steps (9 per tile: odd)
                    const uint32_t d_tmem = tmem + accb * 256u;
                    const uint32_t idesc = (s == NUM_STEPS - 1) ? idesc64 : idesc256;
                    const int n_kb = (s == 0) ? 2 : 4;
                    uint32_t accumulate = 0;
                    if (s == 0) { mbar_wait(bar(BAR_A0_FULL), (uint32_t)it & 1u, err_flag, 55); tc_fence_after(); }
                    for (int kb = 0; kb < n_kb; ++kb)
                        for (int hk = 0; hk < 2; ++hk) {
                            if (s > 0) { mbar_wait(bar(BAR_A_READY + kb * 2 + hk), agen & 1u, err_flag, 56); tc_fence_after(); }
                            kblock(d_tmem, idesc, kb, hk, accumulate);
                        }
                    if (s == 3) {
                        // encoded-points block of layer 5 from the same dZ5, into the first 64 columns of the OTHER buffer: its last
                        // reader (the epilogue of s2) finished before A_READY[7] above, and the epilogue of s3 drains it first
                        uint32_t acc2 = 0;
                        for (int kb = 0; kb < 4; ++kb)
                            for (int hk = 0; hk < 2; ++hk) kblock(tmem + (accb ^ 1u) * 256u, idesc64, kb, hk, acc2);
                    }
                    tc_commit(bar(BAR_ACC_FULL + accb));
                    if (s == NUM_STEPS - 1) tc_commit(bar(BAR_A_FREE));
                    if (s >= 1) ++agen;
                }
            }
        }
    } else if (warp >= 8) {
        // ================= epilogue: 8 warps, warp pair (w, w+4) shares TMEM lane quarter q and splits the columns =================
        const int q = warp & 3;
        const int ch = (warp - 8) >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t acc_uses[2] = {0, 0};
        for (int it = 0; it < my_tiles; ++it) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t row = tile * TILE_M + r;
            const bool live = row < rows;
            float* pe_row = d_pe + row * kPtsChPad + ch * 32;
            for (int s = 0; s < NUM_STEPS; ++s) {
                const int b = (int)(((uint32_t)it * NUM_STEPS + (uint32_t)s) & 1u);
                // fetched BEFORE blocking on the accumulator (no L1 behind the smem carve-out: these are L2 round trips)
                // ReLU mask of this row: byte p of word kb = the 8 elements of physical 16-byte chunk p of K-block kb (mlp_tc2.cu)
                unsigned long long mrow[4] = {~0ull, ~0ull, ~0ull, ~0ull};
                if (s < NUM_STEPS - 1) {
                    const unsigned char* mk = mask_bits + ((size_t)(7 - s) * (size_t)t_alloc + (size_t)tile) * 4096 + (size_t)r * 8;
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) mrow[kb] = __ldg(reinterpret_cast<const unsigned long long*>(mk + kb * 1024));
                }
                const float rr = (s == 0 && live) ? __ldg(d_sigma + row * d_sigma_stride) : 0.0f;
                float4 wq[4];
                if (s == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) wq[j] = __ldg(reinterpret_cast<const float4*>(w_alpha + ch * 16) + j);
                }
                mbar_wait(bar(BAR_ACC_FULL + b), acc_uses[b] & 1u, err_flag, 57);
                ++acc_uses[b];
                tc_fence_after();
                if (s == 3) {           // d pe from layer 5, parked in the other buffer
                    float v[32];
                    tc_ld32(lane_addr + (uint32_t)(b ^ 1) * 256u + (uint32_t)ch * 32u, v);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(pe_row)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
                if (s < NUM_STEPS - 1) {
                    const uint32_t acc_addr = lane_addr + (uint32_t)b * 256u + (uint32_t)ch * 16u;
                    uint32_t va[16], vb[16];
                    tc_ld16_issue(acc_addr, va);
#pragma unroll
                    for (int kh = 0; kh < 8; ++kh) {
                        uint32_t (&cur)[16] = (kh & 1) ? vb : va;
                        uint32_t (&nxt)[16] = (kh & 1) ? va : vb;
                        tc_ld16_wait(cur);
                        // the 64-column block this chunk overwrites (dZ of the previous step) has been copied out
                        if (s >= 1 && (kh & 1) == 0) mbar_wait(bar(BAR_SPILLED + (kh >> 1)), ((uint32_t)it * 8u + (uint32_t)(s - 1)) & 1u, err_flag, 61);
                        float4 wn[4];
                        if (kh < 7) {
                            tc_ld16_issue(acc_addr + (kh + 1) * 32, nxt);
                            if (s == 0) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) wn[j] = __ldg(reinterpret_cast<const float4*>(w_alpha + ch * 16 + (kh + 1) * 32) + j);
                            }
                        }
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(cur[j]);
                        if (s == 0) {   // + d sigma (x) w_alpha (alpha_linear reads h7, model/nerf.py:101)
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 wv = wq[j >> 2];
                                v[j] = fmaf(rr, wv.x, v[j]); v[j + 1] = fmaf(rr, wv.y, v[j + 1]);
                                v[j + 2] = fmaf(rr, wv.z, v[j + 2]); v[j + 3] = fmaf(rr, wv.w, v[j + 3]);
                            }
                        }
                        {
                            const int lc = (kh & 1) * 4 + ch * 2;                                  // logical 16-byte chunk of v[0..7]; v[8..15] is lc + 1
                            const uint32_t b0 = (uint32_t)(mrow[kh >> 1] >> (8 * (lc ^ (r & 7)))) & 0xffu;
                            const uint32_t b1 = (uint32_t)(mrow[kh >> 1] >> (8 * ((lc + 1) ^ (r & 7)))) & 0xffu;
                            const uint32_t bits = b0 | (b1 << 8);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : 0.0f;
                        }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t off = (kh >> 1) * KBLOCK_BYTES + sw128_offset(r, (kh & 1) * 32 + ch * 16 + j * 8);
                            uint4 hi, lo;
                            bwt::split8_bf16_pub(v + 8 * j, hi, lo);
                            *reinterpret_cast<uint4*>(sm + OFF_A_HI + off) = hi;
                            *reinterpret_cast<uint4*>(sm + OFF_A_LO + off) = lo;
                        }
                        tc_fence_before();
                        fence_proxy_async();
                        mbar_arrive(bar(BAR_A_READY + kh));
                        if (s == 0 && kh < 7) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) wq[j] = wn[j];
                        }
                    }
                } else {
                    float v[32];
                    tc_ld32(lane_addr + (uint32_t)b * 256u + (uint32_t)ch * 32u, v);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 o = reinterpret_cast<float4*>(pe_row)[j];
                            o.x += v[4 * j]; o.y += v[4 * j + 1]; o.z += v[4 * j + 2]; o.w += v[4 * j + 3];
                            reinterpret_cast<float4*>(pe_row)[j] = o;
                        }
                    }
                    tc_fence_before();
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

// stage image: element (n, k) of the [N x 32] SW64 tile of (pass, K-block kb, K-half hk) <- wt[step][(k0 + n) * K + kb*64 + hk*32 + k]
__global__ void pack_chain_stream_kernel(const float* const* __restrict__ wt, unsigned char* __restrict__ stream) {
    const int p = blockIdx.y;
    const PassInfo pi = pass_info(p);
    const int i = blockIdx.x;                           // stage within the pass
    if (i >= pass_stages(p)) return;
    const int kbh = i >> 1, lo = i & 1;                 // (kb, hk) pair index, hi/lo
    const float* w = wt[pi.step] + (size_t)pi.k0 * pi.K;
    unsigned char* dst = stream + pass_offset_bytes(p) + (size_t)i * pass_stage_bytes(p);
    for (int e = threadIdx.x; e < pi.N * 32; e += blockDim.x) {
        const int n = e >> 5, k = e & 31;
        const float v = w[(size_t)n * pi.K + kbh * 32 + k];
        uint32_t hb;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hb) : "f"(0.0f), "f"(v));
        const float hi = __uint_as_float(hb << 16);
        uint32_t lb;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lb) : "f"(0.0f), "f"(v - hi));
        *reinterpret_cast<unsigned short*>(dst + sw64_offset(n, k)) = (unsigned short)((lo ? lb : hb) & 0xffffu);
    }
}

}  // namespace dgc

size_t dgrad_chain_stream_bytes() { return dgc::pass_offset_bytes(dgc::NUM_PASSES); }

int pack_dgrad_chain_stream(bnrf_ctx* ctx, int net, cudaStream_t st) {
    NetParams& np = ctx->net[net];
    dgc::pack_chain_stream_kernel<<<dim3(16, dgc::NUM_PASSES), 256, 0, st>>>(np.wt_table, np.dgc_stream);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int launch_dgrad_chain(bnrf_ctx* ctx, int net, const unsigned char* dz9_tiles, const unsigned char* mask_bits, int64_t t_alloc,
                       const float* d_sigma, int64_t d_sigma_stride, int64_t rows, int64_t dz_tile_count, unsigned char* dz_tiles,
                       float* d_pe, cudaStream_t st) {
    using namespace dgc;
    const NetParams& np = ctx->net[net];
    const int64_t tiles64 = bwt::tile_count(rows);
    if (tiles64 > 0x3fffffff) return fail(ctx, BNRF_ERR_ARG, "dgrad chain: too many rows");
    const int tiles = (int)tiles64;
    const int grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    BNRF_CUDA(ctx, cudaFuncSetAttribute(dgrad_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    dgrad_chain_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(np.dgc_stream, dz9_tiles, mask_bits, t_alloc, d_sigma, d_sigma_stride,
                                                              np.w_alpha, rows, tiles, dz_tile_count, dz_tiles, d_pe, ctx->err_flag);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bnrf
