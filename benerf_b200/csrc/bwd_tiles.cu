// Backward-pass GEMMs on tile matrices (bwd_tiles.cuh): no loader warps, no conversions, no transposes.
//
//   tile_dgrad_kernel   dZ_{l-1} = (dZ_l . W_l) (+ rank-1 term) (* ReLU mask)      one launch per linear
//       A = dZ_l tiles (bf16 hi/lo, K-major), B = W_l packed once per optimiser step as bf16 hi/lo K-major blocks
//       (L2-resident), both moved by cp.async.bulk into a 2-stage ring; 3 MMAs per K=16 slice (hi*hi + lo*hi + hi*lo),
//       fp32 accumulation in TMEM (2 x 256 columns: the epilogue of tile i overlaps the MMAs of tile i+1).  The epilogue
//       applies the mask read from the saved activation tiles and writes dZ_{l-1} as a bf16 hi/lo tile matrix again.
//   tile_wgrad_kernel   dW_l += dZ_l^T . H_{l-1},  dB_l += colsum(dZ_l)              ONE launch for all linears of a network
//       the contraction runs over the ROWS of both tile matrices: the same 16 KB blocks are read as MN-major operands
//       (UMMA descriptors: leading byte offset = next 64-column block, stride byte offset = next 8 rows).  Every job
//       (linear) owns a contiguous range of CTAs sized by its HBM bytes; a CTA streams its share of the row tiles through
//       a 3-stage ring of 32-row slices and keeps the whole [M x N] fp32 partial of dW in TMEM (2 x 256 columns), so
//       each operand byte is read from HBM exactly once and dW sees one red.global per CTA and element.  Otherwise idle
//       warps reduce the bias gradient (and alpha_linear's weight gradient) from the staged slices.
//
// Replaces, for the 256-wide linears, the autograd of model/nerf.py:93-112 (train.py:340 loss.backward()).
#include "tc_ptx.cuh"
#include "bwd_tiles.cuh"

namespace bnrf {
namespace bwt {
using namespace tcp;

// kind::f16 instruction descriptor: fp32 accumulate; formats 0 = fp16, 1 = bf16; major 0 = K, 1 = MN
__host__ __device__ constexpr uint32_t make_idesc_x(int M, int N, uint32_t a_fmt, uint32_t b_fmt, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// SWIZZLE_128B MN-major descriptor: 64-element (128 B) atoms along M/N `lbo` bytes apart, 8-row groups along K `sbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void split8_bf16(const float* v, uint4& hi, uint4& lo) { split8_bf16_pub(v, hi, lo); }
__device__ __forceinline__ void split8_f16(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(v[2 * i + 1] - back.y), "f"(v[2 * i] - back.x));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// =====================================================================================================================
namespace dg {
constexpr int THREADS = 320;                 // warp 0 producer, warp 1 MMA issuer, warps 2-9 epilogue
constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = 16384, OFF_B_HI = 32768, OFF_B_LO = 65536, STAGE = 98304;
constexpr int NSTAGE = 2;
constexpr uint32_t OFF_BAR = NSTAGE * STAGE;
constexpr uint32_t SMEM = OFF_BAR + 128 + 1024;
enum { FULL = 0, EMPTY = NSTAGE, ACC_FULL = 2 * NSTAGE, ACC_EMPTY = 2 * NSTAGE + 2, NBAR = 2 * NSTAGE + 4 };
}  // namespace dg

__global__ void __launch_bounds__(dg::THREADS, 1) tile_dgrad_kernel(const __grid_constant__ DgradArgs g, unsigned int* err_flag) {
    using namespace dg;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + OFF_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * NBAR);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(FULL + i), 1); mbar_init(bar(EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(ACC_FULL + i), 1); mbar_init(bar(ACC_EMPTY + i), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int my_tiles = (g.tiles > (int)blockIdx.x) ? (g.tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nkb = g.K / 64;
    const uint32_t b_part = (uint32_t)g.N * 128u;

    if (warp == 0) {
        if (elect_one()) {
            const size_t a_tile = tile_bytes(g.K), a_part = tile_part_bytes(g.K);
            uint32_t cnt = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                const unsigned char* a = g.a_tiles + (size_t)tile * a_tile;
                for (int kb = 0; kb < nkb; ++kb, ++cnt) {
                    const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                    mbar_wait(bar(EMPTY + s), ph ^ 1u, err_flag, 31);
                    const uint32_t st = base + s * STAGE;
                    mbar_expect_tx(bar(FULL + s), 32768u + 2u * b_part);
                    tma_bulk_load(st + OFF_A_HI, a + (size_t)kb * kKbBytes, 16384u, bar(FULL + s));
                    tma_bulk_load(st + OFF_A_LO, a + a_part + (size_t)kb * kKbBytes, 16384u, bar(FULL + s));
                    const unsigned char* b = g.b_img + (size_t)kb * 2 * b_part;
                    tma_bulk_load(st + OFF_B_HI, b, b_part, bar(FULL + s));
                    tma_bulk_load(st + OFF_B_LO, b + b_part, b_part, bar(FULL + s));
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_x(128, g.N, 1, 1, 0, 0);
            uint32_t cnt = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const uint32_t acc = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
                mbar_wait(bar(ACC_EMPTY + acc), (use & 1u) ^ 1u, err_flag, 32);
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * 256u;
                uint32_t accumulate = 0;
                for (int kb = 0; kb < nkb; ++kb, ++cnt) {
                    const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                    mbar_wait(bar(FULL + s), ph, err_flag, 33);
                    tc_fence_after();
                    const uint32_t st = base + s * STAGE;
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16) {
                        const uint64_t ah = make_desc(st + OFF_A_HI + k16 * 32, 0), al = make_desc(st + OFF_A_LO + k16 * 32, 0);
                        const uint64_t bh = make_desc(st + OFF_B_HI + k16 * 32, 0), bl = make_desc(st + OFF_B_LO + k16 * 32, 0);
                        tc_mma_f16(d_tmem, ah, bh, idesc, accumulate);
                        accumulate = 1;
                        tc_mma_f16(d_tmem, al, bh, idesc, 1);
                        tc_mma_f16(d_tmem, ah, bl, idesc, 1);
                    }
                    tc_commit(bar(EMPTY + s));
                }
                tc_commit(bar(ACC_FULL + acc));
            }
        }
    } else {
        // epilogue: warp pair (q, ch) shares TMEM lane quarter q; ch takes 32 of the 64 columns of every K-block
        const int q = warp & 3, ch = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const bool to_tiles = g.epi == DG_TILE || g.epi == DG_TILE_MASKED;
        const bool masked = g.epi == DG_TILE_MASKED;
        const int nkb_out = g.N / 64;
        const size_t o_tile = tile_bytes(g.N), o_part = tile_part_bytes(g.N), m_tile = tile_bytes(256);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int64_t row = (int64_t)tile * kTileRows + r;
            const float rr = (g.r_row && row < g.rows) ? __ldg(g.r_row + row * g.r_stride) : 0.0f;
            const uint32_t acc = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
            const unsigned char* mrow = masked ? g.mask_tiles + (size_t)tile * m_tile + (size_t)r * 128 : nullptr;
            uint4 mk[4];
            if (masked) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mk[i] = __ldg(reinterpret_cast<const uint4*>(mrow + ((uint32_t)((ch * 4 + i) ^ (r & 7)) << 4)));
            }
            mbar_wait(bar(ACC_FULL + acc), use & 1u, err_flag, 34);
            tc_fence_after();
            for (int kb = 0; kb < nkb_out; ++kb) {
                const int c0 = kb * 64 + ch * 32;
                float v[32];
                tc_ld32(lane_addr + acc * 256u + (uint32_t)c0, v);
                uint4 mn[4];
                if (masked && kb + 1 < nkb_out) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        mn[i] = __ldg(reinterpret_cast<const uint4*>(mrow + (size_t)(kb + 1) * kKbBytes + ((uint32_t)((ch * 4 + i) ^ (r & 7)) << 4)));
                }
                if (g.r_row) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 rc = __ldg(reinterpret_cast<const float4*>(g.r_col + c0 + j));
                        v[j] = fmaf(rr, rc.x, v[j]); v[j + 1] = fmaf(rr, rc.y, v[j + 1]);
                        v[j + 2] = fmaf(rr, rc.z, v[j + 2]); v[j + 3] = fmaf(rr, rc.w, v[j + 3]);
                    }
                }
                if (to_tiles) {
                    if (masked) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t w[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                v[8 * i + 2 * e] = (w[e] & 0x00007fffu) ? v[8 * i + 2 * e] : 0.0f;
                                v[8 * i + 2 * e + 1] = (w[e] & 0x7fff0000u) ? v[8 * i + 2 * e + 1] : 0.0f;
                            }
                        }
                    }
                    unsigned char* orow = g.out_tiles + (size_t)tile * o_tile + (size_t)kb * kKbBytes + (size_t)r * 128;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 hi, lo;
                        split8_bf16(v + 8 * i, hi, lo);
                        const uint32_t off = (uint32_t)((ch * 4 + i) ^ (r & 7)) << 4;
                        *reinterpret_cast<uint4*>(orow + off) = hi;
                        *reinterpret_cast<uint4*>(orow + o_part + off) = lo;
                    }
                } else if (row < g.rows) {
                    float4* dst = reinterpret_cast<float4*>(g.out_f32 + row * g.ld_out + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        if (g.epi == DG_F32_ACCUM) { const float4 p = dst[j]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                        dst[j] = o;
                    }
                }
                if (masked && kb + 1 < nkb_out) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) mk[i] = mn[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(ACC_EMPTY + acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

int launch_tile_dgrad(bnrf_ctx* ctx, const DgradArgs& a, cudaStream_t st) {
    if ((a.K != 128 && a.K != 256) || (a.N != 64 && a.N != 256) || a.tiles <= 0) return fail(ctx, BNRF_ERR_ARG, "tile_dgrad: bad shape");
    if ((a.epi == DG_TILE || a.epi == DG_TILE_MASKED) != (a.N == 256)) return fail(ctx, BNRF_ERR_ARG, "tile_dgrad: epilogue / width mismatch");
    const int grid = a.tiles < ctx->sm_count ? a.tiles : ctx->sm_count;
    BNRF_CUDA(ctx, cudaFuncSetAttribute(tile_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dg::SMEM));
    tile_dgrad_kernel<<<grid, dg::THREADS, dg::SMEM, st>>>(a, ctx->err_flag);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

// =====================================================================================================================
namespace wg {
constexpr int THREADS = 320;                 // warp 0 producer, warp 1 MMA issuer, warps 2-5 colsum(dz) + flush, warps 6-9 weighted colsum(h)
constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = 16384, OFF_B_HI = 32768, OFF_B_LO = 49152, STAGE = 65536;
constexpr uint32_t SLICE = 4096;             // 32 rows of one 64-column block
constexpr int NSTAGE = 3;
constexpr uint32_t OFF_BAR = NSTAGE * STAGE;
constexpr uint32_t SMEM = OFF_BAR + 128 + 1024;
enum { FULL = 0, EMPTY = NSTAGE, ACC_FULL = 2 * NSTAGE, NBAR = 2 * NSTAGE + 1 };
}  // namespace wg

__global__ void __launch_bounds__(wg::THREADS, 1) tile_wgrad_kernel(const __grid_constant__ WgradParams p, unsigned int* err_flag) {
    using namespace wg;
    int ji = 0;
    while (ji + 1 < p.n_jobs && (int)blockIdx.x >= p.job[ji + 1].cta0) ++ji;
    const WgradJob& jb = p.job[ji];
    const int c = (int)blockIdx.x - jb.cta0;
    if (c < 0 || c >= jb.ctas) return;
    const int t0 = (int)((int64_t)p.tiles * c / jb.ctas), t1 = (int)((int64_t)p.tiles * (c + 1) / jb.ctas);
    if (t0 >= t1) return;                                    // uniform for the CTA: nothing allocated yet
    const uint32_t iters = (uint32_t)(t1 - t0) * 4u;         // 32-row slices

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + OFF_BAR;
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * NBAR);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const bool aux_b = jb.wrow != nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(FULL + i), 1); mbar_init(bar(EMPTY + i), aux_b ? 9 : 5); }
        mbar_init(bar(ACC_FULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int na = jb.M / 64, nb = jb.N / 64;

    if (warp == 0) {
        // producer: ONE elected thread issues the 4 KB slices (operand, hi/lo part, 64-column block) of a stage back to back (from
        // different lanes every cp.async.bulk sits in a ~150-cycle uniformisation loop, which paced the whole kernel)
        if (elect_one()) {
            for (uint32_t i = 0; i < iters; ++i) {
                const uint32_t s = i % NSTAGE, ph = (i / NSTAGE) & 1u;
                mbar_wait(bar(EMPTY + s), ph ^ 1u, err_flag, 41);
                mbar_expect_tx(bar(FULL + s), (uint32_t)(2 * (na + nb)) * SLICE);
                const size_t t = (size_t)t0 + (i >> 2);
                const uint32_t st = base + s * STAGE;
                const unsigned char* a = jb.dz_tiles + t * tile_bytes(jb.M) + (size_t)(i & 3u) * SLICE;
                const unsigned char* h = jb.h_tiles + t * tile_bytes(jb.N) + (size_t)(i & 3u) * SLICE;
                for (int blk = 0; blk < na; ++blk) {
                    tma_bulk_load(st + OFF_A_HI + (uint32_t)blk * SLICE, a + (size_t)blk * kKbBytes, SLICE, bar(FULL + s));
                    tma_bulk_load(st + OFF_A_LO + (uint32_t)blk * SLICE, a + tile_part_bytes(jb.M) + (size_t)blk * kKbBytes, SLICE, bar(FULL + s));
                }
                for (int blk = 0; blk < nb; ++blk) {
                    tma_bulk_load(st + OFF_B_HI + (uint32_t)blk * SLICE, h + (size_t)blk * kKbBytes, SLICE, bar(FULL + s));
                    tma_bulk_load(st + OFF_B_LO + (uint32_t)blk * SLICE, h + tile_part_bytes(jb.N) + (size_t)blk * kKbBytes, SLICE, bar(FULL + s));
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_x(128, jb.N, 1, 1, 1, 1);      // bf16 x bf16 (kind::f16 rejects mixed formats), both MN-major
            const int mhalves = jb.M / 128;
            for (uint32_t i = 0; i < iters; ++i) {
                const uint32_t s = i % NSTAGE, ph = (i / NSTAGE) & 1u;
                mbar_wait(bar(FULL + s), ph, err_flag, 42);
                tc_fence_after();
                const uint32_t st = base + s * STAGE;
#pragma unroll
                for (int k16 = 0; k16 < 2; ++k16) {
                    const uint64_t bh = make_desc_mn(st + OFF_B_HI + k16 * 2048, SLICE, 1024);
                    const uint64_t bl = make_desc_mn(st + OFF_B_LO + k16 * 2048, SLICE, 1024);
                    for (int mh = 0; mh < mhalves; ++mh) {
                        const uint64_t ah = make_desc_mn(st + OFF_A_HI + mh * 2 * SLICE + k16 * 2048, SLICE, 1024);
                        const uint64_t al = make_desc_mn(st + OFF_A_LO + mh * 2 * SLICE + k16 * 2048, SLICE, 1024);
                        const uint32_t d_tmem = tmem + (uint32_t)mh * 256u;
                        tc_mma_f16(d_tmem, ah, bh, idesc, (i | (uint32_t)k16) ? 1u : 0u);
                        tc_mma_f16(d_tmem, al, bh, idesc, 1);
                        tc_mma_f16(d_tmem, ah, bl, idesc, 1);
                    }
                }
                tc_commit(bar(EMPTY + s));
            }
            tc_commit(bar(ACC_FULL));
        }
    } else if (warp < 6) {
        // bias gradient: warp wa owns the 64-column block wa of dz, a lane two adjacent columns
        const int wa = warp - 2;
        const bool active = jb.dB != nullptr && wa < na;
        float s0 = 0.f, s1 = 0.f;
        for (uint32_t i = 0; i < iters; ++i) {
            const uint32_t s = i % NSTAGE, ph = (i / NSTAGE) & 1u;
            mbar_wait(bar(FULL + s), ph, err_flag, 43);
            if (active) {
                const unsigned char* blk = sm + s * STAGE + OFF_A_HI + wa * SLICE + (lane & 3) * 4;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const uint32_t off = (uint32_t)r * 128u + ((uint32_t)((lane >> 2) ^ (r & 7)) << 4);
                    const uint32_t h = *reinterpret_cast<const uint32_t*>(blk + off);
                    const uint32_t l = *reinterpret_cast<const uint32_t*>(blk + (OFF_A_LO - OFF_A_HI) + off);
                    s0 += __uint_as_float(h << 16) + __uint_as_float(l << 16);
                    s1 += __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(EMPTY + s));
        }
        if (active) {
            atomicAdd(jb.dB + wa * 64 + 2 * lane, s0);
            atomicAdd(jb.dB + wa * 64 + 2 * lane + 1, s1);
        }
        // flush the TMEM partial of dW: lane = output feature, columns = input features
        mbar_wait(bar(ACC_FULL), 0, err_flag, 44);
        tc_fence_after();
        const int q = warp & 3;
        for (int mh = 0; mh < jb.M / 128; ++mh) {
            const int o = mh * 128 + q * 32 + lane;
            float* drow = jb.dW + (size_t)o * jb.ldw + jb.col0;
            const bool vec = (reinterpret_cast<uintptr_t>(drow) & 15) == 0;       // per-thread: depends on o * ldw
            for (int c0 = 0; c0 < jb.N; c0 += 32) {
                float v[32];
                tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)mh * 256u + (uint32_t)c0, v);
                if (jb.col_scale) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= (c0 + j < jb.n_valid) ? __ldg(jb.col_scale + c0 + j) : 0.0f;
                }
                if (vec && c0 + 32 <= jb.n_valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < jb.n_valid) atomicAdd(drow + c0 + j, v[j]);
                }
            }
        }
        tc_fence_before();
    } else if (aux_b) {
        // alpha_linear: dWv[n] += sum_rows wrow[row] * h[row, n]; warp wb owns block wb of h
        const int wb = warp - 6;
        const bool active = wb < nb;
        float b0 = 0.f, b1 = 0.f, bs = 0.f;
        for (uint32_t i = 0; i < iters; ++i) {
            const uint32_t s = i % NSTAGE, ph = (i / NSTAGE) & 1u;
            const int64_t row = ((int64_t)t0 + (i >> 2)) * kTileRows + (int64_t)(i & 3u) * 32 + lane;
            const float wv = row < p.rows ? __ldg(jb.wrow + row * jb.wrow_stride) : 0.0f;
            mbar_wait(bar(FULL + s), ph, err_flag, 45);
            if (active) {
                const unsigned char* blk = sm + s * STAGE + OFF_B_HI + wb * SLICE + (lane & 3) * 4;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const float w = __shfl_sync(0xffffffffu, wv, r);
                    const uint32_t off = (uint32_t)r * 128u + ((uint32_t)((lane >> 2) ^ (r & 7)) << 4);
                    const uint32_t h = *reinterpret_cast<const uint32_t*>(blk + off);
                    const uint32_t l = *reinterpret_cast<const uint32_t*>(blk + (OFF_B_LO - OFF_B_HI) + off);
                    b0 = fmaf(w, __uint_as_float(h << 16) + __uint_as_float(l << 16), b0);
                    b1 = fmaf(w, __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u), b1);
                }
                bs += wv;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(EMPTY + s));
        }
        if (active) {
            atomicAdd(jb.dWv + wb * 64 + 2 * lane, b0);
            atomicAdd(jb.dWv + wb * 64 + 2 * lane + 1, b1);
            if (wb == 0 && jb.dBv) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bs += __shfl_xor_sync(0xffffffffu, bs, o);
                if (lane == 0) atomicAdd(jb.dBv, bs);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

int launch_tile_wgrad(bnrf_ctx* ctx, WgradParams& p, cudaStream_t st) {
    if (p.n_jobs <= 0 || p.n_jobs > kMaxWgradJobs || p.tiles <= 0) return fail(ctx, BNRF_ERR_ARG, "tile_wgrad: bad job list");
    // CTAs per job in proportion to the bytes a row tile of the job moves (operand widths), at least one each
    int64_t total = 0;
    for (int j = 0; j < p.n_jobs; ++j) {
        const WgradJob& jb = p.job[j];
        if ((jb.M != 128 && jb.M != 256) || (jb.N != 64 && jb.N != 256)) return fail(ctx, BNRF_ERR_ARG, "tile_wgrad: bad shape");
        total += jb.M + jb.N;
    }
    const int budget = ctx->sm_count > p.n_jobs ? ctx->sm_count : p.n_jobs;
    int used = 0;
    for (int j = 0; j < p.n_jobs; ++j) {
        int n = (int)((int64_t)budget * (p.job[j].M + p.job[j].N) / total);
        if (n < 1) n = 1;
        if (n > p.tiles) n = p.tiles;
        p.job[j].cta0 = used; p.job[j].ctas = n;
        used += n;
    }
    // hand the rounding remainder to the widest jobs
    for (int j = 0; used < budget && j < p.n_jobs; ++j) {
        if (p.job[j].M + p.job[j].N == 512 && p.job[j].ctas < p.tiles) {
            p.job[j].ctas++; used++;
            for (int k = j + 1; k < p.n_jobs; ++k) p.job[k].cta0++;
        }
    }
    BNRF_CUDA(ctx, cudaFuncSetAttribute(tile_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::SMEM));
    tile_wgrad_kernel<<<used, wg::THREADS, wg::SMEM, st>>>(p, ctx->err_flag);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

// =====================================================================================================================
// fp32 <-> tile matrices (packing of small operands, tests) and the dgrad weight images
__global__ void to_tiles_kernel(const float* __restrict__ src, int64_t rows, int W, int64_t ld, int fmt, unsigned char* __restrict__ tiles,
                                int64_t total_chunks) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total_chunks) return;
    const int cpr = W / 8;
    const int64_t row = e / cpr;
    const int c8 = (int)(e % cpr);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = row < rows ? src[row * ld + c8 * 8 + j] : 0.0f;
    uint4 hi, lo;
    if (fmt == 0) split8_f16(v, hi, lo); else split8_bf16(v, hi, lo);
    const int64_t tile = row / kTileRows;
    const int r = (int)(row % kTileRows);
    unsigned char* dst = tiles + (size_t)tile * tile_bytes(W) + (size_t)(c8 / 8) * kKbBytes + (size_t)r * 128 + ((uint32_t)((c8 & 7) ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + tile_part_bytes(W)) = lo;
}

__global__ void from_tiles_kernel(const unsigned char* __restrict__ tiles, int64_t rows, int W, int fmt, float* __restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * W) return;
    const int64_t row = e / W;
    const int c = (int)(e % W);
    const int64_t tile = row / kTileRows;
    const int r = (int)(row % kTileRows);
    const unsigned char* p = tiles + (size_t)tile * tile_bytes(W) + (size_t)(c / 64) * kKbBytes + sw128_offset(r, c % 64);
    const unsigned short h = *reinterpret_cast<const unsigned short*>(p), l = *reinterpret_cast<const unsigned short*>(p + tile_part_bytes(W));
    if (fmt == 0) dst[e] = __half2float(__ushort_as_half(h)) + __half2float(__ushort_as_half(l));
    else dst[e] = __uint_as_float((uint32_t)h << 16) + __uint_as_float((uint32_t)l << 16);
}

int launch_to_tiles(bnrf_ctx* ctx, const float* src, int64_t rows, int W, int64_t ld, int fmt, unsigned char* tiles, cudaStream_t st) {
    const int64_t chunks = tile_count(rows) * kTileRows * (W / 8);
    to_tiles_kernel<<<(unsigned)ceil_div(chunks, 256), 256, 0, st>>>(src, rows, W, ld, fmt, tiles, chunks);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}
int launch_from_tiles(bnrf_ctx* ctx, const unsigned char* tiles, int64_t rows, int W, int fmt, float* dst, cudaStream_t st) {
    from_tiles_kernel<<<(unsigned)ceil_div(rows * W, 256), 256, 0, st>>>(tiles, rows, W, fmt, dst);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

// img: per 64-wide K-block [hi: N x 128 B][lo: N x 128 B], element (n, k) <- src[n * K + k]; blockIdx.y = image
__global__ void pack_dgrad_images_kernel(const __grid_constant__ DgImageTable t) {
    const DgImageSeg& s = t.seg[blockIdx.y];
    const int cpr = s.K / 8, total = s.N * cpr;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int n = e / cpr, c8 = e % cpr;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = s.src[(size_t)n * s.K + c8 * 8 + j];
        uint4 hi, lo;
        split8_bf16(v, hi, lo);
        unsigned char* dst = s.img + (size_t)(c8 / 8) * 2 * s.N * 128 + (size_t)n * 128 + ((uint32_t)((c8 & 7) ^ (n & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + (size_t)s.N * 128) = lo;
    }
}

int pack_dgrad_images(bnrf_ctx* ctx, const DgImageTable& t, cudaStream_t st) {
    pack_dgrad_images_kernel<<<dim3(8, t.n), 256, 0, st>>>(t);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bwt
}  // namespace bnrf

using namespace bnrf;

extern "C" {

// Debug / test entry points: the two tile kernels on caller-provided fp32 matrices (converted to tile matrices here).
int bnrf_debug_tile_dgrad(bnrf_ctx* ctx, int64_t rows, int K, int N, const float* A, const float* B, const float* mask,
                          const float* r_row, const float* r_col, int accumulate, float* out, void* stream) {
    if (!ctx || !A || !B || !out || rows <= 0) return fail(ctx, BNRF_ERR_ARG, "debug_tile_dgrad: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t tiles = bwt::tile_count(rows);
    unsigned char *a_t = nullptr, *m_t = nullptr, *o_t = nullptr, *img = nullptr;
    BNRF_CUDA(ctx, cudaMallocAsync(&a_t, tiles * bwt::tile_bytes(K), st));
    BNRF_CUDA(ctx, cudaMallocAsync(&img, bwt::dgrad_image_bytes(N, K), st));
    int rc = bwt::launch_to_tiles(ctx, A, rows, K, K, 1, a_t, st);
    if (!rc) { bwt::DgImageTable t{}; t.n = 1; t.seg[0] = bwt::DgImageSeg{B, img, N, K}; rc = bwt::pack_dgrad_images(ctx, t, st); }
    if (!rc && mask) {
        BNRF_CUDA(ctx, cudaMallocAsync(&m_t, tiles * bwt::tile_bytes(256), st));
        rc = bwt::launch_to_tiles(ctx, mask, rows, 256, 256, 1, m_t, st);
    }
    bwt::DgradArgs a{};
    a.a_tiles = a_t; a.K = K; a.b_img = img; a.N = N; a.mask_tiles = m_t; a.r_row = r_row; a.r_stride = 1; a.r_col = r_col;
    a.rows = rows; a.tiles = (int)tiles;
    if (N == 256) {
        BNRF_CUDA(ctx, cudaMallocAsync(&o_t, tiles * bwt::tile_bytes(256), st));
        a.epi = mask ? bwt::DG_TILE_MASKED : bwt::DG_TILE; a.out_tiles = o_t;
        if (!rc) rc = bwt::launch_tile_dgrad(ctx, a, st);
        if (!rc) rc = bwt::launch_from_tiles(ctx, o_t, rows, 256, 1, out, st);
    } else {
        a.epi = accumulate ? bwt::DG_F32_ACCUM : bwt::DG_F32_STORE; a.out_f32 = out; a.ld_out = N;
        if (!rc) rc = bwt::launch_tile_dgrad(ctx, a, st);
    }
    cudaFreeAsync(a_t, st); cudaFreeAsync(img, st);
    if (m_t) cudaFreeAsync(m_t, st);
    if (o_t) cudaFreeAsync(o_t, st);
    return rc;
}

int bnrf_debug_tile_wgrad(bnrf_ctx* ctx, int64_t rows, int M, int N, const float* dz, const float* h, const float* wrow,
                          float* dW, int ldw, int col0, int n_valid, float* dB, float* dWv, float* dBv, void* stream) {
    if (!ctx || !dz || !h || !dW || rows <= 0) return fail(ctx, BNRF_ERR_ARG, "debug_tile_wgrad: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t tiles = bwt::tile_count(rows);
    unsigned char *z_t = nullptr, *h_t = nullptr;
    BNRF_CUDA(ctx, cudaMallocAsync(&z_t, tiles * bwt::tile_bytes(M), st));
    BNRF_CUDA(ctx, cudaMallocAsync(&h_t, tiles * bwt::tile_bytes(N), st));
    int rc = bwt::launch_to_tiles(ctx, dz, rows, M, M, 1, z_t, st);
    if (!rc) rc = bwt::launch_to_tiles(ctx, h, rows, N, N, 1, h_t, st);
    bwt::WgradParams p{};
    p.n_jobs = 1; p.tiles = (int)tiles; p.rows = rows;
    bwt::WgradJob& j = p.job[0];
    j.dz_tiles = z_t; j.M = M; j.h_tiles = h_t; j.N = N; j.dW = dW; j.ldw = ldw; j.col0 = col0; j.n_valid = n_valid; j.dB = dB;
    j.wrow = wrow; j.wrow_stride = 1; j.dWv = dWv; j.dBv = dBv;
    if (!rc) rc = bwt::launch_tile_wgrad(ctx, p, st);
    cudaFreeAsync(z_t, st); cudaFreeAsync(h_t, st);
    return rc;
}

}  // extern "C"
