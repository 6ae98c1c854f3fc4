// K6/K7: alpha compositing and inverse-CDF resampling, one warp per ray.
//
// composite_kernel replaces NeRF.raw2output (model/nerf.py:118-148): ~25 ATen launches over
// [N,S] temporaries become one pass in which each lane owns S/32 consecutive samples and the
// transmittance is an exclusive warp prefix-product (shuffle scan).
// resample_kernel replaces sample_pdf (run_nerf_helpers.py:74-115) and the concat + sort of
// model/nerf.py:322-326: pdf/cdf scan, binary search (searchsorted right=True), lerp, and a
// shared-memory bitonic sort of the merged depths.
//
// torch's CPU cumsum/cumprod accumulate fp32 inputs in double and round each prefix to float
// (ATen ReduceOps cumsum_cpu_kernel, acc_type<float,false> = double); both scans here do the
// same, which keeps weights and cdf within an ulp of the reference whatever the scan order.
#include "common.cuh"
#include "composite.cuh"

namespace bnrf {

constexpr int kWarpsPerBlock = 4;

// MAXK = ceil(S/32) upper bound; lane owns samples [lane*K, lane*K+K) so the scan is order preserving.
template <int C>
__global__ void composite_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                 const float* __restrict__ rays_d, const float* __restrict__ noise, bnrf_rng rng,
                                 uint32_t stream_id, int64_t n_rays, int S, float* __restrict__ rgb_map,
                                 float* __restrict__ disp_map, float* __restrict__ acc_map, float* __restrict__ weights,
                                 float* __restrict__ depth_map, float* __restrict__ sigma) {
    const int lane = threadIdx.x % 32;
    const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + threadIdx.x / 32;
    if (ray >= n_rays) return;
    const int K = (S + 31) / 32;
    const float* rd = rays_d + ray * 3;
    // post-NDC direction norm (model/nerf.py:124)
    const float dn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])), __fmul_rn(rd[2], rd[2])));
    const float* zr = z + ray * S;
    const float* rr = raw + ray * S * (C + 1);
    double trans_local = 1.0;            // product of (1 - alpha + 1e-10) over this lane's samples
    double s_rgb[3] = {0, 0, 0}, s_depth = 0, s_acc = 0;
    // pass 1: per-lane product
    constexpr int MAXK = kMaxSamples / 32;
    float alpha[MAXK];
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const int s = lane * K + k;
        float a = 0.0f;
        if (s < S) {
            const float zi = zr[s];
            float dist = (s + 1 < S) ? __fsub_rn(zr[s + 1], zi) : 1e10f;
            dist = __fmul_rn(dist, dn);
            float nz;
            if (noise) {
                nz = noise[ray * S + s];
            } else {
                uint32_t w[4];
                Philox::draw(rng.seed, rng_offset(rng), rng.ray_base + (uint64_t)ray, (uint32_t)s, stream_id, w);
                nz = Philox::normal(w[0], w[1]);
            }
            const float dens = fmaxf(__fadd_rn(rr[s * (C + 1) + C], nz), 0.0f);     // relu(sigma_raw + noise)
            if (sigma) sigma[ray * S + s] = dens;
            a = __fsub_rn(1.0f, expf(__fmul_rn(-dens, dist)));
            trans_local *= (double)__fadd_rn(__fsub_rn(1.0f, a), 1e-10f);
        }
        alpha[k] = a;
    }
    double total;
    double t_run = warp_excl_scan_mul(trans_local, lane, total);   // transmittance before this lane's first sample
    // pass 2: weights and the three reductions
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const int s = lane * K + k;
        if (s < S) {
            const float a = alpha[k];
            const float T = (float)t_run;                 // cumprod rounds every prefix to fp32
            const float w = __fmul_rn(a, T);
            if (weights) weights[ray * S + s] = w;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float col = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rr[s * (C + 1) + c])));   // sigmoid
                s_rgb[c] += (double)__fmul_rn(w, col);
            }
            s_depth += (double)__fmul_rn(w, zr[s]);
            s_acc += (double)w;
            t_run *= (double)__fadd_rn(__fsub_rn(1.0f, a), 1e-10f);
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) s_rgb[c] = warp_sum(s_rgb[c]);
    s_depth = warp_sum(s_depth);
    s_acc = warp_sum(s_acc);
    if (lane == 0) {
        const float depth = (float)s_depth, acc = (float)s_acc;
        if (rgb_map)
            for (int c = 0; c < C; ++c) rgb_map[ray * C + c] = (float)s_rgb[c];
        if (depth_map) depth_map[ray] = depth;
        if (acc_map) acc_map[ray] = acc;
        if (disp_map) {                                   // 1 / max(1e-10, depth / acc); NaN when acc == 0 (Q14)
            const float r = __fdiv_rn(depth, acc);
            const float m = (r != r) ? r : fmaxf(1e-10f, r);
            disp_map[ray] = __fdiv_rn(1.0f, m);
        }
    }
}

// resample_kernel: one warp per ray around resample_ray (composite.cuh)
__global__ void resample_kernel(const float* __restrict__ z_c, const float* __restrict__ w_c,
                                const float* __restrict__ u_in, bnrf_rng rng, int64_t n_rays, int S, int K,
                                int sort_n, float* __restrict__ z_f) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    const int per_warp = 2 * S + sort_n;
    float* cdf = smem + warp * per_warp;      // [S-1]
    float* bins = cdf + S;                    // [S-1]
    float* buf = bins + S;                    // [sort_n]
    if (ray >= n_rays) return;
    resample_ray(z_c + ray * S, w_c + ray * S, u_in ? u_in + ray * K : nullptr, rng, ray, S, K, sort_n, cdf, bins, buf, z_f + ray * (S + K), lane);
}

int launch_composite(bnrf_ctx* ctx, const float* raw, const float* z, const float* d, const float* noise,
                     const bnrf_rng* rng, uint32_t stream_id, int64_t n, int S, float* rgb, float* disp, float* acc,
                     float* weights, float* depth, float* sigma, cudaStream_t st) {
    if (!raw || !z || !d || n <= 0 || S < 2 || S > kMaxSamples) return fail(ctx, BNRF_ERR_ARG, "composite: bad argument (2 <= S <= %d)", kMaxSamples);
    bnrf_rng r = rng ? *rng : bnrf_rng{};
    const unsigned grid = (unsigned)ceil_div(n, kWarpsPerBlock);
    if (ctx->cfg.channels == 3)
        composite_kernel<3><<<grid, 32 * kWarpsPerBlock, 0, st>>>(raw, z, d, noise, r, stream_id, n, S, rgb, disp, acc, weights, depth, sigma);
    else
        composite_kernel<1><<<grid, 32 * kWarpsPerBlock, 0, st>>>(raw, z, d, noise, r, stream_id, n, S, rgb, disp, acc, weights, depth, sigma);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int launch_resample(bnrf_ctx* ctx, const float* zc, const float* w, const float* u, const bnrf_rng* rng, int64_t n,
                    int S, int K, float* zf, cudaStream_t st) {
    if (!zc || !w || !zf || n <= 0 || S < 3 || K < 1 || S + K > kMaxSamples) return fail(ctx, BNRF_ERR_ARG, "resample: bad argument (S >= 3, S + K <= %d)", kMaxSamples);
    int sort_n = 1;
    while (sort_n < S + K) sort_n <<= 1;
    bnrf_rng r = rng ? *rng : bnrf_rng{};
    const size_t smem = (size_t)kWarpsPerBlock * (2 * S + sort_n) * sizeof(float);
    resample_kernel<<<(unsigned)ceil_div(n, kWarpsPerBlock), 32 * kWarpsPerBlock, smem, st>>>(zc, w, u, r, n, S, K, sort_n, zf);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bnrf
