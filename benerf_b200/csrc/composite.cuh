// Device-side pieces of composite.cu shared with the forward MLP kernel's fused epilogue (mlp_tc3.cu).
#pragma once
#include "common.cuh"

namespace bnrf {

// ------------------------------------------------------------------------------------------
// resample: S coarse depths + weights -> K new depths by inverse CDF over the S-1 mid-points
// (S-2 bins, weights[1:-1]), then sort(concat).  One warp per ray; scratch (shared memory): cdf[S-1], bins[S-1],
// sort buffer of next_pow2(S+K).  zr / wr may live in global or shared memory (the forward MLP kernel calls this on the weights its
// fused compositing has just produced, mlp_tc3.cu).
__device__ inline void resample_ray(const float* zr, const float* wr, const float* u_row /*[K] or NULL*/, const bnrf_rng& rng, int64_t ray,
                                    int S, int K, int sort_n, float* cdf, float* bins, float* buf, float* __restrict__ z_f_row, int lane) {
    const int nb = S - 2;                     // number of pdf bins
    // bins = mid-points of consecutive coarse depths (model/nerf.py:321)
    for (int i = lane; i < S - 1; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(zr[i + 1], zr[i]));
    // weights + 1e-5, their sum (double), pdf, inclusive scan in double rounded per prefix
    const int per = (nb + 31) / 32;
    double local = 0.0;
    for (int k = 0; k < per; ++k) {
        const int i = lane * per + k;
        if (i < nb) local += (double)__fadd_rn(wr[i + 1], 1e-5f);
    }
    const float wsum = (float)warp_sum(local);
    double run = 0.0, tot;
    double pre_local = 0.0;
    for (int k = 0; k < per; ++k) {
        const int i = lane * per + k;
        if (i < nb) pre_local += (double)__fdiv_rn(__fadd_rn(wr[i + 1], 1e-5f), wsum);
    }
    run = warp_excl_scan_add(pre_local, lane, tot);
    if (lane == 0) cdf[0] = 0.0f;
    for (int k = 0; k < per; ++k) {
        const int i = lane * per + k;
        if (i < nb) {
            run += (double)__fdiv_rn(__fadd_rn(wr[i + 1], 1e-5f), wsum);
            cdf[i + 1] = (float)run;
        }
    }
    for (int i = lane; i < S; i += 32) buf[i] = zr[i];
    for (int i = S + K + lane; i < sort_n; i += 32) buf[i] = __int_as_float(0x7f800000);   // +inf padding
    __syncwarp();
    const int nc = S - 1;                     // cdf / bins length
    for (int j = lane; j < K; j += 32) {
        float u;
        if (u_row) {
            u = u_row[j];
        } else {
            uint32_t w[4];
            Philox::draw(rng.seed, rng_offset(rng), rng.ray_base + (uint64_t)ray, (uint32_t)j, kStreamU, w);
            u = Philox::uniform(w[0]);
        }
        // searchsorted(cdf, u, right=True): first index with cdf[idx] > u
        int lo = 0, hi = nc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
        }
        const int below = max(lo - 1, 0), above = min(lo, nc - 1);
        float denom = __fsub_rn(cdf[above], cdf[below]);
        if (denom < 1e-5f) denom = 1.0f;
        const float t = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
        buf[S + j] = __fadd_rn(bins[below], __fmul_rn(t, __fsub_rn(bins[above], bins[below])));
    }
    __syncwarp();
    // bitonic sort of buf[0, sort_n)
    for (int size = 2; size <= sort_n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < sort_n / 2; t += 32) {
                const int i = 2 * t - (t & (stride - 1));     // lower index of the pair
                const int j = i + stride;
                const bool up = ((i & size) == 0);
                const float a = buf[i], b = buf[j];
                if ((a > b) == up) { buf[i] = b; buf[j] = a; }
            }
            __syncwarp();
        }
    }
    for (int i = lane; i < S + K; i += 32) z_f_row[i] = buf[i];
}

}  // namespace bnrf
