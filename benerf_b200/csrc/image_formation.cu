// K8/K9/K10: blur mean, event log-intensity difference, event accumulation.
//
// HBM-bound streaming kernels: every input byte is read exactly once
// with 16-byte coalesced loads, nothing is staged.  Algorithmic bytes:
//   blur   : 4*R*C*(P+1)                 (read P frames, write one)
//   events : 4*R*(C*(B+1) + B)           (read B+1 frames, write B difference maps)
//   scatter: 12*E read + 8 B atomic per event
// Replaces train.py:299-318 (python loop of P slice-adds), train.py:205-236 +
// utils/img_utils.py:13-16 + utils/math_utils.py:4-23 (6 elementwise launches per level) and
// utils/event_utils.py:247-259 (host COO build + H2D + to_dense).
#include "common.cuh"

namespace bnrf {

__device__ inline float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// out[e] = (((x_0[e] + x_1[e]) + ...) + x_{P-1}[e]) / P  over L = R*C contiguous floats per frame
template <int UNROLL>
__global__ void blur_mean_vec4(const float4* __restrict__ rgb, int P, int64_t L4, float inv_den, float4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < L4; e += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int p = 0;
        for (; p + UNROLL <= P; p += UNROLL) {
            float4 v[UNROLL];
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) v[q] = ldg_stream(rgb + (int64_t)(p + q) * L4 + e);
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) {
                acc.x = __fadd_rn(acc.x, v[q].x); acc.y = __fadd_rn(acc.y, v[q].y);
                acc.z = __fadd_rn(acc.z, v[q].z); acc.w = __fadd_rn(acc.w, v[q].w);
            }
        }
        for (; p < P; ++p) {
            const float4 v = ldg_stream(rgb + (int64_t)p * L4 + e);
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
            acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
        }
        out[e] = make_float4(__fdiv_rn(acc.x, inv_den), __fdiv_rn(acc.y, inv_den), __fdiv_rn(acc.z, inv_den), __fdiv_rn(acc.w, inv_den));
    }
}
__global__ void blur_mean_scalar(const float* __restrict__ rgb, int P, int64_t L, float den, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < L; e += stride) {
        float acc = 0.f;
        for (int p = 0; p < P; ++p) acc = __fadd_rn(acc, rgb[(int64_t)p * L + e]);
        out[e] = __fdiv_rn(acc, den);
    }
}

__device__ inline float log_brightness(float x, int mode) {
    if (mode == 0) return logf(__fadd_rn(x, 1e-9f));                       // safe_log
    const float c = __fmul_rn(x, 255.0f);                                   // lin_log, threshold 20
    const float slope = __fdiv_rn(logf(__fadd_rn(20.0f, 1e-9f)), 20.0f);
    return (c < 20.0f) ? __fmul_rn(slope, c) : logf(__fadd_rn(c, 1e-9f));
}
__device__ inline float gray3(float r, float g, float b) {                  // utils/img_utils.py:13-16
    return __fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)), __fmul_rn(b, 0.114f));
}

// C == 3, R % 4 == 0: a thread owns 4 consecutive pixels (48 contiguous bytes per frame).
__global__ void event_logdiff_rgb4(const float4* __restrict__ rgb, int B, int64_t R4, int mode, float4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < R4; g += stride) {
        float prev[4];
        float4 a = ldg_stream(rgb + 3 * g), b = ldg_stream(rgb + 3 * g + 1), c = ldg_stream(rgb + 3 * g + 2);
        for (int f = 0; f <= B; ++f) {
            float4 na, nb, nc;
            if (f < B) {                                                    // prefetch the next frame
                const float4* nx = rgb + (int64_t)(f + 1) * 3 * R4 + 3 * g;
                na = ldg_stream(nx); nb = ldg_stream(nx + 1); nc = ldg_stream(nx + 2);
            }
            float cur[4] = {log_brightness(gray3(a.x, a.y, a.z), mode), log_brightness(gray3(a.w, b.x, b.y), mode),
                            log_brightness(gray3(b.z, b.w, c.x), mode), log_brightness(gray3(c.y, c.z, c.w), mode)};
            if (f > 0)
                out[(int64_t)(f - 1) * R4 + g] = make_float4(__fsub_rn(cur[0], prev[0]), __fsub_rn(cur[1], prev[1]),
                                                             __fsub_rn(cur[2], prev[2]), __fsub_rn(cur[3], prev[3]));
#pragma unroll
            for (int q = 0; q < 4; ++q) prev[q] = cur[q];
            a = na; b = nb; c = nc;
        }
    }
}
__global__ void event_logdiff_generic(const float* __restrict__ rgb, int B, int64_t R, int C, int mode, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += stride) {
        float prev = 0.f;
        for (int f = 0; f <= B; ++f) {
            const float* px = rgb + ((int64_t)f * R + r) * C;
            const float gval = (C == 3) ? gray3(px[0], px[1], px[2]) : px[0];
            const float cur = log_brightness(gval, mode);
            if (f > 0) out[(int64_t)(f - 1) * R + r] = __fsub_rn(cur, prev);
            prev = cur;
        }
    }
}

// ---- backward of the two image-formation operators (train.py:340 through train.py:163-177, 205-318) ----
// d rgb[p][e] = g[e] / P for every pose p
__global__ void blur_mean_backward_kernel(const float* __restrict__ g, int P, int64_t L, float den, float* __restrict__ d_rgb) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < L; e += stride) {
        const float v = g[e] / den;
        for (int p = 0; p < P; ++p) d_rgb[(int64_t)p * L + e] = v;
    }
}
__device__ inline float log_brightness_grad(float x, int mode) {           // d log_brightness / d x
    if (mode == 0) return 1.0f / (x + 1e-9f);
    const float c = x * 255.0f;
    const float slope = logf(20.0f + 1e-9f) / 20.0f;
    return (c < 20.0f) ? slope * 255.0f : 255.0f / (c + 1e-9f);
}
// frame f of rgb contributes +L to difference f-1 and -L to difference f
__global__ void event_logdiff_backward_kernel(const float* __restrict__ rgb, const float* __restrict__ g, int B, int64_t R, int C,
                                              int mode, float* __restrict__ d_rgb) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += stride) {
        for (int f = 0; f <= B; ++f) {
            const float* px = rgb + ((int64_t)f * R + r) * C;
            const float gval = (C == 3) ? gray3(px[0], px[1], px[2]) : px[0];
            float go = 0.f;
            if (f > 0) go += g[(int64_t)(f - 1) * R + r];
            if (f < B) go -= g[(int64_t)f * R + r];
            const float gl = go * log_brightness_grad(gval, mode);
            float* dp = d_rgb + ((int64_t)f * R + r) * C;
            if (C == 3) { dp[0] = gl * 0.299f; dp[1] = gl * 0.587f; dp[2] = gl * 0.114f; }
            else dp[0] = gl;
        }
    }
}

__global__ void accumulate_events_kernel(const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                                         const float* __restrict__ pol, int64_t E, int H, int W, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int xi = x[e], yi = y[e];
        if (xi >= 0 && xi < W && yi >= 0 && yi < H) atomicAdd(out + (int64_t)yi * W + xi, (double)pol[e]);
    }
}

// The same for `bins` consecutive windows of one time-sorted event array: window b = events [bounds[b], bounds[b + 1]) goes to
// image b.  One launch; each thread finds the window of its event by binary search over the (shared-memory) boundaries.
__global__ void accumulate_events_binned_kernel(const int32_t* __restrict__ x, const int32_t* __restrict__ y, const float* __restrict__ pol,
                                                const int64_t* __restrict__ bounds, int bins, int H, int W, double* __restrict__ out) {
    extern __shared__ int64_t s_bounds[];
    for (int i = threadIdx.x; i <= bins; i += blockDim.x) s_bounds[i] = bounds[i];
    __syncthreads();
    const int64_t first = s_bounds[0], last = s_bounds[bins];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t image = (int64_t)H * W;
    for (int64_t e = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < last; e += stride) {
        int lo = 0, hi = bins;                         // largest b with s_bounds[b] <= e (empty windows are skipped by the search)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_bounds[mid] <= e) lo = mid; else hi = mid;
        }
        const int xi = x[e], yi = y[e];
        if (xi >= 0 && xi < W && yi >= 0 && yi < H) atomicAdd(out + (int64_t)lo * image + (int64_t)yi * W + xi, (double)pol[e]);
    }
}

static inline unsigned stream_grid(int64_t work, int threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = ceil_div(work, threads);
    const int64_t cap = (int64_t)sms * 16;          // a multiple of the SM count; grid-stride covers the rest
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace bnrf

using namespace bnrf;

extern "C" int bnrf_blur_mean(const float* rgb, int P, int64_t R, int C, float* out, void* stream) {
    if (!rgb || !out || P <= 0 || R <= 0 || C <= 0) return BNRF_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t L = R * C;
    const bool vec = (L % 4 == 0) && (((uintptr_t)rgb | (uintptr_t)out) % 16 == 0);
    if (vec) blur_mean_vec4<8><<<stream_grid(L / 4, 256), 256, 0, st>>>((const float4*)rgb, P, L / 4, (float)P, (float4*)out);
    else blur_mean_scalar<<<stream_grid(L, 256), 256, 0, st>>>(rgb, P, L, (float)P, out);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

extern "C" int bnrf_event_logdiff(const float* rgb, int B, int64_t R, int C, int log_mode, float* out, void* stream) {
    if (!rgb || !out || B <= 0 || R <= 0 || (C != 1 && C != 3) || (log_mode != 0 && log_mode != 1)) return BNRF_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (C == 3) && (R % 4 == 0) && (((uintptr_t)rgb | (uintptr_t)out) % 16 == 0);
    if (vec) event_logdiff_rgb4<<<stream_grid(R / 4, 256), 256, 0, st>>>((const float4*)rgb, B, R / 4, log_mode, (float4*)out);
    else event_logdiff_generic<<<stream_grid(R, 256), 256, 0, st>>>(rgb, B, R, C, log_mode, out);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

extern "C" int bnrf_blur_mean_backward(const float* g, int P, int64_t R, int C, float* d_rgb, void* stream) {
    if (!g || !d_rgb || P <= 0 || R <= 0 || C <= 0) return BNRF_ERR_ARG;
    blur_mean_backward_kernel<<<stream_grid(R * C, 256), 256, 0, (cudaStream_t)stream>>>(g, P, R * C, (float)P, d_rgb);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

extern "C" int bnrf_event_logdiff_backward(const float* rgb, const float* g, int B, int64_t R, int C, int log_mode, float* d_rgb,
                                           void* stream) {
    if (!rgb || !g || !d_rgb || B <= 0 || R <= 0 || (C != 1 && C != 3) || (log_mode != 0 && log_mode != 1)) return BNRF_ERR_ARG;
    event_logdiff_backward_kernel<<<stream_grid(R, 256), 256, 0, (cudaStream_t)stream>>>(rgb, g, B, R, C, log_mode, d_rgb);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

extern "C" int bnrf_accumulate_events(const int32_t* x, const int32_t* y, const float* pol, int64_t E, int H, int W,
                                      double* out, void* stream) {
    if (E == 0) return BNRF_OK;
    if (!x || !y || !pol || !out || E < 0 || H <= 0 || W <= 0) return BNRF_ERR_ARG;
    accumulate_events_kernel<<<stream_grid(E, 256), 256, 0, (cudaStream_t)stream>>>(x, y, pol, E, H, W, out);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

extern "C" int bnrf_accumulate_events_binned(const int32_t* x, const int32_t* y, const float* pol, int64_t E, const int64_t* bounds,
                                             int bins, int H, int W, double* out, void* stream) {
    if (bins == 0 || E == 0) return BNRF_OK;
    if (!x || !y || !pol || !bounds || !out || E < 0 || bins < 0 || bins > 4096 || H <= 0 || W <= 0) return BNRF_ERR_ARG;
    accumulate_events_binned_kernel<<<stream_grid(E, 256), 256, (size_t)(bins + 1) * sizeof(int64_t), (cudaStream_t)stream>>>(
        x, y, pol, bounds, bins, H, W, out);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}
