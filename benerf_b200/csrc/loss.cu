// The loss block of one training iteration as device code (a15): train.py:163-177 (event pair split, target gather),
// train.py:205-292 (event loss: thresholded for synthetic data, L2-normalised for real data), train.py:294-331 (blur mean
// over the virtual poses + rgb loss), loss/imgloss.py:3-5 (MSE), utils/img_utils.py:13-16, utils/math_utils.py:4-23 -- forward
// AND the gradients w.r.t. the four rendered tensors, so that neither the ~20 elementwise / reduction launches of the eager
// formulation nor an autograd graph over them is needed.
//
//   stage 1  bnrf_training_loss   per event pixel: gray -> log brightness -> difference d (both levels), target t gathered from
//            the accumulated event image; sums of d^2, d t, t^2, (d - thr t)^2 in float64.  Per blur element: mean over the P
//            poses, squared error, and -- being local -- the blur gradients right away.  Thresholded event loss: gradients and
//            the five loss values are finished by the last block of this launch.
//   [ranks]  the normalised event loss divides by L2 norms over the WHOLE ray batch (dim 0): with pixel-sharded ranks the caller
//            all-reduces stats[0..4] (one 40-byte exchange) between the stages
//   stage 2  bnrf_training_loss_finish (normalised loss only): u = d / (|d| + 1e-9), v = t / (|t| + 1e-9),
//            L = c/R sum (u - v)^2;  dL/dd_j = 2c/R [ a (u_j - v_j) - d_j S / (|d| (|d| + 1e-9)^2) ],  S = sum (u_i - v_i) d_i
// The event target is float64 upstream (Q10), so the event terms are float64; the blur terms are float32 means.
#include "common.cuh"

namespace bnrf {
namespace {

__device__ inline float lb_fwd(float x, int mode) {
    if (mode == 0) return logf(__fadd_rn(x, 1e-9f));                       // safe_log
    const float c = __fmul_rn(x, 255.0f);                                   // lin_log, threshold 20
    const float slope = __fdiv_rn(logf(__fadd_rn(20.0f, 1e-9f)), 20.0f);
    return (c < 20.0f) ? __fmul_rn(slope, c) : logf(__fadd_rn(c, 1e-9f));
}
__device__ inline float lb_grad(float x, int mode) {
    if (mode == 0) return 1.0f / (x + 1e-9f);
    const float c = x * 255.0f;
    const float slope = logf(20.0f + 1e-9f) / 20.0f;
    return (c < 20.0f) ? slope * 255.0f : 255.0f / (c + 1e-9f);
}
__device__ inline float gray_of(const float* px, int C) {
    return C == 3 ? __fadd_rn(__fadd_rn(__fmul_rn(px[0], 0.299f), __fmul_rn(px[1], 0.587f)), __fmul_rn(px[2], 0.114f)) : px[0];
}

// stats layout (doubles): 0 sum d^2 fine, 1 sum d t fine, 2 sum d^2 coarse, 3 sum d t coarse, 4 sum t^2,
//                         5 event sq-err fine, 6 event sq-err coarse, 7 blur sq-err fine, 8 blur sq-err coarse; [15] = block counter
constexpr int kStats = 16;

struct LossArgs {
    bnrf_loss_cfg cfg;
    const float *evt_fine, *evt_coarse;      // [2 R_e, C]
    const double* events_accu; const int64_t* idx_evt; int64_t R_e;
    const float *blur_fine, *blur_coarse, *blur_target; int64_t R_b;
    double* stats; float* diff;              // diff: [2][R_e]
    float *d_evt_fine, *d_evt_coarse, *d_blur_fine, *d_blur_coarse;
    double* loss_out;                        // [5]: total, event fine, event coarse, blur fine, blur coarse
};

__device__ inline void block_add(double v, double* dst, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) s += sh[w];
        if (s != 0.0) atomicAdd(dst, s);
    }
}

__device__ inline void write_event_grad(const LossArgs& a, int level, int64_t i, float g) {
    const int C = a.cfg.channels;
    const float* rgb = level ? a.evt_coarse : a.evt_fine;
    float* d = level ? a.d_evt_coarse : a.d_evt_fine;
    if (!d) return;
    const float* p1 = rgb + i * C;
    const float* p2 = rgb + (a.R_e + i) * C;
    const float g1 = -g * lb_grad(gray_of(p1, C), a.cfg.log_mode), g2 = g * lb_grad(gray_of(p2, C), a.cfg.log_mode);
    if (C == 3) {
        d[i * 3] = g1 * 0.299f; d[i * 3 + 1] = g1 * 0.587f; d[i * 3 + 2] = g1 * 0.114f;
        d[(a.R_e + i) * 3] = g2 * 0.299f; d[(a.R_e + i) * 3 + 1] = g2 * 0.587f; d[(a.R_e + i) * 3 + 2] = g2 * 0.114f;
    } else {
        d[i] = g1; d[a.R_e + i] = g2;
    }
}

__device__ inline void finish_losses(const LossArgs& a) {
    const bnrf_loss_cfg& c = a.cfg;
    const double ev_coeff = c.event_threshold > 0.0f ? (double)c.event_coeff_syn : (double)c.event_coeff_real;
    const double ef = c.event_loss ? a.stats[5] / (double)a.R_e * ev_coeff : 0.0;
    const double ec = c.event_loss ? a.stats[6] / (double)a.R_e * ev_coeff : 0.0;
    const double nb = (double)(a.R_b * c.channels);
    // the reference forms the blur terms in float32 (mse of two float32 tensors, then * rgb_coeff)
    const double bf = c.rgb_loss ? (double)((float)(a.stats[7] / nb) * c.rgb_coeff) : 0.0;
    const double bc = c.rgb_loss ? (double)((float)(a.stats[8] / nb) * c.rgb_coeff) : 0.0;
    a.loss_out[1] = ef; a.loss_out[2] = ec; a.loss_out[3] = bf; a.loss_out[4] = bc;
    a.loss_out[0] = (ec + ef) + (bf + bc);          // train.py:236/292 then 331: event (coarse + fine) + rgb (fine + coarse)
}

__global__ void __launch_bounds__(256) loss_stage1_kernel(const LossArgs a) {
    __shared__ double sh[8];
    __shared__ bool last;
    const bnrf_loss_cfg& c = a.cfg;
    const int C = c.channels;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const bool thresholded = c.event_threshold > 0.0f;
    const double thr = (double)c.event_threshold;        // torch.tensor(args.event_threshold) is float32, promoted (train.py:209)
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    if (c.event_loss) {
        for (int64_t i = tid; i < a.R_e; i += stride) {
            const double t = a.events_accu[a.idx_evt[i]];
            s[4] += t * t;
#pragma unroll
            for (int level = 0; level < 2; ++level) {
                const float* rgb = level ? a.evt_coarse : a.evt_fine;
                const float l1 = lb_fwd(gray_of(rgb + i * C, C), c.log_mode), l2 = lb_fwd(gray_of(rgb + (a.R_e + i) * C, C), c.log_mode);
                const float d = __fsub_rn(l2, l1);
                a.diff[level * a.R_e + i] = d;
                s[2 * level] += (double)d * (double)d;
                s[2 * level + 1] += (double)d * t;
                if (thresholded) {
                    const double e = (double)d - t * thr;
                    s[5 + level] += e * e;
                    write_event_grad(a, level, i, (float)(2.0 * (double)c.event_coeff_syn * e / (double)a.R_e));
                }
            }
        }
    } else if (a.R_e > 0) {
        // args.event_loss off: the event render does not enter the loss, its gradient is zero (not "whatever the buffer held")
        for (int64_t i = tid; i < 2 * a.R_e * C; i += stride) {
            if (a.d_evt_fine) a.d_evt_fine[i] = 0.0f;
            if (a.d_evt_coarse) a.d_evt_coarse[i] = 0.0f;
        }
    }
    double sb[2] = {0, 0};
    if (c.rgb_loss) {
        const int64_t L = a.R_b * C;
        const int P = c.n_poses;
        for (int64_t e = tid; e < L; e += stride) {
            const float tgt = a.blur_target[e];
#pragma unroll
            for (int level = 0; level < 2; ++level) {
                const float* rgb = level ? a.blur_coarse : a.blur_fine;
                float acc = 0.f;
                for (int p = 0; p < P; ++p) acc = __fadd_rn(acc, rgb[(int64_t)p * L + e]);     // running sum then one divide (train.py:307-318)
                const float err = __fsub_rn(__fdiv_rn(acc, (float)P), tgt);
                sb[level] += (double)err * (double)err;
                float* d = level ? a.d_blur_coarse : a.d_blur_fine;
                if (d) {
                    const float g = 2.0f * c.rgb_coeff * err / (float)L / (float)P;
                    for (int p = 0; p < P; ++p) d[(int64_t)p * L + e] = g;
                }
            }
        }
    } else if (a.R_b > 0) {
        for (int64_t i = tid; i < (int64_t)c.n_poses * a.R_b * C; i += stride) {
            if (a.d_blur_fine) a.d_blur_fine[i] = 0.0f;
            if (a.d_blur_coarse) a.d_blur_coarse[i] = 0.0f;
        }
    }
    for (int k = 0; k < 7; ++k) block_add(s[k], a.stats + k, sh);
    block_add(sb[0], a.stats + 7, sh);
    block_add(sb[1], a.stats + 8, sh);
    // last block: every sum is complete
    __threadfence();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(reinterpret_cast<unsigned int*>(a.stats + 15), 1u);
        last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0 && (thresholded || !c.event_loss)) {
        __threadfence();
        finish_losses(a);
    }
}

__global__ void __launch_bounds__(256) loss_stage2_kernel(const LossArgs a) {
    __shared__ double sh[8];
    __shared__ bool last;
    const bnrf_loss_cfg& c = a.cfg;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const double tn = sqrt(a.stats[4]), bt = 1.0 / (tn + 1e-9);
    double sq[2] = {0, 0};
    for (int level = 0; level < 2; ++level) {
        const double dd = a.stats[2 * level], dt = a.stats[2 * level + 1];
        const double dn = sqrt(dd), an = 1.0 / (dn + 1e-9);
        const double S = an * dd - bt * dt;
        const double k2 = dn > 0.0 ? S / (dn * (dn + 1e-9) * (dn + 1e-9)) : 0.0;
        const double scale = 2.0 * (double)c.event_coeff_real / (double)a.R_e;
        for (int64_t i = tid; i < a.R_e; i += stride) {
            const double d = (double)a.diff[level * a.R_e + i];
            const double t = a.events_accu[a.idx_evt[i]];
            const double r = an * d - bt * t;
            sq[level] += r * r;
            write_event_grad(a, level, i, (float)(scale * (an * r - d * k2)));
        }
    }
    block_add(sq[0], a.stats + 5, sh);
    block_add(sq[1], a.stats + 6, sh);
    __threadfence();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(reinterpret_cast<unsigned int*>(a.stats + 14), 1u);
        last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        finish_losses(a);
    }
}

int check(const bnrf_loss_cfg* cfg, const void* ws) {
    if (!cfg || !ws) return BNRF_ERR_ARG;
    if ((cfg->channels != 1 && cfg->channels != 3) || cfg->n_poses <= 0 || (cfg->log_mode != 0 && cfg->log_mode != 1)) return BNRF_ERR_ARG;
    return BNRF_OK;
}

LossArgs make_args(const bnrf_loss_cfg* cfg, const float* evt_fine, const float* evt_coarse, const double* events_accu,
                   const int64_t* idx_evt, int64_t R_e, const float* blur_fine, const float* blur_coarse, const float* blur_target,
                   int64_t R_b, void* workspace, float* d_evt_fine, float* d_evt_coarse, float* d_blur_fine, float* d_blur_coarse,
                   double* loss_out) {
    LossArgs a{};
    a.cfg = *cfg;
    a.evt_fine = evt_fine; a.evt_coarse = evt_coarse; a.events_accu = events_accu; a.idx_evt = idx_evt; a.R_e = R_e;
    a.blur_fine = blur_fine; a.blur_coarse = blur_coarse; a.blur_target = blur_target; a.R_b = R_b;
    a.stats = static_cast<double*>(workspace);
    a.diff = reinterpret_cast<float*>(a.stats + kStats);
    a.d_evt_fine = d_evt_fine; a.d_evt_coarse = d_evt_coarse; a.d_blur_fine = d_blur_fine; a.d_blur_coarse = d_blur_coarse;
    a.loss_out = loss_out;
    return a;
}

unsigned grid_for(int64_t work) {
    const int64_t want = ceil_div(work > 0 ? work : 1, 256);
    return (unsigned)(want < 592 ? want : 592);
}

}  // namespace
}  // namespace bnrf

using namespace bnrf;

extern "C" {

size_t bnrf_training_loss_workspace_bytes(int64_t R_e) { return kStats * sizeof(double) + (size_t)(R_e > 0 ? R_e : 0) * 2 * sizeof(float); }

int bnrf_training_loss(const bnrf_loss_cfg* cfg, const float* evt_fine, const float* evt_coarse, const double* events_accu,
                       const int64_t* idx_evt, int64_t R_e, const float* blur_fine, const float* blur_coarse,
                       const float* blur_target, int64_t R_b, void* workspace, float* d_evt_fine, float* d_evt_coarse,
                       float* d_blur_fine, float* d_blur_coarse, double* loss_out, void* stream) {
    int rc = check(cfg, workspace);
    if (rc) return rc;
    if (!loss_out || (cfg->event_loss && (!evt_fine || !evt_coarse || !events_accu || !idx_evt || R_e <= 0)) ||
        (cfg->rgb_loss && (!blur_fine || !blur_coarse || !blur_target || R_b <= 0))) return BNRF_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(workspace, 0, kStats * sizeof(double), st) != cudaSuccess) return BNRF_ERR_CUDA;
    const LossArgs a = make_args(cfg, evt_fine, evt_coarse, events_accu, idx_evt, R_e, blur_fine, blur_coarse, blur_target, R_b, workspace,
                                 d_evt_fine, d_evt_coarse, d_blur_fine, d_blur_coarse, loss_out);
    const int64_t work = (cfg->event_loss ? R_e : 0) > (cfg->rgb_loss ? R_b * cfg->channels : 0) ? R_e : R_b * cfg->channels;
    loss_stage1_kernel<<<grid_for(work), 256, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

int bnrf_training_loss_finish(const bnrf_loss_cfg* cfg, const float* evt_fine, const float* evt_coarse, const double* events_accu,
                              const int64_t* idx_evt, int64_t R_e, int64_t R_b, void* workspace, float* d_evt_fine,
                              float* d_evt_coarse, double* loss_out, void* stream) {
    int rc = check(cfg, workspace);
    if (rc) return rc;
    if (!cfg->event_loss || cfg->event_threshold > 0.0f) return BNRF_OK;           // stage 1 already finished everything
    if (!loss_out || !evt_fine || !evt_coarse || !events_accu || !idx_evt || R_e <= 0) return BNRF_ERR_ARG;
    const LossArgs a = make_args(cfg, evt_fine, evt_coarse, events_accu, idx_evt, R_e, nullptr, nullptr, nullptr, R_b, workspace,
                                 d_evt_fine, d_evt_coarse, nullptr, nullptr, loss_out);
    loss_stage2_kernel<<<grid_for(R_e), 256, 0, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

}  // extern "C"
