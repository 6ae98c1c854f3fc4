// Fused optimiser tail of the training step: the reference runs three torch.optim.Adam objects over
// 48 + 48 + 2 tensors, five learning-rate updates and a zero_grad per iteration (train.py:343-394,
// model/optimize.py:36-55).  With parameters, gradients and both moments laid out as flat fp32 buffers
// (benerf_b200/train.py) the whole tail is ONE elementwise launch: gradient averaging over the data-parallel ranks
// (the 1/world of the all-reduce), Adam with bias correction exactly as torch.optim.Adam (no weight decay, no amsgrad),
// a per-group learning rate (the host evaluates the exponential decay), and clearing the gradient buffer for the next step.
#include "common.cuh"

namespace bnrf {

struct AdamGroups { bnrf_adam_group g[8]; int n; };

__global__ void adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                                 const __grid_constant__ AdamGroups groups, float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                                 float grad_scale, int zero_grads) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float lr = 0.f;
        bool active = false;
#pragma unroll 1
        for (int k = 0; k < groups.n; ++k)
            if (i >= groups.g[k].begin && i < groups.g[k].end) { lr = groups.g[k].lr; active = groups.g[k].active != 0; }
        if (active) {
            const float gr = g[i] * grad_scale;
            const float mi = beta1 * m[i] + (1.0f - beta1) * gr;          // exp_avg.lerp_(grad, 1 - beta1)
            const float vi = beta2 * v[i] + (1.0f - beta2) * gr * gr;     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            m[i] = mi; v[i] = vi;
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            p[i] = p[i] - (lr / bc1) * (mi / denom);
        }
        if (zero_grads) g[i] = 0.0f;
    }
}

struct AdamSched { bnrf_adam_sched_group g[8]; int n; };

__global__ void adam_step_sched_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                                       const __grid_constant__ AdamSched groups, uint64_t* __restrict__ step_dev, double decay_steps,
                                       float beta1, float beta2, float eps, float grad_scale, int zero_grads,
                                       unsigned int* __restrict__ blocks_done) {
    __shared__ float s_lr[8];
    __shared__ float s_bc1, s_bc2_sqrt;
    if (threadIdx.x < groups.n) {
        const uint64_t gs = *step_dev;                       // global_step of this iteration
        const double e = gs == 0 ? 0.0 : (double)(gs - 1) / decay_steps;
        s_lr[threadIdx.x] = (float)((double)groups.g[threadIdx.x].lr0 * pow((double)groups.g[threadIdx.x].decay_rate, e));
    } else if (threadIdx.x == 32) {
        const double step = (double)(*step_dev + 1);
        s_bc1 = (float)(1.0 - pow((double)beta1, step));
        s_bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
    }
    __syncthreads();
    const float bc1 = s_bc1, bc2_sqrt = s_bc2_sqrt;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float lr = 0.f;
        bool active = false;
#pragma unroll 1
        for (int k = 0; k < groups.n; ++k)
            if (i >= groups.g[k].begin && i < groups.g[k].end) { lr = s_lr[k]; active = groups.g[k].active != 0; }
        if (active) {
            const float gr = g[i] * grad_scale;
            const float mi = beta1 * m[i] + (1.0f - beta1) * gr;
            const float vi = beta2 * v[i] + (1.0f - beta2) * gr * gr;
            m[i] = mi; v[i] = vi;
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            p[i] = p[i] - (lr / bc1) * (mi / denom);
        }
        if (zero_grads) g[i] = 0.0f;
    }
    if (blocks_done) {
        // global_step += 1 inside this launch: every block has read the counter before the barrier above, so the last block to
        // get here may advance it (atomicInc wraps the arrival count back to zero for the next launch)
        __syncthreads();
        if (threadIdx.x == 0 && atomicInc(blocks_done, gridDim.x - 1) == gridDim.x - 1) *step_dev += 1;
    }
}

__global__ void step_advance_kernel(uint64_t* step_dev) { *step_dev += 1; }

}  // namespace bnrf

using namespace bnrf;

extern "C" int bnrf_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                              const bnrf_adam_group* groups, int n_groups, int64_t step, float beta1, float beta2, float eps,
                              float grad_scale, int zero_grads, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || !groups || n_groups <= 0 || n_groups > 8 || step < 1) return BNRF_ERR_ARG;
    AdamGroups gs{};
    gs.n = n_groups;
    for (int k = 0; k < n_groups; ++k) gs.g[k] = groups[k];
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + 255) / 256, cap = (int64_t)sms * 8;
    adam_step_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        params, grads, exp_avg, exp_avg_sq, n, gs, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale, zero_grads);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

extern "C" int bnrf_adam_step_sched(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                    const bnrf_adam_sched_group* groups, int n_groups, uint64_t* step_dev, double decay_steps,
                                    float beta1, float beta2, float eps, float grad_scale, int zero_grads, uint64_t* advance_scratch,
                                    void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || !groups || n_groups <= 0 || n_groups > 8 || !step_dev || decay_steps <= 0) return BNRF_ERR_ARG;
    AdamSched gs{};
    gs.n = n_groups;
    for (int k = 0; k < n_groups; ++k) gs.g[k] = groups[k];
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + 255) / 256, cap = (int64_t)sms * 8;
    adam_step_sched_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        params, grads, exp_avg, exp_avg_sq, n, gs, step_dev, decay_steps, beta1, beta2, eps, grad_scale, zero_grads,
        reinterpret_cast<unsigned int*>(advance_scratch));
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}

extern "C" int bnrf_step_advance(uint64_t* step_dev, void* stream) {
    if (!step_dev) return BNRF_ERR_ARG;
    step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    return cudaGetLastError() == cudaSuccess ? BNRF_OK : BNRF_ERR_CUDA;
}
