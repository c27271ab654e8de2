// A18 (optimiser part): Adam over a list of fp32 tensors in one launch.  Replaces torch.optim.Adam as the reference
// configures it (train.py:217-226 + utils/__init__.py:33-45: Adam(eps = 1e-8, betas (0.9, 0.999), weight_decay from the
// config, amsgrad off) -- 48 tensors of the two MLPs, 1.18 M parameters) in the training step: torch's multi-tensor
// kernel takes ~80 us per network for 2.4 MB of parameters; this one is HBM-bound (16 B read + 12 B written per
// parameter) and graph-capturable: the step count and the learning rate live in device memory.
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with g += weight_decay * p first (torch's L2 form), t = the incremented step -- the arithmetic of
// torch.optim.Adam's single-tensor path, operation for operation.
#include "common.cuh"

#define ADAM_MAX_TENSORS 64
struct AdamList {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    int64_t start[ADAM_MAX_TENSORS + 1];     // prefix sums of the tensor sizes
    int n;
};

__global__ void __launch_bounds__(256)
adam_step_kernel(AdamList L, float* __restrict__ step, const float* __restrict__ lr_dev, float lr_host,
                 float beta1, float beta2, float omb1, float omb2, float eps, float weight_decay,
                 unsigned int* __restrict__ done)
{
    // every block reads the step count before any block can publish the incremented one (the last block to finish does)
    const float t = step[0] + 1.0f;
    const float lr = lr_dev ? lr_dev[0] : lr_host;
    const float bc1 = 1.0f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, t));
    const float step_size = lr / bc1;
    const int64_t total = L.start[L.n];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;                                   // tensor holding flat element i (binary search over <= 64 segments)
        for (int s = 32; s > 0; s >>= 1) if (k + s < L.n && L.start[k + s] <= i) k += s;
        const int64_t j = i - L.start[k];
        float p = L.p[k][j];
        float g = L.g[k][j];
        if (weight_decay != 0.0f) g = g + weight_decay * p;
        // omb1 / omb2 = 1 - beta rounded from double on the host, as torch passes them (1.0f - 0.999f is off by 1.3e-5)
        const float m = L.m[k][j] + (g - L.m[k][j]) * omb1;                    // lerp_(grad, 1 - beta1)
        const float v = L.v[k][j] * beta2 + omb2 * g * g;                      // mul_(beta2).addcmul_(g, g, 1 - beta2)
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p = p - step_size * (m / denom);                                       // addcdiv_(m, denom, -step_size)
        L.p[k][j] = p; L.m[k][j] = m; L.v[k][j] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) { step[0] = t; *done = 0u; }
    }
}

extern "C" int an_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                            const int64_t* sizes, int n_tensors, float* step, const float* lr_dev, float lr,
                            float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float eps,
                            float weight_decay, unsigned int* done_counter, void* stream)
{
    if (!params || !grads || !exp_avg || !exp_avg_sq || !sizes || !step || !done_counter) return AN_ERR_ARG;
    if (n_tensors <= 0 || n_tensors > ADAM_MAX_TENSORS) return AN_ERR_UNSUPPORTED;
    AdamList L;
    L.n = n_tensors;
    L.start[0] = 0;
    for (int i = 0; i < n_tensors; ++i) {
        if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || sizes[i] <= 0) return AN_ERR_ARG;
        L.p[i] = params[i]; L.g[i] = grads[i]; L.m[i] = exp_avg[i]; L.v[i] = exp_avg_sq[i];
        L.start[i + 1] = L.start[i] + sizes[i];
    }
    const int64_t total = L.start[n_tensors];
    const int threads = 256;
    const int64_t want = (total + threads - 1) / threads;
    const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
    adam_step_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(L, step, lr_dev, lr, beta1, beta2, one_minus_beta1,
                                                                     one_minus_beta2, eps, weight_decay, done_counter);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
