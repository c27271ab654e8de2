// A18 (render losses): mse(rgb) coarse + fine and lambda * l1(alpha) coarse + fine of train.py:228-262 with their
// gradients, one launch -- in torch this is ~25 microsecond-sized launches (mse, mean, abs, sign, mul, add, fill ...)
// forward and backward inside a step that is otherwise a handful of kernels.
//   loss = mean((rgb_c - t)^2) + mean((rgb_f - t)^2) + lam * (mean|a_c - ta| + mean|a_f - ta|)
// Up to 32 CTAs, one thread per ray; every CTA stores its four partial sums, the CTA that arrives last adds them in CTA
// order -> the value does not depend on scheduling.  The gradients d loss / d input are written by the same pass (the
// backward of the autograd node only scales them by the incoming gradient).
#include "common.cuh"

namespace {

constexpr int LOSS_MAX_CTAS = 32;
struct LossWs { unsigned int arrived; unsigned int pad[3]; float part[LOSS_MAX_CTAS][4]; };     // caller's scratch, zero before the first use

__global__ void __launch_bounds__(1024)
render_loss_kernel(const float* __restrict__ rgb_c, const float* __restrict__ rgb_f, const float* __restrict__ acc_c,
                   const float* __restrict__ acc_f, const float* __restrict__ tgt_rgb, const float* __restrict__ tgt_acc,
                   int64_t n_rays, float lam, float* __restrict__ terms, LossWs* __restrict__ ws,
                   float* __restrict__ g_rgb_c, float* __restrict__ g_rgb_f, float* __restrict__ g_acc_c, float* __restrict__ g_acc_f)
{
    __shared__ float part[4][32];
    __shared__ bool last;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float k_rgb = 2.0f / (float)(3 * n_rays), k_acc = lam / (float)n_rays;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t e = r * 3 + c;
            const float t = tgt_rgb[e];
            const float dc = rgb_c[e] - t;
            s[0] += dc * dc; g_rgb_c[e] = k_rgb * dc;
            if (rgb_f) { const float df = rgb_f[e] - t; s[1] += df * df; g_rgb_f[e] = k_rgb * df; }
        }
        const float t = tgt_acc[r];
        const float dc = acc_c[r] - t;
        s[2] += fabsf(dc); g_acc_c[r] = dc > 0.f ? k_acc : (dc < 0.f ? -k_acc : 0.f);
        if (acc_f) { const float df = acc_f[r] - t; s[3] += fabsf(df); g_acc_f[r] = df > 0.f ? k_acc : (df < 0.f ? -k_acc : 0.f); }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float w = warp_sum(s[k]);
        if (lane == 0) part[k][warp] = w;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float w = warp_sum(lane < (int)(blockDim.x >> 5) ? part[k][lane] : 0.f);
            if (lane == 0) ws->part[blockIdx.x][k] = w;
        }
        if (lane == 0) {
            __threadfence();
            last = atomicAdd(&ws->arrived, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (unsigned b = 0; b < gridDim.x; ++b)
            for (int k = 0; k < 4; ++k) t[k] += ((volatile float(*)[4])ws->part)[b][k];
        const float m0 = t[0] / (float)(3 * n_rays), m1 = t[1] / (float)(3 * n_rays), m2 = t[2] / (float)n_rays, m3 = t[3] / (float)n_rays;
        terms[0] = m0; terms[1] = m1; terms[2] = m2; terms[3] = m3;
        terms[4] = m0 + m1 + lam * (m2 + m3);
        ws->arrived = 0;                          // ready for the next launch on this scratch (stream-ordered)
    }
}

}  // namespace

extern "C" int64_t an_render_loss_ws_bytes(void) { return (int64_t)sizeof(LossWs); }

extern "C" int an_render_loss(const float* rgb_coarse, const float* rgb_fine, const float* acc_coarse, const float* acc_fine,
                              const float* tgt_rgb, const float* tgt_acc, int64_t n_rays, float lambda_alphas, float* terms, void* ws,
                              float* g_rgb_coarse, float* g_rgb_fine, float* g_acc_coarse, float* g_acc_fine, void* stream)
{
    if (!rgb_coarse || !acc_coarse || !tgt_rgb || !tgt_acc || !terms || !ws || !g_rgb_coarse || !g_acc_coarse || n_rays <= 0) return AN_ERR_ARG;
    if ((rgb_fine && !g_rgb_fine) || (acc_fine && !g_acc_fine) || (!rgb_fine) != (!acc_fine)) return AN_ERR_ARG;
    int64_t ctas = (n_rays + 1023) / 1024;
    if (ctas > LOSS_MAX_CTAS) ctas = LOSS_MAX_CTAS;
    render_loss_kernel<<<(unsigned)ctas, 1024, 0, (cudaStream_t)stream>>>(rgb_coarse, rgb_fine, acc_coarse, acc_fine, tgt_rgb, tgt_acc, n_rays,
                                                              lambda_alphas, terms, (LossWs*)ws, g_rgb_coarse, g_rgb_fine, g_acc_coarse, g_acc_fine);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
