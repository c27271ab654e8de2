// A1/A2 ray generation (+ body-space transform) and A3 stratified sampling.
// Reference semantics: datasets/anim_nerf_dataset.py:56-85 (gen_ray_directions/gen_rays),
// models/anim_nerf.py:128-137 (ray part of convert_to_body_model_space),
// models/volume_rendering.py:29-56 (sample_coarse).  HBM-bound elementwise kernels: one
// thread per ray (32 B out, written as two float4) / per sample.
#include "common.cuh"

__global__ void raygen_kernel(const float* __restrict__ c2w, const float* __restrict__ focal,
                              const float* __restrict__ center, const int32_t* __restrict__ pix,
                              const float* __restrict__ ginv, int B, int R, int H, int W,
                              float near_, float far_, float4* __restrict__ rays)
{
    const int64_t total = (int64_t)B * R;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / R);
        const int r = (int)(i - (int64_t)b * R);
        int row, col;
        if (pix) { row = pix[2 * i]; col = pix[2 * i + 1]; }
        else     { row = r / W;      col = r - row * W; }
        const float* C = c2w + b * 12;
        const float fx = focal[2 * b], fy = focal[2 * b + 1];
        const float cx = center[2 * b], cy = center[2 * b + 1];
        float dx = ((float)col - cx) / fx, dy = -((float)row - cy) / fy, dz = -1.0f;
        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
        dx /= nrm; dy /= nrm; dz /= nrm;
        float d0 = dx * C[0] + dy * C[1] + dz * C[2];
        float d1 = dx * C[4] + dy * C[5] + dz * C[6];
        float d2 = dx * C[8] + dy * C[9] + dz * C[10];
        float o0 = C[3], o1 = C[7], o2 = C[11];
        float nr = near_, fr = far_;
        if (ginv) {
            const float* G = ginv + b * 16;
            const float p0 = G[0] * o0 + G[1] * o1 + G[2] * o2 + G[3];
            const float p1 = G[4] * o0 + G[5] * o1 + G[6] * o2 + G[7];
            const float p2 = G[8] * o0 + G[9] * o1 + G[10] * o2 + G[11];
            const float e0 = G[0] * d0 + G[1] * d1 + G[2] * d2;
            const float e1 = G[4] * d0 + G[5] * d1 + G[6] * d2;
            const float e2 = G[8] * d0 + G[9] * d1 + G[10] * d2;
            o0 = p0; o1 = p1; o2 = p2; d0 = e0; d1 = e1; d2 = e2;
            const float cam = sqrtf(o0 * o0 + o1 * o1 + o2 * o2);
            nr = fmaxf(near_, cam - 1.0f);
            fr = fminf(far_, cam + 1.0f);
        }
        rays[2 * i]     = make_float4(o0, o1, o2, d0);
        rays[2 * i + 1] = make_float4(d1, d2, nr, fr);
    }
}

__global__ void sample_coarse_kernel(const float* __restrict__ rays, int64_t n_rays, int Kc,
                                     float perturb, const float* __restrict__ noise_u,
                                     uint64_t seed, float* __restrict__ z)
{
    const int64_t total = n_rays * Kc;
    const float step = 1.0f / (float)Kc;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ray = e / Kc;
        const int i = (int)(e - ray * Kc);
        const float nr = __ldg(rays + ray * 8 + 6), fr = __ldg(rays + ray * 8 + 7);
        const float t = (float)i * step;
        float zi = nr * (1.0f - t) + fr * t;
        if (perturb > 0.0f) {
            const float tm = (float)(i - 1) * step, tp = (float)(i + 1) * step;
            const float zm = nr * (1.0f - tm) + fr * tm;
            const float zp = nr * (1.0f - tp) + fr * tp;
            const float lo = (i == 0) ? zi : 0.5f * (zi + zm);
            const float hi = (i == Kc - 1) ? zi : 0.5f * (zp + zi);
            const float u = noise_u ? __ldg(noise_u + e) : philox_u01(seed, (uint64_t)e);
            zi = lo + (hi - lo) * (perturb * u);
        }
        z[e] = zi;
    }
}

extern "C" int an_raygen_fwd(const float* c2w, const float* focal, const float* center,
                             const int32_t* pix, const float* ginv, int B, int R, int H, int W,
                             float near_, float far_, float* rays, void* stream)
{
    if (!c2w || !focal || !center || !rays || B <= 0 || R <= 0) return AN_ERR_ARG;
    if (!pix && (H <= 0 || W <= 0 || (int64_t)H * W != R)) return AN_ERR_ARG;
    if (((uintptr_t)rays) & 15) return AN_ERR_ALIGN;
    const int64_t total = (int64_t)B * R;
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads < 148 * 16 ? (total + threads - 1) / threads : 148 * 16);
    raygen_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(c2w, focal, center, pix, ginv, B, R, H, W,
                                                                  near_, far_, (float4*)rays);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int an_sample_coarse_fwd(const float* rays, int64_t n_rays, int Kc, float perturb,
                                    const float* noise_u, uint64_t seed, float* z, void* stream)
{
    if (!rays || !z || n_rays <= 0 || Kc <= 0) return AN_ERR_ARG;
    const int64_t total = n_rays * Kc;
    const int threads = 256;
    const int64_t want = (total + threads - 1) / threads;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    sample_coarse_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(rays, n_rays, Kc, perturb, noise_u, seed, z);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
