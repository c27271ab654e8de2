// A1/A2 ray generation (+ body-space transform) and A3 stratified sampling.
// Reference semantics: datasets/anim_nerf_dataset.py:56-85 (gen_ray_directions/gen_rays),
// models/anim_nerf.py:128-137 (ray part of convert_to_body_model_space),
// models/volume_rendering.py:29-56 (sample_coarse).  HBM-bound elementwise kernels: one
// thread per ray (32 B out, written as two float4) / per sample.
#include "common.cuh"

#include "raygen.cuh"

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
        an_make_ray(c2w + b * 12, focal + 2 * b, center + 2 * b, ginv ? ginv + b * 16 : nullptr, row, col,
                    near_, far_, rays + 2 * i);
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
