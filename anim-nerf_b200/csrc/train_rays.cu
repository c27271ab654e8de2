// training-ray sampling on the device.  Replaces, per training step, the reference's CPU
// dataloader work of datasets/anim_nerf_dataset.py:235-262 (__getitem__): full-frame gen_rays (262 144 rays built,
// 1024 kept), get_pixelcoords' draws (:10-54, 'foreground_pixel': `fore_rate` of the pixels from the eroded
// silhouette, the rest from the band between the 64-px and the fore_erode-px dilations, drawn with replacement,
// foreground first), and the gathers of rgb / alpha at those pixels (:259-261) with the white-background
// composite (:244-245).  The frames stay resident in HBM as uint8; the per-frame candidate lists (the two
// morphological masks, which depend only on the frame) are built once at load time (host class
// DeviceFrameStore).  One thread per sampled pixel: draw -> pixel -> colour/alpha gather -> ray (the same
// an_make_ray as an_raygen_fwd, incl. the body-space transform).  HBM-bound and tiny: 64 B written per ray.
#include "common.cuh"
#include "raygen.cuh"

__global__ void sample_training_rays_kernel(
    const uint8_t* __restrict__ images, const uint8_t* __restrict__ masks,
    const int32_t* __restrict__ fg_list, const int32_t* __restrict__ fg_off,
    const int32_t* __restrict__ bg_list, const int32_t* __restrict__ bg_off,
    const int32_t* __restrict__ frame_ids, const float* __restrict__ c2w, const float* __restrict__ focal,
    const float* __restrict__ center, const float* __restrict__ ginv, int B, int n, int n_fg, int H, int W,
    float near_, float far_, int white_bkgd, int with_background, const int32_t* __restrict__ sel, uint64_t seed,
    float4* __restrict__ rays, float* __restrict__ rgbs, float* __restrict__ alphas, int32_t* __restrict__ pix)
{
    const int64_t total = (int64_t)B * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / n);
        const int j = (int)(i - (int64_t)b * n);
        const int f = frame_ids[b];
        const bool fore = j < n_fg;
        const int32_t* list = fore ? fg_list : bg_list;
        const int32_t lo = fore ? fg_off[f] : bg_off[f];
        const int32_t cnt = (fore ? fg_off[f + 1] : bg_off[f + 1]) - lo;
        int32_t k;
        if (sel) k = sel[i];                                   // parity mode: the reference's np.random.choice draws
        else {
            k = (int32_t)(philox_u01(seed, (uint64_t)i) * (float)cnt);
        }
        k = k < 0 ? 0 : (k >= cnt ? cnt - 1 : k);
        const int32_t p = cnt > 0 ? list[lo + k] : 0;          // empty lists are rejected on the host
        const int row = p / W, col = p - row * W;
        const int64_t px = ((int64_t)f * H + row) * W + col;
        const float m = (float)masks[px] / 255.0f;             // mask / 255.  (:201)
        float c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = (float)images[px * 3 + a] / 255.0f;      // img / 255.   (:200)
            // separate roundings, as torch evaluates img * mask + (1 - mask): no FMA contraction (bit-exact colours)
            if (!with_background) v = __fmul_rn(v, m);                                   // (:203-204)
            if (white_bkgd) v = __fadd_rn(__fmul_rn(v, m), __fsub_rn(1.0f, m));          // (:244-245)
            c[a] = v;
        }
        rgbs[i * 3] = c[0]; rgbs[i * 3 + 1] = c[1]; rgbs[i * 3 + 2] = c[2];
        alphas[i] = m;
        if (pix) { pix[2 * i] = row; pix[2 * i + 1] = col; }
        an_make_ray(c2w + b * 12, focal + 2 * b, center + 2 * b, ginv ? ginv + b * 16 : nullptr, row, col,
                    near_, far_, rays + 2 * i);
    }
}

extern "C" int an_sample_training_rays_fwd(const uint8_t* images, const uint8_t* masks,
                                           const int32_t* fg_list, const int32_t* fg_off,
                                           const int32_t* bg_list, const int32_t* bg_off,
                                           const int32_t* frame_ids, const float* c2w, const float* focal,
                                           const float* center, const float* ginv, int B, int n, int n_fg, int H, int W,
                                           float near_, float far_, int white_bkgd, int with_background,
                                           const int32_t* sel, uint64_t seed,
                                           float* rays, float* rgbs, float* alphas, int32_t* pix, void* stream)
{
    if (!images || !masks || !fg_list || !fg_off || !bg_list || !bg_off || !frame_ids || !c2w || !focal || !center ||
        !rays || !rgbs || !alphas)
        return AN_ERR_ARG;
    if (B <= 0 || n <= 0 || n_fg < 0 || n_fg > n || H <= 0 || W <= 0) return AN_ERR_ARG;
    if (((uintptr_t)rays) & 15) return AN_ERR_ALIGN;
    const int64_t total = (int64_t)B * n;
    const int threads = 256;
    const int64_t want = (total + threads - 1) / threads;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    sample_training_rays_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        images, masks, fg_list, fg_off, bg_list, bg_off, frame_ids, c2w, focal, center, ginv, B, n, n_fg, H, W,
        near_, far_, white_bkgd, with_background, sel, seed, (float4*)rays, rgbs, alphas, pix);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
