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

// Stratified depth of sample i (models/volume_rendering.py:39-54), every operation rounded separately in the order
// torch evaluates it (near*(1-t) + far*t; lower + (upper-lower)*perturb*rand): no FMA contraction, so every kernel
// that samples (the stand-alone and the fused one) produces the same bits, whatever the surrounding code.
__device__ __forceinline__ float coarse_lin(float nr, float fr, float t) {
    return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, t)), __fmul_rn(fr, t));
}
__device__ __forceinline__ float coarse_depth(float nr, float fr, int i, int Kc, float perturb, float u)
{
    const float step = 1.0f / (float)Kc;
    float zi = coarse_lin(nr, fr, (float)i * step);
    if (perturb > 0.0f) {
        const float zm = coarse_lin(nr, fr, (float)(i - 1) * step), zp = coarse_lin(nr, fr, (float)(i + 1) * step);
        const float lo = (i == 0) ? zi : __fmul_rn(0.5f, __fadd_rn(zi, zm));
        const float hi = (i == Kc - 1) ? zi : __fmul_rn(0.5f, __fadd_rn(zp, zi));
        zi = __fadd_rn(lo, __fmul_rn(__fmul_rn(__fsub_rn(hi, lo), perturb), u));
    }
    return zi;
}

__global__ void sample_coarse_kernel(const float* __restrict__ rays, int64_t n_rays, int Kc,
                                     float perturb, const float* __restrict__ noise_u,
                                     uint64_t seed, float* __restrict__ z)
{
    const int64_t total = n_rays * Kc;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ray = e / Kc;
        const int i = (int)(e - ray * Kc);
        const float nr = __ldg(rays + ray * 8 + 6), fr = __ldg(rays + ray * 8 + 7);
        float u = 0.f;
        if (perturb > 0.0f) u = noise_u ? __ldg(noise_u + e) : philox_u01(seed, (uint64_t)e);
        z[e] = coarse_depth(nr, fr, i, Kc, perturb, u);
    }
}

// ------------------------------------------------------------------ fused: ray generation / body-space transform + stratified sampling
// One launch for A1/A2/A3 (north_star: "get_rays and stratified sampling become one fused ray-gen + sample kernel"):
// lane per ray, then the warp over each ray's samples.  The ray comes from the camera (pixel list or full grid) or from given
// world-space rays (the reference's training batches carry rays, train.py:172), is taken to the body's root frame
// with the near/far clamp (models/anim_nerf.py:128-137) and sampled (models/volume_rendering.py:29-56).  (Round 2 rebuilt the ray per sample,
// thread per sample: 9 warp-instructions per sample, 0.17 ms for a 512x512 frame.)
__device__ __forceinline__ void body_ray(const float* __restrict__ c2w, const float* __restrict__ focal,
                                         const float* __restrict__ center, const int32_t* __restrict__ pix,
                                         const float* __restrict__ rays_world, const float* __restrict__ ginv,
                                         int64_t ray, int R, int W, float near_, float far_, float4* out)
{
    const int b = (int)(ray / R);
    const float* G = ginv ? ginv + b * 16 : nullptr;
    if (rays_world) {
        const float4 r0 = __ldg((const float4*)rays_world + ray * 2), r1 = __ldg((const float4*)rays_world + ray * 2 + 1);
        float o0 = r0.x, o1 = r0.y, o2 = r0.z, d0 = r0.w, d1 = r1.x, d2 = r1.y, nr = r1.z, fr = r1.w;
        if (G) {
            const float p0 = G[0] * o0 + G[1] * o1 + G[2] * o2 + G[3];
            const float p1 = G[4] * o0 + G[5] * o1 + G[6] * o2 + G[7];
            const float p2 = G[8] * o0 + G[9] * o1 + G[10] * o2 + G[11];
            const float e0 = G[0] * d0 + G[1] * d1 + G[2] * d2;
            const float e1 = G[4] * d0 + G[5] * d1 + G[6] * d2;
            const float e2 = G[8] * d0 + G[9] * d1 + G[10] * d2;
            o0 = p0; o1 = p1; o2 = p2; d0 = e0; d1 = e1; d2 = e2;
            const float cam = sqrtf(o0 * o0 + o1 * o1 + o2 * o2);
            nr = fmaxf(nr, cam - 1.0f);
            fr = fminf(fr, cam + 1.0f);
        }
        out[0] = make_float4(o0, o1, o2, d0);
        out[1] = make_float4(d1, d2, nr, fr);
        return;
    }
    const int r = (int)(ray - (int64_t)b * R);
    int row, col;
    if (pix) { row = pix[2 * ray]; col = pix[2 * ray + 1]; }
    else     { row = r / W;        col = r - row * W; }
    an_make_ray(c2w + b * 12, focal + 2 * b, center + 2 * b, G, row, col, near_, far_, out);
}

__global__ void rays_sample_kernel(const float* __restrict__ c2w, const float* __restrict__ focal,
                                   const float* __restrict__ center, const int32_t* __restrict__ pix,
                                   const float* __restrict__ rays_world, const float* __restrict__ ginv,
                                   int B, int R, int W, int Kc, float near_, float far_, float perturb,
                                   const float* __restrict__ noise_u, uint64_t seed,
                                   float4* __restrict__ rays_body, float* __restrict__ z, int G)
{
    // a warp takes G (32) rays at a time: lane l builds ray l once (and writes it), then the warp walks the 32 rays with
    // the lanes over the samples (near'/far' by shuffle) -- every store coalesced, no per-sample ray rebuild
    const int lane = threadIdx.x & 31;
    const int64_t n_rays = (int64_t)B * R;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // G rays per warp and round (a power of two <= 32: 32 for full frames, 4 for a training batch of a few thousand
    // rays, which would otherwise leave most SMs without a warp)
    for (int64_t base = warp0 * G; base < n_rays; base += nwarps * G) {
        const int64_t ray = base + lane;
        float4 rb[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        if (lane < G && ray < n_rays) {
            body_ray(c2w, focal, center, pix, rays_world, ginv, ray, R, W, near_, far_, rb);
            rays_body[2 * ray] = rb[0]; rays_body[2 * ray + 1] = rb[1];
        }
        const int cnt = (int)(n_rays - base < G ? n_rays - base : G);
        for (int r = 0; r < cnt; ++r) {
            const float nr = __shfl_sync(0xffffffffu, rb[1].z, r), fr = __shfl_sync(0xffffffffu, rb[1].w, r);
            const int64_t e0 = (base + r) * Kc;
            for (int i = lane; i < Kc; i += 32) {
                float u = 0.f;
                if (perturb > 0.0f) u = noise_u ? __ldg(noise_u + e0 + i) : philox_u01(seed, (uint64_t)(e0 + i));
                z[e0 + i] = coarse_depth(nr, fr, i, Kc, perturb, u);
            }
        }
    }
}

// Backward of the fused op with respect to ginv (the only differentiable input on the training path: the SMPL root
// transform, models/anim_nerf.py:131): warp per ray, lanes over the samples.  With z_i = near'(1 - t_i) + far' t_i
// (also under perturbation: both stratum ends are affine in near', far'), near' = max(near, |o'| - 1),
// far' = min(far, |o'| + 1), o' = G o + t, d' = G d:
//   g_near' = g_rays[6] + sum_i g_z_i (1 - t_i),  g_far' = g_rays[7] + sum_i g_z_i t_i,  t_i = (z_i - near')/(far' - near')
//   g_o' = g_rays[0:3] + ([clamp near active] g_near' + [clamp far active] g_far') o'/|o'|,  g_d' = g_rays[3:6]
//   g_G.R += g_o' o^T + g_d' d^T,  g_G.t += g_o'   (atomically into g_ginv (B,4,4), zeroed by the launcher)
__global__ void __launch_bounds__(256)
rays_sample_bwd_kernel(const float* __restrict__ c2w, const float* __restrict__ focal, const float* __restrict__ center,
                       const int32_t* __restrict__ pix, const float* __restrict__ rays_world,
                       const float* __restrict__ rays_body, const float* __restrict__ z,
                       const float* __restrict__ g_rays, const float* __restrict__ g_z,
                       int B, int R, int W, int Kc, float near_, float far_, float* __restrict__ g_ginv)
{
    const int lane = threadIdx.x & 31;
    const int64_t n_rays = (int64_t)B * R;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // every warp owns a contiguous range of rays (almost always of one frame) and keeps the frame's 12 sums in
    // registers: one set of atomics per warp and frame instead of one per ray
    const int64_t per = (n_rays + nwarps - 1) / nwarps;
    const int64_t r_begin = warp0 * per, r_end = (warp0 + 1) * per < n_rays ? (warp0 + 1) * per : n_rays;
    float acc[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) acc[e] = 0.f;
    int cur_b = -1;
    for (int64_t ray = r_begin; ray < r_end; ++ray) {
        const int b = (int)(ray / R);
        if (b != cur_b) {
            if (lane == 0 && cur_b >= 0) {
#pragma unroll
                for (int e = 0; e < 12; ++e) { atomicAdd(g_ginv + cur_b * 16 + e, acc[e]); acc[e] = 0.f; }
            }
            cur_b = b;
        }
        const float4 b0 = __ldg((const float4*)rays_body + ray * 2), b1 = __ldg((const float4*)rays_body + ray * 2 + 1);
        const float nr = b1.z, fr = b1.w;
        const float inv = 1.0f / (fr - nr);
        float sn = 0.f, sf = 0.f;
        if (g_z) {
            for (int i = lane; i < Kc; i += 32) {
                const float g = __ldg(g_z + ray * Kc + i), t = (__ldg(z + ray * Kc + i) - nr) * inv;
                sn += g * (1.0f - t); sf += g * t;
            }
            sn = warp_sum(sn); sf = warp_sum(sf);
        }
        if (lane == 0) {
            // world-space ray: given, or regenerated from the camera (without the root transform)
            float4 w[2];
            body_ray(c2w, focal, center, pix, rays_world, nullptr, ray, R, W, near_, far_, w);
            const float o0 = w[0].x, o1 = w[0].y, o2 = w[0].z, d0 = w[0].w, d1 = w[1].x, d2 = w[1].y, nw = w[1].z, fw = w[1].w;
            const float4 g0 = __ldg((const float4*)g_rays + ray * 2), g1 = __ldg((const float4*)g_rays + ray * 2 + 1);
            const float gn = g1.z + sn, gf = g1.w + sf;
            const float cam = sqrtf(b0.x * b0.x + b0.y * b0.y + b0.z * b0.z);
            float s = 0.f;
            if (cam - 1.0f > nw) s += gn;          // near' = cam - 1
            if (cam + 1.0f < fw) s += gf;          // far'  = cam + 1
            s = cam > 0.f ? s / cam : 0.f;
            const float go0 = g0.x + s * b0.x, go1 = g0.y + s * b0.y, go2 = g0.z + s * b0.z;
            const float gd0 = g0.w, gd1 = g1.x, gd2 = g1.y;
            acc[0] += go0 * o0 + gd0 * d0; acc[1] += go0 * o1 + gd0 * d1; acc[2] += go0 * o2 + gd0 * d2; acc[3] += go0;
            acc[4] += go1 * o0 + gd1 * d0; acc[5] += go1 * o1 + gd1 * d1; acc[6] += go1 * o2 + gd1 * d2; acc[7] += go1;
            acc[8] += go2 * o0 + gd2 * d0; acc[9] += go2 * o1 + gd2 * d1; acc[10] += go2 * o2 + gd2 * d2; acc[11] += go2;
        }
    }
    if (lane == 0 && cur_b >= 0) {
#pragma unroll
        for (int e = 0; e < 12; ++e) atomicAdd(g_ginv + cur_b * 16 + e, acc[e]);
    }
}

// Ray-side gradients of one render pass from the per-point gradients (x = o + z d): warp per ray, lanes over samples.
//   g_o = sum_k g_x_k,  g_d = sum_k z_k g_x_k,  g_z_k = g_z_comp_k + g_x_k . d   (g_x read at valid samples only)
// writes g_rays (n_rays, 8) = [g_o, g_d, 0, g_far_comp] and g_z (n_rays, K).
__global__ void __launch_bounds__(256)
ray_point_grad_kernel(const float* __restrict__ rays, const float* __restrict__ z, const uint8_t* __restrict__ valid,
                      const float* __restrict__ g_xyz, const float* __restrict__ g_z_comp, const float* __restrict__ g_far_comp,
                      int64_t n_rays, int K, float4* __restrict__ g_rays, float* __restrict__ g_z)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ray = warp0; ray < n_rays; ray += nwarps) {
        const float4 r0 = __ldg((const float4*)rays + ray * 2), r1 = __ldg((const float4*)rays + ray * 2 + 1);
        const float d0 = r0.w, d1 = r1.x, d2 = r1.y;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        for (int k = lane; k < K; k += 32) {
            const int64_t gid = ray * K + k;
            float gz = g_z_comp ? __ldg(g_z_comp + gid) : 0.f;
            if (valid[gid]) {
                const float x0 = g_xyz[gid * 3], x1 = g_xyz[gid * 3 + 1], x2 = g_xyz[gid * 3 + 2], zz = __ldg(z + gid);
                a0 += x0; a1 += x1; a2 += x2;
                c0 += zz * x0; c1 += zz * x1; c2 += zz * x2;
                gz += x0 * d0 + x1 * d1 + x2 * d2;
            }
            g_z[gid] = gz;
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
        if (lane == 0) {
            g_rays[2 * ray] = make_float4(a0, a1, a2, c0);
            g_rays[2 * ray + 1] = make_float4(c1, c2, 0.f, g_far_comp ? __ldg(g_far_comp + ray) : 0.f);
        }
    }
}

static int launch_blocks(int64_t work_items, int threads)
{
    const int64_t want = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)an_num_sms() * 16;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

extern "C" int an_rays_sample_fwd(const float* c2w, const float* focal, const float* center, const int32_t* pix,
                                  const float* rays_world, const float* ginv, int B, int R, int H, int W, int Kc,
                                  float near_, float far_, float perturb, const float* noise_u, uint64_t seed,
                                  float* rays_body, float* z, void* stream)
{
    if (!rays_body || !z || B <= 0 || R <= 0 || Kc <= 0) return AN_ERR_ARG;
    if (!rays_world && (!c2w || !focal || !center)) return AN_ERR_ARG;
    if (!rays_world && !pix && (H <= 0 || W <= 0 || (int64_t)H * W != R)) return AN_ERR_ARG;
    if ((((uintptr_t)rays_body) | ((uintptr_t)rays_world)) & 15) return AN_ERR_ALIGN;
    const int G = (int64_t)B * R >= 131072 ? 32 : 4;
    rays_sample_kernel<<<launch_blocks((int64_t)B * R * (32 / G), 128), 128, 0, (cudaStream_t)stream>>>(     // one warp per G rays
        c2w, focal, center, pix, rays_world, ginv, B, R, W, Kc, near_, far_, perturb, noise_u, seed, (float4*)rays_body, z, G);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int an_rays_sample_bwd(const float* c2w, const float* focal, const float* center, const int32_t* pix,
                                  const float* rays_world, const float* rays_body, const float* z,
                                  const float* g_rays_body, const float* g_z, int B, int R, int H, int W, int Kc,
                                  float near_, float far_, float* g_ginv, void* stream)
{
    if (!rays_body || !z || !g_rays_body || !g_ginv || B <= 0 || R <= 0 || Kc <= 0) return AN_ERR_ARG;
    if (!rays_world && (!c2w || !focal || !center)) return AN_ERR_ARG;
    if ((((uintptr_t)rays_body) | ((uintptr_t)rays_world) | ((uintptr_t)g_rays_body)) & 15) return AN_ERR_ALIGN;
    cudaError_t e = cudaMemsetAsync(g_ginv, 0, (size_t)B * 16 * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    rays_sample_bwd_kernel<<<launch_blocks((int64_t)B * R * 2, 256), 256, 0, (cudaStream_t)stream>>>(    // 16 rays per warp
        c2w, focal, center, pix, rays_world, rays_body, z, g_rays_body, g_z, B, R, W, Kc, near_, far_, g_ginv);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int an_ray_point_grad(const float* rays, const float* z, const uint8_t* valid, const float* g_xyz,
                                 const float* g_z_comp, const float* g_far_comp, int64_t n_rays, int K,
                                 float* g_rays, float* g_z, void* stream)
{
    if (!rays || !z || !valid || !g_xyz || !g_rays || !g_z || n_rays <= 0 || K <= 0) return AN_ERR_ARG;
    if ((((uintptr_t)rays) | ((uintptr_t)g_rays)) & 15) return AN_ERR_ALIGN;
    ray_point_grad_kernel<<<launch_blocks(n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        rays, z, valid, g_xyz, g_z_comp, g_far_comp, n_rays, K, (float4*)g_rays, g_z);
    AN_CHECK_LAUNCH();
    return AN_OK;
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
