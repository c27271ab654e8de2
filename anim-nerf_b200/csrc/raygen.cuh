// One ray from a pixel: datasets/anim_nerf_dataset.py:56-85 (gen_ray_directions/gen_rays: pixel (col, row), no
// half-pixel offset, camera direction [(col-cx)/fx, -(row-cy)/fy, -1] normalised, rotated by c2w[:, :3]) fused
// with the ray part of models/anim_nerf.py:128-137 (root-frame transform, near/far clamp) when G != NULL.
// Shared by the ray-generation kernel and the training-ray sampler so both produce identical bits.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ void an_make_ray(const float* __restrict__ C, const float* __restrict__ focal,
                                            const float* __restrict__ center, const float* __restrict__ G,
                                            int row, int col, float near_, float far_, float4* __restrict__ out)
{
    const float fx = focal[0], fy = focal[1];
    const float cx = center[0], cy = center[1];
    float dx = ((float)col - cx) / fx, dy = -((float)row - cy) / fy, dz = -1.0f;
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= nrm; dy /= nrm; dz /= nrm;
    float d0 = dx * C[0] + dy * C[1] + dz * C[2];
    float d1 = dx * C[4] + dy * C[5] + dz * C[6];
    float d2 = dx * C[8] + dy * C[9] + dz * C[10];
    float o0 = C[3], o1 = C[7], o2 = C[11];
    float nr = near_, fr = far_;
    if (G) {
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
    out[0] = make_float4(o0, o1, o2, d0);
    out[1] = make_float4(d1, d2, nr, fr);
}
