// Shared helpers for the animnerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/animnerf_b200.h"

#define AN_CHECK_LAUNCH()                         \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

static inline int an_num_sms() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Counter-based RNG (Philox-4x32-10), one 128-bit block per (seed, counter): no state, so a
// kernel can draw U[0,1) for element i as philox(seed, i).  Written out here instead of using
// curand_kernel.h to keep the stream definition explicit and stable.
__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 ctr) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) {   // [0,1) with 24 bits
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float philox_u01(uint64_t seed, uint64_t i) {
    uint4 r = philox4x32_10(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                            make_uint4((uint32_t)(i >> 2), (uint32_t)(i >> 34), 0x616e696du, 0u));
    uint32_t v = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
    return u01(v);
}
