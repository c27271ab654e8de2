// Micro-benchmark: issue rate of back-to-back tcgen05.mma.cta_group::2 (M = 256, K = 16) by form / N / operand reuse.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I anim-nerf_b200/csrc tools/micro/mma_rate.cu -o tools/_variants/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace tc;

__device__ __forceinline__ void umma_pair_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(bd), "r"(id), "r"(acc) : "memory");
}

// mode bit0: TS form; nb = distinct 8 KB-spaced B buffers cycled per group; na = accumulators cycled per group;
// kper = MMAs per group (commit after each group); wait_every = groups between barrier waits
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
rate_kernel(int N, int ts, int nb, int na, int kper, int groups, int uni, int flags, long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t raw = smem_u32(smem);
    const uint32_t sb = (raw + 1023u) & ~1023u;
    const uint32_t bar = sb + 200 * 1024, slot = bar + 64;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    for (uint32_t i = threadIdx.x; i < 200 * 256; i += blockDim.x) ((uint32_t*)(smem + (sb - raw)))[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 1) tmem_alloc_pair(slot, 512);
    tc_fence_before(); cluster_sync_all(); tc_fence_after();
    const uint32_t tm = *(volatile uint32_t*)(smem + (sb - raw) + 200 * 1024 + 64);
    if (uni && warp == 0 && rank == 0) {
        // whole warp runs the loop (warp-uniform operands), one elected lane issues
        const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
        uint32_t phase = 0;
        long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t ga = (flags & 1) ? 0u : (uint32_t)(g & (na - 1)), gb = (flags & 1) ? 0u : (uint32_t)(g & (nb - 1));
            const uint32_t acc = N > 128 ? tm + (ts ? 256u : ga * 256u) : tm + 256 + ga * 128u;
            const uint32_t wb = sb + 32768 + gb * 16384u;
            const uint32_t always = (flags & 2) ? 1u : 0u;
            if (elect_one()) {
                for (int k = 0; k < kper; ++k) {
                    if (ts) umma_pair_ts(acc, tm + (uint32_t)(k & 3) * 8u, make_desc(wb + (k & 3) * 32u, 16, 1024), idesc, (k > 0) | always);
                    else umma_pair(acc, make_desc(sb + (k & 3) * 32u, 16, 1024), make_desc(wb + (k & 3) * 32u, 16, 1024), idesc, (k > 0) | always);
                }
            }
            __syncwarp();
            if ((g & 15) == 15 || g == groups - 1) {
                if (elect_one()) umma_commit_pair(bar);
                __syncwarp();
                mbar_wait(bar, phase); phase ^= 1u;
            }
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
    } else if (warp == 0 && (threadIdx.x & 31) == 0 && rank == 0) {
        const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
        uint32_t phase = 0;
        long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t acc = N > 128 ? tm + (ts ? 256u : (uint32_t)(g % na) * 256u) : tm + 256 + (uint32_t)(g % na) * 128u;
            const uint32_t wb = sb + 32768 + (uint32_t)(g % nb) * 16384u;
            for (int k = 0; k < kper; ++k) {
                if (ts) umma_pair_ts(acc, tm + (uint32_t)(k & 3) * 8u, make_desc(wb + (k & 3) * 32u, 16, 1024), idesc, k > 0);
                else umma_pair(acc, make_desc(sb + (k & 3) * 32u, 16, 1024), make_desc(wb + (k & 3) * 32u, 16, 1024), idesc, k > 0);
            }
            if ((g & 15) == 15 || g == groups - 1) {
                umma_commit_pair(bar);
                mbar_wait(bar, phase); phase ^= 1u;
            }
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    } else if (warp == 0 && (threadIdx.x & 31) == 0) {
        // peer: consume the multicast commits
        uint32_t phase = 0;
        for (int g = 0; g < groups; ++g)
            if ((g & 15) == 15 || g == groups - 1) { mbar_wait(bar, phase); phase ^= 1u; }
    }
    tc_fence_before(); cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tm, 512);
}

int main()
{
    long long* out; cudaMalloc(&out, 8);
    const int smem = 202 * 1024 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    struct { int N, ts, nb, na, kper; } cfg[] = {
        {256, 0, 4, 2, 4}, {256, 0, 4, 2, 16}, {128, 0, 4, 2, 4}, {128, 1, 4, 2, 4}, {128, 1, 4, 2, 16}, {128, 1, 4, 2, 1}};
    for (int flags = 0; flags < 4; ++flags)
    for (int uni = 1; uni < 2; ++uni)
    for (auto c : cfg) {
        const int groups = 512;
        for (int rep = 0; rep < 2; ++rep) {
            rate_kernel<<<148, 128, smem>>>(c.N, c.ts, c.nb, c.na, c.kper, groups, uni, flags, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("flags=%d uni=%d N=%3d %s nb=%d na=%d kper=%2d : %7.1f clk / MMA (ideal %d)\n", flags, uni, c.N, c.ts ? "TS" : "SS", c.nb, c.na, c.kper,
               (double)h / (groups * c.kper), c.N / 2);
    }
    return 0;
}
