// Microbenchmark: TMEM read (tcgen05.ld) / write (tcgen05.st) throughput per SM vs number of warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../anim-nerf_b200/csrc/tc_common.cuh"
using namespace tc;

__device__ __forceinline__ void use(uint32_t (&v)[32], uint32_t& acc) {
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("xor.b32 %0, %0, %1;" : "+r"(acc) : "r"(v[i]));
}
template <int MODE>   // 0: ld x32 with wait each, 1: ld x32 two in flight, 2: st x32, 3: ld+st+cvt+sts like the epilogue
__global__ void k(int iters, long long* out, uint32_t* sink)
{
    __shared__ uint32_t slot;
    __shared__ uint4 sm[2048];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 256;
    uint32_t va[32], vb[32];
    uint32_t acc = 0;
    for (int i = 0; i < 32; ++i) { va[i] = i; vb[i] = i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int b = 0; b < 8; ++b) { tmem_ld32(tb + b * 32, va); tmem_ld_wait(); use(va, acc); }
        } else if (MODE == 1) {
            tmem_ld32(tb, va);
#pragma unroll
            for (int b = 0; b < 8; b += 2) {
                tmem_ld_wait(); tmem_ld32(tb + (b + 1) * 32, vb); use(va, acc);
                tmem_ld_wait(); if (b + 2 < 8) tmem_ld32(tb + (b + 2) * 32, va); use(vb, acc);
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int b = 0; b < 8; ++b) { va[0] = acc + b; tmem_st32(tb + b * 32, va); }
            tmem_st_wait();
        } else {
            tmem_ld32(tb, va);
#pragma unroll
            for (int b = 0; b < 8; b += 2) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t (&v)[32] = h ? vb : va;
                    tmem_ld_wait();
                    if (b + h + 1 < 8) tmem_ld32(tb + (b + h + 1) * 32, h ? va : vb);
                    uint32_t w[16];
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) w[kk] = pack_relu_bf16(__uint_as_float(v[2 * kk]), __uint_as_float(v[2 * kk + 1]));
#pragma unroll
                    for (int u = 0; u < 4; ++u) sm[(u * blockDim.x + threadIdx.x) & 2047] = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
                }
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[threadIdx.x] = acc + sm[threadIdx.x].x;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

template <int MODE> void run(const char* name, int warps)
{
    long long* d; uint32_t* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4096);
    const int iters = 200;
    k<MODE><<<148, warps * 32>>>(iters, d, s);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * 8 * 4096 * warps;      // per SM
    printf("%-28s warps %2d  clk %8lld  %.1f B/clk/SM  (%s)\n", name, warps, h[0], bytes / h[0], cudaGetErrorString(e));
    cudaFree(d); cudaFree(s);
}

int main()
{
    for (int w : {4, 8, 16}) run<0>("ld x32, wait each", w);
    for (int w : {4, 8, 16}) run<1>("ld x32, 2 in flight", w);
    for (int w : {4, 8, 16}) run<2>("st x32", w);
    for (int w : {4, 8, 16}) run<3>("ld + cvt.relu + sts", w);
    return 0;
}
