// Ordered compaction of the valid-point flags: cidx = ids of the valid points in ascending order, count = how many.
// (A11: the MLP kernels, an_knn_unpose_bwd and the backward all walk this list.)  The KNN kernels can also append valid
// ids themselves with a warp-aggregated atomic, but then the list order depends on warp scheduling, and with it the
// composition of the MLP's 128-row tiles and the fp32 summation order of the weight gradient: last-bit differences
// from run to run that Adam amplifies (tools/probe_determinism.py).  Here the order is a function of the flags only --
// two small launches over 1 B/point: per-range counts, then every range sums its predecessors' counts (<= 1024 of
// them) and writes its ids.  Ascending ids also make the MLP's gathers and scatters walk memory forward.
#include "common.cuh"

namespace {

constexpr int CP_THREADS = 256;
constexpr int CP_CHUNK = CP_THREADS * 16;        // flags per block iteration (one uint4 per thread)
constexpr int CP_MAX_RANGES = 1024;

__device__ __forceinline__ int flags16(const uint8_t* __restrict__ valid, int64_t i, int64_t n, uint32_t (&w)[4])
{
    // 16 flags starting at i (i is a multiple of 16, the array 16-byte aligned); beyond n: zero
    if (i + 16 <= n) { const uint4 v = *(const uint4*)(valid + i); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
    else {
        w[0] = w[1] = w[2] = w[3] = 0u;
        for (int k = 0; k < 16 && i + k < n; ++k) w[k >> 2] |= (uint32_t)(valid[i + k] != 0) << (8 * (k & 3));
    }
    int c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { w[k] = (w[k] | (w[k] >> 1) | (w[k] >> 2) | (w[k] >> 3)) & 0x01010101u; c += __popc(w[k]); }   // any non-zero byte counts (flags are 0/1)
    return c;
}

__device__ __forceinline__ int block_sum(int v, int* sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int k = 0; k < CP_THREADS / 32; ++k) t += sh[k];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_count_kernel(const uint8_t* __restrict__ valid, int64_t n, int64_t range, int32_t* __restrict__ counts)
{
    __shared__ int sh[CP_THREADS / 32];
    const int64_t r0 = (int64_t)blockIdx.x * range, r1 = min(n, r0 + range);
    int c = 0;
    uint32_t w[4];
    for (int64_t i = r0 + (int64_t)threadIdx.x * 16; i < r1; i += CP_CHUNK) c += flags16(valid, i, r1, w);
    const int t = block_sum(c, sh);
    if (threadIdx.x == 0) counts[blockIdx.x] = t;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_write_kernel(const uint8_t* __restrict__ valid, int64_t n, int64_t range, const int32_t* __restrict__ counts,
                     int32_t* __restrict__ cidx, int32_t* __restrict__ count)
{
    __shared__ int sh[CP_THREADS / 32];
    __shared__ int wsum[CP_THREADS / 32];
    int pre = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += CP_THREADS) pre += counts[k];
    int base = block_sum(pre, sh);                       // ids of all earlier ranges come first
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *count = base + counts[blockIdx.x];
    const int64_t r0 = (int64_t)blockIdx.x * range, r1 = min(n, r0 + range);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t c0 = r0; c0 < r1; c0 += CP_CHUNK) {
        const int64_t i = c0 + (int64_t)threadIdx.x * 16;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        const int c = i < r1 ? flags16(valid, i, r1, w) : 0;
        int inc = c;                                      // inclusive scan over the block
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < CP_THREADS / 32; ++k) { const int v = wsum[k]; if (k < warp) woff += v; total += v; }
        int o = base + woff + inc - c;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if ((w[k >> 2] >> (8 * (k & 3))) & 1u) cidx[o++] = (int32_t)(i + k);
        base += total;
        __syncthreads();
    }
}

}  // namespace

extern "C" int64_t an_compact_ws_bytes(void) { return (int64_t)CP_MAX_RANGES * 4; }

extern "C" int an_compact_valid(const uint8_t* valid, int64_t n, int32_t* cidx, int32_t* count, void* ws, void* stream)
{
    if (!valid || !cidx || !count || !ws || n <= 0 || n > 0x7fffffffLL) return AN_ERR_ARG;
    if (((uintptr_t)valid) & 15) return AN_ERR_ALIGN;
    int64_t range = CP_CHUNK;
    while ((n + range - 1) / range > CP_MAX_RANGES) range *= 2;
    const unsigned n_ranges = (unsigned)((n + range - 1) / range);
    compact_count_kernel<<<n_ranges, CP_THREADS, 0, (cudaStream_t)stream>>>(valid, n, range, (int32_t*)ws);
    AN_CHECK_LAUNCH();
    compact_write_kernel<<<n_ranges, CP_THREADS, 0, (cudaStream_t)stream>>>(valid, n, range, (const int32_t*)ws, cidx, count);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
