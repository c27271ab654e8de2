// Weight repack: fp32 nn.Linear (out,in) tensors -> bf16 UMMA-ready swizzled chunk images
// (forward W and backward W^T), fp32 bias block, fused head-layer block, flat fp32 copy.  Runs once
// per optimiser step (2.4 MB read, ~5 MB written per net); checkpoints keep the reference's
// state-dict layout because the fp32 nn.Parameters stay the source of truth.
//
// mlp_fuse_kernel forms the head layer of mlp_layout.cuh in fp32: W' = W_dir . W_final (128x256),
// b' = W_dir . b_final + b_dir (reference models/nerf.py:173 then :150: no activation in between).
#include "common.cuh"
#include "mlp_layout.cuh"
#include <cuda_bf16.h>

struct PackPtrs { const float* w[mlp::NLIN]; const float* b[mlp::NLIN]; };

// grid 129 x 1024 threads: CTA i < 128 -> row i of W' (thread (q, j): the quarter q of sum_k Wd[i][k] Wf[k][j], the
// four quarters added through shared memory in a fixed order); CTA 128 -> b'.  (One thread per output with a 256-step
// dependent loop was latency-bound: 21 us per net and step.)
__global__ void __launch_bounds__(1024)
mlp_fuse_kernel(PackPtrs P, uint8_t* __restrict__ packed)
{
    using namespace mlp;
    float* fused = (float*)(packed + FUSED_OFF);
    const float* __restrict__ Wf = P.w[8];
    const float* __restrict__ Wd = P.w[9];
    __shared__ float s_row[256];
    __shared__ float s_part[4][256];
    const int i = blockIdx.x, j = threadIdx.x & 255, q = threadIdx.x >> 8;
    if (i < 128) {
        if (q == 0) s_row[j] = Wd[i * 256 + j];
        __syncthreads();
        float acc = 0.f;
#pragma unroll 16
        for (int k = q * 64; k < q * 64 + 64; ++k) acc += s_row[k] * __ldg(Wf + k * 256 + j);     // coalesced over j
        s_part[q][j] = acc;
        __syncthreads();
        if (q == 0) fused[i * 256 + j] = ((s_part[0][j] + s_part[1][j]) + s_part[2][j]) + s_part[3][j];
    } else {
        // b'[jj] = b_dir[jj] + sum_k Wd[jj][k] b_final[k]: warp per output, lanes over k
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int jj = warp; jj < 128; jj += 32) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < 8; ++m) acc += Wd[jj * 256 + lane + 32 * m] * P.b[8][lane + 32 * m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) fused[128 * 256 + jj] = acc + P.b[9][jj];
        }
    }
}

// value of the forward B operand of GEMM layer g at (output row r, input column c)
__device__ __forceinline__ float fwd_weight(const PackPtrs& P, const float* fused, int g, int r, int c)
{
    using namespace mlp;
    if (g < 8) return P.w[g][(int64_t)r * lin_in(g) + c];
    if (g == 8) return r < 128 ? fused[r * 256 + c] : (r == 128 ? P.w[10][c] : 0.f);
    return r < 3 ? P.w[11][r * 128 + c] : 0.f;
}

// PACK_SPLIT CTAs per chunk image (fwd: 36, bwd: 40) and per bias slab (9) -- a thread converts 8 elements instead of 64:
// the kernel is a chain of dependent load -> store rounds, not bandwidth --, then CTAs for the fp32 blocks
#define PACK_SPLIT 8
__global__ void __launch_bounds__(256)
mlp_pack_kernel(PackPtrs P, uint8_t* __restrict__ packed)
{
    using namespace mlp;
    const float* fused = (const float*)(packed + FUSED_OFF);
    constexpr int IMG_JOBS = FWD_CHUNKS + FWD_SLABS + BWD_CHUNKS;
    int job = blockIdx.x;
    int part = 0, nparts = 1;
    if (job < IMG_JOBS * PACK_SPLIT) { part = job % PACK_SPLIT; nparts = PACK_SPLIT; job /= PACK_SPLIT; }
    else job -= IMG_JOBS * (PACK_SPLIT - 1);
    const int e0 = part * blockDim.x + threadIdx.x, estep = nparts * blockDim.x;
    if (job < FWD_CHUNKS) {
        int g = 0, kc = job;
        while (kc >= g_chunks(g)) { kc -= g_chunks(g); ++g; }
        const int N = g_N(g);
        // source column range of this chunk
        int c0, ncol;
        if (g == 0) { c0 = 0; ncol = 63; }
        else if (g == 4) { if (kc == 0) { c0 = 0; ncol = 63; } else { c0 = 63 + 64 * (kc - 1); ncol = 64; } }
        else { c0 = 64 * kc; ncol = 64; }
        uint8_t* dst = packed + fwd_chunk_off(g, kc);
        for (int e = e0; e < N * 64; e += estep) {
            const int r = e >> 6, c = e & 63;
            const float v = (c < ncol) ? fwd_weight(P, fused, g, r, c0 + c) : 0.0f;
            *(__nv_bfloat16*)(dst + img_off(r, c)) = __float2bfloat16_rn(v);
        }
        return;
    }
    job -= FWD_CHUNKS;
    if (job < FWD_SLABS) {        // bias slab of GEMM layer g: k = 15 carries the bias, everything else is zero
        const int g = job, N = g_N(g);
        uint8_t* dst = packed + fwd_bias_off(g);
        for (int e = e0; e < N * 16; e += estep) {
            const int r = e >> 4, k = e & 15;
            float v = 0.f;
            if (k == 15) v = g < 8 ? P.b[g][r] : (r < 128 ? fused[128 * 256 + r] : 0.f);
            *(__nv_bfloat16*)(dst + slab_off(r, k)) = __float2bfloat16_rn(v);
        }
        return;
    }
    job -= FWD_SLABS;
    if (job < BWD_CHUNKS) {
        int s = 0, kc = job;
        while (kc >= bs_chunks(s)) { kc -= bs_chunks(s); ++s; }
        const int rows = bs_rows(s);
        uint8_t* dst = packed + bwd_chunk_off(s, kc);
        for (int e = e0; e < rows * 64; e += estep) {
            const int r = e >> 6, c = e & 63;          // r = input feature (row of W^T), c = output feature in chunk
            float v = 0.f;
            if (s == 0) { if (c < 3) v = P.w[11][c * 128 + r]; }                     // rgb^T
            else if (s == 1) {                                                        // head^T
                const int o = kc * 64 + c;
                if (o < 128) v = fused[o * 256 + r]; else if (o == 128) v = P.w[10][r];
            } else {
                const int g = bs_layer(s), in = lin_in(g), out = lin_out(g);
                const int o = kc * 64 + c;
                if (r < bs_in_valid(s) && o < out) v = P.w[g][(int64_t)o * in + bs_in0(s) + r];
            }
            *(__nv_bfloat16*)(dst + img_off(r, c)) = __float2bfloat16_rn(v);
        }
        return;
    }
    job -= BWD_CHUNKS;
    if (job == 0) {
        float* sm = (float*)(packed + SMALL_OFF);
        for (int e = threadIdx.x; e < SMALL_FLOATS; e += blockDim.x) {
            const int g = e >> 8, c = e & 255;
            float v = 0.f;
            if (g < 8) v = P.b[g][c];
            else if (g == 8) v = c < 128 ? fused[128 * 256 + c] : (c == 128 ? P.b[10][0] : 0.f);
            else v = c < 3 ? P.b[11][c] : 0.f;
            sm[e] = v;
        }
        return;
    }
    // flat fp32 copy: remaining CTAs stride over all linears
    const int nflat = gridDim.x - IMG_JOBS * PACK_SPLIT - 1;
    float* flat = (float*)(packed + FLAT_OFF);
    for (int id = 0; id < NLIN; ++id) {
        const int64_t nw = (int64_t)lin_out(id) * lin_in(id);
        for (int64_t e = (int64_t)(job - 1) * blockDim.x + threadIdx.x; e < nw; e += (int64_t)nflat * blockDim.x)
            flat[flat_w_off(id) + e] = P.w[id][e];
        for (int64_t e = (int64_t)(job - 1) * blockDim.x + threadIdx.x; e < lin_out(id); e += (int64_t)nflat * blockDim.x)
            flat[flat_b_off(id) + e] = P.b[id][e];
    }
}

// Chain rule through the fused head layer, once per backward call: from dW' (128x256), db' (128) in the
// scratch tail of the gradient vector to  dW_final = W_dir^T dW',  dW_dir = dW' W_final^T + db' b_final^T,
// db_final = W_dir^T db',  db_dir = db'.   grid (256 + 128 + 1) x 256 threads; reads the flat fp32 copy.
__global__ void __launch_bounds__(1024)
mlp_unfuse_grad_kernel(const uint8_t* __restrict__ packed, float* __restrict__ g)
{
    using namespace mlp;
    const float* flat = (const float*)(packed + FLAT_OFF);
    const float* Wf = flat + flat_w_off(8);
    const float* bf = flat + flat_b_off(8);
    const float* Wd = flat + flat_w_off(9);
    const float* dWp = g + GRAD_FUSED_W;
    const float* dbp = g + GRAD_FUSED_B;
    __shared__ float s_v[256];
    __shared__ float s_part[4][256];
    const int blk = blockIdx.x, j = threadIdx.x & 255, q = threadIdx.x >> 8;     // 1024 threads: 4 quarters x 256 columns
    if (blk < 256) {                     // dW_final[k = blk][j] = sum_i Wd[i][k] dW'[i][j]: quarter q sums i in [32 q, 32 q + 32)
        const int k = blk;
        if (threadIdx.x < 128) s_v[threadIdx.x] = Wd[threadIdx.x * 256 + k];
        __syncthreads();
        float acc = 0.f;
#pragma unroll 16
        for (int i = q * 32; i < q * 32 + 32; ++i) acc += s_v[i] * dWp[i * 256 + j];
        s_part[q][j] = acc;
        __syncthreads();
        if (q == 0) g[flat_w_off(8) + k * 256 + j] += ((s_part[0][j] + s_part[1][j]) + s_part[2][j]) + s_part[3][j];
    } else if (blk < 384) {              // dW_dir[i][k] = sum_jj dW'[i][jj] Wf[k][jj] + db'[i] bf[k]
        // warp per output k (32 warps), lanes over jj: every W_final row is read as coalesced 128-byte segments
        const int i = blk - 256;
        if (threadIdx.x < 256) s_v[threadIdx.x] = dWp[i * 256 + threadIdx.x];
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float sv[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) sv[m] = s_v[lane + 32 * m];
        const float dbi = dbp[i];
        for (int k = warp; k < 256; k += 32) {
            const float* wrow = Wf + (int64_t)k * 256;
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < 8; ++m) acc += sv[m] * wrow[lane + 32 * m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) g[flat_w_off(9) + i * 256 + k] += acc + dbi * bf[k];
        }
    } else if (threadIdx.x < 256) {      // db_final[k = j] = sum_i Wd[i][k] db'[i];  db_dir = db'
        float acc = 0.f;
#pragma unroll 16
        for (int i = 0; i < 128; ++i) acc += Wd[i * 256 + j] * dbp[i];
        g[flat_b_off(8) + j] += acc;
        if (j < 128) g[flat_b_off(9) + j] += dbp[j];
    }
}

int mlp_unfuse_grad_launch(const void* packed, float* g_params, cudaStream_t stream)
{
    mlp_unfuse_grad_kernel<<<385, 1024, 0, stream>>>((const uint8_t*)packed, g_params);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int64_t an_mlp_packed_bytes(void) { return mlp::PACKED_BYTES; }
extern "C" int64_t an_mlp_grad_floats(void) { return mlp::GRAD_FLOATS; }

extern "C" int an_mlp_pack(const float* const* w_host, const float* const* b_host, void* packed, void* stream)
{
    if (!w_host || !b_host || !packed) return AN_ERR_ARG;
    if (((uintptr_t)packed) & 1023) return AN_ERR_ALIGN;
    PackPtrs P;
    for (int i = 0; i < mlp::NLIN; ++i) {
        if (!w_host[i] || !b_host[i]) return AN_ERR_ARG;
        P.w[i] = w_host[i]; P.b[i] = b_host[i];
    }
    mlp_fuse_kernel<<<129, 1024, 0, (cudaStream_t)stream>>>(P, (uint8_t*)packed);
    AN_CHECK_LAUNCH();
    const int blocks = (mlp::FWD_CHUNKS + mlp::FWD_SLABS + mlp::BWD_CHUNKS) * PACK_SPLIT + 1 + 64;
    mlp_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, (uint8_t*)packed);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
