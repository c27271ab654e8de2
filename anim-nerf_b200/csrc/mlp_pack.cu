// Weight repack: fp32 nn.Linear (out,in) tensors -> bf16 UMMA-ready swizzled chunk images
// (forward W and backward W^T), fp32 bias/head block, flat fp32 copy.  Runs once per
// optimiser step (2.4 MB read, 4.8 MB written per net); checkpoints keep the reference's
// state-dict layout because the fp32 nn.Parameters stay the source of truth.
#include "common.cuh"
#include "mlp_layout.cuh"
#include <cuda_bf16.h>

struct PackPtrs { const float* w[mlp::NLIN]; const float* b[mlp::NLIN]; };

// one CTA per chunk image (fwd: 38, bwd: 42), then CTAs for the fp32 blocks
__global__ void __launch_bounds__(256)
mlp_pack_kernel(PackPtrs P, uint8_t* __restrict__ packed)
{
    using namespace mlp;
    int job = blockIdx.x;
    if (job < FWD_CHUNKS) {
        int g = 0, kc = job;
        while (kc >= g_chunks(g)) { kc -= g_chunks(g); ++g; }
        const int N = g_N(g), in = lin_in(g);
        // source column range of this chunk
        int c0, ncol;
        if (g == 0) { c0 = 0; ncol = 63; }
        else if (g == 4) { if (kc == 0) { c0 = 0; ncol = 63; } else { c0 = 63 + 64 * (kc - 1); ncol = 64; } }
        else { c0 = 64 * kc; ncol = 64; }
        uint8_t* dst = packed + fwd_chunk_off(g, kc);
        const float* Wg = P.w[g];
        for (int e = threadIdx.x; e < N * 64; e += blockDim.x) {
            const int r = e >> 6, c = e & 63;
            const float v = (c < ncol) ? Wg[(int64_t)r * in + c0 + c] : 0.0f;
            *(__nv_bfloat16*)(dst + img_off(r, c)) = __float2bfloat16_rn(v);
        }
        return;
    }
    job -= FWD_CHUNKS;
    if (job < BWD_CHUNKS) {
        int s = 0, kc = job;
        while (kc >= bs_chunks(s)) { kc -= bs_chunks(s); ++s; }
        const int g = bs_layer(s), rows = bs_rows(s), in = lin_in(g), out = lin_out(g);
        const int in0 = bs_in0(s), nvalid = bs_in_valid(s);
        uint8_t* dst = packed + bwd_chunk_off(s, kc);
        const float* Wg = P.w[g];
        for (int e = threadIdx.x; e < rows * 64; e += blockDim.x) {
            const int r = e >> 6, c = e & 63;          // r = input feature (row of W^T), c = output feature in chunk
            const int o = kc * 64 + c;
            const float v = (r < nvalid && o < out) ? Wg[(int64_t)o * in + in0 + r] : 0.0f;
            *(__nv_bfloat16*)(dst + img_off(r, c)) = __float2bfloat16_rn(v);
        }
        return;
    }
    job -= BWD_CHUNKS;
    if (job == 0) {
        float* sm = (float*)(packed + SMALL_OFF);
        for (int e = threadIdx.x; e < SMALL_FLOATS; e += blockDim.x) {
            float v = 0.f;
            if (e < SM_WS) { const int g = e >> 8, c = e & 255; v = (c < lin_out(g)) ? P.b[g][c] : 0.f; }
            else if (e < SM_BS) v = P.w[10][e - SM_WS];
            else if (e < SM_WR) v = (e == SM_BS) ? P.b[10][0] : 0.f;
            else if (e < SM_BR) v = P.w[11][e - SM_WR];
            else v = (e - SM_BR < 3) ? P.b[11][e - SM_BR] : 0.f;
            sm[e] = v;
        }
        return;
    }
    // flat fp32 copy: remaining CTAs stride over all linears
    const int nflat = gridDim.x - FWD_CHUNKS - BWD_CHUNKS - 1;
    float* flat = (float*)(packed + FLAT_OFF);
    for (int id = 0; id < NLIN; ++id) {
        const int64_t nw = (int64_t)lin_out(id) * lin_in(id);
        for (int64_t e = (int64_t)(job - 1) * blockDim.x + threadIdx.x; e < nw; e += (int64_t)nflat * blockDim.x)
            flat[flat_w_off(id) + e] = P.w[id][e];
        for (int64_t e = (int64_t)(job - 1) * blockDim.x + threadIdx.x; e < lin_out(id); e += (int64_t)nflat * blockDim.x)
            flat[flat_b_off(id) + e] = P.b[id][e];
    }
}

extern "C" int64_t an_mlp_packed_bytes(void) { return mlp::PACKED_BYTES; }
extern "C" int64_t an_mlp_grad_floats(void) { return mlp::FLAT_FLOATS; }

extern "C" int an_mlp_pack(const float* const* w_host, const float* const* b_host, void* packed, void* stream)
{
    if (!w_host || !b_host || !packed) return AN_ERR_ARG;
    if (((uintptr_t)packed) & 1023) return AN_ERR_ALIGN;
    PackPtrs P;
    for (int i = 0; i < mlp::NLIN; ++i) {
        if (!w_host[i] || !b_host[i]) return AN_ERR_ARG;
        P.w[i] = w_host[i]; P.b[i] = b_host[i];
    }
    const int blocks = mlp::FWD_CHUNKS + mlp::BWD_CHUNKS + 1 + 64;
    mlp_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, (uint8_t*)packed);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
