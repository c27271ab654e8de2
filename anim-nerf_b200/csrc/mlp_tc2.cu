// A9-A11 forward, second kernel organisation ("pipe"): ONE 128-row tile per CTA iteration, two TMEM accumulators
// that alternate between consecutive layers, and chunk-level hand-over between the epilogue of layer g and the MMAs
// of layer g+1.  Same arithmetic, images, stash and outputs as mlp_fwd_tc_kernel (mlp_tc.cu); selected with impl = 2.
//
// Why: the two-tile ping-pong kernel needs 2 x (64 KB activations + 16 KB encoding) of shared memory, which leaves
// room for only two 32 KB weight slots; in training mode the L2 -> smem weight stream then falls behind the tensor
// core (the ring is latency-bound).  With one tile resident the ring has FOUR slots, all eight epilogue
// warps work on the same tile (two threads per row, 128 columns each), and the tensor core is kept busy across the
// layer boundary by starting layer g+1 on K-chunk kc as soon as the epilogue of layer g has produced that 64-column
// chunk of the activation image (the other accumulator is free, so nothing has to drain first).
//   warp 0      weight producer, 4-stage ring of 32 KB chunk images
//   warp 1      MMA issuer: per K-chunk waits for (a) the A chunk (encoding barrier / per-chunk activation barrier) and
//               (b) the weight slot; accumulator = layer parity; commit per slot and per layer
//   warps 2-9   epilogue: warp w owns TMEM lanes 32*(w%4).. (rows) and columns 128*((w-2)/4).. ; per 32-column block:
//               tcgen05.ld -> (mask) -> cvt.relu.bf16x2 -> swizzled smem image; after every second block the 64-column
//               chunk is published (fence.proxy.async + mbarrier arrive) and, when training, its 4 KB row slab is
//               streamed to the stash.  While draining layer g the accumulator is re-initialised (tcgen05.st) with the
//               bias of the layer that writes it next (g+2).
#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc_common.cuh"

namespace {
constexpr int THREADS = 320;
constexpr int NSTAGE = 4;
constexpr uint32_t SM_ACT = 0;                          // [4 chunks][128 rows x 128 B]
constexpr uint32_t SM_ENC = 65536;                      // [128 rows x 128 B]
constexpr uint32_t SM_WST = 81920;                      // [NSTAGE][32 KB]
constexpr uint32_t SM_BAR = SM_WST + NSTAGE * 32768;    // 212992
constexpr uint32_t SM_BIASBUF = SM_BAR + 256;           // [2 column halves][2 layer parities][128 fp32]
constexpr uint32_t SM_BYTES = SM_BIASBUF + 2048;
constexpr uint32_t SM_ALLOC = SM_BYTES + 1024;          // slack for manual 1024-B alignment
static_assert(SM_ALLOC <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");
}  // namespace

template <bool TRAIN>
__global__ void __launch_bounds__(THREADS, 1)
mlp_fwd_pipe_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ xyz_cano,
                    const int32_t* __restrict__ cidx, const int32_t* __restrict__ count, int64_t n_max,
                    float* __restrict__ sigma_out, float* __restrict__ rgb_out, uint8_t* __restrict__ stash)
{
    using namespace mlp;
    using namespace tc;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t bar_full = sbase + SM_BAR;            // [NSTAGE]  TMA -> MMA
    const uint32_t bar_empty = sbase + SM_BAR + 32;      // [NSTAGE]  MMA -> TMA
    const uint32_t bar_actr = sbase + SM_BAR + 64;       // [4 chunks] epilogue -> MMA: activation chunk written (128 arrivals)
    const uint32_t bar_encr = sbase + SM_BAR + 96;       // encoding image written (128 arrivals)
    const uint32_t bar_accr = sbase + SM_BAR + 104;      // [2] MMA -> epilogue: accumulator of a layer complete
    const uint32_t tmem_slot = sbase + SM_BAR + 128;

    int64_t n = n_max;
    if (cidx) { const int64_t c = *count; n = c < n_max ? c : n_max; }
    const int64_t num_tiles = ((n + 255) / 256) * 2;     // the backward kernels walk whole 256-point pairs of tiles
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int c = 0; c < 4; ++c) mbar_init(bar_actr + 8 * c, 128);
        mbar_init(bar_encr, 128);
        mbar_init(bar_accr, 1); mbar_init(bar_accr + 8, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *(volatile uint32_t*)(sgen + SM_BAR + 128);

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                for (int g = 0; g < NG; ++g) {
                    const uint32_t bytes = g_chunk_bytes(g);
                    for (int kc = 0; kc < g_chunks(g); ++kc, ++it) {
                        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                        mbar_expect_tx(bar_full + 8 * s, bytes);
                        bulk_g2s(sbase + SM_WST + s * 32768u, packed + fwd_chunk_off(g, kc), bytes, bar_full + 8 * s);
                    }
                }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t it = 0, itc = 0;            // weight-slot counter, tile counter of this CTA
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++itc)
                for (int g = 0; g < NG; ++g) {
                    const uint32_t idesc = make_idesc_bf16(128, g_N(g), 0, 0);
                    const uint32_t acc = tmem_base + (uint32_t)(g & 1) * 256u;
                    for (int kc = 0; kc < g_chunks(g); ++kc, ++it) {
                        const bool from_enc = (g == 0) || (g == 4 && kc == 0);
                        const int ac = (g == 4) ? kc - 1 : kc;
                        // the A chunk: the encoding image of this tile, or 64 columns of layer g-1's output -- the
                        // (itc*9 + g-1)-th completion of that chunk's barrier (nine image-producing layers per tile)
                        if (from_enc) { if (g == 0) mbar_wait(bar_encr, itc & 1u); }
                        else mbar_wait(bar_actr + 8 * ac, (itc * 9u + (uint32_t)(g - 1)) & 1u);
                        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                        mbar_wait(bar_full + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t wb = sbase + SM_WST + s * 32768u;
                        const uint32_t ab = from_enc ? (sbase + SM_ENC) : (sbase + SM_ACT + ac * 16384u);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma(acc, make_desc(ab + k * 32u, 16, 1024), make_desc(wb + k * 32u, 16, 1024), idesc, 1u);   // starts from the bias
                        umma_commit(bar_empty + 8 * s);
                    }
                    umma_commit(bar_accr + 8 * (g & 1));
                }
        }
    } else {
        // ------------------------------------------------------------ epilogue (2 threads = 1 row: column halves)
        const int e = threadIdx.x - 64;
        const int ch = e >> 7;                       // column half: blocks 4*ch .. 4*ch+3, activation chunks 2*ch, 2*ch+1
        const int q = warp & 3;                      // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const uint32_t sw = (uint32_t)(row & 7);
        uint8_t* act_row = sgen + SM_ACT + row * 128;
        uint8_t* enc_row = sgen + SM_ENC + row * 128;
        const uint32_t act_s = sbase + SM_ACT;
        const uint32_t enc_s = sbase + SM_ENC;
        const uint32_t tm_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const float* small = (const float*)(packed + SMALL_OFF);

        {   // accumulator 0 starts from b_0, accumulator 1 from b_1 (later: re-initialised while draining)
            uint32_t b0[32];
#pragma unroll 1
            for (int a = 0; a < 2; ++a)
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    const int blk = ch * 4 + j;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const uint4 b4 = __ldg((const uint4*)(small + SM_BIAS + a * 256 + blk * 32) + c4);
                        b0[4 * c4] = b4.x; b0[4 * c4 + 1] = b4.y; b0[4 * c4 + 2] = b4.z; b0[4 * c4 + 3] = b4.w;
                    }
                    tmem_st32(tm_lane + a * 256 + blk * 32, b0);
                }
            tmem_st_wait();
            tc_fence_before();
        }

        uint32_t itc = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++itc) {
            const int64_t p = tile * 128 + row;
            const bool in = p < n;
            const int64_t id = in ? (cidx ? (int64_t)cidx[p] : p) : 0;
            uint8_t* st_tile = TRAIN ? stash + tile * ST_TILE : nullptr;
            if (ch == 0) {   // positional encoding -> bf16 K-major image (64 columns, last one zero), by the first column half
                float x[3] = {0.f, 0.f, 0.f};
                if (in) { x[0] = xyz_cano[id * 3]; x[1] = xyz_cano[id * 3 + 1]; x[2] = xyz_cano[id * 3 + 2]; }
                float ev[64];
                float s[3], c[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) { ev[a] = x[a]; sincosf(x[a], &s[a], &c[a]); }
#pragma unroll
                for (int k = 0; k < 10; ++k) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        ev[3 + 6 * k + a] = s[a]; ev[6 + 6 * k + a] = c[a];
                        const float s2 = 2.f * s[a] * c[a], c2 = 1.f - 2.f * s[a] * s[a];
                        s[a] = s2; c[a] = c2;
                    }
                }
                ev[63] = 0.f;
                if (TRAIN) {      // this warp's stash stores of the previous tile must have drained its rows
                    if (lane == 0) bulk_wait_read0();
                    __syncwarp();
                }
#pragma unroll
                for (uint32_t u = 0; u < 8; ++u) {
                    uint4 v;
                    v.x = pack_bf16(ev[8 * u], ev[8 * u + 1]); v.y = pack_bf16(ev[8 * u + 2], ev[8 * u + 3]);
                    v.z = pack_bf16(ev[8 * u + 4], ev[8 * u + 5]); v.w = pack_bf16(ev[8 * u + 6], ev[8 * u + 7]);
                    *(uint4*)(enc_row + ((u ^ sw) << 4)) = v;
                }
                fence_proxy_async();
                if (TRAIN) {
                    __syncwarp();
                    if (lane == 0) { bulk_s2g(st_tile + ST_ENC + q * 4096, enc_s + q * 4096u, 4096); bulk_commit(); }
                }
                mbar_arrive(bar_encr);
            }

            for (int g = 0; g < NG; ++g) {
                const uint32_t tm = tm_lane + (uint32_t)(g & 1) * 256u;
                const int nld = g < 8 ? 8 : (g == 8 ? 5 : 1);        // 32-column blocks of this layer's output
                // bias of the layer that writes this accumulator next: g+2, wrapping into the next tile
                const int bl = (g + 2) % NG;
                const float bmine = __ldg(small + SM_BIAS + bl * 256 + ch * 128 + (e & 127));
                // staging buffer alternates with the layer parity: a thread that is one layer ahead (it has passed the
                // previous layer's named barrier, which every thread of the half reached after its reads two layers
                // back) never overwrites values a slower thread of the half is still reading
                uint8_t* bias_s = sgen + SM_BIASBUF + ch * 1024 + (g & 1) * 512;
                // this warp's rows of the previous image are being stored to the stash: drained before they are
                // overwritten below; waited for ahead of the accumulator wait, off the critical path
                if (TRAIN && g > 0 && lane == 0) bulk_wait_read0();
                mbar_wait(bar_accr + 8 * (g & 1), (itc * 5u + (uint32_t)(g >> 1)) & 1u);
                tc_fence_after();
                ((float*)bias_s)[e & 127] = bmine;
                named_bar_sync(1 + ch, 128);
                bool arrived0 = false;                              // this half's first chunk published?
                uint32_t va[32], vb[32];
                if (ch * 4 < nld) tmem_ld32(tm + ch * 128, va);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int blk = ch * 4 + j;
                    uint32_t (&v)[32] = (j & 1) ? vb : va;
                    if (blk < nld) {
                        tmem_ld_wait();
                        if (j < 3 && blk + 1 < nld) tmem_ld32(tm + (blk + 1) * 32, (j & 1) ? va : vb);   // prefetch the next block
                    }
                    {   // bias for the next writer of these columns
                        uint32_t bq[32];
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const uint4 b4 = *((const uint4*)(bias_s + j * 128) + c4);
                            bq[4 * c4] = b4.x; bq[4 * c4 + 1] = b4.y; bq[4 * c4 + 2] = b4.z; bq[4 * c4 + 3] = b4.w;
                        }
                        tmem_st32(tm + blk * 32, bq);
                    }
                    if (blk >= nld) continue;
                    if (g == 9) {                 // rgb head: columns 0..2 (bias already in the accumulator)
                        if (in) {
                            rgb_out[id * 3] = 1.f / (1.f + __expf(-__uint_as_float(v[0])));
                            rgb_out[id * 3 + 1] = 1.f / (1.f + __expf(-__uint_as_float(v[1])));
                            rgb_out[id * 3 + 2] = 1.f / (1.f + __expf(-__uint_as_float(v[2])));
                        }
                        continue;
                    }
                    if (g == 8 && blk == 4) {     // density head: column 128 of the head layer, raw
                        if (in) sigma_out[id] = __uint_as_float(v[0]);
                        continue;
                    }
                    if (TRAIN) {                  // 1-bit ReLU mask from the sign bits (bit (31-c) <-> column c)
                        uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            n0 = __funnelshift_l(v[c], n0, 1); n1 = __funnelshift_l(v[8 + c], n1, 1);
                            n2 = __funnelshift_l(v[16 + c], n2, 1); n3 = __funnelshift_l(v[24 + c], n3, 1);
                        }
                        const uint32_t neg = (((n0 * 256u + n1) * 256u + n2) * 256u) + n3;
                        if (g <= 7) *(uint32_t*)(st_tile + ST_MASK + g * 4096 + blk * 512 + row * 4) = ~neg;
                        else *(uint32_t*)(st_tile + ST_CMASK + blk * 512 + row * 4) = ~neg;
                    }
                    uint32_t w[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) w[k] = pack_relu_bf16(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
                    uint8_t* dst = act_row + (blk >> 1) * 16384;
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u)
                        *(uint4*)(dst + ((((uint32_t)(blk & 1) * 4 + u) ^ sw) << 4)) =
                            make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
                    if (blk & 1) {
                        // this thread's part of the 64-column chunk (blk >> 1) is written: publish it to the MMA issuer
                        // (layer g+1 starts on this K-chunk at once) and, when training, stream the warp's 4 KB row slab
                        fence_proxy_async();
                        if (TRAIN) {
                            __syncwarp();
                            if (lane == 0) {
                                const uint32_t off = (uint32_t)(blk >> 1) * 16384u + (uint32_t)q * 4096u;
                                bulk_s2g(st_tile + (g == 8 ? ST_C : ST_H + (int64_t)g * 65536) + off, act_s + off, 4096);
                                bulk_commit();
                            }
                        }
                        // the half's first chunk is published at once; its second (last) one after the TMEM fence
                        // below, so that "layer g+1's accumulator is complete" implies every thread has finished and
                        // fenced its TMEM reads and bias writes of layer g (layer g+2 accumulates onto them)
                        if (j == 1) { mbar_arrive(bar_actr + 8 * (blk >> 1)); arrived0 = true; }
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                if (g < 9) {                      // (chunks the layer did not produce still complete their phase)
                    if (!arrived0) mbar_arrive(bar_actr + 8 * (2 * ch));
                    mbar_arrive(bar_actr + 8 * (2 * ch + 1));
                }
            }
        }
        if (TRAIN && lane == 0) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

int mlp_fwd_pipe_launch(const void* packed, const float* xyz_cano, const int32_t* cidx, const int32_t* count,
                        int64_t n_max, float* sigma, float* rgb, void* stash, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(mlp_fwd_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    const int64_t tiles = ((n_max + 255) / 256) * 2;
    const int sms = an_num_sms();
    const int grid = (int)(tiles < sms ? tiles : sms);
    if (stash)
        mlp_fwd_pipe_kernel<true><<<grid, THREADS, SM_ALLOC, stream>>>(
            (const uint8_t*)packed, xyz_cano, cidx, count, n_max, sigma, rgb, (uint8_t*)stash);
    else
        mlp_fwd_pipe_kernel<false><<<grid, THREADS, SM_ALLOC, stream>>>(
            (const uint8_t*)packed, xyz_cano, cidx, count, n_max, sigma, rgb, nullptr);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
