// EXPERIMENT (not built into the library; kept as the record of a measured negative result).
// To try it: copy into anim-nerf_b200/csrc/, declare mlp_fwd_ts_launch in mlp_tc.cu and call it from an_mlp_fwd when
// stash == NULL.  Parity-green on the inference tests; 1.79 ms per 2^20 points against 0.81 ms for mlp_tc.cu: with one
// 128-row tile per CTA every CTA re-streams the whole weight set per 128 rows (twice mlp_tc.cu's L2 -> SM traffic, four
// times mlp_bwd.cu's), and a half-layer's MMAs (N = 128) overlap only a quarter of the other half's epilogue.
//
// A9-A11, second organisation of the MLP forward: activations never touch shared memory.
// Reference: models/embedding.py:22-39 + models/nerf.py:129-175.
//
// mlp_tc.cu keeps each tile's activation image in shared memory (the A operand of the next layer's MMAs) and is
// bound by shared-memory bandwidth: per 128-cycle MMA a CTA moves 4 KB A reads + 4 KB B reads + 4 KB weight-ring writes
// + 4 KB epilogue image stores.  Here the A operand lives in TENSOR MEMORY (TS form of tcgen05.mma): the epilogue
// writes the next layer's input with tcgen05.st, the MMAs read it from TMEM, and shared memory carries the weights only
// (2 KB read + 2 KB written per 64-cycle MMA = half the pipe).
//
// Persistent CTA pairs (cta_group::2, M = 256): ONE 128-row tile per CTA and iteration.  TMEM (512 columns):
//   A0 [0,128) / A1 [128,256)   bf16 activations, 256 features = 128 columns; layer g reads A[g&1], its epilogue
//                               writes A[(g+1)&1]
//   ACC0 [256,384) / ACC1 [384,512)   fp32 accumulators of the two N = 128 halves of a layer's output
// A layer runs as two half-layers (output features [0,128) and [128,256)): while the tensor cores run half 1, the
// epilogue drains half 0 (all eight epilogue warps on one half: thread = (row, 64 columns)), and the next layer's
// first K-chunks (input features 0..127 = half 0's output) start while half 1 is being drained.
//   warp 0      weight producer (both CTAs): this CTA's 64 rows of every (half-layer, K-chunk) image / bias slab into
//               a 16-stage ring of 8 KB
//   warp 1      leader: MMA issuer (TS form for hidden-layer inputs, SS form for the encoding chunks and the bias
//               step, whose A operand is the encoding image in shared memory); peer: stage relay
//   warps 2-9   epilogue: tcgen05.ld -> ReLU + bf16 pack -> tcgen05.st into the next layer's A columns
// Heads as in mlp_tc.cu: the fused final+colour layer is half 0 (N = 128) + a 16-wide half 1 whose column 0 is the
// density; the rgb head is a 16-wide GEMM on c.  The bias enters as one K = 16 MMA per half-layer (bias slab).
#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc_common.cuh"

#ifdef AN_MLP_TRACE
// debug timeline (tools/trace_ts.py): CTA 0, one region of 64 Ki entries per role, plain stores
__device__ unsigned long long* g_trace_ts = nullptr;
#define TRACE_DECL unsigned int tr_n = 0
#define TRACE(role, ev, a, b)                                                                                      \
    do {                                                                                                           \
        if (blockIdx.x == 0 && g_trace_ts && tr_n < 65535u) {                                                      \
            g_trace_ts[(role) * 65536 + 1 + tr_n] = ((unsigned long long)clock64() << 24) | ((unsigned long long)(role) << 20) | ((ev) << 16) | ((a) << 8) | (b); \
            g_trace_ts[(role) * 65536] = ++tr_n;                                                                   \
        }                                                                                                          \
    } while (0)
#else
#define TRACE_DECL
#define TRACE(role, ev, a, b)
#endif

namespace {

constexpr int THREADS = 320;
constexpr int NST = 16;                              // ring stages
constexpr uint32_t STG = 8192;                       // this CTA's 64 rows x 128 B of a chunk image
constexpr uint32_t SM_ENC = 0;                       // [2 iterations][128 rows x 128 B]
constexpr uint32_t SM_WST = 32768;                   // [NST][8 KB]
constexpr uint32_t SM_BAR = SM_WST + NST * STG;      // 163840
constexpr uint32_t SM_BYTES = SM_BAR + 512;
constexpr uint32_t SM_ALLOC = SM_BYTES + 1024;

constexpr uint32_t TM_A = 0, TM_ACC = 256;

__device__ __forceinline__ void umma_pair_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// half-layers of GEMM layer g and their output rows: g < 8: two halves of 128; g = 8: 128 (colour features) + 16
// (density + pad); g = 9: 16 (rgb + pad)
__device__ __forceinline__ int n_halves(int g) { return g == 9 ? 1 : 2; }
__device__ __forceinline__ int half_rows(int g, int h) { return g < 8 ? 128 : (g == 8 ? (h == 0 ? 128 : 16) : 16); }
__device__ __forceinline__ int half_row0(int g, int h) { return h == 0 ? 0 : 128; }
__device__ __forceinline__ bool half_has_bias(int g, int h) { return g < 8 || (g == 8 && h == 0); }

}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
mlp_fwd_ts_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ xyz_cano,
                  const int32_t* __restrict__ cidx, const int32_t* __restrict__ count, int64_t n_max,
                  float* __restrict__ sigma_out, float* __restrict__ rgb_out)
{
    using namespace mlp;
    using namespace tc;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    const uint32_t bar_full = sbase + SM_BAR;              // [NST]
    const uint32_t bar_empty = sbase + SM_BAR + 128;       // [NST]
    const uint32_t bar_acc = sbase + SM_BAR + 256;         // [2] MMA -> epilogue: half h's accumulators complete
    const uint32_t bar_a = sbase + SM_BAR + 272;           // [2] epilogue (both CTAs) -> MMA: half h drained + its A columns written
    const uint32_t bar_enc = sbase + SM_BAR + 288;         // epilogue (both CTAs) -> MMA: encoding image written, accumulators free
    const uint32_t tmem_slot = sbase + SM_BAR + 304;

    int64_t n = n_max;
    if (cidx) { const int64_t c = *count; n = c < n_max ? c : n_max; }
    const int64_t num_iters = (n + 255) / 256;             // 128 rows per CTA and iteration

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(bar_full + 8 * s, rank == 0 ? 2 : 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int h = 0; h < 2; ++h) { mbar_init(bar_acc + 8 * h, 1); mbar_init(bar_a + 8 * h, 512); }
        mbar_init(bar_enc, 512);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *(volatile uint32_t*)(sgen + SM_BAR + 304);

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer
        if (lane == 0) {
            uint32_t it = 0;
            TRACE_DECL;
            for (int64_t iter = pair; iter < num_iters; iter += npairs)
                for (int g = 0; g < NG; ++g)
                    for (int h = 0; h < n_halves(g); ++h) {
                        const uint32_t rows = (uint32_t)half_rows(g, h) >> 1;                   // this CTA's rows of the half
                        const uint32_t r0 = (uint32_t)half_row0(g, h) + rank * rows;
                        const int ns = g_chunks(g) + (half_has_bias(g, h) ? 1 : 0);
                        for (int kc = 0; kc < ns; ++kc, ++it) {
                            const bool slab = kc == g_chunks(g);
                            const uint32_t bytes = rows * (slab ? 32u : 128u);
                            const uint8_t* src = packed + (slab ? fwd_bias_off(g) + r0 * 32u : fwd_chunk_off(g, kc) + r0 * 128u);
                            const uint32_t s = it % NST, ph = (it / NST) & 1u;
                            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                            mbar_expect_tx(bar_full + 8 * s, bytes);
                            bulk_g2s(sbase + SM_WST + s * STG, src, bytes, bar_full + 8 * s);
                            TRACE(0, 0, g, h * 8 + kc);
                        }
                    }
        }
    } else if (warp == 1 && rank != 0) {
        // ------------------------------------------------------------ peer: stage relay
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t iter = pair; iter < num_iters; iter += npairs)
                for (int g = 0; g < NG; ++g)
                    for (int h = 0; h < n_halves(g); ++h) {
                        const int ns = g_chunks(g) + (half_has_bias(g, h) ? 1 : 0);
                        for (int kc = 0; kc < ns; ++kc, ++it) {
                            const uint32_t s = it % NST, ph = (it / NST) & 1u;
                            mbar_wait(bar_full + 8 * s, ph);
                            mbar_arrive_remote(bar_full + 8 * s, 0);
                        }
                    }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ leader: MMA issuer
        // (whole warp runs the loop so the operands stay warp-uniform; one elected lane issues, see tc::elect_one)
        {
            TRACE_DECL;
            const bool tr = lane == 0; (void)tr;
            uint32_t it = 0, enc_phase = 0, a_phase = 0;        // a_phase: bit h = parity of bar_a[h]
            for (int64_t iter = pair, li = 0; iter < num_iters; iter += npairs, ++li) {
                const uint32_t enc_s = sbase + SM_ENC + (uint32_t)(li & 1) * 16384u;
                if (tr) TRACE(1, 5, 0, 0);
                mbar_wait(bar_enc, enc_phase); enc_phase ^= 1u;      // encoding image ready, both accumulators drained
                if (tr) TRACE(1, 6, 0, 0);
                tc_fence_after();
                for (int g = 0; g < NG; ++g) {
                    const uint32_t a_cur = tmem_base + TM_A + (uint32_t)(g & 1) * 128u;
                    uint32_t got = 0;                                // bit h: layer g-1's bar_a[h] observed
                    for (int h = 0; h < n_halves(g); ++h) {
                        const uint32_t idesc = make_idesc_bf16(256, half_rows(g, h), 0, 0);
                        const uint32_t acc = tmem_base + TM_ACC + (uint32_t)h * 128u;
                        const int nc = g_chunks(g);
                        for (int kc = 0; kc < nc; ++kc, ++it) {
                            const bool from_enc = (g == 0) || (g == 4 && kc == 0);
                            const int fc = (g == 4) ? kc - 1 : kc;             // 64-feature chunk of the previous layer's output
                            if (!from_enc) {
                                const int ph_ = fc >> 1;                         // produced by half fc/2 of layer g-1
                                if (!((got >> ph_) & 1u)) {
                                    if (tr) TRACE(1, 0, g, ph_);
                                    mbar_wait(bar_a + 8 * ph_, (a_phase >> ph_) & 1u); a_phase ^= 1u << ph_; got |= 1u << ph_;
                                    tc_fence_after();
                                    if (tr) TRACE(1, 1, g, ph_);
                                }
                            }
                            const uint32_t s = it % NST, ph = (it / NST) & 1u;
                            if (tr) TRACE(1, 3, g, h * 8 + kc);
                            mbar_wait(bar_full + 8 * s, ph);
                            tc_fence_after();
                            if (tr) TRACE(1, 2, g, h * 8 + kc);
                            const uint32_t wb = sbase + SM_WST + s * STG;
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint32_t accum = (kc > 0 || k > 0) ? 1u : 0u;
                                    if (from_enc)
                                        umma_pair(acc, make_desc(enc_s + k * 32u, 16, 1024), make_desc(wb + k * 32u, 16, 1024), idesc, accum);
                                    else
                                        umma_pair_ts(acc, a_cur + (uint32_t)fc * 32u + (uint32_t)k * 8u, make_desc(wb + k * 32u, 16, 1024), idesc, accum);
                                }
                                umma_commit_pair(bar_empty + 8 * s);
                            }
                            __syncwarp();
                        }
                        if (half_has_bias(g, h)) {
                            const uint32_t s = it % NST, ph = (it / NST) & 1u;
                            mbar_wait(bar_full + 8 * s, ph);
                            tc_fence_after();
                            if (elect_one()) {
                                umma_pair(acc, make_desc(enc_s + 96u, 16, 1024), make_desc_noswz(sbase + SM_WST + s * STG, 128, 256), idesc, 1u);
                                umma_commit_pair(bar_empty + 8 * s);
                            }
                            __syncwarp();
                            ++it;
                        }
                        if (elect_one()) umma_commit_pair(bar_acc + 8 * h);
                        __syncwarp();
                        if (tr) TRACE(1, 4, g, h);
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: thread = (row, 64 columns of the half)
        const int q = warp & 3, cg = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const float* small = (const float*)(packed + SMALL_OFF);
        const float b_sigma = __ldg(small + SM_BIAS + 8 * 256 + 128);
        const float b_rgb0 = __ldg(small + SM_BIAS + 9 * 256), b_rgb1 = __ldg(small + SM_BIAS + 9 * 256 + 1),
                    b_rgb2 = __ldg(small + SM_BIAS + 9 * 256 + 2);
        const bool tracer = lane == 0 && (warp == 2 || warp == 6);
        const int trole = 2 + cg;
        TRACE_DECL;
        uint32_t acc_phase = 0;                       // bit h = parity of bar_acc[h]

        for (int64_t iter = pair, li = 0; iter < num_iters; iter += npairs, ++li) {
            const int64_t p = (iter * 2 + rank) * 128 + row;
            const bool in = p < n;
            const int64_t id = in ? (cidx ? (int64_t)cidx[p] : p) : 0;
            if (cg == 0) {      // positional encoding -> bf16 K-major image (64 columns; the last one multiplies the bias slabs)
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
                ev[63] = 1.f;
                uint8_t* enc_row = sgen + SM_ENC + (li & 1) * 16384 + row * 128;
#pragma unroll
                for (uint32_t u = 0; u < 8; ++u) {
                    uint4 v;
                    v.x = pack_bf16(ev[8 * u], ev[8 * u + 1]); v.y = pack_bf16(ev[8 * u + 2], ev[8 * u + 3]);
                    v.z = pack_bf16(ev[8 * u + 4], ev[8 * u + 5]); v.w = pack_bf16(ev[8 * u + 6], ev[8 * u + 7]);
                    *(uint4*)(enc_row + ((u ^ sw) << 4)) = v;
                }
                fence_proxy_async();
            }
            tc_fence_before();                       // this thread's TMEM reads of the previous iteration are done
            mbar_arrive_remote(bar_enc, 0);

            for (int g = 0; g < NG; ++g) {
                const uint32_t a_nxt = tlane + TM_A + (uint32_t)((g + 1) & 1) * 128u;
                for (int h = 0; h < n_halves(g); ++h) {
                    if (tracer) TRACE(trole, 0, g, h);
                    mbar_wait(bar_acc + 8 * h, (acc_phase >> h) & 1u); acc_phase ^= 1u << h;
                    tc_fence_after();
                    if (tracer) TRACE(trole, 1, g, h);
                    const uint32_t acc = tlane + TM_ACC + (uint32_t)h * 128u;
                    if (g == 9) {                    // rgb head: columns 0..2
                        if (cg == 0) {
                            uint32_t v[32];
                            tmem_ld32(acc, v); tmem_ld_wait();
                            if (in) {
                                rgb_out[id * 3] = 1.f / (1.f + __expf(-(__uint_as_float(v[0]) + b_rgb0)));
                                rgb_out[id * 3 + 1] = 1.f / (1.f + __expf(-(__uint_as_float(v[1]) + b_rgb1)));
                                rgb_out[id * 3 + 2] = 1.f / (1.f + __expf(-(__uint_as_float(v[2]) + b_rgb2)));
                            }
                        }
                        continue;                    // the next iteration's bar_enc arrive covers "accumulators drained"
                    }
                    if (g == 8 && h == 1) {          // density: column 0 of the 16-wide half
                        if (cg == 0) {
                            uint32_t v[32];
                            tmem_ld32(acc, v); tmem_ld_wait();
                            if (in) sigma_out[id] = __uint_as_float(v[0]) + b_sigma;
                        }
                        continue;
                    }
                    uint32_t va[32], vb[32], w[32];
                    tmem_ld32(acc + (uint32_t)cg * 64u, va);
                    tmem_ld32(acc + (uint32_t)cg * 64u + 32u, vb);
                    tmem_ld_wait();
                    if (tracer) TRACE(trole, 2, g, h);
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        w[k] = pack_relu_bf16(__uint_as_float(va[2 * k]), __uint_as_float(va[2 * k + 1]));
                        w[16 + k] = pack_relu_bf16(__uint_as_float(vb[2 * k]), __uint_as_float(vb[2 * k + 1]));
                    }
                    // features [128 h + 64 cg, +64) of the next layer's input = 32 packed columns
                    tmem_st32(a_nxt + (uint32_t)h * 64u + (uint32_t)cg * 32u, w);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive_remote(bar_a + 8 * h, 0);
                    if (tracer) TRACE(trole, 3, g, h);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

#ifdef AN_MLP_TRACE
extern "C" int an_debug_trace_ts(void* buf) { return (int)cudaMemcpyToSymbol(g_trace_ts, &buf, sizeof(buf)); }
#endif

int mlp_fwd_ts_launch(const void* packed, const float* xyz_cano, const int32_t* cidx, const int32_t* count,
                      int64_t n_max, float* sigma, float* rgb, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    const int64_t iters = (n_max + 255) / 256;
    const int pairs = an_num_sms() / 2;
    const int grid = 2 * (int)(iters < pairs ? iters : pairs);
    mlp_fwd_ts_kernel<<<grid, THREADS, SM_ALLOC, stream>>>((const uint8_t*)packed, xyz_cano, cidx, count, n_max, sigma, rgb);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
