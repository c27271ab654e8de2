// A9-A11: positional encoding + 8x256 NeRF MLP forward on the 5th-gen tensor cores
// (tcgen05.mma.cta_group::2, accumulators in TMEM, weights streamed by TMA bulk copies, activations
// never leave the SM).  Reference: models/embedding.py:22-39 + models/nerf.py:129-175.
//
// Persistent CTA PAIRS (cluster of 2 = the two SMs of a TPC), one pair per TPC.  A pair iteration owns
// 512 compacted (valid) points: CTA rank r holds two 128-row tiles (points r*256 .. r*256+255).  Every
// MMA is a cta_group::2 instruction with M = 256: tile t of both CTAs against one weight chunk, of which
// each CTA stages only its half of the rows (N/2) -- per SM that halves the B-operand shared-memory
// reads and the L2 -> shared-memory weight stream compared with one CTA per tile.  The two tiles of a
// CTA ping-pong: the tensor cores run layer g of tile t (both CTAs) while the epilogue of tile 1-t runs
// on the CUDA cores.
//   warp 0      weight producer (both CTAs): cp.async.bulk of this CTA's half of every pre-swizzled
//               64-wide K-chunk image / bias slab into a 4-stage ring of 16 KB (full/empty mbarriers);
//               runs ahead across layers and tiles.
//   warp 1      leader CTA: MMA issuer -- one elected thread issues tcgen05.mma (M=256, N=256|144|16,
//               K=16), 4 K-steps per chunk; tcgen05.commit (multicast to both CTAs) releases the stage /
//               publishes the tile's accumulators.  Peer CTA: relays "my half of the stage has landed"
//               to the leader's full barrier (a bulk copy can only signal an mbarrier of its own CTA).
//               Owns the TMEM allocation (512 columns = 2 x 128x256 fp32 per CTA).
//   warps 2-9   epilogue, one thread per row: tcgen05.ld 32 columns at a time, ReLU + bf16 pack in one
//               cvt, st.shared into the K-major SWIZZLE_128B image that is the next layer's A operand
//               (in place: the layer's MMAs have completed); the peer's threads arrive on the leader's
//               barrier through the cluster.  The bias is NOT added here: it enters as one more K = 16
//               MMA per layer against the encoding image's K-step 3, whose pad column holds 1.0
//               (mlp_layout.cuh: bias slab) -- no per-thread bias loads, no accumulator re-initialisation.
//               Both heads run on the tensor cores too: the density is one more output column of the
//               fused final+colour layer, the rgb head a 16-wide GEMM; their biases are added in fp32.
//               The encoding (sin/cos by double-angle recurrence from one sincosf per coordinate;
//               inputs stay fp32 until after the encoding) is the prologue.
// Training mode (stash != NULL) additionally streams every activation image to HBM with TMA
// bulk stores plus 1-bit ReLU masks; the backward kernels consume them (mlp_bwd.cu).
//
// Roofline: tensor-bound.  1 179 904 algorithmic FLOP per point (the reference's 12 linears; the
// fused head layer executes 1 118 208 of them, the bias steps add 9 x 8192); per 512-point iteration
// the pair issues (36 chunks x 4 + 9) x 2 tiles MMAs of M = 256.  Algorithmic HBM bytes per point:
// 16 B in (id + xyz), 16 B out.
#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc_common.cuh"

namespace {

constexpr int THREADS = 320;
constexpr int NSTAGE = 4;
constexpr uint32_t STAGE_BYTES = 16384;            // this CTA's half of a 64-wide chunk of a 256-row weight image
constexpr uint32_t SM_ACT = 0;                     // [2 tiles][4 chunks][128 rows x 128 B]
constexpr uint32_t SM_ENC = 131072;                // [2 tiles][128 rows x 128 B]
constexpr uint32_t SM_WST = 163840;                // [NSTAGE][16 KB]
constexpr uint32_t SM_BAR = SM_WST + NSTAGE * STAGE_BYTES;   // 229376
constexpr uint32_t SM_BYTES = SM_BAR + 128;
constexpr uint32_t SM_ALLOC = SM_BYTES + 1024;     // slack for manual 1024-B alignment
static_assert(SM_ALLOC <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");

#ifdef AN_MLP_TRACE
// debug timeline (tools/trace_mlp.py): CTA 0, one region of 64 Ki entries per role, plain stores
// (clock << 24 | role << 20 | ev << 16 | a << 8 | b); entry 0 of a region = its event count
__device__ unsigned long long* g_trace = nullptr;
#define TRACE_DECL unsigned int tr_n = 0
#define TRACE(role, ev, a, b)                                                                                      \
    do {                                                                                                           \
        if (blockIdx.x == 0 && g_trace && tr_n < 65535u) {                                                         \
            g_trace[(role) * 65536 + 1 + tr_n] = ((unsigned long long)clock64() << 24) | ((unsigned long long)(role) << 20) | ((ev) << 16) | ((a) << 8) | (b); \
            g_trace[(role) * 65536] = ++tr_n;                                                                      \
        }                                                                                                          \
    } while (0)
#else
#define TRACE_DECL
#define TRACE(role, ev, a, b)
#endif

// ring steps of GEMM layer g: its weight chunks, then (g < 9) the bias slab
__device__ __forceinline__ int ring_steps(int g) { return mlp::g_chunks(g) + (g < 9 ? 1 : 0); }

}  // namespace

// MODE 0 = inference, 1 = training (activation stash), 2 = tangent (forward-mode derivative of the
// trunk: tau_l = m_l * (W_l tau_{l-1}), no biases, ReLU masks m_l read from the primal stash `pstash`,
// input tau_0 = d enc(x) . tvec; writes the tau images to `stash` in the activation layout -- the weight
// gradient of a loss on d sigma/d xyz is then the ordinary wgrad kernel on (tau images, dY images),
// see an_mlp_fwd_tangent).  The tangent runs the 8 trunk layers + the fused head layer (for tau_sigma).
// With tscale (a per-point scalar c) the same pass produces T_l = tau_l + c * X_l (X = the primal activations):
// T_l = m_l * (W_l T_{l-1} + c b_l), T_0 = tau_0 + c enc(x) -- the bias column of the encoding image holds c
// instead of 1, so the bias step adds c * bias.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
mlp_fwd_tc_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ xyz_cano,
                  const int32_t* __restrict__ cidx, const int32_t* __restrict__ count, int64_t n_max,
                  float* __restrict__ sigma_out, float* __restrict__ rgb_out, uint8_t* __restrict__ stash,
                  const float* __restrict__ tvec, const uint8_t* __restrict__ pstash, const float* __restrict__ tscale)
{
    using namespace mlp;
    using namespace tc;
    constexpr bool TRAIN = MODE >= 1;          // writes the image stash
    constexpr bool TAN = MODE == 2;
    constexpr int NGT = TAN ? 9 : NG;          // layer steps per iteration
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();             // 0 = leader (issues the MMAs), 1 = peer
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    const uint32_t bar_full = sbase + SM_BAR;            // [NSTAGE]  TMA (+ peer relay) -> MMA
    const uint32_t bar_empty = sbase + SM_BAR + 32;      // [NSTAGE]  MMA -> TMA (both CTAs)
    const uint32_t bar_act = sbase + SM_BAR + 64;        // [2 tiles] epilogue of both CTAs -> MMA (leader's copy: 256 arrivals)
    const uint32_t bar_acc = sbase + SM_BAR + 80;        // [2 tiles] MMA -> epilogue (tcgen05.commit, both CTAs)
    const uint32_t tmem_slot = sbase + SM_BAR + 96;

    int64_t n = n_max;
    if (cidx) { const int64_t c = *count; n = c < n_max ? c : n_max; }
    const int64_t num_iters = (n + PAIR_POINTS - 1) / PAIR_POINTS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, rank == 0 ? 2 : 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(bar_act + 8 * t, 256); mbar_init(bar_acc + 8 * t, 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();          // barriers of both CTAs initialised and TMEM allocated before any remote arrive / MMA
    tc_fence_after();
    const uint32_t tmem_base = *(volatile uint32_t*)(sgen + SM_BAR + 96);

    if (warp == 0) {
        // ------------------------------------------------------------ weight producer (this CTA's half of every B operand)
        {   // whole warp runs the loop, one elected lane issues (uniform operands, see tc::elect_one)
            TRACE_DECL;
            uint32_t it = 0;
            const uint64_t keep = l2_policy_keep();
            for (int64_t iter = pair; iter < num_iters; iter += npairs)
                for (int g = 0; g < NGT; ++g) {
                    const int nc = g_chunks(g), ns = ring_steps(g);
                    for (int t = 0; t < 2; ++t)
                        for (int kc = 0; kc < ns; ++kc, ++it) {
                            const uint32_t bytes = (kc < nc ? g_chunk_bytes(g) : g_bias_bytes(g)) >> 1;
                            const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                            if (lane == 0) TRACE(0, 0, g, t * 8 + kc);
                            if (elect_one()) {
                                mbar_expect_tx(bar_full + 8 * s, bytes);
                                bulk_g2s_hint(sbase + SM_WST + s * STAGE_BYTES, packed + fwd_chunk_off(g, kc) + rank * bytes, bytes, bar_full + 8 * s, keep);
                            }
                            __syncwarp();
                        }
                }
        }
    } else if (warp == 1 && rank != 0) {
        // ------------------------------------------------------------ peer: tell the leader when this CTA's half of a stage has landed
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t iter = pair; iter < num_iters; iter += npairs)
                for (int g = 0; g < NGT; ++g) {
                    const int ns = 2 * ring_steps(g);
                    for (int i = 0; i < ns; ++i, ++it) {
                        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                        mbar_wait(bar_full + 8 * s, ph);
                        mbar_arrive_remote(bar_full + 8 * s, 0);
                    }
                }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ leader: MMA issuer for both CTAs
        // (whole warp runs the loop so the operands stay warp-uniform; one elected lane issues, see elect_one())
        {
            TRACE_DECL;
            const bool tr = lane == 0; (void)tr;
            uint32_t it = 0, act_phase = 0;
            for (int64_t iter = pair; iter < num_iters; iter += npairs)
                for (int g = 0; g < NGT; ++g) {
                    const uint32_t idesc = make_idesc_bf16(256, g_N(g), 0, 0);
                    const int nc = g_chunks(g), ns = ring_steps(g);
                    for (int t = 0; t < 2; ++t) {
                        if (tr) TRACE(1, 0, g, t);
                        mbar_wait(bar_act + 8 * t, act_phase);
                        if (tr) TRACE(1, 1, g, t);
                        tc_fence_after();
                        for (int kc = 0; kc < ns; ++kc, ++it) {
                            const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                            if (tr) TRACE(1, 3, g, t * 8 + kc);
                            mbar_wait(bar_full + 8 * s, ph);
                            if (tr) TRACE(1, 2, g, t * 8 + kc);
                            tc_fence_after();
                            const uint32_t wb = sbase + SM_WST + s * STAGE_BYTES;
                            if (kc < nc) {
                                const bool from_enc = (g == 0) || (g == 4 && kc == 0);
                                const int ac = (g == 4) ? kc - 1 : kc;
                                const uint32_t ab = from_enc ? (sbase + SM_ENC + t * 16384u)
                                                             : (sbase + SM_ACT + t * 65536u + ac * 16384u);
                                if (elect_one()) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        umma_pair(tmem_base + t * 256u, make_desc(ab + k * 32u, 16, 1024),
                                                  make_desc(wb + k * 32u, 16, 1024), idesc, (kc > 0 || k > 0) ? 1u : 0u);
                                    umma_commit_pair(bar_empty + 8 * s);   // stage free (in both CTAs) once these MMAs retire
                                }
                            } else if (elect_one()) {
                                // bias step: A = K-step 3 of the encoding image (column 63 = 1), B = the slab (k = 15 = bias)
                                umma_pair(tmem_base + t * 256u, make_desc(sbase + SM_ENC + t * 16384u + 96u, 16, 1024),
                                          make_desc_noswz(wb, 128, 256), idesc, 1u);
                                umma_commit_pair(bar_empty + 8 * s);
                            }
                            __syncwarp();
                        }
                        if (elect_one()) umma_commit_pair(bar_acc + 8 * t);   // accumulators of (layer g, tile t) complete in both CTAs
                        __syncwarp();
                    }
                    act_phase ^= 1u;
                }
        }
    } else {
        // ------------------------------------------------------------ epilogue (1 thread = 1 row)
        const int e = threadIdx.x - 64;
        const int t = e >> 7;
        const int q = warp & 3;                      // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const uint32_t sw = (uint32_t)(row & 7);
        uint8_t* act_row = sgen + SM_ACT + t * 65536 + row * 128;
        uint8_t* enc_row = sgen + SM_ENC + t * 16384 + row * 128;
        const uint32_t act_s = sbase + SM_ACT + t * 65536u;
        const uint32_t enc_s = sbase + SM_ENC + t * 16384u;
        const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
        const uint32_t my_act = bar_act + 8 * t, my_acc = bar_acc + 8 * t;
        const float* small = (const float*)(packed + SMALL_OFF);
        const uint64_t stream_pol = l2_policy_stream(); (void)stream_pol;
        const float b_sigma = __ldg(small + SM_BIAS + 8 * 256 + 128);
        const float b_rgb0 = __ldg(small + SM_BIAS + 9 * 256), b_rgb1 = __ldg(small + SM_BIAS + 9 * 256 + 1),
                    b_rgb2 = __ldg(small + SM_BIAS + 9 * 256 + 2);
        const bool leader = (e & 127) == 0;          // trace only
        uint32_t acc_phase = 0;
        TRACE_DECL;

        for (int64_t iter = pair; iter < num_iters; iter += npairs) {
            const int64_t tile = (iter * 2 + rank) * 2 + t;          // = compact point index / 128
            const int64_t p = tile * 128 + row;
            const bool in = p < n;
            const int64_t id = in ? (cidx ? (int64_t)cidx[p] : p) : 0;
            uint8_t* st_tile = TRAIN ? stash + tile * ST_TILE : nullptr;
            const uint8_t* pst_tile = TAN ? pstash + tile * ST_TILE : nullptr;
            float x[3] = {0.f, 0.f, 0.f};
            float tv[3] = {0.f, 0.f, 0.f};
            if (in) { x[0] = xyz_cano[id * 3]; x[1] = xyz_cano[id * 3 + 1]; x[2] = xyz_cano[id * 3 + 2]; }
            float tsc = 0.f;           // tangent mode: weight of the primal activations in the images (0: pure tangent)
            if (TAN && in) {
                tv[0] = tvec[id * 3]; tv[1] = tvec[id * 3 + 1]; tv[2] = tvec[id * 3 + 2];
                if (tscale) tsc = tscale[id];
            }
            {   // positional encoding -> bf16 K-major image (64 columns; the last one multiplies the bias slabs)
                float ev[64];
                float s[3], c[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) { ev[a] = TAN ? tv[a] + tsc * x[a] : x[a]; sincosf(x[a], &s[a], &c[a]); }
                float fk = 1.f;
#pragma unroll
                for (int k = 0; k < 10; ++k) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        if (TAN) {     // d/dx [sin(2^k x), cos(2^k x)] . tv
                            ev[3 + 6 * k + a] = fk * c[a] * tv[a] + tsc * s[a]; ev[6 + 6 * k + a] = -fk * s[a] * tv[a] + tsc * c[a];
                        } else {
                            ev[3 + 6 * k + a] = s[a]; ev[6 + 6 * k + a] = c[a];
                        }
                        const float s2 = 2.f * s[a] * c[a], c2 = 1.f - 2.f * s[a] * s[a];
                        s[a] = s2; c[a] = c2;
                    }
                    fk *= 2.f;
                }
                ev[63] = TAN ? tsc : 1.f;
                if (TRAIN) {      // this warp's TMA stores of the previous iteration must have drained its rows
                    if (elect_one()) bulk_wait_read0();
                    __syncwarp();
                }
#pragma unroll
                for (uint32_t u = 0; u < 8; ++u) {
                    uint4 v;
                    v.x = pack_bf16(ev[8 * u], ev[8 * u + 1]); v.y = pack_bf16(ev[8 * u + 2], ev[8 * u + 3]);
                    v.z = pack_bf16(ev[8 * u + 4], ev[8 * u + 5]); v.w = pack_bf16(ev[8 * u + 6], ev[8 * u + 7]);
                    *(uint4*)(enc_row + ((u ^ sw) << 4)) = v;
                }
            }
            fence_proxy_async();
            if (TRAIN) {          // every warp streams its own 32 rows (4 KB, contiguous in the image) to the stash
                __syncwarp();
                if (elect_one()) { bulk_s2g_hint(st_tile + ST_ENC + q * 4096, enc_s + q * 4096u, 4096, stream_pol); bulk_commit(); }
            }
            mbar_arrive_remote(my_act, 0);

            for (int g = 0; g < NGT; ++g) {
                const int nld = g < 8 ? 8 : (g == 8 ? 5 : 1);      // accumulator blocks (32 columns) to drain
                uint32_t pmw[8];               // tangent: the primal's ReLU masks of this layer ([block][row] words)
                if (TAN) {
#pragma unroll
                    for (int cb = 0; cb < 8; ++cb)
                        pmw[cb] = g <= 7 ? __ldg((const uint32_t*)(pst_tile + ST_MASK + g * 4096 + cb * 512 + row * 4)) : 0xffffffffu;
                }
                if (leader) TRACE(2 + t, 0, g, 0);
                // this warp's rows of the layer g-1 image are being stored: those bulk stores must have drained before
                // the rows are overwritten below (after the accumulator wait); waited for here, off the critical path
                if (TRAIN && g > 0) {
                    if (elect_one()) bulk_wait_read0();
                    __syncwarp();
                }
                mbar_wait(my_acc, acc_phase); acc_phase ^= 1u;
                if (leader) TRACE(2 + t, 1, g, 0);
                tc_fence_after();
                if (leader) TRACE(2 + t, 2, g, 0);
                uint32_t va[32], vb[32];
                tmem_ld32(tm, va);
#pragma unroll 1
                for (int cb = 0; cb < nld; cb += 2) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int blk = cb + half;
                        if (blk >= nld) break;
                        uint32_t (&v)[32] = half ? vb : va;
                        tmem_ld_wait();
                        if (blk + 1 < nld) tmem_ld32(tm + (blk + 1) * 32, half ? va : vb);   // prefetch the next block
                        if (!TAN && g == 9) {         // rgb head: columns 0..2
                            if (in) {
                                rgb_out[id * 3] = 1.f / (1.f + __expf(-(__uint_as_float(v[0]) + b_rgb0)));
                                rgb_out[id * 3 + 1] = 1.f / (1.f + __expf(-(__uint_as_float(v[1]) + b_rgb1)));
                                rgb_out[id * 3 + 2] = 1.f / (1.f + __expf(-(__uint_as_float(v[2]) + b_rgb2)));
                            }
                            continue;
                        }
                        if (g == 8 && blk == 4) {     // density head: column 128 of the head layer, raw
                            if (in && (!TAN || sigma_out)) sigma_out[id] = __uint_as_float(v[0]) + (TAN ? tsc * b_sigma : b_sigma);
                            continue;
                        }
                        if (TAN) {                    // tau = mask * (W tau_prev): the primal's ReLU pattern, no clamp
                            const uint32_t m = pmw[blk];
#pragma unroll
                            for (int c = 0; c < 32; ++c) v[c] = ((m >> mask_bit_of_col(c)) & 1u) ? v[c] : 0u;
                        }
                        if (TRAIN && !TAN) {
                            // 1-bit ReLU mask from the sign bits: one funnel shift per column; bit (31-c) of the
                            // word <-> column c of the block (tc::mask_bit_of_col)
                            // (four independent 8-column chains, then merged, so the shifts are not one serial chain)
                            uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                n0 = __funnelshift_l(v[c], n0, 1); n1 = __funnelshift_l(v[8 + c], n1, 1);
                                n2 = __funnelshift_l(v[16 + c], n2, 1); n3 = __funnelshift_l(v[24 + c], n3, 1);
                            }
                            const uint32_t neg = (((n0 * 256u + n1) * 256u + n2) * 256u) + n3;
                            if (g <= 7) *(uint32_t*)(st_tile + ST_MASK + g * 4096 + blk * 512 + row * 4) = ~neg;     // [block][row]: coalesced
                            else *(uint32_t*)(st_tile + ST_CMASK + blk * 512 + row * 4) = ~neg;
                        }
                        uint32_t w[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            w[k] = TAN ? pack_bf16(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]))
                                       : pack_relu_bf16(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
                        uint8_t* dst = act_row + (blk >> 1) * 16384;
#pragma unroll
                        for (uint32_t u = 0; u < 4; ++u)
                            *(uint4*)(dst + ((((uint32_t)(blk & 1) * 4 + u) ^ sw) << 4)) =
                                make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
                        if (TRAIN && (blk & 1)) {
                            // a 64-column chunk of the image is complete for this warp's 32 rows: stream those 4 KB to
                            // the stash now (h_{g+1}, or c after the head layer) -- small stores spread over the epilogue
                            // instead of one 64 KB burst behind a 128-thread barrier at its end
                            fence_proxy_async();
                            __syncwarp();
                            if (elect_one()) {
                                const uint32_t off = (uint32_t)(blk >> 1) * 16384u + (uint32_t)q * 4096u;
                                bulk_s2g_hint(st_tile + (g == 8 ? ST_C : ST_H + (int64_t)g * 65536) + off, act_s + off, 4096, stream_pol);
                                bulk_commit();
                            }
                        }
                    }
                }
                if (leader) TRACE(2 + t, 4, g, 0);
                tc_fence_before();                // TMEM reads done before the MMAs that follow the arrive overwrite the accumulators
                if (g < 9) {      // (tangent mode ends at g = 8: its c image is stored too, so every image of the stash is defined)
                    fence_proxy_async();
                    if (g < NGT - 1) mbar_arrive_remote(my_act, 0);
                }
                if (leader) TRACE(2 + t, 3, g, 0);
            }
        }
        if (TRAIN && elect_one()) bulk_wait0();
    }
    tc_fence_before();
    cluster_sync_all();          // the peer's shared memory / TMEM stay alive until the leader's last MMA has retired
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

#ifdef AN_MLP_TRACE
extern "C" int an_debug_trace_fwd(void* buf) {
    return (int)cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
}
#endif

extern "C" int64_t an_mlp_stash_bytes(int64_t n_max)
{
    if (n_max <= 0) return 0;
    return mlp::n_tiles_for(n_max) * mlp::ST_TILE;
}

// persistent CTA pairs: one pair per TPC, never more pairs than 512-point iterations
static inline int pair_grid(int64_t n_max)
{
    const int64_t iters = (n_max + mlp::PAIR_POINTS - 1) / mlp::PAIR_POINTS;
    const int pairs = an_num_sms() / 2;
    return 2 * (int)(iters < pairs ? iters : pairs);
}

extern "C" int an_mlp_fwd(const void* packed, const float* xyz_cano, const int32_t* cidx, const int32_t* count,
                          int64_t n_max, float* sigma, float* rgb, void* stash, void* stream)
{
    if (!packed || !xyz_cano || !sigma || !rgb || n_max <= 0) return AN_ERR_ARG;
    if (cidx && !count) return AN_ERR_ARG;
    if (((uintptr_t)packed) & 1023) return AN_ERR_ALIGN;
    if (stash && (((uintptr_t)stash) & 127)) return AN_ERR_ALIGN;
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(mlp_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    const int grid = pair_grid(n_max);
    if (stash)
        mlp_fwd_tc_kernel<1><<<grid, THREADS, SM_ALLOC, (cudaStream_t)stream>>>(
            (const uint8_t*)packed, xyz_cano, cidx, count, n_max, sigma, rgb, (uint8_t*)stash, nullptr, nullptr, nullptr);
    else
        mlp_fwd_tc_kernel<0><<<grid, THREADS, SM_ALLOC, (cudaStream_t)stream>>>(
            (const uint8_t*)packed, xyz_cano, cidx, count, n_max, sigma, rgb, nullptr, nullptr, nullptr, nullptr);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

// Forward-mode tangent of the trunk (the normal-smoothness regulariser differentiates
// d sigma/d xyz with respect to the weights -- torch double backward in the reference, nerf.py:177-190).
// For a loss L(s), s = d sigma/d xyz, with v = dL/ds per point:  dL/dW_l = (m_l * delta_l) tau_{l-1}^T where
// delta_l = d sigma/d h_l are the dY images of an an_mlp_bwd_dgrad run with g_sigma = 1, g_rgb = 0 over the
// same points, and tau_l = m_l * (W_l tau_{l-1}), tau_0 = (d enc/d xyz) v is what this kernel writes to
// `tstash` (activation-stash layout): an_mlp_bwd_wgrad(packed, tstash, dY scratch) then yields the weight
// gradients (its bias outputs are not gradients of L -- biases do not enter d sigma/d xyz -- the caller
// zeroes them).  pstash = the primal forward's stash over the same compacted points (ReLU masks).
// tsigma (ids) receives w_sigma . tau_8 when non-NULL.
// tscale (ids) non-NULL: c = dL/d sigma per point.  Because the sigma-only activation-gradient chain is linear in its
// per-point scalar seed (delta' = c * delta), the first-order term  sum_p c_p delta_p X_p^T  folds into the same
// product: the kernel writes T_l = tau_l + c X_l (T_l = m_l * (W_l T_{l-1} + c b_l), T_0 = tau_0 + c enc(x)), and
// an_mlp_bwd_wgrad_scaled(packed, tstash, delta scratch, bias_scale = c) yields the whole gradient of L(sigma, s) --
// weights from delta T^T, biases from sum_p c_p delta_p -- in one pass instead of tangent + wgrad + dgrad + wgrad.
extern "C" int an_mlp_fwd_tangent(const void* packed, const float* xyz_cano, const float* tvec, const float* tscale,
                                  const void* pstash, const int32_t* cidx, const int32_t* count, int64_t n_max,
                                  float* tsigma, void* tstash, void* stream)
{
    if (!packed || !xyz_cano || !tvec || !pstash || !tstash || n_max <= 0) return AN_ERR_ARG;
    if (cidx && !count) return AN_ERR_ARG;
    if (((uintptr_t)packed) & 1023) return AN_ERR_ALIGN;
    if ((((uintptr_t)pstash) & 127) || (((uintptr_t)tstash) & 127)) return AN_ERR_ALIGN;
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    const int grid = pair_grid(n_max);
    mlp_fwd_tc_kernel<2><<<grid, THREADS, SM_ALLOC, (cudaStream_t)stream>>>(
        (const uint8_t*)packed, xyz_cano, cidx, count, n_max, tsigma, nullptr, (uint8_t*)tstash, tvec,
        (const uint8_t*)pstash, tscale);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
