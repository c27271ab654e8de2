// A17 (MLP part): backward of the positional encoding + 8x256 MLP on tcgen05/TMEM.
// Autograd of models/nerf.py:129-175 + models/embedding.py:22-39.  Three kernels:
//
//  1. mlp_bwd_dgrad_kernel -- same skeleton as the forward (persistent CTA pairs, cta_group::2 MMAs of
//     M = 256 over tile t of both CTAs, 2 x 128-row tiles per CTA ping-ponging, each CTA streaming its
//     half of every W^T chunk image by TMA, accumulators in TMEM).  The activation-gradient image dY
//     stays in shared memory as the next layer's A operand; ReLU masks come from the forward's
//     1-bit stash.  Both heads run on the tensor cores (mlp_layout.cuh): step 0 is the rgb head
//     (K = 16: d rgb_pre -> d c), step 1 the fused head layer (K = 144: [d c_pre | d sigma] -> d h8).
//     The encoding gradients (two N=64 GEMMs: layer 5's encoding part and layer 1) are folded
//     to d(xyz) analytically in the epilogue.  Every pre-activation gradient image is also
//     streamed to HBM scratch (TMA bulk store) for the weight-gradient kernel.
//  2. mlp_bwd_wgrad_kernel -- dW_g = dY_g^T X_g as K-streaming GEMMs over the points: both
//     operands are the row-major point images already in HBM, consumed as MN-major UMMA operands
//     (no transposes).  One CTA owns one (layer, point-split) and keeps the full dW_g tile
//     (2 x 128 x N fp32) in TMEM across its whole K loop; bias gradients are column sums of the
//     dY image taken from shared memory by the otherwise idle warps.  Every CTA stores its partial
//     dW / db into its own slice of a workspace; mlp_wgrad_reduce_kernel then adds the slices to the
//     gradient in a fixed order -- no floating-point atomics, so the gradient is reproducible bit for bit.  HBM-bound by construction
//     (128 FLOP/B < ridge 214 FLOP/B): ~9.4 KB read per point.  The head-layer and rgb jobs run
//     transposed (A = X, B = dY) because their dY is narrower than one UMMA M.
//  3. mlp_unfuse_grad_kernel (mlp_pack.cu) -- chain rule from the fused head layer's dW', db' to
//     xyz_encoding_final / dir_encoding.
#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc_common.cuh"

namespace {
constexpr int THREADS = 320;
constexpr int NSTAGE = 6;
constexpr uint32_t STAGE_BYTES = 16384;                          // this CTA's half of a W^T chunk image
constexpr uint32_t SM_ACT = 0;                                   // [2][4][16 KB]
constexpr uint32_t SM_WST = 131072;                              // [NSTAGE][16 KB]
constexpr uint32_t SM_BAR = SM_WST + NSTAGE * STAGE_BYTES;
constexpr uint32_t SM_BYTES = SM_BAR + 192;
constexpr uint32_t SM_ALLOC = SM_BYTES + 1024;
static_assert(SM_ALLOC <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");

// order in which the backward consumes the W^T steps of mlp_layout.cuh (s6 = layer-5 encoding
// part runs before s5 so both read the same dY image and share one accumulator region)
__device__ __forceinline__ int step_of(int i) {
    const int order[11] = {0, 1, 2, 3, 4, 6, 5, 7, 8, 9, 10};
    return order[i];
}
}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
mlp_bwd_dgrad_kernel(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ stash,
                     const float* __restrict__ xyz_cano, const float* __restrict__ rgb,
                     const int32_t* __restrict__ cidx, const int32_t* __restrict__ count, int64_t n_max,
                     const float* __restrict__ g_sigma, const float* __restrict__ g_rgb,
                     float* __restrict__ g_xyz, uint8_t* __restrict__ dy)
{
    using namespace mlp;
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + SM_BAR, bar_empty = sbase + SM_BAR + 64;
    const uint32_t bar_act = sbase + SM_BAR + 128, bar_acc = sbase + SM_BAR + 144, tmem_slot = sbase + SM_BAR + 160;
    const bool want_gx = g_xyz != nullptr;
    const uint32_t rank = cluster_ctarank();             // 0 = leader (issues the MMAs), 1 = peer
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

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
    const uint32_t tmem_base = *(volatile uint32_t*)(sgen + SM_BAR + 160);

    // roles as in the forward (mlp_tc.cu): producer in both CTAs, MMA issuer in the leader / stage relay in the
    // peer, epilogue warps; the tiles of a CTA ping-pong: MMAs of one tile overlap the other tile's epilogue
    if (warp == 0) {
        {   // whole warp runs the loop, one elected lane issues (tc::elect_one)
            uint32_t it = 0;
            const uint64_t keep = l2_policy_keep();
            for (int64_t iter = pair; iter < num_iters; iter += npairs)
                for (int i = 0; i < 11; ++i) {
                    const int s = step_of(i);
                    if (!want_gx && (s == 6 || s == 10)) continue;
                    const uint32_t bytes = bs_chunk_bytes(s) >> 1;       // this CTA's half of the rows
                    // one load per chunk: both tiles of the CTA run their MMAs against the same staged copy (a step has
                    // at most 4 chunks, the ring 6 stages: the next step's first chunks prefetch meanwhile)
                    for (int kc = 0; kc < bs_chunks(s); ++kc, ++it) {
                        const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
                        if (elect_one()) {
                            mbar_expect_tx(bar_full + 8 * st, bytes);
                            bulk_g2s_hint(sbase + SM_WST + st * STAGE_BYTES, packed + bwd_chunk_off(s, kc) + rank * bytes, bytes, bar_full + 8 * st, keep);
                        }
                        __syncwarp();
                    }
                }
        }
    } else if (warp == 1 && rank != 0) {
        if (lane == 0) {          // peer: tell the leader when this CTA's half of a stage has landed
            uint32_t it = 0;
            for (int64_t iter = pair; iter < num_iters; iter += npairs)
                for (int i = 0; i < 11; ++i) {
                    const int s = step_of(i);
                    if (!want_gx && (s == 6 || s == 10)) continue;
                    for (int j = 0; j < bs_chunks(s); ++j, ++it) {
                        const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                        mbar_wait(bar_full + 8 * st, ph);
                        mbar_arrive_remote(bar_full + 8 * st, 0);
                    }
                }
        }
    } else if (warp == 1) {
        {          // leader: MMA issuer for both CTAs (whole warp runs the loop, one elected lane issues: tc::elect_one)
            uint32_t it = 0, act_phase = 0;
            for (int64_t iter = pair; iter < num_iters; iter += npairs)
                for (int i = 0; i < 11; ++i) {
                    const int s = step_of(i);
                    if (!want_gx && (s == 6 || s == 10)) continue;
                    const uint32_t idesc = make_idesc_bf16(256, bs_rows(s), 0, 0);
                    const int nc = bs_chunks(s);
                    for (int t = 0; t < 2; ++t) {
                        mbar_wait(bar_act + 8 * t, act_phase);
                        tc_fence_after();
                        for (int kc = 0; kc < nc; ++kc) {
                            const uint32_t i2 = it + kc, st = i2 % NSTAGE, ph = (i2 / NSTAGE) & 1u;
                            if (t == 0) { mbar_wait(bar_full + 8 * st, ph); tc_fence_after(); }       // tile 1 reuses the staged chunk
                            const uint32_t wb = sbase + SM_WST + st * STAGE_BYTES;
                            const uint32_t ab = sbase + SM_ACT + t * 65536u + kc * 16384u;
                            const int nk = bs_ksteps(s, kc);
                            if (elect_one()) {
                                for (int k = 0; k < nk; ++k)
                                    umma_pair(tmem_base + t * 256u, make_desc(ab + k * 32u, 16, 1024),
                                              make_desc(wb + k * 32u, 16, 1024), idesc, (kc > 0 || k > 0) ? 1u : 0u);
                                if (t == 1) umma_commit_pair(bar_empty + 8 * st);      // both tiles done with the stage
                            }
                            __syncwarp();
                        }
                        if (elect_one()) umma_commit_pair(bar_acc + 8 * t);
                        __syncwarp();
                    }
                    it += nc;
                    act_phase ^= 1u;
                }
        }
    } else {
        const int e = threadIdx.x - 64;
        const int t = e >> 7;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t sw = (uint32_t)(row & 7);
        uint8_t* act_row = sgen + SM_ACT + t * 65536 + row * 128;
        const uint32_t act_s = sbase + SM_ACT + t * 65536u;
        const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
        const uint32_t my_act = bar_act + 8 * t, my_acc = bar_acc + 8 * t;
        const uint64_t stream_pol = l2_policy_stream();
        uint32_t acc_phase = 0;

        for (int64_t iter = pair; iter < num_iters; iter += npairs) {
            const int64_t tile = (iter * 2 + rank) * 2 + t;          // = compact point index / 128
            const int64_t p = tile * 128 + row;
            const bool in = p < n;
            const int64_t id = in ? (cidx ? (int64_t)cidx[p] : p) : 0;
            const uint8_t* st_tile = stash + tile * ST_TILE;
            uint8_t* dy_tile = dy + tile * DY_TILE;
            float gs = 0.f, dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
            if (in) {
                gs = g_sigma[id];
                const float a0 = rgb[id * 3], a1 = rgb[id * 3 + 1], a2 = rgb[id * 3 + 2];
                dp0 = g_rgb[id * 3] * a0 * (1.f - a0); dp1 = g_rgb[id * 3 + 1] * a1 * (1.f - a1); dp2 = g_rgb[id * 3 + 2] * a2 * (1.f - a2);
            }
            // this warp's bulk stores of the previous iteration must have finished reading its rows of the image
            if (elect_one()) bulk_wait_read0();
            __syncwarp();
            {   // d rgb_pre -> A image of step 0 (K = 16: units 0,1 of chunk 0; columns 0..2 carry data)
                *(uint4*)(act_row + ((0u ^ sw) << 4)) = make_uint4(pack_bf16(dp0, dp1), pack_bf16(dp2, 0.f), 0u, 0u);
                *(uint4*)(act_row + ((1u ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            }
            fence_proxy_async();
            __syncwarp();         // every warp streams its own 32 rows (4 KB, contiguous in the image) to the scratch
            if (elect_one()) { bulk_s2g_hint(dy_tile + DY_RGB + q * 4096, act_s + q * 4096u, 4096, stream_pol); bulk_commit(); }
            mbar_arrive_remote(my_act, 0);

            // encoding derivative factors (same double-angle recurrence as the forward)
            float gx[3] = {0.f, 0.f, 0.f};
            float x[3] = {0.f, 0.f, 0.f};
            if (want_gx && in) { x[0] = xyz_cano[id * 3]; x[1] = xyz_cano[id * 3 + 1]; x[2] = xyz_cano[id * 3 + 2]; }

            for (int i = 0; i < 11; ++i) {
                const int s = step_of(i);
                if (!want_gx && (s == 6 || s == 10)) continue;
                // which ReLU mask applies to this step's output, and where the dY image goes; the mask words are
                // fetched now, so that their (HBM) latency hides behind the accumulator wait
                int mask_layer = -1;            // index into the stash's h masks (0..7 = h1..h8); step 0 uses the c mask
                int64_t dy_off = 0;
                if (s == 0) { dy_off = DY_HEAD; }
                else if (s >= 1 && s <= 4) { mask_layer = 8 - s; dy_off = DY_H + (int64_t)(8 - s) * 65536; }
                else if (s == 5) { mask_layer = 3; dy_off = DY_H + 3 * 65536; }
                else { mask_layer = 9 - s; dy_off = DY_H + (int64_t)(9 - s) * 65536; }     // s = 7,8,9 -> h3,h2,h1
                uint32_t mw[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                if (s != 6 && s != 10) {
                    if (mask_layer >= 0) {
                        const uint8_t* mp = st_tile + ST_MASK + mask_layer * 4096 + row * 4;     // [block][row] words
#pragma unroll
                        for (int cb = 0; cb < 8; ++cb) mw[cb] = __ldg((const uint32_t*)(mp + cb * 512));
                    } else {                        // step 0: d c (128 columns) masked by [c > 0]
#pragma unroll
                        for (int cb = 0; cb < 4; ++cb) mw[cb] = __ldg((const uint32_t*)(st_tile + ST_CMASK + cb * 512 + row * 4));
                    }
                }
                if (s != 6 && s != 10) {
                    // this warp's bulk stores of the previous image must have drained before this step overwrites its
                    // rows; waited for here (the overwrite itself happens after the accumulator wait, i.e. after the
                    // MMAs that read the image as their A operand), off the critical path
                    if (elect_one()) bulk_wait_read0();
                    __syncwarp();
                }
                mbar_wait(my_acc, acc_phase); acc_phase ^= 1u;
                tc_fence_after();
                if (s == 6 || s == 10) {
                    // d(enc) (64 columns) -> d(xyz): enc = [x, sin(2^k x), cos(2^k x)]_k
                    float de[64];
                    {
                        uint32_t v[32];
                        tmem_ld32(tm, v); tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) de[c] = __uint_as_float(v[c]);
                        tmem_ld32(tm + 32, v); tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) de[32 + c] = __uint_as_float(v[c]);
                    }
                    float sn[3], cs[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) { sincosf(x[a], &sn[a], &cs[a]); gx[a] += de[a]; }
                    float f = 1.f;
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            gx[a] += f * (de[3 + 6 * k + a] * cs[a] - de[6 + 6 * k + a] * sn[a]);
                            const float s2 = 2.f * sn[a] * cs[a], c2 = 1.f - 2.f * sn[a] * sn[a];
                            sn[a] = s2; cs[a] = c2;
                        }
                        f *= 2.f;
                    }
                    tc_fence_before();
                    if (s == 10) {
                        if (in) { g_xyz[id * 3] = gx[0]; g_xyz[id * 3 + 1] = gx[1]; g_xyz[id * 3 + 2] = gx[2]; }
                    } else {
                        mbar_arrive_remote(my_act, 0);  // A image unchanged; accumulator region is free again
                    }
                    continue;
                }
                const int ncb = s == 0 ? 4 : 8;
                uint32_t va[32], vb[32];
                tmem_ld32(tm, va);
#pragma unroll
                for (int cb = 0; cb < 8; ++cb) {
                    if (cb >= ncb) break;
                    uint32_t (&v)[32] = (cb & 1) ? vb : va;
                    tmem_ld_wait();
                    if (cb + 1 < ncb) tmem_ld32(tm + (cb + 1) * 32, (cb & 1) ? va : vb);     // prefetch the next block
                    float f[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] = __uint_as_float(v[c]);
                    const uint32_t m = mw[cb];
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] = ((m >> mask_bit_of_col(c)) & 1u) ? f[c] : 0.f;
                    uint8_t* dst = act_row + (cb >> 1) * 16384;
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) {
                        uint4 o;
                        o.x = pack_bf16(f[8 * u], f[8 * u + 1]); o.y = pack_bf16(f[8 * u + 2], f[8 * u + 3]);
                        o.z = pack_bf16(f[8 * u + 4], f[8 * u + 5]); o.w = pack_bf16(f[8 * u + 6], f[8 * u + 7]);
                        *(uint4*)(dst + ((((uint32_t)(cb & 1) * 4 + u) ^ sw) << 4)) = o;
                    }
                    if (cb & 1) {
                        // a 64-column chunk of the dY image is complete for this warp's 32 rows: stream those 4 KB to
                        // the scratch now -- small stores spread over the epilogue, no 128-thread barrier
                        fence_proxy_async();
                        __syncwarp();
                        if (elect_one()) {
                            const uint32_t off = (uint32_t)(cb >> 1) * 16384u + (uint32_t)q * 4096u;
                            bulk_s2g_hint(dy_tile + dy_off + off, act_s + off, 4096, stream_pol);
                            bulk_commit();
                        }
                    }
                }
                if (s == 0) {   // third chunk of the head layer's dY: column 0 = d sigma (K-step 0 = units 0,1)
                    *(uint4*)(act_row + 2 * 16384 + ((0u ^ sw) << 4)) = make_uint4(pack_bf16(gs, 0.f), 0u, 0u, 0u);
                    *(uint4*)(act_row + 2 * 16384 + ((1u ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                }
                tc_fence_before();
                fence_proxy_async();
                if (s == 0) {   // the d sigma chunk of the head layer's dY
                    __syncwarp();
                    if (elect_one()) {
                        const uint32_t off = 2u * 16384u + (uint32_t)q * 4096u;
                        bulk_s2g_hint(dy_tile + dy_off + off, act_s + off, 4096, stream_pol);
                        bulk_commit();
                    }
                }
                if (!(s == 9 && !want_gx)) mbar_arrive_remote(my_act, 0);      // last step has no consumer MMA
            }
        }
        if (elect_one()) bulk_wait0();
    }
    tc_fence_before();
    cluster_sync_all();          // the peer's shared memory / TMEM stay alive until the leader's last MMA has retired
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// ------------------------------------------------------------------------------ wgrad
namespace {
// Job table.  Normal job:  dW (out x in) = dY^T X:  A = dY image (M = out features, from the dY scratch),
// B = X image (N = in features, from the forward stash); accumulator lane = out feature, column = in feature.
// Transposed job (swap = 1, the two head layers, whose dY is narrower than one UMMA M):  dW^T = X^T dY:
// A = X image (stash), B = dY image (scratch, N = head width); lane = in feature, column = out feature.
struct WJob {
    int64_t a_off; int a_chunks;        // A operand: image offset inside its tile, 64-column chunks
    int64_t b_off; int b_chunks; int N; // B operand, UMMA N
    int swap;                           // 0: A = dY (scratch), B = X (stash);  1: A = X (stash), B = dY (scratch)
    int kind;                           // 0: trunk linear `lin` (columns col0..col0+ncol_valid), 1: fused head layer, 2: rgb head
    int lin; int col0; int ncol_valid; int bias;
};
__device__ __forceinline__ WJob wjob(int j) {
    using namespace mlp;
    WJob w;
    switch (j) {
        case 0:  w = {DY_H + 0 * 65536, 4, ST_ENC, 1, 64, 0, 0, 0, 0, 63, 1}; break;                 // L1: X = enc
        case 1:  w = {DY_H + 1 * 65536, 4, ST_H + 0 * 65536, 4, 256, 0, 0, 1, 0, 256, 1}; break;    // L2: X = h1
        case 2:  w = {DY_H + 2 * 65536, 4, ST_H + 1 * 65536, 4, 256, 0, 0, 2, 0, 256, 1}; break;
        case 3:  w = {DY_H + 3 * 65536, 4, ST_H + 2 * 65536, 4, 256, 0, 0, 3, 0, 256, 1}; break;
        case 4:  w = {DY_H + 4 * 65536, 4, ST_ENC, 1, 64, 0, 0, 4, 0, 63, 0}; break;                 // L5 encoding columns
        case 5:  w = {DY_H + 4 * 65536, 4, ST_H + 3 * 65536, 4, 256, 0, 0, 4, 63, 256, 1}; break;   // L5 hidden columns: X = h4
        case 6:  w = {DY_H + 5 * 65536, 4, ST_H + 4 * 65536, 4, 256, 0, 0, 5, 0, 256, 1}; break;
        case 7:  w = {DY_H + 6 * 65536, 4, ST_H + 5 * 65536, 4, 256, 0, 0, 6, 0, 256, 1}; break;
        case 8:  w = {DY_H + 7 * 65536, 4, ST_H + 6 * 65536, 4, 256, 0, 0, 7, 0, 256, 1}; break;    // L8: X = h7
        case 9:  w = {ST_H + 7 * 65536, 4, DY_HEAD, 3, HEAD_N, 1, 1, 8, 0, 129, 1}; break;           // head layer: X = h8, dY = [d c_pre | d sigma]
        default: w = {ST_C, 2, DY_RGB, 1, RGB_N, 1, 2, 11, 0, 3, 1}; break;                          // rgb head: X = c
    }
    return w;
}
constexpr int NJOBS = 11;
constexpr int WG_THREADS = 320;
constexpr int WG_STAGES = 3;
constexpr uint32_t WG_STAGE_BYTES = 65536;               // A half (<=32 KB) + B half (<=32 KB): 64 points
constexpr uint32_t WG_BAR = WG_STAGES * WG_STAGE_BYTES;
constexpr uint32_t WG_ALLOC = WG_BAR + 128 + 1024;
}  // namespace

// grid = (splits, NJOBS).  CTA (sp, j) accumulates dW_j over tiles sp, sp+splits, ... and stores the result into
// slice sp of `partial` ([splits][GRAD_FLOATS], indexed like the gradient vector).
__global__ void __launch_bounds__(WG_THREADS, 1)
mlp_bwd_wgrad_kernel(const uint8_t* __restrict__ stash, const uint8_t* __restrict__ dy,
                     const int32_t* __restrict__ count, int has_count, int64_t n_max,
                     float* __restrict__ partial, const float* __restrict__ bias_scale)
{
    using namespace mlp;
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + WG_BAR, bar_empty = sbase + WG_BAR + 32, bar_cs = sbase + WG_BAR + 64;
    const uint32_t bar_done = sbase + WG_BAR + 96, tmem_slot = sbase + WG_BAR + 104;

    int64_t n = n_max;
    if (has_count) { const int64_t c = *count; n = c < n_max ? c : n_max; }
    const int64_t n_tiles = n_tiles_for(n);              // tiles written by the forward / dgrad pairs (zero dY rows beyond n)
    float* __restrict__ g_params = partial + (int64_t)blockIdx.x * GRAD_FLOATS;     // this split's slice
    const WJob job = wjob(blockIdx.y);
    const int M_halves = (job.a_chunks + 1) / 2;          // 128 accumulator lanes per half
    const int N = job.N;                                  // accumulator columns per half
    const uint32_t a_half_bytes = (uint32_t)job.a_chunks * 8192u, b_half_bytes = (uint32_t)job.b_chunks * 8192u;
    const uint8_t* a_base = job.swap ? stash : dy;
    const uint8_t* b_base = job.swap ? dy : stash;
    const int64_t a_tile = job.swap ? ST_TILE : DY_TILE, b_tile = job.swap ? DY_TILE : ST_TILE;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_cs + 8 * s, 256);
        }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *(volatile uint32_t*)(sgen + WG_BAR + 104);

    // stage layout: A chunks (a_chunks x 8 KB: 64 points x 128 B each) then B chunks at +32 KB
    if (warp == 0) {
        {   // whole warp runs the loop, one elected lane issues (tc::elect_one)
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int half = 0; half < 2; ++half, ++it) {
                    const uint32_t st = it % WG_STAGES, ph = (it / WG_STAGES) & 1u;
                    mbar_wait(bar_empty + 8 * st, ph ^ 1u);
                    mbar_wait(bar_cs + 8 * st, ph ^ 1u);          // column-sum readers are done with the stage too
                    const uint32_t dst = sbase + st * WG_STAGE_BYTES;
                    const uint8_t* asrc = a_base + tile * a_tile + job.a_off + half * 8192;
                    const uint8_t* bsrc = b_base + tile * b_tile + job.b_off + half * 8192;
                    if (elect_one()) {
                        mbar_expect_tx(bar_full + 8 * st, a_half_bytes + b_half_bytes);
                        for (int c = 0; c < job.a_chunks; ++c) bulk_g2s(dst + c * 8192u, asrc + c * 16384, 8192, bar_full + 8 * st);
                        for (int c = 0; c < job.b_chunks; ++c) bulk_g2s(dst + 32768u + c * 8192u, bsrc + c * 16384, 8192, bar_full + 8 * st);
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        {   // MMA issuer: whole warp runs the loop, one elected lane issues (tc::elect_one)
            uint32_t it = 0;
            const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);           // both operands MN-major
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int half = 0; half < 2; ++half, ++it) {
                    const uint32_t st = it % WG_STAGES, ph = (it / WG_STAGES) & 1u;
                    mbar_wait(bar_full + 8 * st, ph);
                    tc_fence_after();
                    const uint32_t ab = sbase + st * WG_STAGE_BYTES, bb = ab + 32768u;
                    if (elect_one()) {
                        for (int h = 0; h < M_halves; ++h)
#pragma unroll
                            for (int k = 0; k < 4; ++k)      // 16 points per UMMA K-step = 2 x (8 rows x 128 B)
                                umma(tmem_base + h * 256u,
                                     make_desc(ab + h * 16384u + k * 2048u, 8192, 1024),
                                     make_desc(bb + k * 2048u, 8192, 1024), idesc, (it > 0 || k > 0) ? 1u : 0u);
                        umma_commit(bar_empty + 8 * st);
                    }
                    __syncwarp();
                }
            if (elect_one()) umma_commit(bar_done);
            __syncwarp();
        }
    } else {
        // warps 2..9: bias gradient = column sums of the dY image (thread = dY column), then the
        // final TMEM -> global accumulation.
        const int e = threadIdx.x - 64;                  // 0..255 = dY column
        float bsum = 0.f;
        const int dy_cols = job.swap ? job.ncol_valid : job.a_chunks * 64;
        const bool do_bias = job.bias && e < dy_cols;
        const uint32_t dy_img = job.swap ? 32768u : 0u;   // the dY image is the B operand of a transposed job
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int half = 0; half < 2; ++half, ++it) {
                const uint32_t st = it % WG_STAGES, ph = (it / WG_STAGES) & 1u;
                mbar_wait(bar_full + 8 * st, ph);
                if (do_bias) {
                    const uint8_t* img = sgen + st * WG_STAGE_BYTES + dy_img + (e >> 6) * 8192;
                    const int c = e & 63;
                    if (bias_scale) {      // db = sum_p c_p dY_p: per-point weights in compact order (an_mlp_bwd_wgrad_scaled)
                        const int64_t p0 = tile * 128 + half * 64;
#pragma unroll 8
                        for (int r = 0; r < 64; ++r) {
                            const uint16_t raw16 = *(const uint16_t*)(img + r * 128 + ((((c >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1)));
                            const float sc = (p0 + r < n) ? __ldg(bias_scale + p0 + r) : 0.f;
                            bsum += sc * __uint_as_float((uint32_t)raw16 << 16);
                        }
                    } else {
#pragma unroll 8
                        for (int r = 0; r < 64; ++r) {
                            const uint16_t raw16 = *(const uint16_t*)(img + r * 128 + ((((c >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1)));
                            bsum += __uint_as_float((uint32_t)raw16 << 16);
                        }
                    }
                }
                mbar_arrive(bar_cs + 8 * st);
            }
        if (do_bias && n_tiles > blockIdx.x) {
            float* dst;
            if (job.kind == 0) dst = g_params + flat_b_off(job.lin) + e;
            else if (job.kind == 1) dst = e < 128 ? g_params + GRAD_FUSED_B + e : g_params + flat_b_off(10);   // db', d b_sigma
            else dst = g_params + flat_b_off(11) + e;
            *dst = bsum;
        }
        // drain the accumulators
        if (n_tiles > (int64_t)blockIdx.x) {
            mbar_wait(bar_done, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int grp = (warp - 2) >> 2;                         // two warp groups split the column blocks
            const int nblk = (N + 31) / 32;
            const int in_dim = lin_in(job.lin);
            float* Wg = g_params + flat_w_off(job.lin);
            for (int h = 0; h < M_halves; ++h) {
                const int lane_row = h * 128 + q * 32 + lane;
                for (int cb = grp; cb < nblk; cb += 2) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * 256u + cb * 32u, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int col = cb * 32 + c;
                        if (col >= job.ncol_valid) continue;
                        float* dst;
                        if (job.kind == 0) dst = Wg + (int64_t)lane_row * in_dim + job.col0 + col;          // lane = out, col = in
                        else if (job.kind == 1)                                                              // lane = in (h8), col = out
                            dst = col < 128 ? g_params + GRAD_FUSED_W + col * 256 + lane_row : g_params + flat_w_off(10) + lane_row;
                        else dst = g_params + flat_w_off(11) + col * 128 + lane_row;                         // lane = in (c), col = rgb channel
                        *dst = __uint_as_float(v[c]);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

// g[i] += sum over the active splits (in order) of partial[s][i], for every entry the job table writes: the weights
// and biases of the eight trunk linears, the density head, the rgb head and the fused head layer's scratch tail;
// xyz_encoding_final / dir_encoding (flat ids 8, 9) come from the chain rule kernel that follows.
// One thread per four consecutive entries: the (at most REDUCE_MAX_SPLITS) slice loads of a thread are all in flight at once
// (the per-entry loop over a run-time slice count was a chain of dependent rounds: 22 us per launch for 33 MB).
#define REDUCE_MAX_SPLITS 16
__global__ void __launch_bounds__(256)
mlp_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, const int32_t* __restrict__ count, int has_count,
                        int64_t n_max, float* __restrict__ g)
{
    using namespace mlp;
    int64_t n = n_max;
    if (has_count) { const int64_t c = *count; n = c < n_max ? c : n_max; }
    const int64_t n_tiles = n_tiles_for(n);
    const int active = (int)(n_tiles < splits ? n_tiles : splits);     // splits beyond the tile count wrote nothing
    const int64_t skip0 = flat_w_off(8), skip1 = flat_w_off(10);
    const bool vec_ok = active <= REDUCE_MAX_SPLITS && (GRAD_FLOATS & 3) == 0 && ((((uintptr_t)g) | ((uintptr_t)partial)) & 15) == 0;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i < GRAD_FLOATS; i += (int64_t)gridDim.x * blockDim.x * 4) {
        if (vec_ok && i + 3 < GRAD_FLOATS && (i + 3 < skip0 || i >= skip1)) {
            float4 v[REDUCE_MAX_SPLITS];
#pragma unroll
            for (int s = 0; s < REDUCE_MAX_SPLITS; ++s)
                if (s < active) v[s] = __ldg((const float4*)(partial + (int64_t)s * GRAD_FLOATS + i));
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int s = 0; s < REDUCE_MAX_SPLITS; ++s)
                if (s < active) { acc.x += v[s].x; acc.y += v[s].y; acc.z += v[s].z; acc.w += v[s].w; }
            float4 o = *(float4*)(g + i);
            o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
            *(float4*)(g + i) = o;
        } else {
            for (int64_t e = i; e < i + 4 && e < GRAD_FLOATS; ++e) {
                if (e >= skip0 && e < skip1) continue;
                float acc = 0.f;
                for (int s = 0; s < active; ++s) acc += partial[(int64_t)s * GRAD_FLOATS + e];
                g[e] += acc;
            }
        }
    }
}

int mlp_unfuse_grad_launch(const void* packed, float* g_params, cudaStream_t stream);

static inline int wgrad_splits() { const int s = an_num_sms() / NJOBS; return s < 1 ? 1 : s; }    // 13 on a 148-SM part -> 143 CTAs, one wave

extern "C" int64_t an_mlp_wgrad_ws_bytes(void) { return (int64_t)wgrad_splits() * mlp::GRAD_FLOATS * 4; }

extern "C" int64_t an_mlp_bwd_scratch_bytes(int64_t n_max)
{
    if (n_max <= 0) return 0;
    return mlp::n_tiles_for(n_max) * mlp::DY_TILE;
}

static int bwd_check(const void* packed, const void* stash, const void* scratch, int64_t n_max,
                     const int32_t* cidx, const int32_t* count)
{
    if (!stash || !scratch || n_max <= 0) return AN_ERR_ARG;
    if (cidx && !count) return AN_ERR_ARG;
    if ((packed && (((uintptr_t)packed) & 1023)) || (((uintptr_t)stash) & 127) || (((uintptr_t)scratch) & 127)) return AN_ERR_ALIGN;
    return AN_OK;
}

extern "C" int an_mlp_bwd_dgrad(const void* packed, const void* stash, const float* xyz_cano, const float* rgb,
                                const int32_t* cidx, const int32_t* count, int64_t n_max,
                                const float* g_sigma, const float* g_rgb, float* g_xyz_cano,
                                void* scratch, void* stream)
{
    if (!packed || !xyz_cano || !rgb || !g_sigma || !g_rgb) return AN_ERR_ARG;
    int rc = bwd_check(packed, stash, scratch, n_max, cidx, count);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(mlp_bwd_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    const int64_t iters = (n_max + mlp::PAIR_POINTS - 1) / mlp::PAIR_POINTS;
    const int pairs = an_num_sms() / 2;
    const int grid = 2 * (int)(iters < pairs ? iters : pairs);      // persistent CTA pairs, one per TPC
    mlp_bwd_dgrad_kernel<<<grid, THREADS, SM_ALLOC, (cudaStream_t)stream>>>(
        (const uint8_t*)packed, (const uint8_t*)stash, xyz_cano, rgb, cidx, count, n_max, g_sigma, g_rgb,
        g_xyz_cano, (uint8_t*)scratch);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

static int wgrad_launch(const void* packed, const void* stash, const void* scratch, const int32_t* cidx,
                        const int32_t* count, int64_t n_max, float* g_params, void* wgrad_ws, const float* bias_scale, void* stream)
{
    if (!g_params || !packed || !wgrad_ws) return AN_ERR_ARG;
    if (((uintptr_t)wgrad_ws) & 15) return AN_ERR_ALIGN;
    int rc = bwd_check(packed, stash, scratch, n_max, cidx, count);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(mlp_bwd_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_ALLOC);
    if (e != cudaSuccess) return (int)e;
    const int64_t tiles = mlp::n_tiles_for(n_max);
    int64_t splits = wgrad_splits();
    if (splits > tiles) splits = tiles;
    dim3 wgrid((unsigned)splits, NJOBS);
    mlp_bwd_wgrad_kernel<<<wgrid, WG_THREADS, WG_ALLOC, (cudaStream_t)stream>>>(
        (const uint8_t*)stash, (const uint8_t*)scratch, count, cidx ? 1 : 0, n_max, (float*)wgrad_ws, bias_scale);
    AN_CHECK_LAUNCH();
    mlp_wgrad_reduce_kernel<<<(unsigned)((mlp::GRAD_FLOATS / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)wgrad_ws, (int)splits, count, cidx ? 1 : 0, n_max, g_params);
    AN_CHECK_LAUNCH();
    // chain rule through the fused head layer: dW', db' -> xyz_encoding_final / dir_encoding gradients
    return mlp_unfuse_grad_launch(packed, g_params, (cudaStream_t)stream);
}

extern "C" int an_mlp_bwd_wgrad(const void* packed, const void* stash, const void* scratch, const int32_t* cidx,
                                const int32_t* count, int64_t n_max, float* g_params, void* wgrad_ws, void* stream)
{
    return wgrad_launch(packed, stash, scratch, cidx, count, n_max, g_params, wgrad_ws, nullptr, stream);
}

// same, with the bias gradients weighted per point: db = sum_p bias_scale[p] dY_p (p in compact order, n_max entries);
// the weight gradients are unchanged (dY^T X).  Pairs with an_mlp_fwd_tangent(tscale) -- see there.
extern "C" int an_mlp_bwd_wgrad_scaled(const void* packed, const void* stash, const void* scratch, const int32_t* cidx,
                                       const int32_t* count, int64_t n_max, const float* bias_scale, float* g_params,
                                       void* wgrad_ws, void* stream)
{
    if (!bias_scale) return AN_ERR_ARG;
    return wgrad_launch(packed, stash, scratch, cidx, count, n_max, g_params, wgrad_ws, bias_scale, stream);
}

extern "C" int an_mlp_bwd(const void* packed, const void* stash, const float* xyz_cano, const float* rgb,
                          const int32_t* cidx, const int32_t* count, int64_t n_max,
                          const float* g_sigma, const float* g_rgb, float* g_params, float* g_xyz_cano,
                          void* scratch, void* wgrad_ws, void* stream)
{
    int rc = an_mlp_bwd_dgrad(packed, stash, xyz_cano, rgb, cidx, count, n_max, g_sigma, g_rgb, g_xyz_cano, scratch, stream);
    if (rc) return rc;
    return an_mlp_bwd_wgrad(packed, stash, scratch, cidx, count, n_max, g_params, wgrad_ws, stream);
}
