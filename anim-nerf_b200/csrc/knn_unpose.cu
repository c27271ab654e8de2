// A5-A8: K-nearest-SMPL-vertex search fused with the inverse-skinning blend ("unpose").
// Reference: knn_cuda.KNN(k=4) call at models/anim_nerf.py:158-159, get_neighbs :153-178,
// unpose :180-192, batch_index_select/batch_transform :24-39, point generation
// models/volume_rendering.py:117-120.
//
// Arithmetic contract for the neighbour search (bit-exact indices vs the oracle):
//   d2 = ((qx-vx)^2 + (qy-vy)^2) + (qz-vz)^2, every operation rounded to fp32 (no FMA
//   contraction: __fmul_rn/__fadd_rn), neighbours ordered by (d2, vertex index) ascending,
//   dist = sqrt_rn(d2).
//
// Two search kernels share one epilogue:
//   mode 0  exhaustive: the frame's vertex table (6890 x float4 = 110 KB) is staged once per
//           CTA in shared memory and every thread scans it with warp-broadcast LDS.128 reads.
//   mode 1  grid-pruned: a per-frame uniform grid (cell ~ dis_threshold/3) built by
//           an_vertex_grid_build, with a dilated occupancy flag per cell.  A query with no vertex
//           within dis_threshold is invalid *exactly* (valid needs d_min < threshold, SURVEY
//           App. A) and skips everything; otherwise a ball-pruned row scan of the 7^3 cell box
//           returns the exact 4-NN whenever the 4th neighbour lies within the threshold, else the
//           thread falls back to an exhaustive scan of the global table.  The (d2,index) order
//           key makes the result independent of the scan order: both modes return identical bits.
// The epilogue (confidence from skinning-weight L1 distance, exp(-dist) weights, blend of the
// 4 neighbours' 3x4 observation->canonical transforms, affine apply, validity) reads the
// L2-resident tables (lbs 661 KB, ober2cano 441 KB per frame) with 128-bit loads.
// The search is fp32-ALU bound, not HBM bound: algorithmic HBM bytes per point are 12 B in
// (or 0 when generated from the ray) + 16 B out (+32 B idx/qw when training).
#include "common.cuh"
#include <math_constants.h>

#define KNN_THREADS 256
#define KNN_B0_SCALE 1.25f   // must match the cell the host passes: 3*cell >= KNN_B0_SCALE * dis_threshold
#define GRID_MAXC (AN_GRID_MAX_DIM * AN_GRID_MAX_DIM * AN_GRID_MAX_DIM)

struct GridHeader { float ox, oy, oz, cell; int nx, ny, nz; float flag_r; };    // flag_r > 0: flags = "some vertex within flag_r of the cell's box"

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
#define GRID_R 3                                    // search box radius in cells: GRID_R * cell >= dis_threshold
#define GRID_EXT (AN_GRID_MAX_DIM + 2 * GRID_R)     // occupancy flags cover the grid dilated by GRID_R cells
#define GRID_FLAG_BYTES ((GRID_EXT * GRID_EXT * GRID_EXT + 15) / 16 * 16)
static inline int64_t grid_frame_bytes(int V) {
    return align_up(sizeof(GridHeader), 16) + align_up((int64_t)(GRID_MAXC + 1) * 4, 16) +
           align_up((int64_t)GRID_MAXC * 4, 16) + GRID_FLAG_BYTES + align_up((int64_t)V * 16, 16);
}
#define GRID_OFF_START 32
#define GRID_OFF_COUNT (GRID_OFF_START + ((GRID_MAXC + 1) * 4 + 15) / 16 * 16)
#define GRID_OFF_FLAGS (GRID_OFF_COUNT + GRID_MAXC * 4)
#define GRID_OFF_SORTED (GRID_OFF_FLAGS + GRID_FLAG_BYTES)

struct Best4 { float d[4]; int i[4]; };

__device__ __forceinline__ void best_init(Best4& b) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { b.d[j] = CUDART_INF_F; b.i[j] = 0x7fffffff; }
}
// ascending-index scan: strict '<' keeps the lowest index on ties
__device__ __forceinline__ void best_push_ordered(Best4& b, float d2, int idx) {
    if (d2 < b.d[3]) {
        if (d2 < b.d[2]) {
            b.d[3] = b.d[2]; b.i[3] = b.i[2];
            if (d2 < b.d[1]) {
                b.d[2] = b.d[1]; b.i[2] = b.i[1];
                if (d2 < b.d[0]) { b.d[1] = b.d[0]; b.i[1] = b.i[0]; b.d[0] = d2; b.i[0] = idx; }
                else { b.d[1] = d2; b.i[1] = idx; }
            } else { b.d[2] = d2; b.i[2] = idx; }
        } else { b.d[3] = d2; b.i[3] = idx; }
    }
}
__device__ __forceinline__ bool key_less(float d2, int idx, float bd, int bi) {
    return d2 < bd || (d2 == bd && idx < bi);
}
// arbitrary-order scan: explicit (d2, index) key
__device__ __forceinline__ void best_push_any(Best4& b, float d2, int idx) {
    if (key_less(d2, idx, b.d[3], b.i[3])) {
        if (key_less(d2, idx, b.d[2], b.i[2])) {
            b.d[3] = b.d[2]; b.i[3] = b.i[2];
            if (key_less(d2, idx, b.d[1], b.i[1])) {
                b.d[2] = b.d[1]; b.i[2] = b.i[1];
                if (key_less(d2, idx, b.d[0], b.i[0])) { b.d[1] = b.d[0]; b.i[1] = b.i[0]; b.d[0] = d2; b.i[0] = idx; }
                else { b.d[1] = d2; b.i[1] = idx; }
            } else { b.d[2] = d2; b.i[2] = idx; }
        } else { b.d[3] = d2; b.i[3] = idx; }
    }
}
__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, float vx, float vy, float vz) {
    const float dx = __fsub_rn(qx, vx), dy = __fsub_rn(qy, vy), dz = __fsub_rn(qz, vz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Query point gid from one of three sources: explicit points `xyz`; rays (B*R, 8) + depths z (B*R, K); or -- both xyz and
// z NULL -- a LATTICE (extract_mesh.py:27-35,152-156): `rays` then points to [cx, cy, cz, (float)nj, x[nj], z[K], y[ni]],
// the point is lattice[i][j][k] = (x[j], y[i], z[k]) + centre, gid = (i * nj + j) * K + k.
__device__ __forceinline__ void load_query(const float* __restrict__ xyz, const float* __restrict__ rays,
                                           const float* __restrict__ z, int64_t gid, int K,
                                           float& qx, float& qy, float& qz)
{
    if (xyz) { qx = xyz[gid * 3]; qy = xyz[gid * 3 + 1]; qz = xyz[gid * 3 + 2]; }
    else if (z) {
        const int64_t ray = gid / K;                 // rays are (B*R, 8), z is (B*R, K)
        const float4 r0 = __ldg((const float4*)rays + ray * 2), r1 = __ldg((const float4*)rays + ray * 2 + 1);
        const float zz = z[gid];
        // o + z*d with the product rounded first (torch evaluates mul and add separately; no FMA)
        qx = __fadd_rn(r0.x, __fmul_rn(zz, r0.w)); qy = __fadd_rn(r0.y, __fmul_rn(zz, r1.x)); qz = __fadd_rn(r0.z, __fmul_rn(zz, r1.y));
    } else {
        const int nj = (int)__ldg(rays + 3);
        const int64_t t = gid / K;
        const int k = (int)(gid - t * K);
        const int64_t i = t / nj;
        const int j = (int)(t - i * nj);
        // the reference adds the centre to the fp32 lattice point (extract_mesh.py:156): one fp32 add per coordinate
        qx = __fadd_rn(__ldg(rays + 4 + j), __ldg(rays));
        qy = __fadd_rn(__ldg(rays + 4 + nj + K + i), __ldg(rays + 1));
        qz = __fadd_rn(__ldg(rays + 4 + nj + k), __ldg(rays + 2));
    }
}

struct UnposeOut {
    float* xyz_cano; uint8_t* valid; int32_t* idx; float* dist; float* qw;
    float* sigma; float* rgb; int32_t* cidx; int32_t* count;
};

// Result of the blend for one query: validity, canonical point, neighbour distances and blend weights.
struct Unposed { bool valid; float xc0, xc1, xc2; float dd[4]; float q[4]; };

// Blend shared by the search kernels.  `have4`: best holds the exact 4-NN (distances are emitted);
// `found`: ... and the nearest one is within the threshold, so the point may be valid (blend evaluated).
__device__ __forceinline__ void unpose_compute(const Best4& best, bool have4, bool found, float qx, float qy, float qz,
                                               int b, int V, int J, const float* __restrict__ ober2cano,
                                               const float* __restrict__ lbsw, float thr, bool active, Unposed& u)
{
    bool valid = false;
    float xc0 = 0.f, xc1 = 0.f, xc2 = 0.f;
    float dd[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (active && have4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dd[j] = __fsqrt_rn(best.d[j]);
    }
    if (active && found) {
        // confidence: skinning rows of neighbour j vs neighbour 0
        const float* w0 = lbsw + (int64_t)best.i[0] * J;
        float l1[4] = {0.f, 0.f, 0.f, 0.f};
        if ((J & 3) == 0 && (((uintptr_t)lbsw) & 15) == 0) {      // 16-byte rows: 128-bit loads
            for (int c = 0; c < J; c += 4) {
                const float4 a = __ldg((const float4*)(w0 + c));
#pragma unroll
                for (int j = 1; j < 4; ++j) {
                    const float4 v = __ldg((const float4*)(lbsw + (int64_t)best.i[j] * J + c));
                    l1[j] += fabsf(v.x - a.x); l1[j] += fabsf(v.y - a.y); l1[j] += fabsf(v.z - a.z); l1[j] += fabsf(v.w - a.w);
                }
            }
        } else {
            for (int c = 0; c < J; ++c) {
                const float a = __ldg(w0 + c);
#pragma unroll
                for (int j = 1; j < 4; ++j) l1[j] += fabsf(__ldg(lbsw + (int64_t)best.i[j] * J + c) - a);
            }
        }
        float qs = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float conf = (expf(-l1[j] / 0.02f) > 0.9f) ? 1.0f : 0.0f;   // weight_std=0.1 -> 2*std^2
            q[j] = expf(-dd[j]) * conf;
            qs += q[j];
        }
        float dbar = 0.f;
        float m[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) m[e] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            q[j] = q[j] / qs;
            dbar += q[j] * dd[j];
            const float4* M = (const float4*)(ober2cano + ((int64_t)b * V + best.i[j]) * 16);
            const float4 r0 = __ldg(M), r1 = __ldg(M + 1), r2 = __ldg(M + 2);
            m[0] += q[j] * r0.x; m[1] += q[j] * r0.y; m[2] += q[j] * r0.z; m[3] += q[j] * r0.w;
            m[4] += q[j] * r1.x; m[5] += q[j] * r1.y; m[6] += q[j] * r1.z; m[7] += q[j] * r1.w;
            m[8] += q[j] * r2.x; m[9] += q[j] * r2.y; m[10] += q[j] * r2.z; m[11] += q[j] * r2.w;
        }
        valid = dbar < thr;
        xc0 = m[0] * qx + m[1] * qy + m[2] * qz + m[3];
        xc1 = m[4] * qx + m[5] * qy + m[6] * qz + m[7];
        xc2 = m[8] * qx + m[9] * qy + m[10] * qz + m[11];
    }
    u.valid = valid; u.xc0 = xc0; u.xc1 = xc1; u.xc2 = xc2;
#pragma unroll
    for (int j = 0; j < 4; ++j) { u.dd[j] = dd[j]; u.q[j] = q[j]; }
}

// Outputs of one query + the warp-aggregated compaction of the valid ids (every lane of the warp must call it).
__device__ __forceinline__ void unpose_write(const Unposed& u, const Best4& best, bool have4, int64_t gid,
                                             const UnposeOut& o, bool active)
{
    const bool valid = u.valid;
    if (active) {
        o.xyz_cano[gid * 3] = u.xc0; o.xyz_cano[gid * 3 + 1] = u.xc1; o.xyz_cano[gid * 3 + 2] = u.xc2;
        o.valid[gid] = valid ? 1 : 0;
        if (o.idx) {
            int4 v = have4 ? make_int4(best.i[0], best.i[1], best.i[2], best.i[3]) : make_int4(-1, -1, -1, -1);
            ((int4*)o.idx)[gid] = v;
        }
        if (o.dist) ((float4*)o.dist)[gid] = make_float4(u.dd[0], u.dd[1], u.dd[2], u.dd[3]);
        if (o.qw) ((float4*)o.qw)[gid] = make_float4(u.q[0], u.q[1], u.q[2], u.q[3]);
        if (!valid) {
            if (o.sigma) o.sigma[gid] = -1e5f;
            if (o.rgb) { o.rgb[gid * 3] = 0.f; o.rgb[gid * 3 + 1] = 0.f; o.rgb[gid * 3 + 2] = 0.f; }
        }
    }
    if (o.cidx) {   // warp-aggregated compaction of valid point ids
        const unsigned mask = __ballot_sync(0xffffffffu, valid);
        if (mask) {
            const int lane = threadIdx.x & 31;
            const int leader = __ffs(mask) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(o.count, __popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid) o.cidx[base + __popc(mask & ((1u << lane) - 1))] = (int32_t)gid;
        }
    }
}

__device__ __forceinline__ void unpose_epilogue(const Best4& best, bool have4, bool found, float qx, float qy, float qz,
                                                int64_t gid, int b, int V, int J,
                                                const float* __restrict__ ober2cano,
                                                const float* __restrict__ lbsw, float thr,
                                                const UnposeOut& o, bool active)
{
    Unposed u;
    unpose_compute(best, have4, found, qx, qy, qz, b, V, J, ober2cano, lbsw, thr, active, u);
    unpose_write(u, best, have4, gid, o, active);
}

// ------------------------------------------------------------------ mode 0: exhaustive
__global__ void __launch_bounds__(KNN_THREADS)
knn_unpose_brute_kernel(const float* __restrict__ xyz, const float* __restrict__ rays,
                        const float* __restrict__ z, int K, int64_t N,
                        const float* __restrict__ verts, int V, const float* __restrict__ ober2cano,
                        const float* __restrict__ lbsw, int J, float thr, UnposeOut o)
{
    extern __shared__ float4 s_v[];
    const int b = blockIdx.y;
    const float* vb = verts + (int64_t)b * V * 3;
    for (int v = threadIdx.x; v < V; v += blockDim.x)
        s_v[v] = make_float4(vb[v * 3], vb[v * 3 + 1], vb[v * 3 + 2], 0.f);
    __syncthreads();
    const int64_t per = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (N + per - 1) / per;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t n = it * per + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        const bool active = n < N;
        const int64_t gid = (int64_t)b * N + (active ? n : 0);
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (active) load_query(xyz, rays, z, gid, K, qx, qy, qz);
        Best4 best; best_init(best);
        if (active) {
#pragma unroll 4
            for (int v = 0; v < V; ++v) {
                const float4 p = s_v[v];
                best_push_ordered(best, dist2_rn(qx, qy, qz, p.x, p.y, p.z), v);
            }
        }
        unpose_epilogue(best, true, true, qx, qy, qz, gid, b, V, J, ober2cano, lbsw, thr, o, active);
    }
}

// ------------------------------------------------------------------ grid build
__global__ void __launch_bounds__(1024)
vertex_grid_build_kernel(const float* __restrict__ verts, int V, float cell_in, float flag_r, char* __restrict__ ws,
                         int64_t frame_bytes)
{
    __shared__ float s_red[6][32];
    __shared__ GridHeader s_h;
    __shared__ int s_scan[1024];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* vb = verts + (int64_t)b * V * 3;
    char* base = ws + (int64_t)b * frame_bytes;
    GridHeader* hdr = (GridHeader*)base;
    int* cell_start = (int*)(base + GRID_OFF_START);
    int* counts = (int*)(base + GRID_OFF_COUNT);
    float4* sorted = (float4*)(base + GRID_OFF_SORTED);

    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int v = tid; v < V; v += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float x = vb[v * 3 + a]; lo[a] = fminf(lo[a], x); hi[a] = fmaxf(hi[a], x); }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { s_red[a][wid] = lo[a]; s_red[3 + a][wid] = hi[a]; }
    }
    __syncthreads();
    if (tid == 0) {
        float l[3], h[3];
        for (int a = 0; a < 3; ++a) {
            l[a] = s_red[a][0]; h[a] = s_red[3 + a][0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { l[a] = fminf(l[a], s_red[a][w]); h[a] = fmaxf(h[a], s_red[3 + a][w]); }
        }
        float cell = cell_in;
        const float ext = fmaxf(h[0] - l[0], fmaxf(h[1] - l[1], h[2] - l[2]));
        if (ext / cell > (float)(AN_GRID_MAX_DIM - 1)) cell = ext / (float)(AN_GRID_MAX_DIM - 1);
        s_h.ox = l[0]; s_h.oy = l[1]; s_h.oz = l[2]; s_h.cell = cell;
        s_h.nx = min(AN_GRID_MAX_DIM, (int)floorf((h[0] - l[0]) / cell) + 1);
        s_h.ny = min(AN_GRID_MAX_DIM, (int)floorf((h[1] - l[1]) / cell) + 1);
        s_h.nz = min(AN_GRID_MAX_DIM, (int)floorf((h[2] - l[2]) / cell) + 1);
        s_h.flag_r = flag_r;
        *hdr = s_h;
    }
    __syncthreads();
    const GridHeader h = s_h;
    const int ncell = h.nx * h.ny * h.nz;
    for (int c = tid; c < ncell; c += blockDim.x) counts[c] = 0;
    __syncthreads();
    for (int v = tid; v < V; v += blockDim.x) {
        const int cx = min(h.nx - 1, max(0, (int)floorf((vb[v * 3] - h.ox) / h.cell)));
        const int cy = min(h.ny - 1, max(0, (int)floorf((vb[v * 3 + 1] - h.oy) / h.cell)));
        const int cz = min(h.nz - 1, max(0, (int)floorf((vb[v * 3 + 2] - h.oz) / h.cell)));
        atomicAdd(&counts[(cz * h.ny + cy) * h.nx + cx], 1);
    }
    __syncthreads();
    // exclusive scan over ncell counts: thread t owns cells [t*per, (t+1)*per)
    const int per = (ncell + blockDim.x - 1) / blockDim.x;
    int local = 0;
    for (int c = tid * per; c < min(ncell, (tid + 1) * per); ++c) local += counts[c];
    s_scan[tid] = local;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
        const int add = tid >= o ? s_scan[tid - o] : 0;
        __syncthreads();
        s_scan[tid] += add;
        __syncthreads();
    }
    int run = s_scan[tid] - local;
    for (int c = tid * per; c < min(ncell, (tid + 1) * per); ++c) { cell_start[c] = run; run += counts[c]; }
    if (tid == 0) cell_start[ncell] = V;
    __syncthreads();
    for (int v = tid; v < V; v += blockDim.x) {
        const float x = vb[v * 3], y = vb[v * 3 + 1], zc = vb[v * 3 + 2];
        const int cx = min(h.nx - 1, max(0, (int)floorf((x - h.ox) / h.cell)));
        const int cy = min(h.ny - 1, max(0, (int)floorf((y - h.oy) / h.cell)));
        const int cz = min(h.nz - 1, max(0, (int)floorf((zc - h.oz) / h.cell)));
        const int c = (cz * h.ny + cy) * h.nx + cx;
        const int pos = cell_start[c] + atomicSub(&counts[c], 1) - 1;
        sorted[pos] = make_float4(x, y, zc, __int_as_float(v));
    }
}

// Validity pre-filter for every cell of the extended grid (grid dilated by GRID_R cells): can a query inside this cell have
// a vertex within the threshold?  flag_r > 0: exact test on the cell's BOX -- some vertex of the neighbourhood lies within
// flag_r of the box (a query q in the box with |q - v| < flag_r implies that, so a zero flag proves d_min >= flag_r for
// every query of the cell).  flag_r <= 0: the coarser test "any vertex in the 7^3-cell neighbourhood" (valid for
// thresholds up to GRID_R cells).  One warp per extended cell, all frames in one launch (grid.y = frame).
__global__ void __launch_bounds__(256)
vertex_grid_flags_kernel(char* __restrict__ ws, int64_t frame_bytes)
{
    char* base = ws + (int64_t)blockIdx.y * frame_bytes;
    const GridHeader h = *(const GridHeader*)base;
    const int* __restrict__ cell_start = (const int*)(base + GRID_OFF_START);
    const float4* __restrict__ sorted = (const float4*)(base + GRID_OFF_SORTED);
    uint8_t* flags = (uint8_t*)(base + GRID_OFF_FLAGS);
    const int ex_n = h.nx + 2 * GRID_R, ey_n = h.ny + 2 * GRID_R, ez_n = h.nz + 2 * GRID_R;
    const int lane = threadIdx.x & 31;
    const float flag_r = h.flag_r;
    const float rr = flag_r * 1.0001f + 1e-6f, r2 = rr * rr;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    // one warp per extended cell (lanes share its candidates), warps stride over the cells of the frame
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < ex_n * ey_n * ez_n; e += n_warps) {
    const int ex = e % ex_n, ey = (e / ex_n) % ey_n, ez = e / (ex_n * ey_n);
    // extended index = grid index + GRID_R; neighbourhood of grid cell g is [g-R, g+R] = [e-2R, e]
    const int x0 = max(ex - 2 * GRID_R, 0), x1 = min(ex, h.nx - 1);
    const float bx0 = h.ox + (float)(ex - GRID_R) * h.cell, by0 = h.oy + (float)(ey - GRID_R) * h.cell,
                bz0 = h.oz + (float)(ez - GRID_R) * h.cell;
    bool any = false;                                                     // warp-uniform
    if (x0 <= x1)
        for (int rbase = 0; rbase < (2 * GRID_R + 1) * (2 * GRID_R + 1) && !any; rbase += 32) {
            // every lane looks up one (gz, gy) row of the neighbourhood (its two loads run in parallel across the warp) ...
            const int r = rbase + lane;
            const int gz = ez - 2 * GRID_R + r / (2 * GRID_R + 1), gy = ey - 2 * GRID_R + r % (2 * GRID_R + 1);
            int s0 = 0, e1 = 0;
            if (r < (2 * GRID_R + 1) * (2 * GRID_R + 1) && gz >= 0 && gz < h.nz && gy >= 0 && gy < h.ny) {
                // gap between this row's (y,z) cell and the box: whole cells in between
                const float gy_ = fmaxf(0.f, (float)(abs(gy - (ey - GRID_R)) - 1)) * h.cell;
                const float gz_ = fmaxf(0.f, (float)(abs(gz - (ez - GRID_R)) - 1)) * h.cell;
                if (!(flag_r > 0.f) || gy_ * gy_ + gz_ * gz_ < r2) {
                    const int row = (gz * h.ny + gy) * h.nx;
                    s0 = __ldg(cell_start + row + x0); e1 = __ldg(cell_start + row + x1 + 1);
                }
            }
            unsigned rows = __ballot_sync(0xffffffffu, e1 > s0);
            if (!(flag_r > 0.f)) { any = rows != 0u; continue; }
            // ... then the warp scans the vertices of the non-empty rows, 32 at a time, until one lies within the radius
            while (rows && !any) {
                const int src = __ffs(rows) - 1;
                rows &= rows - 1;
                const int rs0 = __shfl_sync(0xffffffffu, s0, src), re1 = __shfl_sync(0xffffffffu, e1, src);
                for (int p0 = rs0; p0 < re1 && !any; p0 += 32) {
                    bool hit = false;
                    if (p0 + lane < re1) {
                        const float4 v = __ldg(sorted + p0 + lane);
                        const float dx = fmaxf(fmaxf(bx0 - v.x, v.x - (bx0 + h.cell)), 0.f);
                        const float dy = fmaxf(fmaxf(by0 - v.y, v.y - (by0 + h.cell)), 0.f);
                        const float dz = fmaxf(fmaxf(bz0 - v.z, v.z - (bz0 + h.cell)), 0.f);
                        hit = dx * dx + dy * dy + dz * dz < r2;
                    }
                    any = __any_sync(0xffffffffu, hit);
                }
            }
        }
    if (lane == 0) flags[e] = (uint8_t)any;
    }
}

// ------------------------------------------------------------------ mode 1: grid-pruned
// Two kernels.
//  knn_classify_kernel  one thread per query: generate the point, look up the dilated occupancy
//      flag of its cell (cell ~ 1.25*dis_threshold/3, search box 7x7x7 cells).  A query with no
//      vertex within the box radius is invalid *exactly* (valid needs d_min < threshold, SURVEY
//      App. A): it gets its final outputs here.  Every other query is appended (x,y,z,id) to a
//      compact work list, 32 consecutive queries per warp-aggregated append, so consecutive list
//      entries are (almost always) consecutive samples of one ray.
//      Seeded pass (fine pass of VolumeRenderer.forward, models/volume_rendering.py:199-207): half of
//      the sorted samples ARE coarse samples (same ray, same depth bits => the same point), so their
//      4-NN are read back from the coarse pass's idx table, the four distances re-evaluated (same
//      arithmetic => same bits) and the epilogue runs right here, without any search.
//  knn_search_kernel    one thread per listed query, a warp pulling 32 consecutive entries at a time.
//      Every lane first gets an upper bound on its 4th-neighbour distance:
//        warm   (seeded pass) the 4-NN of the nearest coarse sample of the same ray are four real
//               vertices: the largest of their distances to this query bounds d4 -- typically within
//               half a coarse step (1.5 cm) of the truth;
//        cold   the query's own cell, scanned first, usually yields four candidates;
//        then the bounds travel along the warp by the triangle inequality
//               d4(q) <= d4(q') + |q - q'| in doubling strides (a min-plus scan in both directions).
//      The exact search is then a ball-pruned walk of the 7x7 rows of the cell box, nearest ring first:
//      a row is skipped when its slab is farther than the bound, its x range is trimmed to the ball,
//      and the bound drops to the current 4th-best distance as candidates arrive.
// Exactness: everything skipped is provably farther than the final 4th neighbour whenever that
// neighbour lies within the query's initial bound and the box (margins absorb fp32 rounding);
// otherwise the warp rescans the whole table cooperatively.  The (d2,index) order key makes the
// result independent of the scan order: both modes return identical bits.
struct QueryWs { unsigned int n_work; unsigned int next_chunk; unsigned int pad[2]; unsigned long long stats[4]; };   // then float4 work[B*N]

// Seeds from an earlier pass over the same rays (all NULL/0 when absent): src[g] < Kc says query g is
// coarse sample src[g] of its ray, nn[g] is the coarse sample nearest in depth, idx is the earlier
// pass's (B*R*Kc, 4) neighbour table.  With xc / valid (and qw when this pass emits qw) the results of a shared sample
// are copied from the earlier pass instead of being re-blended (same inputs, same arithmetic: same bits).
struct SeedIn { const uint8_t* src; const uint8_t* nn; const int32_t* idx; int Kc;
                const float* xc; const uint8_t* valid; const float* qw; };      // earlier pass's xyz_cano / valid / qw (optional)

__device__ __forceinline__ void warp_merge4(Best4& lb, Best4& g)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned kd = __float_as_uint(lb.d[0]);                  // d2 >= 0: uint order == float order
        const unsigned mn = __reduce_min_sync(0xffffffffu, kd);
        const unsigned ci = (kd == mn) ? (unsigned)lb.i[0] : 0xffffffffu;
        const unsigned mi = __reduce_min_sync(0xffffffffu, ci);
        g.d[k] = __uint_as_float(mn); g.i[k] = (int)mi;
        if (kd == mn && (unsigned)lb.i[0] == mi) {
            lb.d[0] = lb.d[1]; lb.i[0] = lb.i[1]; lb.d[1] = lb.d[2]; lb.i[1] = lb.i[2];
            lb.d[2] = lb.d[3]; lb.i[2] = lb.i[3]; lb.d[3] = CUDART_INF_F; lb.i[3] = 0x7fffffff;
        }
    }
}

// Rows (dy,dz) of the 7x7 cell box, nearest ring first: ring k = max(|dy|,|dz|).
__constant__ int8_t c_row_dy[49] = {0, -1, 0, 1, -1, 1, -1, 0, 1, -2, -1, 0, 1, 2, -2, 2, -2, 2, -2, 2, -2, -1, 0, 1, 2, -3, -2, -1, 0, 1, 2, 3, -3, 3, -3, 3, -3, 3, -3, 3, -3, 3, -3, -2, -1, 0, 1, 2, 3};
__constant__ int8_t c_row_dz[49] = {0, -1, -1, -1, 0, 0, 1, 1, 1, -2, -2, -2, -2, -2, -1, -1, 0, 0, 1, 1, 2, 2, 2, 2, 2, -3, -3, -3, -3, -3, -3, -3, -2, -2, -1, -1, 0, 0, 1, 1, 2, 2, 3, 3, 3, 3, 3, 3, 3};
__constant__ int8_t c_row_ring[49] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3};
static_assert(GRID_R == 3, "row tables are generated for a 7x7 box");
#define N_ROWS ((2 * GRID_R + 1) * (2 * GRID_R + 1))

// the four seed vertices of a query: distances re-evaluated with the contract's arithmetic
__device__ __forceinline__ void seed_distances(const float* __restrict__ vb, const int4 si, float qx, float qy, float qz,
                                               float d[4])
{
    const int ids[4] = {si.x, si.y, si.z, si.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float* v = vb + (int64_t)ids[j] * 3;
        d[j] = dist2_rn(qx, qy, qz, __ldg(v), __ldg(v + 1), __ldg(v + 2));
    }
}

#ifndef AN_KNN_CLS_MINB
#define AN_KNN_CLS_MINB 4
#endif
__global__ void __launch_bounds__(KNN_THREADS, AN_KNN_CLS_MINB)
knn_classify_kernel(const float* __restrict__ xyz, const float* __restrict__ rays, const float* __restrict__ z,
                    int K, int64_t N, int64_t total, const float* __restrict__ verts, int V,
                    const char* __restrict__ ws, int64_t frame_bytes, QueryWs* __restrict__ qws,
                    const float* __restrict__ ober2cano, const float* __restrict__ lbsw, int J, float thr,
                    SeedIn sd, UnposeOut o)
{
    float4* __restrict__ work = (float4*)(qws + 1);
    const int lane = threadIdx.x & 31;
    const float thr2 = thr * thr * (1.0f + 1e-5f);
    // shared samples take the seeding pass's result as it is when that pass handed over everything this pass emits
    const bool copy = sd.idx && sd.xc && sd.valid && !o.dist && (!o.qw || sd.qw);
    for (int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~31ll; g0 < total; g0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gid = g0 + lane;
        const bool active = gid < total;
        bool maybe = false, same = false, have4 = false, found = false;
        float qx = 0.f, qy = 0.f, qz = 0.f;
        Best4 sb; best_init(sb);
        const int b = active ? (int)(gid / N) : 0;
        int64_t cg = 0;                                   // shared sample: its index in the seeding pass
        if (active) {
            load_query(xyz, rays, z, gid, K, qx, qy, qz);
            if (sd.idx) {
                const int s = sd.src[gid];
                if (s < sd.Kc) {          // this sample IS coarse sample s of its ray: reuse that pass's neighbours
                    same = true;
                    cg = (gid / K) * sd.Kc + s;
                    // (a copied sample needs its neighbour record only when this pass emits indices: every inference
                    //  frame skips this load -- a third of the kernel's reads and one level of its dependent-load chain)
                    const int4 si = (copy && !o.idx) ? make_int4(-1, -1, -1, -1) : __ldg((const int4*)sd.idx + cg);
                    if (si.x >= 0) {
                        have4 = true;
                        sb.i[0] = si.x; sb.i[1] = si.y; sb.i[2] = si.z; sb.i[3] = si.w;
                        if (!copy) {
                            seed_distances(verts + (int64_t)b * V * 3, si, qx, qy, qz, sb.d);
                            found = sb.d[0] < thr2;
                        }
                    }
                }
            }
            if (!same) {
                const GridHeader h = *(const GridHeader*)(ws + (int64_t)b * frame_bytes);
                const uint8_t* __restrict__ flags = (const uint8_t*)(ws + (int64_t)b * frame_bytes + GRID_OFF_FLAGS);
                const float inv_cell = 1.0f / h.cell;
                const int ex = (int)floorf((qx - h.ox) * inv_cell) + GRID_R, ey = (int)floorf((qy - h.oy) * inv_cell) + GRID_R,
                          ez = (int)floorf((qz - h.oz) * inv_cell) + GRID_R;
                const int ex_n = h.nx + 2 * GRID_R, ey_n = h.ny + 2 * GRID_R, ez_n = h.nz + 2 * GRID_R;
                maybe = ex >= 0 && ex < ex_n && ey >= 0 && ey < ey_n && ez >= 0 && ez < ez_n;
                // (flags built for a smaller radius than this call's threshold cannot reject anything)
                if (maybe && !(h.flag_r > 0.f && thr > h.flag_r)) maybe = __ldg(flags + ((int64_t)ez * ey_n + ey) * ex_n + ex) != 0;
                if (!maybe) {            // no vertex within the box radius (>= threshold): final outputs of an invalid point
                    o.xyz_cano[gid * 3] = 0.f; o.xyz_cano[gid * 3 + 1] = 0.f; o.xyz_cano[gid * 3 + 2] = 0.f;
                    o.valid[gid] = 0;
                    if (o.idx) ((int4*)o.idx)[gid] = make_int4(-1, -1, -1, -1);
                    if (o.dist) ((float4*)o.dist)[gid] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (o.qw) ((float4*)o.qw)[gid] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (o.sigma) o.sigma[gid] = -1e5f;
                    if (o.rgb) { o.rgb[gid * 3] = 0.f; o.rgb[gid * 3 + 1] = 0.f; o.rgb[gid * 3 + 2] = 0.f; }
                }
            }
        }
        if (sd.idx) {    // (warp-uniform) samples shared with the seeding pass finish here
            Unposed u;
            if (copy) {  // ... with that pass's own result for the same point
                u.valid = false; u.xc0 = u.xc1 = u.xc2 = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) { u.dd[j] = 0.f; u.q[j] = 0.f; }
                if (active && same) {
                    u.valid = sd.valid[cg] != 0;
                    u.xc0 = sd.xc[cg * 3]; u.xc1 = sd.xc[cg * 3 + 1]; u.xc2 = sd.xc[cg * 3 + 2];
                    if (o.qw) { const float4 q4 = __ldg((const float4*)sd.qw + cg); u.q[0] = q4.x; u.q[1] = q4.y; u.q[2] = q4.z; u.q[3] = q4.w; }
                }
            } else
                unpose_compute(sb, have4, found, qx, qy, qz, b, V, J, ober2cano, lbsw, thr, active && same, u);
            unpose_write(u, sb, have4, gid, o, active && same);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, maybe);
        if (mask) {
            const int leader = __ffs(mask) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&qws->n_work, (unsigned)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (maybe) work[base + __popc(mask & ((1u << lane) - 1))] = make_float4(qx, qy, qz, __int_as_float((int)gid));
        }
    }
}

// The grid search keeps its four best as packed 64-bit keys, (bits of d2) << 32 | vertex index: d2 >= +0 and never NaN
// for finite inputs, so unsigned order of the bits == float order and ONE 64-bit compare is the (d2, index) order of the
// contract above (a NaN d2 packs above the +inf sentinel and is never taken, as with the float compare).
struct Key4 { unsigned long long k[4]; };
__device__ __forceinline__ unsigned long long key_pack(float d2, int idx) {
    return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)idx;
}
__device__ __forceinline__ float key_d2(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ void key_init(Key4& b) {
#pragma unroll
    for (int j = 0; j < 4; ++j) b.k[j] = key_pack(CUDART_INF_F, 0x7fffffff);
}
// Insert a key known to beat the current 4th entry: overwrite it, then three compare-exchanges restore the order --
// no data-dependent branches inside.
__device__ __forceinline__ void key_insert_sorted(Key4& b, unsigned long long key) {
    b.k[3] = key;
#pragma unroll
    for (int k = 3; k > 0; --k) {
        const bool sw = b.k[k] < b.k[k - 1];
        const unsigned long long lo = sw ? b.k[k] : b.k[k - 1], hi = sw ? b.k[k - 1] : b.k[k];
        b.k[k - 1] = lo; b.k[k] = hi;
    }
}
// A vertex may be met twice (a seed, then again in its cell): equal keys never pass the strict
// comparison against slot 3, slots 0-2 are checked by index (the key's low word).
__device__ __forceinline__ bool key_accepts(const Key4& b, unsigned long long key) {
    const unsigned i = (unsigned)key;
    return key < b.k[3] && i != (unsigned)b.k[0] && i != (unsigned)b.k[1] && i != (unsigned)b.k[2];
}
__device__ __forceinline__ void key_unpack(const Key4& b, Best4& o) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { o.d[j] = key_d2(b.k[j]); o.i[j] = (int)(unsigned)b.k[j]; }
}

// One trip of a lane's candidate walk: up to four table entries [p, min(p + 4, pe)) (all loads in flight before the
// first use), then a per-lane loop that takes the trip's closest remaining candidate while it can still enter the
// list: the insert code runs once per ACCEPTED candidate of the busiest lane instead of once per candidate slot, and
// every lane inside the loop is inserting.  `dedup`: the list may already hold some of these vertices (seeds).
// Returns true when the list changed (the caller tightens its bound).
template <bool DEDUP>
__device__ __forceinline__ bool key_trip4(Key4& mine, const float4* __restrict__ sorted, int p, int pe,
                                          float qx, float qy, float qz)
{
    const int last = pe - 1;
    const float4 v0 = __ldg(sorted + p), v1 = __ldg(sorted + min(p + 1, last)),
                 v2 = __ldg(sorted + min(p + 2, last)), v3 = __ldg(sorted + min(p + 3, last));
    float d0 = dist2_rn(qx, qy, qz, v0.x, v0.y, v0.z);
    float d1 = p + 1 < pe ? dist2_rn(qx, qy, qz, v1.x, v1.y, v1.z) : CUDART_INF_F;
    float d2 = p + 2 < pe ? dist2_rn(qx, qy, qz, v2.x, v2.y, v2.z) : CUDART_INF_F;
    float d3 = p + 3 < pe ? dist2_rn(qx, qy, qz, v3.x, v3.y, v3.z) : CUDART_INF_F;
    float lim = key_d2(mine.k[3]);
    float m = fminf(fminf(d0, d1), fminf(d2, d3));
    if (!(m <= lim)) return false;             // cheap reject first: almost every trip fails it late in the walk
    do {
        const bool s0 = d0 == m, s1 = !s0 && d1 == m, s2 = !s0 && !s1 && d2 == m;
        const float vw = s0 ? v0.w : (s1 ? v1.w : (s2 ? v2.w : v3.w));
        d0 = s0 ? CUDART_INF_F : d0; d1 = s1 ? CUDART_INF_F : d1; d2 = s2 ? CUDART_INF_F : d2;
        d3 = (s0 || s1 || s2) ? d3 : CUDART_INF_F;
        const unsigned long long key = key_pack(m, __float_as_int(vw));
        if (DEDUP ? key_accepts(mine, key) : key < mine.k[3]) key_insert_sorted(mine, key);
        lim = key_d2(mine.k[3]);
        m = fminf(fminf(d0, d1), fminf(d2, d3));
    } while (m <= lim && m < CUDART_INF_F);
    return true;
}

#define SEARCH_THREADS 128

// One row of the cell box for one query: slab gap^2 and the candidate range [s0, e1) of the cell-sorted
// table after trimming the x extent to the ball of radius sqrt(Bm).  Returns false when nothing is left.
struct RowCtx { float qx, fy, fz, ox, cell, inv_cell, Bm; int cx, cy, cz, nx, ny, nz; };
__device__ __forceinline__ bool row_range(const RowCtx& c, int r, const int* __restrict__ cell_start,
                                          float& g2, int& row, int& x0, int& x1)
{
    const int dy = c_row_dy[r], dz = c_row_dz[r];
    const int gy = c.cy + dy, gz = c.cz + dz;
    // distance from the query to the row's slab per axis: below it, above it, or inside (0) -- branch-free
    const float gapy = fmaxf(fmaxf((float)dy * c.cell - c.fy, c.fy - (float)(dy + 1) * c.cell), 0.f);
    const float gapz = fmaxf(fmaxf((float)dz * c.cell - c.fz, c.fz - (float)(dz + 1) * c.cell), 0.f);
    g2 = gapy * gapy + gapz * gapz;
    if (!(gz >= 0 && gz < c.nz && gy >= 0 && gy < c.ny && g2 <= c.Bm)) return false;
    const float w = sqrtf(fmaxf(c.Bm - g2, 0.f)) + 1e-3f * c.cell;
    x0 = max(c.cx - GRID_R, (int)floorf((c.qx - w - c.ox) * c.inv_cell));
    x1 = min(c.cx + GRID_R, (int)floorf((c.qx + w - c.ox) * c.inv_cell));
    x0 = max(x0, 0); x1 = min(x1, c.nx - 1);
    row = (gz * c.ny + gy) * c.nx;
    return x0 <= x1;
}

// Rows are listed in lockstep into a CAP-entry shared-memory list per thread, then one flat loop in which every
// lane that still has a candidate evaluates it; the list is refilled (with the tightened bound) until the rows
// run out.  (A variant with nested per-lane loops and no shared memory was measured and dropped: 1.6x slower.)
// 8 CTAs of 128 threads per SM (64 registers): the search is bound by L2 gather latency (69 % long-scoreboard
// stalls at 24 resident warps), 32 resident warps hide more of it: 0.71 -> 0.60 ms coarse, 0.82 -> 0.69 ms fine pass
#ifndef AN_KNN_MINB
#define AN_KNN_MINB 8
#endif
#ifndef SEARCH_CAP
#define SEARCH_CAP 16          // row-list entries per thread and round (8 and 50 measured: within 3 %)
#endif
#ifndef SEARCH_DRAIN
#define SEARCH_DRAIN 8         // lanes left at which the warp drains the remaining lists cooperatively
#endif
#ifdef AN_KNN_STATS            // candidate statistics for tools/bench_knn.py (variant build only)
#define KNN_STAT(x) x
#else
#define KNN_STAT(x)
#endif
__global__ void __launch_bounds__(SEARCH_THREADS, AN_KNN_MINB)
knn_search_kernel(int K, int64_t N, const float* __restrict__ verts, int V, const char* __restrict__ ws,
                  int64_t frame_bytes, QueryWs* __restrict__ qws, const float* __restrict__ ober2cano,
                  const float* __restrict__ lbsw, int J, float thr, SeedIn sd, UnposeOut o)
{
    constexpr int CAP = SEARCH_CAP;
    // per-thread list of candidate ranges, layout [entry][thread] (bank = thread: conflict-free for any
    // per-lane entry index): .x = start | end << 16 (positions in the cell-sorted table), .y = row slab gap^2
    extern __shared__ uint2 s_ent[];
    const float4* __restrict__ work = (const float4*)(qws + 1);
    const int lane = threadIdx.x & 31;
    const unsigned n_work = qws->n_work;
    const unsigned n_chunks = (n_work + 31) / 32;
    const float thr2 = thr * thr * (1.0f + 1e-5f);   // prune only what is invalid beyond rounding doubt
    uint2* __restrict__ my_ent = s_ent + threadIdx.x;
    KNN_STAT(unsigned n_cand = 0; unsigned n_iter = 0; unsigned n_redo = 0;)
    for (;;) {
        unsigned chunk = 0;
        if (lane == 0) chunk = atomicAdd(&qws->next_chunk, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n_chunks) break;
        const unsigned e = chunk * 32 + lane;
        const bool active = e < n_work;
        float qx = 0.f, qy = 0.f, qz = 0.f;
        int64_t gid = 0;
        if (active) { const float4 w = __ldg(work + e); qx = w.x; qy = w.y; qz = w.z; gid = (int64_t)__float_as_int(w.w); }
        const int b = (int)(gid / N);
        const char* base = ws + (int64_t)b * frame_bytes;
        const GridHeader h = *(const GridHeader*)base;
        const int* __restrict__ cell_start = (const int*)(base + GRID_OFF_START);
        const float4* __restrict__ sorted = (const float4*)(base + GRID_OFF_SORTED);
        const float inv_cell = 1.0f / h.cell;
        const float box_r = GRID_R * h.cell * (1.0f - 1e-4f);
        const float box_r2 = box_r * box_r;
        // cold bound: slightly beyond the threshold so that a 4th neighbour a little farther than a
        // valid nearest one (d0 < thr <= d4) is still found without the exhaustive fallback
        float B = fminf(box_r2, thr * thr * (KNN_B0_SCALE * KNN_B0_SCALE));
        const int cx = (int)floorf((qx - h.ox) * inv_cell), cy = (int)floorf((qy - h.oy) * inv_cell),
                  cz = (int)floorf((qz - h.oz) * inv_cell);
        const bool own = active && cx >= 0 && cx < h.nx && cy >= 0 && cy < h.ny && cz >= 0 && cz < h.nz;
        Key4 mine; key_init(mine);
        bool warm = false;
        if (active && sd.idx) {       // seeds: the 4-NN of the nearest coarse sample of this ray
            const int4 si = __ldg((const int4*)sd.idx + (gid / K) * sd.Kc + sd.nn[gid]);
            if (si.x >= 0) {
                warm = true;
                float d[4];
                seed_distances(verts + (int64_t)b * V * 3, si, qx, qy, qz, d);
                key_insert_sorted(mine, key_pack(d[0], si.x)); key_insert_sorted(mine, key_pack(d[1], si.y));
                key_insert_sorted(mine, key_pack(d[2], si.z)); key_insert_sorted(mine, key_pack(d[3], si.w));
            }
        }
        const bool did_own = own && !warm;
        {   // cold lanes -- own cell: usually yields four candidates and a tight bound
            int p = 0, pe = 0;
            if (did_own) { const int c = (cz * h.ny + cy) * h.nx + cx; p = __ldg(cell_start + c); pe = __ldg(cell_start + c + 1); }
            KNN_STAT(n_cand += (unsigned)(pe - p);)
            for (; p < pe; p += 4) key_trip4<false>(mine, sorted, p, pe, qx, qy, qz);
        }
        {   // bounds travel along the warp (neighbouring lanes are neighbouring samples of a ray):
            // d4(q) <= d4(q') + |q - q'|, doubling strides in both directions
            float u = active ? sqrtf(fminf(key_d2(mine.k[3]), B)) : CUDART_INF_F;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
#pragma unroll
                for (int dir = -1; dir <= 1; dir += 2) {
                    const int src = lane + dir * s;
                    const int srcc = min(31, max(0, src));
                    const float ou = __shfl_sync(0xffffffffu, u, srcc);
                    const float ox = __shfl_sync(0xffffffffu, qx, srcc), oy = __shfl_sync(0xffffffffu, qy, srcc), oz = __shfl_sync(0xffffffffu, qz, srcc);
                    const int ob = __shfl_sync(0xffffffffu, b, srcc);
                    if (src == srcc && ob == b) {           // inactive lanes carry u = inf
                        const float sx = qx - ox, sy = qy - oy, sz = qz - oz;
                        u = fminf(u, (ou + sqrtf(sx * sx + sy * sy + sz * sz)) * (1.0f + 1e-4f));
                    }
                }
            }
            B = fminf(B, u * u * (1.0f + 1e-4f));
        }
        float Bm = fminf(fminf(B, key_d2(mine.k[3])) * 1.001f, box_r2 * 1.01f);
        RowCtx rc;
        rc.qx = qx; rc.fy = qy - (h.oy + cy * h.cell); rc.fz = qz - (h.oz + cz * h.cell);
        rc.ox = h.ox; rc.cell = h.cell; rc.inv_cell = inv_cell; rc.cx = cx; rc.cy = cy; rc.cz = cz;
        rc.nx = h.nx; rc.ny = h.ny; rc.nz = h.nz;
        {
            int r = 0;                                // warp-uniform row cursor
            bool done = !active;                      // this lane's ball ends before the current ring
            for (;;) {
                // list the next rows that intersect the ball (at most CAP entries per lane per round)
                int n_ent = 0;
                for (; r < N_ROWS; ++r) {
                    if (__any_sync(0xffffffffu, n_ent > CAP - 2)) break;
                    const int ring = c_row_ring[r];
                    if (ring >= 2) { const float rg = (float)(ring - 1) * h.cell; if (rg * rg > Bm) done = true; }
                    if (!__any_sync(0xffffffffu, !done)) { r = N_ROWS; break; }
                    float g2; int row, x0, x1;
                    rc.Bm = Bm;
                    if (!done && row_range(rc, r, cell_start, g2, row, x0, x1)) {
                        const int s0 = __ldg(cell_start + row + x0), e1 = __ldg(cell_start + row + x1 + 1);
                        if (r == 0 && did_own && cx >= x0 && cx <= x1) {     // the own cell was scanned above: split around it
                            const int e0 = __ldg(cell_start + row + cx), s1 = __ldg(cell_start + row + cx + 1);
                            if (e0 > s0) { my_ent[n_ent * SEARCH_THREADS] = make_uint2((unsigned)s0 | ((unsigned)e0 << 16), __float_as_uint(g2)); ++n_ent; }
                            if (e1 > s1) { my_ent[n_ent * SEARCH_THREADS] = make_uint2((unsigned)s1 | ((unsigned)e1 << 16), __float_as_uint(g2)); ++n_ent; }
                        } else if (e1 > s0) {
                            my_ent[n_ent * SEARCH_THREADS] = make_uint2((unsigned)s0 | ((unsigned)e1 << 16), __float_as_uint(g2)); ++n_ent;
                        }
                    }
                }
                // one flat loop over the listed candidates: a lane whose range is used up pops its next entry
                // (rows that fell outside the shrunken ball are skipped whole), then every lane that still has
                // a candidate evaluates it -- the distance / insert code runs convergent across the warp.
                int k = 0, p = 0, pe = 0;
                bool live = n_ent > 0;
                for (;;) {
                    if (live && p >= pe) {
                        live = false;
                        while (k < n_ent) {
                            const uint2 en = my_ent[k * SEARCH_THREADS]; ++k;
                            if (__uint_as_float(en.y) <= Bm) { p = (int)(en.x & 0xffffu); pe = (int)(en.x >> 16); live = true; break; }
                        }
                    }
                    const unsigned live_mask = __ballot_sync(0xffffffffu, live);
                    if (__popc(live_mask) <= SEARCH_DRAIN) break;      // few lanes left: drain them cooperatively below
                    KNN_STAT(++n_iter;)
                    if (live) {
                        KNN_STAT(n_cand += (unsigned)min(4, pe - p);)
                        if (key_trip4<true>(mine, sorted, p, pe, qx, qy, qz)) Bm = fminf(Bm, key_d2(mine.k[3]) * 1.001f);
                        p += 4;
                    }
                }
                // The lanes still holding candidates (long lists: far queries with a wide ball) are drained one
                // at a time by the whole warp: 32 consecutive table entries per step (coalesced), the step's
                // minimum found by REDUX and handed to the owning lane while it beats that lane's 4th best.
                __syncwarp();     // the drain reads OTHER lanes' row lists from shared memory: order it after their writes
                unsigned heavy = __ballot_sync(0xffffffffu, live);
                while (heavy) {
                    const int L = __ffs(heavy) - 1;
                    heavy &= heavy - 1;
                    const float ux = __shfl_sync(0xffffffffu, qx, L), uy = __shfl_sync(0xffffffffu, qy, L), uz = __shfl_sync(0xffffffffu, qz, L);
                    int up = __shfl_sync(0xffffffffu, p, L), upe = __shfl_sync(0xffffffffu, pe, L), uk = __shfl_sync(0xffffffffu, k, L);
                    const int un = __shfl_sync(0xffffffffu, n_ent, L);
                    float uBm = __shfl_sync(0xffffffffu, Bm, L), bd3 = __shfl_sync(0xffffffffu, key_d2(mine.k[3]), L);
                    const int ub = __shfl_sync(0xffffffffu, b, L);
                    const float4* __restrict__ usorted = (const float4*)(ws + (int64_t)ub * frame_bytes + GRID_OFF_SORTED);
                    const uint2* __restrict__ uent = s_ent + (threadIdx.x & ~31) + L;
                    for (;;) {
                        if (up >= upe) {                          // (warp-uniform) next listed range still inside the ball
                            bool got = false;
                            while (uk < un) {
                                const uint2 en = uent[uk * SEARCH_THREADS]; ++uk;
                                if (__uint_as_float(en.y) <= uBm) { up = (int)(en.x & 0xffffu); upe = (int)(en.x >> 16); got = true; break; }
                            }
                            if (!got) break;
                        }
                        const int pos = up + lane;
                        const bool in = pos < upe;
                        up += 32;
                        KNN_STAT(++n_iter; n_cand += in ? 1u : 0u;)
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (in) v = __ldg(usorted + pos);
                        unsigned key = in ? __float_as_uint(dist2_rn(ux, uy, uz, v.x, v.y, v.z)) : 0x7f800000u;   // d2 >= 0: uint order == float order
                        for (;;) {
                            const unsigned mn = __reduce_min_sync(0xffffffffu, key);
                            if (mn >= 0x7f800000u || __uint_as_float(mn) > bd3) break;    // ties with the 4th best go to the exact test
                            const int w = __ffs(__ballot_sync(0xffffffffu, key == mn)) - 1;
                            const int vi = __shfl_sync(0xffffffffu, __float_as_int(v.w), w);
                            const unsigned long long k64 = ((unsigned long long)mn << 32) | (unsigned)vi;
                            if (lane == L && key_accepts(mine, k64)) key_insert_sorted(mine, k64);
                            if (lane == w) key = 0x7f800000u;
                            bd3 = __shfl_sync(0xffffffffu, key_d2(mine.k[3]), L);
                        }
                        uBm = fminf(uBm, bd3 * 1.001f);
                    }
                    if (lane == L) Bm = fminf(Bm, key_d2(mine.k[3]) * 1.001f);
                }
                if (r >= N_ROWS) break;
                __syncwarp();     // ... and the next round's list writes after those reads
            }
        }
        // the walk saw every vertex within sqrt(min(B, box_r2)): four of them => the 4-NN are exact
        Best4 best;
        key_unpack(mine, best);
        const bool exact4 = active && best.d[3] <= B && best.d[3] <= box_r2;
        const bool near_ = active && best.d[0] < thr2;          // may be valid: needs the exact 4-NN
        bool have4 = exact4;
        const bool redo = near_ && !exact4;            // rare: 4th neighbour not provably inside the scanned ball
        unsigned redo_mask = __ballot_sync(0xffffffffu, redo);
        KNN_STAT(n_redo += redo ? 1 : 0;)
        while (redo_mask) {                            // exhaustive rescan, warp-cooperative
            const int qi = __ffs(redo_mask) - 1;
            redo_mask &= redo_mask - 1;
            const float ux = __shfl_sync(0xffffffffu, qx, qi), uy = __shfl_sync(0xffffffffu, qy, qi), uz = __shfl_sync(0xffffffffu, qz, qi);
            const int ub = __shfl_sync(0xffffffffu, b, qi);
            const float4* __restrict__ srt = (const float4*)(ws + (int64_t)ub * frame_bytes + GRID_OFF_SORTED);
            Best4 lb; best_init(lb);
            for (int p = lane; p < V; p += 32) {
                const float4 v = __ldg(srt + p);
                best_push_any(lb, dist2_rn(ux, uy, uz, v.x, v.y, v.z), __float_as_int(v.w));
            }
            Best4 g;
            warp_merge4(lb, g);
            if (lane == qi) { best = g; have4 = true; }
        }
        unpose_epilogue(best, have4, have4 && near_, qx, qy, qz, gid, b, V, J, ober2cano, lbsw, thr, o, active);
    }
    KNN_STAT(atomicAdd(&qws->stats[0], (unsigned long long)n_cand); atomicAdd(&qws->stats[1], (unsigned long long)n_iter);
             atomicAdd(&qws->stats[2], (unsigned long long)n_redo);)
}

// ------------------------------------------------------------------ backward
__global__ void __launch_bounds__(256)
knn_unpose_bwd_kernel(const float* __restrict__ g_xc, const int32_t* __restrict__ cidx,
                      const int32_t* __restrict__ count, const float* __restrict__ xyz,
                      const float* __restrict__ rays, const float* __restrict__ z, int K, int64_t N,
                      int V, const int32_t* __restrict__ idx, const float* __restrict__ qw,
                      const float* __restrict__ ober2cano, float* __restrict__ g_o2c,
                      float* __restrict__ g_xyz)
{
    const int n_valid = *count;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_valid; p += gridDim.x * blockDim.x) {
        const int64_t gid = cidx[p];
        const int b = (int)(gid / N);
        float qx, qy, qz;
        load_query(xyz, rays, z, gid, K, qx, qy, qz);
        const float g0 = g_xc[gid * 3], g1 = g_xc[gid * 3 + 1], g2 = g_xc[gid * 3 + 2];
        const int4 nb = ((const int4*)idx)[gid];
        const float4 q4 = ((const float4*)qw)[gid];
        const int ni[4] = {nb.x, nb.y, nb.z, nb.w};
        const float qq[4] = {q4.x, q4.y, q4.z, q4.w};
        float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (qq[j] == 0.0f) continue;
            const int64_t rec = ((int64_t)b * V + ni[j]) * 16;
            const float4* M = (const float4*)(ober2cano + rec);
            const float4 r0 = __ldg(M), r1 = __ldg(M + 1), r2 = __ldg(M + 2);
            gx += qq[j] * (r0.x * g0 + r1.x * g1 + r2.x * g2);
            gy += qq[j] * (r0.y * g0 + r1.y * g1 + r2.y * g2);
            gz += qq[j] * (r0.z * g0 + r1.z * g1 + r2.z * g2);
            if (g_o2c) {          // one 128-bit reduction per matrix row (REDG.ADD.F32x4) instead of four scalar ones
                float4* G = (float4*)(g_o2c + rec);
                const float a0 = qq[j] * g0, a1 = qq[j] * g1, a2 = qq[j] * g2;
                atomicAdd(G + 0, make_float4(a0 * qx, a0 * qy, a0 * qz, a0));
                atomicAdd(G + 1, make_float4(a1 * qx, a1 * qy, a1 * qz, a1));
                atomicAdd(G + 2, make_float4(a2 * qx, a2 * qy, a2 * qz, a2));
            }
        }
        if (g_xyz) { g_xyz[gid * 3] = gx; g_xyz[gid * 3 + 1] = gy; g_xyz[gid * 3 + 2] = gz; }
    }
}

// ------------------------------------------------------------------ C ABI
extern "C" int64_t an_vertex_grid_bytes(int B, int V) { return B > 0 && V > 0 ? (int64_t)B * grid_frame_bytes(V) : 0; }

extern "C" int an_vertex_grid_build(const float* verts, int B, int V, float cell, float flag_radius, void* ws, void* stream)
{
    if (!verts || !ws || B <= 0 || V <= 0 || !(cell > 0.f)) return AN_ERR_ARG;
    if (flag_radius > (float)GRID_R * cell) return AN_ERR_ARG;          // the flags look GRID_R cells around a cell
    if (((uintptr_t)ws) & 15) return AN_ERR_ALIGN;
    vertex_grid_build_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(verts, V, cell, flag_radius, (char*)ws, grid_frame_bytes(V));
    AN_CHECK_LAUNCH();
    vertex_grid_flags_kernel<<<dim3(B >= 8 ? 128 : 512, B), 256, 0, (cudaStream_t)stream>>>(
        (char*)ws, grid_frame_bytes(V));
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int64_t an_knn_query_ws_bytes(int B, int64_t N)
{
    return (B > 0 && N > 0) ? (int64_t)sizeof(QueryWs) + (int64_t)B * N * (int64_t)sizeof(float4) : 0;
}

static int launch_search(int sms, int64_t total, cudaStream_t st, int K, int64_t N, const float* verts, int V,
                         const char* grid_ws, int64_t frame_bytes, QueryWs* qws, const float* ober2cano,
                         const float* lbsw, int J, float thr, SeedIn sd, UnposeOut o)
{
    const int smem = SEARCH_CAP * SEARCH_THREADS * 8;
    // persistent search CTAs pull 32-query chunks from the work list (its length is device-side)
    int64_t sb = (total + SEARCH_THREADS - 1) / SEARCH_THREADS;
    if (sb > (int64_t)sms * 8) sb = (int64_t)sms * 8;
    knn_search_kernel<<<(unsigned)sb, SEARCH_THREADS, smem, st>>>(
        K, N, verts, V, grid_ws, frame_bytes, qws, ober2cano, lbsw, J, thr, sd, o);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int an_knn_unpose_fwd(const float* xyz, const float* rays, const float* z, int B, int R, int K,
                                 int64_t N, const float* verts, int V, const void* grid_ws, void* query_ws,
                                 const float* ober2cano, const float* lbs_weights, int J,
                                 float dis_threshold, int mode,
                                 const uint8_t* seed_src, const uint8_t* seed_nn, const int32_t* seed_idx, int seed_Kc,
                                 const float* seed_xyz_cano, const uint8_t* seed_valid, const float* seed_qw,
                                 float* xyz_cano, uint8_t* valid, int32_t* idx, float* dist, float* qw,
                                 float* sigma, float* rgb, int32_t* cidx, int32_t* count, void* stream)
{
    if (!verts || !ober2cano || !lbs_weights || !xyz_cano || !valid || B <= 0 || N <= 0 || V < 4 || J <= 0) return AN_ERR_ARG;
    if (!xyz && (!rays || K <= 0)) return AN_ERR_ARG;
    if (!xyz && z && (int64_t)R * K != N) return AN_ERR_ARG;
    if (!xyz && !z && (R != 0 || B != 1 || N % K)) return AN_ERR_ARG;         // lattice mode (an_knn_unpose_lattice_fwd)
    if (cidx && !count) return AN_ERR_ARG;
    if ((int64_t)B * N > 0x7fffffffLL) return AN_ERR_UNSUPPORTED;        // compact ids are int32
    if ((((uintptr_t)ober2cano) | ((uintptr_t)rays) | ((uintptr_t)idx) | ((uintptr_t)dist) | ((uintptr_t)qw)) & 15) return AN_ERR_ALIGN;
    if (seed_idx) {                                                       // seeds refer to samples of the same rays
        if (xyz || !seed_src || !seed_nn || seed_Kc <= 0 || seed_Kc > 255 || mode != 1) return AN_ERR_ARG;
        if (((uintptr_t)seed_idx) & 15) return AN_ERR_ALIGN;
    }
    UnposeOut o{xyz_cano, valid, idx, dist, qw, sigma, rgb, cidx, count};
    const int sms = an_num_sms();
    int64_t bx = (N + KNN_THREADS - 1) / KNN_THREADS;
    if (mode == 0) {
        const size_t smem = (size_t)V * sizeof(float4);
        if (smem > 200 * 1024) return AN_ERR_UNSUPPORTED;
        cudaError_t e = cudaFuncSetAttribute(knn_unpose_brute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const int64_t cap = (sms * 2 + B - 1) / B;        // staging the table costs 110 KB/CTA: keep CTAs few and fat
        if (bx > cap) bx = cap;
        dim3 grid((unsigned)bx, (unsigned)B);
        knn_unpose_brute_kernel<<<grid, KNN_THREADS, smem, (cudaStream_t)stream>>>(
            xyz, rays, z, K, N, verts, V, ober2cano, lbs_weights, J, dis_threshold, o);
    } else if (mode == 1) {
        if (!grid_ws || !query_ws) return AN_ERR_ARG;
        if (((uintptr_t)query_ws) & 15) return AN_ERR_ALIGN;
        if (V > 65535) return AN_ERR_UNSUPPORTED;                          // candidate ranges are packed 16+16 bits
        QueryWs* qws = (QueryWs*)query_ws;
        cudaStream_t st = (cudaStream_t)stream;
        cudaError_t e = cudaMemsetAsync(qws, 0, sizeof(QueryWs), st);
        if (e != cudaSuccess) return (int)e;
        if (seed_qw && (((uintptr_t)seed_qw) & 15)) return AN_ERR_ALIGN;
        const SeedIn sd{seed_idx ? seed_src : nullptr, seed_idx ? seed_nn : nullptr, seed_idx, seed_Kc,
                        seed_idx ? seed_xyz_cano : nullptr, seed_idx ? seed_valid : nullptr, seed_idx ? seed_qw : nullptr};
        const int64_t total = (int64_t)B * N;
        int64_t cb = (total + KNN_THREADS - 1) / KNN_THREADS;
        if (cb > (int64_t)sms * 32) cb = (int64_t)sms * 32;
        knn_classify_kernel<<<(unsigned)cb, KNN_THREADS, 0, st>>>(
            xyz, rays, z, K, N, total, verts, V, (const char*)grid_ws, grid_frame_bytes(V), qws, ober2cano, lbs_weights, J,
            dis_threshold, sd, o);
        AN_CHECK_LAUNCH();
        return launch_search(sms, total, st, K, N, verts, V, (const char*)grid_ws, grid_frame_bytes(V), qws, ober2cano,
                             lbs_weights, J, dis_threshold, sd, o);
    } else return AN_ERR_ARG;
    AN_CHECK_LAUNCH();
    return AN_OK;
}

// A5-A8 over the density lattice of extract_mesh.py (cfg4): the query points are generated in the kernel from the three
// axes and the centre instead of being materialised (12 B/point written and read back, plus the torch outer product).
extern "C" int an_knn_unpose_lattice_fwd(const float* lattice, int ni, int nj, int nk, const float* verts, int V,
                                         const void* grid_ws, void* query_ws, const float* ober2cano,
                                         const float* lbs_weights, int J, float dis_threshold,
                                         float* xyz_cano, uint8_t* valid, float* sigma, float* rgb,
                                         int32_t* cidx, int32_t* count, void* stream)
{
    if (!lattice || ni <= 0 || nj <= 0 || nk <= 0 || nj >= (1 << 24)) return AN_ERR_ARG;
    const int64_t N = (int64_t)ni * nj * nk;
    return an_knn_unpose_fwd(nullptr, lattice, nullptr, 1, 0, nk, N, verts, V, grid_ws, query_ws, ober2cano, lbs_weights, J,
                             dis_threshold, 1, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, xyz_cano, valid, nullptr,
                             nullptr, nullptr, sigma, rgb, cidx, count, stream);
}

extern "C" int an_knn_unpose_bwd(const float* g_xyz_cano, const int32_t* cidx, const int32_t* count,
                                 const float* xyz, const float* rays, const float* z, int B, int R, int K,
                                 int64_t N, int V, const int32_t* idx, const float* qw, const float* ober2cano,
                                 float* g_ober2cano, float* g_xyz, void* stream)
{
    if (!g_xyz_cano || !cidx || !count || !idx || !qw || !ober2cano || B <= 0 || N <= 0) return AN_ERR_ARG;
    if (!xyz && (!rays || !z || K <= 0)) return AN_ERR_ARG;
    if ((((uintptr_t)g_ober2cano) | ((uintptr_t)ober2cano) | ((uintptr_t)idx) | ((uintptr_t)qw)) & 15) return AN_ERR_ALIGN;
    knn_unpose_bwd_kernel<<<an_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(
        g_xyz_cano, cidx, count, xyz, rays, z, K, N, V, idx, qw, ober2cano, g_ober2cano, g_xyz);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
