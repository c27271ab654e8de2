// marching cubes of the density lattice on the device (reference: extract_mesh.py:165,
// `mcubes.marching_cubes(-sigmas, 0.)` -- PyMCubes, a CPU library; see anim-nerf_b200/mesh.py for what is pinned).
//
// A thread owns one lattice point ("site") s = (i, j, k), k fastest, and with it
//   * the three lattice edges leaving s in +x, +y, +z: a vertex sits on an edge whose end values straddle the isovalue
//     (inside: value < iso), at s + t e, t = (iso - v0) / (v1 - v0);
//   * the cell whose corner 0 is s (when i, j, k are not on the upper boundary): its triangles come from the 256-entry
//     table, three cube-edge ids per triangle.
// Pass 1 counts (vertices, triangles) per site, scans them inside each 1024-site block, stores the site's vertex offset
// within the block (the triangle offsets are recomputed in pass 2 by the same scan) and the block totals.  A single-CTA
// scan turns the block totals into block offsets.  Pass 2 writes the vertices and, for every triangle corner, looks up
// the index of the vertex on that cube edge: offset of the edge's owner site + rank of the edge among the owner's edges.
// Vertex and face order are functions of the lattice only (no atomics): the mesh is reproducible bit for bit.
// HBM-bound integer/float work: the lattice is read twice (each value reused by 8 cells / 6 edges from L1/L2), 2 B of
// scratch per site are written and read.
#include "common.cuh"

namespace {

constexpr int MC_BLOCK = 1024;        // lattice points per block (the unit of voff / block_counts)
constexpr int MC_THREADS = 256;       // each thread owns MC_PER consecutive points
constexpr int MC_PER = MC_BLOCK / MC_THREADS;

struct Site { int nv, nt, cfg; bool hx, hy, hz; float v0, vx, vy, vz; };

__device__ __forceinline__ Site classify(const float* __restrict__ vol, int nx, int ny, int nz, float iso, int64_t s,
                                         const int8_t* __restrict__ tri)
{
    Site r{};
    const int64_t n = (int64_t)nx * ny * nz;
    if (s >= n) return r;
    const int k = (int)(s % nz), j = (int)((s / nz) % ny), i = (int)(s / ((int64_t)nz * ny));
    const int64_t sx = (int64_t)ny * nz, sy = nz;
    const bool ix = i + 1 < nx, iy = j + 1 < ny, iz = k + 1 < nz;
    r.v0 = vol[s];
    const bool in0 = r.v0 < iso;
    r.vx = ix ? vol[s + sx] : r.v0; r.vy = iy ? vol[s + sy] : r.v0; r.vz = iz ? vol[s + 1] : r.v0;
    r.hx = ix && ((r.vx < iso) != in0); r.hy = iy && ((r.vy < iso) != in0); r.hz = iz && ((r.vz < iso) != in0);
    r.nv = (int)r.hx + (int)r.hy + (int)r.hz;
    if (ix && iy && iz) {
        // corners: 0 (0,0,0) 1 (1,0,0) 2 (1,1,0) 3 (0,1,0) 4 (0,0,1) 5 (1,0,1) 6 (1,1,1) 7 (0,1,1)
        int c = in0 ? 1 : 0;
        c |= (r.vx < iso) ? 2 : 0;
        c |= (vol[s + sx + sy] < iso) ? 4 : 0;
        c |= (r.vy < iso) ? 8 : 0;
        c |= (r.vz < iso) ? 16 : 0;
        c |= (vol[s + sx + 1] < iso) ? 32 : 0;
        c |= (vol[s + sx + sy + 1] < iso) ? 64 : 0;
        c |= (vol[s + sy + 1] < iso) ? 128 : 0;
        r.cfg = c;
        if (c != 0 && c != 255) {
            const int8_t* t = tri + c * 16;
            int m = 0;
            while (m < 15 && t[m] >= 0) m += 3;
            r.nt = m / 3;
        }
    }
    return r;
}

// exclusive scan of (a, b) over the block; totals to every thread
__device__ __forceinline__ void block_scan2(int a, int b, int& ea, int& eb, int& ta, int& tb)
{
    __shared__ int wa[32], wb[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int ya = __shfl_up_sync(0xffffffffu, ia, d), yb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) { ia += ya; ib += yb; }
    }
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        int xa = wa[lane], xb = wb[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ya = __shfl_up_sync(0xffffffffu, xa, d), yb = __shfl_up_sync(0xffffffffu, xb, d);
            if (lane >= d) { xa += ya; xb += yb; }
        }
        wa[lane] = xa; wb[lane] = xb;
    }
    __syncthreads();
    const int pa = warp ? wa[warp - 1] : 0, pb = warp ? wb[warp - 1] : 0;
    ea = pa + ia - a; eb = pb + ib - b;
    ta = wa[31]; tb = wb[31];
    __syncthreads();
}
// (warp totals of warps that do not exist read as zero: wa/wb are filled for blockDim.x / 32 warps, the second-level scan
//  runs over all 32 slots, so blocks of fewer than 1024 threads must clear the tail first)
__device__ __forceinline__ void block_scan2_small(int a, int b, int& ea, int& eb, int& ta, int& tb)
{
    __shared__ int wa[32], wb[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int ya = __shfl_up_sync(0xffffffffu, ia, d), yb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) { ia += ya; ib += yb; }
    }
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    int pa = 0, pb = 0, sa = 0, sb = 0;
    for (int k = 0; k < nw; ++k) { const int xa = wa[k], xb = wb[k]; if (k < warp) { pa += xa; pb += xb; } sa += xa; sb += xb; }
    ea = pa + ia - a; eb = pb + ib - b; ta = sa; tb = sb;
    __syncthreads();
}

__global__ void __launch_bounds__(MC_THREADS)
mc_count_kernel(const float* __restrict__ vol, int nx, int ny, int nz, float iso, const int8_t* __restrict__ tri,
                uint16_t* __restrict__ voff, int32_t* __restrict__ counts)
{
    const int64_t s0 = (int64_t)blockIdx.x * MC_BLOCK + (int64_t)threadIdx.x * MC_PER;
    const int64_t n = (int64_t)nx * ny * nz;
    int nv[MC_PER], sv = 0, st = 0;
#pragma unroll
    for (int q = 0; q < MC_PER; ++q) {
        const Site r = classify(vol, nx, ny, nz, iso, s0 + q, tri);
        nv[q] = r.nv; sv += r.nv; st += r.nt;
    }
    int ev, et, tv, tt;
    block_scan2_small(sv, st, ev, et, tv, tt);
    if (tv) {            // blocks without a vertex leave their voff untouched: no triangle refers to them
#pragma unroll
        for (int q = 0; q < MC_PER; ++q) { if (s0 + q < n) voff[s0 + q] = (uint16_t)ev; ev += nv[q]; }
    }
    if (threadIdx.x == 0) { counts[2 * blockIdx.x] = tv; counts[2 * blockIdx.x + 1] = tt; }
}

// in-place exclusive scan of the (vertex, triangle) block totals by one CTA; totals[0..1] = grand totals
__global__ void __launch_bounds__(1024)
mc_scan_kernel(int32_t* __restrict__ counts, int64_t n_blocks, int64_t* __restrict__ totals)
{
    __shared__ long long carry[2];
    if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t b = base + threadIdx.x;
        const int a = b < n_blocks ? counts[2 * b] : 0, c = b < n_blocks ? counts[2 * b + 1] : 0;
        int ea, ec, ta, tc;
        block_scan2(a, c, ea, ec, ta, tc);
        const long long ca = carry[0], cc = carry[1];
        // offsets stay below 2^31: the driver (mesh.py) allocates int32 faces; checked on the host against the totals
        if (b < n_blocks) { counts[2 * b] = (int32_t)(ca + ea); counts[2 * b + 1] = (int32_t)(cc + ec); }
        __syncthreads();
        if (threadIdx.x == 0) { carry[0] = ca + ta; carry[1] = cc + tc; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; }
}

// owner site offset (di, dj, dk) and axis of the 12 cube edges
__constant__ int8_t c_edge_owner[12][4] = {
    {0, 0, 0, 0}, {1, 0, 0, 1}, {0, 1, 0, 0}, {0, 0, 0, 1}, {0, 0, 1, 0}, {1, 0, 1, 1},
    {0, 1, 1, 0}, {0, 0, 1, 1}, {0, 0, 0, 2}, {1, 0, 0, 2}, {1, 1, 0, 2}, {0, 1, 0, 2}};

__global__ void __launch_bounds__(MC_THREADS)
mc_emit_kernel(const float* __restrict__ vol, int nx, int ny, int nz, float iso, const int8_t* __restrict__ tri,
               const uint16_t* __restrict__ voff, const int32_t* __restrict__ block_off, const int64_t* __restrict__ totals,
               float* __restrict__ verts, int32_t* __restrict__ faces)
{
    const int64_t n = (int64_t)nx * ny * nz;
    const int64_t n_blocks = (n + MC_BLOCK - 1) / MC_BLOCK;
    // this block's totals from the offsets: nothing to emit -> nothing to classify (most blocks of a lattice)
    const int v0 = block_off[2 * blockIdx.x], t0 = block_off[2 * blockIdx.x + 1];
    const int64_t v1 = blockIdx.x + 1 < n_blocks ? block_off[2 * blockIdx.x + 2] : totals[0];
    const int64_t t1 = blockIdx.x + 1 < n_blocks ? block_off[2 * blockIdx.x + 3] : totals[1];
    if (v1 == v0 && t1 == t0) return;
    const int64_t s0 = (int64_t)blockIdx.x * MC_BLOCK + (int64_t)threadIdx.x * MC_PER;
    Site r[MC_PER];
    int sv = 0, st = 0;
#pragma unroll
    for (int q = 0; q < MC_PER; ++q) { r[q] = classify(vol, nx, ny, nz, iso, s0 + q, tri); sv += r[q].nv; st += r[q].nt; }
    int ev, et, tv, tt;
    block_scan2_small(sv, st, ev, et, tv, tt);
    const int64_t sx = (int64_t)ny * nz, sy = nz;
#pragma unroll
    for (int q = 0; q < MC_PER; ++q) {
        const int64_t s = s0 + q;
        if (s >= n) break;
        const int k = (int)(s % nz), j = (int)((s / nz) % ny), i = (int)(s / ((int64_t)nz * ny));
        const Site& c = r[q];
        if (c.nv) {
            float* v = verts + ((int64_t)v0 + ev) * 3;
            if (c.hx) { v[0] = (float)i + (iso - c.v0) / (c.vx - c.v0); v[1] = (float)j; v[2] = (float)k; v += 3; }
            if (c.hy) { v[0] = (float)i; v[1] = (float)j + (iso - c.v0) / (c.vy - c.v0); v[2] = (float)k; v += 3; }
            if (c.hz) { v[0] = (float)i; v[1] = (float)j; v[2] = (float)k + (iso - c.v0) / (c.vz - c.v0); }
            ev += c.nv;
        }
        if (c.nt) {
            int32_t* f = faces + ((int64_t)t0 + et) * 3;
            const int8_t* t = tri + c.cfg * 16;
            for (int m = 0; m < 3 * c.nt; ++m) {
                const int e = t[m];
                const int di = c_edge_owner[e][0], dj = c_edge_owner[e][1], dk = c_edge_owner[e][2], axis = c_edge_owner[e][3];
                const int64_t o = s + di * sx + dj * sy + dk;
                int rank = 0;
                if (axis > 0) {       // the owner's earlier edges (x, then y) exist when they are inside the lattice and straddle the isovalue
                    const bool oin = vol[o] < iso;
                    if (i + di + 1 < nx) rank += ((vol[o + sx] < iso) != oin) ? 1 : 0;
                    if (axis > 1 && j + dj + 1 < ny) rank += ((vol[o + sy] < iso) != oin) ? 1 : 0;
                }
                f[m] = block_off[2 * (o / MC_BLOCK)] + (int32_t)voff[o] + rank;
            }
            et += c.nt;
        }
    }
}

}  // namespace

extern "C" int an_mc_count(const float* volume, int nx, int ny, int nz, float iso, const int8_t* tri_table,
                           uint16_t* voff, int32_t* block_counts, void* stream)
{
    if (!volume || !tri_table || !voff || !block_counts || nx < 2 || ny < 2 || nz < 2) return AN_ERR_ARG;
    const int64_t n = (int64_t)nx * ny * nz;
    mc_count_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_THREADS, 0, (cudaStream_t)stream>>>(
        volume, nx, ny, nz, iso, tri_table, voff, block_counts);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int an_mc_scan(int32_t* block_counts, int64_t n_blocks, int64_t* totals, void* stream)
{
    if (!block_counts || !totals || n_blocks <= 0) return AN_ERR_ARG;
    mc_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(block_counts, n_blocks, totals);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int an_mc_emit(const float* volume, int nx, int ny, int nz, float iso, const int8_t* tri_table,
                          const uint16_t* voff, const int32_t* block_offsets, const int64_t* totals, float* vertices,
                          int32_t* faces, void* stream)
{
    if (!volume || !tri_table || !voff || !block_offsets || !totals || !vertices || !faces || nx < 2 || ny < 2 || nz < 2) return AN_ERR_ARG;
    const int64_t n = (int64_t)nx * ny * nz;
    mc_emit_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_THREADS, 0, (cudaStream_t)stream>>>(
        volume, nx, ny, nz, iso, tri_table, voff, block_offsets, totals, vertices, faces);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
