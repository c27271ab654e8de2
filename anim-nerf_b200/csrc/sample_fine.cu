// A13/A14 inverse-CDF resampling + sort-merge of coarse and fine depths, one warp per ray.
// Reference: models/volume_rendering.py:59-97 (sample_fine) and :199-207:
//   bins = mid-points of z_coarse (Kc-1);  p = w[1:Kc-1] + 1e-5;  pdf = p/sum p;
//   cdf = [0, cumsum(pdf)] (Kc-1);  u = linspace(0,1,Kf) (det) | U[0,1);
//   ind = #{m: cdf[m] <= u}  (searchsorted right=True);  below = max(ind-1,0);
//   above = min(ind, Kc-2);  den = cdf[above]-cdf[below], den<1e-5 -> 1;
//   z_f = bins[below] + (u-cdf[below])/den*(bins[above]-bins[below]);
//   z_all = sort(cat(z_coarse, z_f)).
// The sort is a rank merge: the coarse depths are already ascending and the Kf fine depths are put in
// (value, draw index) order by a warp bitonic sort of packed 64-bit keys in shared memory, so each
// element's output slot is (own rank) + (#elements of the other list before it), found by binary search
// (O(K log^2 K) per ray; the all-pairs count it replaces was 80 % of the kernel).  HBM traffic per ray:
// (Kc + Kc) * 4 B read, (Kc+Kf) * 5 B written.
#include "common.cuh"

#define SF_WARPS 4
#define SF_MAXK 256

__device__ __forceinline__ int upper_bound_smem(const float* a, int n, float v) {
    // #{m < n : a[m] <= v}
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// order-preserving map of a float to an unsigned key, and (value, index) keys for the fine depths
__device__ __forceinline__ unsigned ord_bits(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_value(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ascending bitonic sort of n (a power of two, >= 64) keys in shared memory by one warp
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* key, int n, int lane) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (n >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // lower index of the pair (bit j clear)
                const int l = i | j;
                const unsigned long long a = key[i], b = key[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { key[i] = b; key[l] = a; }
            }
            __syncwarp();
        }
    }
}

__global__ void searchsorted_right_kernel(const float* __restrict__ cdf, const float* __restrict__ u,
                                          int64_t n_rows, int M, int F, int32_t* __restrict__ inds)
{
    const int64_t total = n_rows * F;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / F;
        const float* c = cdf + row * M;
        const float v = u[e];
        int lo = 0, hi = M;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(c + mid) <= v) lo = mid + 1; else hi = mid;
        }
        inds[e] = lo;
    }
}

__global__ void __launch_bounds__(SF_WARPS * 32)
sample_fine_merge_kernel(const float* __restrict__ weights, const float* __restrict__ z_coarse,
                         const float* __restrict__ u_in, int64_t n_rays, int Kc, int Kf, int det,
                         uint64_t seed, float* __restrict__ z_fine, float* __restrict__ z_all,
                         uint8_t* __restrict__ src, uint8_t* __restrict__ nn_coarse)
{
    // dynamic shared memory sized to the actual sample counts (sf_smem_bytes): per warp the sort keys, then
    // coarse depths, bins and cdf (Kc floats each, Kc rounded up to even)
    extern __shared__ unsigned long long s_dyn[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int nsort = 64;                 // fine depths padded to a power of two for the sort
    while (nsort < Kf) nsort <<= 1;
    const int kce = (Kc + 1) & ~1;
    unsigned long long* key = s_dyn + (size_t)wid * (nsort + 3 * kce / 2);
    float* zc = (float*)(key + nsort); float* bins = zc + kce; float* cdf = bins + kce;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nb = Kc - 1;          // bins / cdf entries
    const int np = Kc - 2;          // pdf entries
    for (int64_t ray = warp0; ray < n_rays; ray += nwarps) {
        const float* wr = weights + ray * Kc;
        for (int i = lane; i < Kc; i += 32) zc[i] = z_coarse[ray * Kc + i];
        __syncwarp();
        for (int i = lane; i < nb; i += 32) bins[i] = 0.5f * (zc[i] + zc[i + 1]);
        // pdf normaliser
        float part = 0.f;
        for (int m = lane; m < np; m += 32) part += wr[m + 1] + 1e-5f;
        const float total = warp_sum(part);
        // inclusive scan of pdf in rounds of 32 with carry
        float carry = 0.f;
        if (lane == 0) cdf[0] = 0.f;
        for (int base = 0; base < np; base += 32) {
            const int m = base + lane;
            float v = (m < np) ? (wr[m + 1] + 1e-5f) / total : 0.f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float n = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += n;
            }
            v += carry;
            if (m < np) cdf[m + 1] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
        __syncwarp();
        // draw + invert
        const float step = (Kf > 1) ? 1.0f / (float)(Kf - 1) : 0.f;
        for (int j = lane; j < Kf; j += 32) {
            float u;
            if (u_in) u = u_in[ray * Kf + j];
            else if (det) u = (j < Kf / 2) ? step * (float)j : 1.0f - step * (float)(Kf - 1 - j);  // torch.linspace
            else u = philox_u01(seed, (uint64_t)(ray * Kf + j));
            const int ind = upper_bound_smem(cdf, nb, u);
            const int below = max(ind - 1, 0), above = min(ind, nb - 1);
            float den = cdf[above] - cdf[below];
            if (den < 1e-5f) den = 1.0f;
            const float v = bins[below] + (u - cdf[below]) / den * (bins[above] - bins[below]);
            key[j] = ((unsigned long long)ord_bits(v) << 32) | (unsigned)j;
            if (z_fine) z_fine[ray * Kf + j] = v;
        }
        for (int j = Kf + lane; j < nsort; j += 32) key[j] = ~0ull;
        __syncwarp();
        {   // evenly spaced u (perturb = 0: every inference frame) invert to ascending depths: skip the sort then
            bool asc = true;
            for (int j = lane; j + 1 < Kf; j += 32) asc = asc && key[j] < key[j + 1];
            if (!__all_sync(0xffffffffu, asc)) warp_bitonic_sort(key, nsort, lane);
        }
        // rank merge (coarse first on ties; equal fine depths in draw order).  The two sides compare differently -- packed
        // keys here, floats below -- which agree for every pair of depths except (+0.0, -0.0); depths are >= near > 0.
        const int Ka = Kc + Kf;
        for (int i = lane; i < Kc; i += 32) {
            const float a = zc[i];
            const unsigned long long ka = (unsigned long long)ord_bits(a) << 32;
            int lo = 0, hi = Kf;                                // #fine depths < a
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (key[mid] < ka) lo = mid + 1; else hi = mid;
            }
            const int pos = i + lo;
            z_all[ray * Ka + pos] = a;
            if (src) src[ray * Ka + pos] = (uint8_t)i;
            if (nn_coarse) nn_coarse[ray * Ka + pos] = (uint8_t)i;
        }
        for (int r = lane; r < Kf; r += 32) {
            const unsigned long long kb = key[r];
            const float b = ord_value((unsigned)(kb >> 32));
            const int j = (int)(unsigned)kb;
            const int nc = upper_bound_smem(zc, Kc, b);     // coarse depths <= b: samples nc-1 and nc bracket b
            const int cnt = nc + r;
            z_all[ray * Ka + cnt] = b;
            if (src) src[ray * Ka + cnt] = (uint8_t)(Kc + j);
            if (nn_coarse) {
                int nn = nc == 0 ? 0 : (nc >= Kc ? Kc - 1 : ((b - zc[nc - 1] <= zc[nc] - b) ? nc - 1 : nc));
                nn_coarse[ray * Ka + cnt] = (uint8_t)nn;
            }
        }
        __syncwarp();
    }
}

extern "C" int an_searchsorted_right(const float* cdf, const float* u, int64_t n_rows, int M, int F,
                                     int32_t* inds, void* stream)
{
    if (!cdf || !u || !inds || n_rows <= 0 || M <= 0 || F <= 0) return AN_ERR_ARG;
    const int64_t total = n_rows * F;
    const int64_t want = (total + 255) / 256;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    searchsorted_right_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cdf, u, n_rows, M, F, inds);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int an_sample_fine_merge_fwd(const float* weights, const float* z_coarse, const float* u,
                                        int64_t n_rays, int Kc, int Kf, int det, uint64_t seed,
                                        float* z_fine, float* z_all, uint8_t* src, uint8_t* nn_coarse, void* stream)
{
    if (!weights || !z_coarse || !z_all || n_rays <= 0 || Kc < 3 || Kf <= 0) return AN_ERR_ARG;
    if (Kc > SF_MAXK || Kf > SF_MAXK || Kc + Kf > 256) return AN_ERR_UNSUPPORTED;
    int nsort = 64;
    while (nsort < Kf) nsort <<= 1;
    const int smem = SF_WARPS * (nsort * 8 + 3 * ((Kc + 1) & ~1) * 4);
    // grid = what is resident at once, so the grid-stride loop gives every CTA the same share of rays
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sample_fine_merge_kernel, SF_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 8;
    const int64_t want = (n_rays + SF_WARPS - 1) / SF_WARPS;
    const int64_t cap = (int64_t)an_num_sms() * per_sm;
    const int blocks = (int)(want < cap ? want : cap);
    sample_fine_merge_kernel<<<blocks, SF_WARPS * 32, smem, (cudaStream_t)stream>>>(
        weights, z_coarse, u, n_rays, Kc, Kf, det, seed, z_fine, z_all, src, nn_coarse);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
