// A13/A14 inverse-CDF resampling + sort-merge of coarse and fine depths, one warp per ray.
// Reference: models/volume_rendering.py:59-97 (sample_fine) and :199-207:
//   bins = mid-points of z_coarse (Kc-1);  p = w[1:Kc-1] + 1e-5;  pdf = p/sum p;
//   cdf = [0, cumsum(pdf)] (Kc-1);  u = linspace(0,1,Kf) (det) | U[0,1);
//   ind = #{m: cdf[m] <= u}  (searchsorted right=True);  below = max(ind-1,0);
//   above = min(ind, Kc-2);  den = cdf[above]-cdf[below], den<1e-5 -> 1;
//   z_f = bins[below] + (u-cdf[below])/den*(bins[above]-bins[below]);
//   z_all = sort(cat(z_coarse, z_f)).
// The sort is a rank merge: the coarse depths are already ascending, so each element's output
// slot is (own index) + (#elements of the other list before it), found by binary search /
// short scans in shared memory -- no generic sort network.  HBM traffic per ray:
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
    __shared__ float s_zc[SF_WARPS][SF_MAXK];
    __shared__ float s_bins[SF_WARPS][SF_MAXK];
    __shared__ float s_cdf[SF_WARPS][SF_MAXK];
    __shared__ float s_zf[SF_WARPS][SF_MAXK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* zc = s_zc[wid]; float* bins = s_bins[wid]; float* cdf = s_cdf[wid]; float* zf = s_zf[wid];
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
            zf[j] = v;
            if (z_fine) z_fine[ray * Kf + j] = v;
        }
        __syncwarp();
        // rank merge (coarse first on ties)
        const int Ka = Kc + Kf;
        for (int i = lane; i < Kc; i += 32) {
            const float a = zc[i];
            int cnt = 0;
            for (int f = 0; f < Kf; ++f) cnt += (zf[f] < a);
            const int pos = i + cnt;
            z_all[ray * Ka + pos] = a;
            if (src) src[ray * Ka + pos] = (uint8_t)i;
            if (nn_coarse) nn_coarse[ray * Ka + pos] = (uint8_t)i;
        }
        for (int j = lane; j < Kf; j += 32) {
            const float b = zf[j];
            const int nc = upper_bound_smem(zc, Kc, b);     // coarse depths <= b: samples nc-1 and nc bracket b
            int cnt = nc;
            for (int f = 0; f < Kf; ++f) cnt += (zf[f] < b) || (zf[f] == b && f < j);
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
    const int64_t want = (n_rays + SF_WARPS - 1) / SF_WARPS;
    const int64_t cap = (int64_t)an_num_sms() * 16;
    const int blocks = (int)(want < cap ? want : cap);
    sample_fine_merge_kernel<<<blocks, SF_WARPS * 32, 0, (cudaStream_t)stream>>>(
        weights, z_coarse, u, n_rays, Kc, Kf, det, seed, z_fine, z_all, src, nn_coarse);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
