/* CPU oracle: exhaustive k-nearest-vertex search.  TEST INFRASTRUCTURE ONLY --
 * nothing under anim-nerf_b200/ links or calls this.
 *
 * Restates the contract of the reference's only native dependency, KNN_CUDA v0.2
 * (unlimblue/KNN_CUDA; not vendored in the reference, call site
 * models/anim_nerf.py:82-83,158-159): for each query the k smallest Euclidean
 * distances over all vertices, ascending, lowest vertex index first on ties.
 * Arithmetic contract: fp32, d2 = ((qx-vx)^2 + (qy-vy)^2) + (qz-vz)^2
 * with every operation individually rounded (compile with -ffp-contract=off),
 * dist = sqrtf(d2).  Parity-unpinned against the real KNN_CUDA (absent offline);
 * pinned against torch.cdist(compute_mode=donot_use_mm...).topk via tests/golden.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC knn_oracle.c -o libknn_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>

#define KMAX 16

void knn_oracle(const float* verts, int64_t V, const float* xyz, int64_t N, int k,
                float* dist_out, int32_t* idx_out)
{
    if (k > KMAX) k = KMAX;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        const float qx = xyz[3 * n], qy = xyz[3 * n + 1], qz = xyz[3 * n + 2];
        float bd[KMAX];
        int32_t bi[KMAX];
        for (int j = 0; j < k; ++j) { bd[j] = INFINITY; bi[j] = -1; }
        for (int64_t v = 0; v < V; ++v) {
            const float dx = qx - verts[3 * v], dy = qy - verts[3 * v + 1], dz = qz - verts[3 * v + 2];
            const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
            const float s = xx + yy;
            const float d2 = s + zz;
            if (d2 < bd[k - 1]) {            /* strict: an equal distance never displaces an earlier index */
                int j = k - 1;
                while (j > 0 && d2 < bd[j - 1]) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
                bd[j] = d2; bi[j] = (int32_t)v;
            }
        }
        for (int j = 0; j < k; ++j) { dist_out[n * k + j] = sqrtf(bd[j]); idx_out[n * k + j] = bi[j]; }
    }
}
