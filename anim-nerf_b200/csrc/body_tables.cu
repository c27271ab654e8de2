// A16 + vertex part of A2 + clac_ober2cano_transform: the per-frame tables of the rendering path
// built by two kernels instead of ~250 torch launches.
// Reference: smplx/lbs.py:152-251 (lbs), :298-330 (batch_rodrigues), :348-420
// (batch_rigid_transform), smplx/body_models.py:289-387 (SMPL.forward; transl added into A and T),
// models/anim_nerf.py:108-126 (set_body_model), :128-145 (convert_to_body_model_space),
// :147-151 (clac_ober2cano_transform).
//
//   body_joints_kernel   one warp per (frame, posed|template): Rodrigues of the 24 joint rotations,
//       rest joints J = J_template + J_shapedirs . betas (the regressor applied to the linear shape
//       model once at load time), kinematic chain, relative transforms A_j (3x4, transl included),
//       pose feature (R_j - I, 207 values) and, for the posed body, the inverse root transform.
//   body_tables_kernel   one thread per (frame, vertex), posed and template in the same thread:
//       shape/pose blend-shape offsets, T = sum_j W_vj A_j, posed vertex, root-frame conversion,
//       closed-form affine inverse (adjugate / determinant: blended transforms are not rigid),
//       translation shift by the offset differences, ober2cano = T_template . T^-1.
// Outputs feed the KNN/unpose kernels directly: verts (B,V,3) in the root frame, ober2cano
// (B,V,4,4), ginv (B,4,4) for the ray generator, template vertices (B,V,3).
// HBM/L2-bound gather work: per frame it reads posedirs (207 x 3V fp32 = 17 MB, L2 resident across
// frames) and writes 88 B per vertex.
//
// Backward (an_body_tables_bwd; the reference's shipped configuration optimises the SMPL parameters,
// config.py:34 optim_body_params, train.py:141-145,330-331): gradients reach the builder through ober2cano
// (scatter of the blend backward, an_knn_unpose_bwd) and through ginv (the ray transform); posed vertices carry
// none (the neighbour search runs under no_grad, models/anim_nerf.py:157-159).
//   body_tables_bwd_kernel  one thread per (frame, vertex): recomputes the vertex's forward chain, pulls the
//       gradient back through ober2cano = T_tmpl . inv(ginv . T) (+ offset shift) to the blended transform, the
//       offsets and ginv, and reduces per block (warp shuffles, then one atomicAdd per value and block) into a
//       per-frame accumulator: dA (24 x 12), d pose-feature (207), d betas (10), d transl (3), d ginv (12).
//   body_joints_bwd_kernel  one warp per frame: root inverse, A_j -> world transforms, the kinematic chain in
//       reverse, rest joints (-> betas through the regressed shape directions), Rodrigues.
// Template-body parameters receive no gradient (they come from the batch, train.py:176-181).
#include "common.cuh"

#define NJ 24
#define NFEAT 207          // (NJ-1)*9
#define NBETA 10
#define JWS_FLOATS 512     // per (frame, which): A[24][12] = 288 | feat[207] -> 495 | pad

namespace {

__device__ __forceinline__ void affine_inverse_3x4(const float* T, float* I)
{
    // columns c0,c1,c2 of the 3x3 block; rows of the inverse = cross products / det
    const float a00 = T[0], a01 = T[1], a02 = T[2], a10 = T[4], a11 = T[5], a12 = T[6], a20 = T[8], a21 = T[9], a22 = T[10];
    // r0 = c1 x c2, r1 = c2 x c0, r2 = c0 x c1   (c_k = column k)
    const float r00 = a11 * a22 - a21 * a12, r01 = a21 * a02 - a01 * a22, r02 = a01 * a12 - a11 * a02;
    const float r10 = a12 * a20 - a22 * a10, r11 = a22 * a00 - a02 * a20, r12 = a02 * a10 - a12 * a00;
    const float r20 = a10 * a21 - a20 * a11, r21 = a20 * a01 - a00 * a21, r22 = a00 * a11 - a10 * a01;
    const float det = a00 * r00 + a10 * r01 + a20 * r02;
    const float t0 = T[3], t1 = T[7], t2 = T[11];
    I[0] = r00 / det; I[1] = r01 / det; I[2] = r02 / det;
    I[4] = r10 / det; I[5] = r11 / det; I[6] = r12 / det;
    I[8] = r20 / det; I[9] = r21 / det; I[10] = r22 / det;
    I[3] = -(I[0] * t0 + I[1] * t1 + I[2] * t2);
    I[7] = -(I[4] * t0 + I[5] * t1 + I[6] * t2);
    I[11] = -(I[8] * t0 + I[9] * t1 + I[10] * t2);
}

// C = A . B for 3x4 affine transforms (implicit last row [0,0,0,1])
__device__ __forceinline__ void affine_mul(const float* A, const float* B, float* C)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float s = A[r * 4] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c];
            if (c == 3) s += A[r * 4 + 3];
            C[r * 4 + c] = s;
        }
    }
}

}  // namespace

// grid = 2*B warps (one CTA of 32 threads each): blockIdx.x = b*2 + which (0 posed, 1 template)
__global__ void __launch_bounds__(32)
body_joints_kernel(const float* __restrict__ betas, const float* __restrict__ pose, const float* __restrict__ transl,
                   const float* __restrict__ betas_t, const float* __restrict__ pose_t, const float* __restrict__ transl_t,
                   int Bt, const float* __restrict__ J_template, const float* __restrict__ J_shapedirs,
                   const int32_t* __restrict__ parents, float* __restrict__ jws, float* __restrict__ ginv)
{
    __shared__ float s_loc[NJ][12];     // local transforms [R | rel]
    __shared__ float s_G[NJ][12];       // world transforms
    __shared__ float s_J[NJ][3];
    const int b = blockIdx.x >> 1, which = blockIdx.x & 1, lane = threadIdx.x;
    const int bs = which ? (Bt == 1 ? 0 : b) : b;
    const float* be = (which ? betas_t : betas) + bs * NBETA;
    const float* po = (which ? pose_t : pose) + bs * NJ * 3;
    const float* tr = which ? transl_t : transl;
    float* out = jws + (size_t)blockIdx.x * JWS_FLOATS;

    if (lane < NJ) {
        const int j = lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s = J_template[j * 3 + c];
#pragma unroll
            for (int l = 0; l < NBETA; ++l) s += J_shapedirs[(j * 3 + c) * NBETA + l] * be[l];
            s_J[j][c] = s;
        }
        // Rodrigues with the reference's angle = |r + 1e-8| (smplx/lbs.py:316)
        const float rx = po[j * 3], ry = po[j * 3 + 1], rz = po[j * 3 + 2];
        const float ex = rx + 1e-8f, ey = ry + 1e-8f, ez = rz + 1e-8f;
        const float ang = sqrtf(ex * ex + ey * ey + ez * ez);
        const float x = rx / ang, y = ry / ang, z = rz / ang;
        float sn, cs;
        sincosf(ang, &sn, &cs);
        const float oc = 1.0f - cs;
        // R = I + sin K + (1-cos) K^2,  K = [[0,-z,y],[z,0,-x],[-y,x,0]]
        float R[9];
        R[0] = 1.0f + oc * (-(z * z) - y * y); R[1] = -sn * z + oc * (x * y);       R[2] = sn * y + oc * (x * z);
        R[3] = sn * z + oc * (x * y);          R[4] = 1.0f + oc * (-(z * z) - x * x); R[5] = -sn * x + oc * (y * z);
        R[6] = -sn * y + oc * (x * z);         R[7] = sn * x + oc * (y * z);        R[8] = 1.0f + oc * (-(y * y) - x * x);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) s_loc[j][r * 4 + c] = R[r * 3 + c];
        if (j > 0) {
#pragma unroll
            for (int e = 0; e < 9; ++e) out[288 + (j - 1) * 9 + e] = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
        }
    }
    __syncwarp();
    if (lane < NJ) {
        const int j = lane, p = parents[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) s_loc[j][c * 4 + 3] = (j == 0 || p < 0) ? s_J[j][c] : s_J[j][c] - s_J[p][c];
    }
    __syncwarp();
    // kinematic chain: parents precede children (SMPL kintree); 12 lanes compute one element each
    if (lane < 12) s_G[0][lane] = s_loc[0][lane];
    __syncwarp();
    for (int j = 1; j < NJ; ++j) {
        const int p = parents[j];
        if (lane < 12) {
            const int r = lane >> 2, c = lane & 3;
            float s = s_G[p][r * 4] * s_loc[j][c] + s_G[p][r * 4 + 1] * s_loc[j][4 + c] + s_G[p][r * 4 + 2] * s_loc[j][8 + c];
            if (c == 3) s += s_G[p][r * 4 + 3];
            s_G[j][lane] = s;
        }
        __syncwarp();
    }
    // A_j = G_j with translation - R_G J_j, plus transl
    if (lane < NJ) {
        const int j = lane;
        float A[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) A[e] = s_G[j][e];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float corr = A[r * 4] * s_J[j][0] + A[r * 4 + 1] * s_J[j][1] + A[r * 4 + 2] * s_J[j][2];
            A[r * 4 + 3] = A[r * 4 + 3] - corr;
            if (tr) A[r * 4 + 3] += tr[bs * 3 + r];
        }
#pragma unroll
        for (int e = 0; e < 12; ++e) out[j * 12 + e] = A[e];
        if (j == 0 && which == 0) {
            float I[12];
            affine_inverse_3x4(A, I);
#pragma unroll
            for (int e = 0; e < 12; ++e) ginv[b * 16 + e] = I[e];
            ginv[b * 16 + 12] = 0.f; ginv[b * 16 + 13] = 0.f; ginv[b * 16 + 14] = 0.f; ginv[b * 16 + 15] = 1.f;
        }
    }
}

// grid (ceil(V/128), B), 128 threads: one vertex per thread, posed and template together
__global__ void __launch_bounds__(128)
body_tables_kernel(const float* __restrict__ betas, const float* __restrict__ betas_t, int Bt,
                   const float* __restrict__ transl, const float* __restrict__ transl_t,
                   const float* __restrict__ v_template, const float* __restrict__ shapedirs,
                   const float* __restrict__ posedirs, const float* __restrict__ lbsw, int V,
                   const float* __restrict__ jws, const float* __restrict__ ginv,
                   float* __restrict__ verts, float* __restrict__ o2c, float* __restrict__ verts_tmpl)
{
    __shared__ float s_A[2][NJ * 12];
    __shared__ float s_feat[2][NFEAT + 1];
    __shared__ float s_beta[2][NBETA];
    __shared__ float s_g[12];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int bt = Bt == 1 ? 0 : b;
    for (int e = tid; e < 2 * JWS_FLOATS; e += blockDim.x) {
        const int w = e / JWS_FLOATS, k = e - w * JWS_FLOATS;
        const float val = jws[((size_t)b * 2 + w) * JWS_FLOATS + k];
        if (k < 288) s_A[w][k] = val;
        else if (k < 288 + NFEAT) s_feat[w][k - 288] = val;
    }
    if (tid < NBETA) { s_beta[0][tid] = betas[b * NBETA + tid]; s_beta[1][tid] = betas_t[bt * NBETA + tid]; }
    if (tid < 12) s_g[tid] = ginv[b * 16 + tid];
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + tid;
    if (v >= V) return;

    float so[2][3], pof[2][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int l = 0; l < NBETA; ++l) {
            const float d = __ldg(shapedirs + ((size_t)v * 3 + c) * NBETA + l);
            s0 += s_beta[0][l] * d; s1 += s_beta[1][l] * d;
        }
        so[0][c] = s0; so[1][c] = s1;
        pof[0][c] = 0.f; pof[1][c] = 0.f;
    }
    {
        const float* pd = posedirs + (size_t)v * 3;
        const size_t stride = (size_t)V * 3;
#pragma unroll 3
        for (int k = 0; k < NFEAT; ++k) {
            const float d0 = __ldg(pd + k * stride), d1 = __ldg(pd + k * stride + 1), d2 = __ldg(pd + k * stride + 2);
            const float f0 = s_feat[0][k], f1 = s_feat[1][k];
            pof[0][0] += f0 * d0; pof[0][1] += f0 * d1; pof[0][2] += f0 * d2;
            pof[1][0] += f1 * d0; pof[1][1] += f1 * d1; pof[1][2] += f1 * d2;
        }
    }
    float T[2][12];
#pragma unroll
    for (int e = 0; e < 12; ++e) { T[0][e] = 0.f; T[1][e] = 0.f; }
    for (int j = 0; j < NJ; ++j) {
        const float w = __ldg(lbsw + (size_t)v * NJ + j);
#pragma unroll
        for (int e = 0; e < 12; ++e) { T[0][e] += w * s_A[0][j * 12 + e]; T[1][e] += w * s_A[1][j * 12 + e]; }
    }
    // SMPL.forward adds transl to T *after* the blend (T += [0|transl]); A already carries transl, and the
    // skinning weights sum to 1 only approximately: rebuild the reference's T = sum_j w_j (A_j - transl) + transl
    {
        float wsum = 0.f;
        for (int j = 0; j < NJ; ++j) wsum += __ldg(lbsw + (size_t)v * NJ + j);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (transl) T[0][r * 4 + 3] += (1.0f - wsum) * transl[b * 3 + r];
            if (transl_t) T[1][r * 4 + 3] += (1.0f - wsum) * transl_t[bt * 3 + r];
        }
    }
    const float vt0 = v_template[v * 3], vt1 = v_template[v * 3 + 1], vt2 = v_template[v * 3 + 2];
    float vw[2][3];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const float p0 = vt0 + so[w][0] + pof[w][0], p1 = vt1 + so[w][1] + pof[w][1], p2 = vt2 + so[w][2] + pof[w][2];
#pragma unroll
        for (int r = 0; r < 3; ++r) vw[w][r] = T[w][r * 4] * p0 + T[w][r * 4 + 1] * p1 + T[w][r * 4 + 2] * p2 + T[w][r * 4 + 3];
    }
    // root frame: verts_b = ginv [v;1], T_b = ginv . T
    const size_t o = (size_t)b * V + v;
#pragma unroll
    for (int r = 0; r < 3; ++r)
        verts[o * 3 + r] = s_g[r * 4] * vw[0][0] + s_g[r * 4 + 1] * vw[0][1] + s_g[r * 4 + 2] * vw[0][2] + s_g[r * 4 + 3];
    if (verts_tmpl) { verts_tmpl[o * 3] = vw[1][0]; verts_tmpl[o * 3 + 1] = vw[1][1]; verts_tmpl[o * 3 + 2] = vw[1][2]; }
    float Tb[12], Ti[12], M[12];
    affine_mul(s_g, T[0], Tb);
    affine_inverse_3x4(Tb, Ti);
#pragma unroll
    for (int r = 0; r < 3; ++r) Ti[r * 4 + 3] += (so[1][r] - so[0][r]) + (pof[1][r] - pof[0][r]);
    affine_mul(T[1], Ti, M);
    float4* dst = (float4*)(o2c + o * 16);
    dst[0] = make_float4(M[0], M[1], M[2], M[3]);
    dst[1] = make_float4(M[4], M[5], M[6], M[7]);
    dst[2] = make_float4(M[8], M[9], M[10], M[11]);
    dst[3] = make_float4(0.f, 0.f, 0.f, 1.f);
}

// ------------------------------------------------------------------------------ backward
#define ACC_A 0            // dA[24][12]
#define ACC_FEAT 288       // d feat[207]
#define ACC_BETA 495       // d betas[10] (blend-shape part)
#define ACC_TRANSL 505     // d transl[3] (skinning-weight remainder part)
#define ACC_GINV 508       // d ginv[12]
#define ACC_FLOATS 520

namespace {
// gX (3x4) of X = inverse(Y^-1) i.e. Y = inv(X) for affine transforms: gX = -Y^T gY Y^T restricted to the top rows
__device__ __forceinline__ void affine_inverse_bwd(const float* Y, const float* gY, float* gX)
{
    // S = gY.R Y.R^T + gY.t Y.t^T (3x3);  gX.R = -Y.R^T S;  gX.t = -Y.R^T gY.t
    float S[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            S[r * 3 + c] = gY[r * 4] * Y[c * 4] + gY[r * 4 + 1] * Y[c * 4 + 1] + gY[r * 4 + 2] * Y[c * 4 + 2] + gY[r * 4 + 3] * Y[c * 4 + 3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            gX[r * 4 + c] = -(Y[r] * S[c] + Y[4 + r] * S[3 + c] + Y[8 + r] * S[6 + c]);
        gX[r * 4 + 3] = -(Y[r] * gY[3] + Y[4 + r] * gY[7] + Y[8 + r] * gY[11]);
    }
}
}  // namespace

__global__ void __launch_bounds__(128)
body_tables_bwd_kernel(const float* __restrict__ transl, const float* __restrict__ shapedirs, const float* __restrict__ posedirs,
                       const float* __restrict__ lbsw, int V, const float* __restrict__ jws,
                       const float* __restrict__ ginv, const float* __restrict__ g_o2c, float* __restrict__ acc)
{
    __shared__ float s_A[2][NJ * 12];
    __shared__ float s_g[12];
    __shared__ float s_part[4][ACC_FLOATS];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < 2 * 288; e += blockDim.x) {
        const int w = e / 288, k = e - w * 288;
        s_A[w][k] = jws[((size_t)b * 2 + w) * JWS_FLOATS + k];
    }
    if (tid < 12) s_g[tid] = ginv[b * 16 + tid];
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + tid;
    const bool in = v < V;
    const int vv = in ? v : V - 1;
    float* part = s_part[warp];

    // ---- forward chain of this vertex (see body_tables_kernel): T (posed, 3x4), rotation of T_tmpl
    float T0[12], T1[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) { T0[e] = 0.f; T1[e] = 0.f; }
    float wsum = 0.f;
    for (int j = 0; j < NJ; ++j) {
        const float w = __ldg(lbsw + (size_t)vv * NJ + j);
        wsum += w;
#pragma unroll
        for (int e = 0; e < 12; ++e) { T0[e] += w * s_A[0][j * 12 + e]; T1[e] += w * s_A[1][j * 12 + e]; }
    }
    if (transl) {       // the reference's T = sum_j w_j (A_j - transl) + transl (see body_tables_kernel)
#pragma unroll
        for (int r = 0; r < 3; ++r) T0[r * 4 + 3] += (1.0f - wsum) * transl[b * 3 + r];
    }
    float Tb[12], Ti[12];
    affine_mul(s_g, T0, Tb);
    affine_inverse_3x4(Tb, Ti);                   // pre-shift inverse
    // ---- gradient of M = T1 . Ti' (Ti' = Ti with t += shift)
    float gM[12];
    {
        const float4* src = (const float4*)(g_o2c + ((size_t)b * V + vv) * 16);
        const float4 r0 = in ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f), r1 = in ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f),
                     r2 = in ? __ldg(src + 2) : make_float4(0.f, 0.f, 0.f, 0.f);
        gM[0] = r0.x; gM[1] = r0.y; gM[2] = r0.z; gM[3] = r0.w; gM[4] = r1.x; gM[5] = r1.y; gM[6] = r1.z; gM[7] = r1.w;
        gM[8] = r2.x; gM[9] = r2.y; gM[10] = r2.z; gM[11] = r2.w;
    }
    float gTi[12];                                // T1.R^T gM (rotation and translation columns alike)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) gTi[r * 4 + c] = T1[r] * gM[c] + T1[4 + r] * gM[4 + c] + T1[8 + r] * gM[8 + c];
    const float gsh[3] = {-gTi[3], -gTi[7], -gTi[11]};        // d shape offset = d pose offset of the posed body
    float gTb[12], gT0[12];
    affine_inverse_bwd(Ti, gTi, gTb);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) gT0[r * 4 + c] = s_g[r] * gTb[c] + s_g[4 + r] * gTb[4 + c] + s_g[8 + r] * gTb[8 + c];

    // ---- block reduction: every warp sums its 32 vertices per output into its own row of s_part
    // d ginv: rotation gTb.R T0.R^T + gTb.t T0.t^T, translation gTb.t
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float x = warp_sum(gTb[r * 4] * T0[c * 4] + gTb[r * 4 + 1] * T0[c * 4 + 1] + gTb[r * 4 + 2] * T0[c * 4 + 2] + gTb[r * 4 + 3] * T0[c * 4 + 3]);
            if (lane == 0) part[ACC_GINV + r * 4 + c] = x;
        }
        const float x = warp_sum(gTb[r * 4 + 3]);
        if (lane == 0) part[ACC_GINV + r * 4 + 3] = x;
    }
    {   // d transl (remainder term) and d betas
        const float rem = 1.0f - wsum;
#pragma unroll
        for (int r = 0; r < 3; ++r) { const float x = warp_sum(rem * gT0[r * 4 + 3]); if (lane == 0) part[ACC_TRANSL + r] = x; }
        for (int l = 0; l < NBETA; ++l) {
            float c = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) c += __ldg(shapedirs + ((size_t)vv * 3 + k) * NBETA + l) * gsh[k];
            c = warp_sum(c);
            if (lane == 0) part[ACC_BETA + l] = c;
        }
    }
    for (int j = 0; j < NJ; ++j) {      // dA_j = w_vj gT0
        const float w = __ldg(lbsw + (size_t)vv * NJ + j);
#pragma unroll
        for (int e = 0; e < 12; ++e) { const float x = warp_sum(w * gT0[e]); if (lane == 0) part[ACC_A + j * 12 + e] = x; }
    }
    {   // d feat_k = posedirs[k][v] . gsh
        const float* pd = posedirs + (size_t)vv * 3;
        const size_t stride = (size_t)V * 3;
#pragma unroll 3
        for (int k = 0; k < NFEAT; ++k) {
            float c = __ldg(pd + k * stride) * gsh[0] + __ldg(pd + k * stride + 1) * gsh[1] + __ldg(pd + k * stride + 2) * gsh[2];
            c = warp_sum(c);
            if (lane == 0) part[ACC_FEAT + k] = c;
        }
    }
    __syncthreads();
    for (int e = tid; e < ACC_FLOATS; e += blockDim.x)
        atomicAdd(acc + (size_t)b * ACC_FLOATS + e, s_part[0][e] + s_part[1][e] + s_part[2][e] + s_part[3][e]);
}

// grid = B warps.  Recomputes the posed body's joint chain, then walks it backwards.
__global__ void __launch_bounds__(32)
body_joints_bwd_kernel(const float* __restrict__ betas, const float* __restrict__ pose, const float* __restrict__ transl,
                       const float* __restrict__ J_template, const float* __restrict__ J_shapedirs,
                       const int32_t* __restrict__ parents, const float* __restrict__ acc, const float* __restrict__ g_ginv_ext,
                       float* __restrict__ g_betas, float* __restrict__ g_pose, float* __restrict__ g_transl)
{
    __shared__ float s_loc[NJ][12];     // local transforms [R | rel]
    __shared__ float s_G[NJ][12];       // world transforms
    __shared__ float s_J[NJ][3];
    __shared__ float s_gG[NJ][12];
    __shared__ float s_gL[NJ][12];
    __shared__ float s_gJ[NJ][3];
    const int b = blockIdx.x, lane = threadIdx.x;
    const float* be = betas + b * NBETA;
    const float* po = pose + b * NJ * 3;
    const float* ac = acc + (size_t)b * ACC_FLOATS;
    float rx = 0.f, ry = 0.f, rz = 0.f, ang = 1.f, x = 0.f, y = 0.f, z = 0.f, sn = 0.f, cs = 1.f;
    if (lane < NJ) {
        const int j = lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s = J_template[j * 3 + c];
#pragma unroll
            for (int l = 0; l < NBETA; ++l) s += J_shapedirs[(j * 3 + c) * NBETA + l] * be[l];
            s_J[j][c] = s;
        }
        rx = po[j * 3]; ry = po[j * 3 + 1]; rz = po[j * 3 + 2];
        const float ex = rx + 1e-8f, ey = ry + 1e-8f, ez = rz + 1e-8f;
        ang = sqrtf(ex * ex + ey * ey + ez * ez);
        x = rx / ang; y = ry / ang; z = rz / ang;
        sincosf(ang, &sn, &cs);
        const float oc = 1.0f - cs;
        s_loc[j][0] = 1.0f + oc * (-(z * z) - y * y); s_loc[j][1] = -sn * z + oc * (x * y);        s_loc[j][2] = sn * y + oc * (x * z);
        s_loc[j][4] = sn * z + oc * (x * y);          s_loc[j][5] = 1.0f + oc * (-(z * z) - x * x); s_loc[j][6] = -sn * x + oc * (y * z);
        s_loc[j][8] = -sn * y + oc * (x * z);         s_loc[j][9] = sn * x + oc * (y * z);         s_loc[j][10] = 1.0f + oc * (-(y * y) - x * x);
    }
    __syncwarp();
    if (lane < NJ) {
        const int j = lane, p = parents[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) s_loc[j][c * 4 + 3] = (j == 0 || p < 0) ? s_J[j][c] : s_J[j][c] - s_J[p][c];
    }
    __syncwarp();
    if (lane < 12) s_G[0][lane] = s_loc[0][lane];
    __syncwarp();
    for (int j = 1; j < NJ; ++j) {
        const int p = parents[j];
        if (lane < 12) {
            const int r = lane >> 2, c = lane & 3;
            float s = s_G[p][r * 4] * s_loc[j][c] + s_G[p][r * 4 + 1] * s_loc[j][4 + c] + s_G[p][r * 4 + 2] * s_loc[j][8 + c];
            if (c == 3) s += s_G[p][r * 4 + 3];
            s_G[j][lane] = s;
        }
        __syncwarp();
    }
    // ---- dA_j (+ the root inverse's contribution to dA_0) -> d world transform, d rest joint, d transl
    float gt[3] = {0.f, 0.f, 0.f};
    if (lane < NJ) {
        const int j = lane;
        float gA[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) gA[e] = ac[ACC_A + j * 12 + e];
        if (j == 0) {
            float A0[12], Y[12], gY[12], gX[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) A0[e] = s_G[0][e];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                A0[r * 4 + 3] -= A0[r * 4] * s_J[0][0] + A0[r * 4 + 1] * s_J[0][1] + A0[r * 4 + 2] * s_J[0][2];
                if (transl) A0[r * 4 + 3] += transl[b * 3 + r];
            }
            affine_inverse_3x4(A0, Y);
#pragma unroll
            for (int e = 0; e < 12; ++e) gY[e] = ac[ACC_GINV + e] + (g_ginv_ext ? g_ginv_ext[b * 16 + e] : 0.f);
            affine_inverse_bwd(Y, gY, gX);
#pragma unroll
            for (int e = 0; e < 12; ++e) gA[e] += gX[e];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) s_gG[j][r * 4 + c] = gA[r * 4 + c] - gA[r * 4 + 3] * s_J[j][c];
            s_gG[j][r * 4 + 3] = gA[r * 4 + 3];
            gt[r] = gA[r * 4 + 3];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            s_gJ[j][c] = -(s_G[j][c] * gA[3] + s_G[j][4 + c] * gA[7] + s_G[j][8 + c] * gA[11]);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) gt[r] = warp_sum(gt[r]);
    __syncwarp();
    // ---- kinematic chain in reverse: G_j = G_p . L_j
    for (int j = NJ - 1; j >= 1; --j) {
        const int p = parents[j];
        float upd = 0.f;
        if (lane < 12) {
            const int r = lane >> 2, c = lane & 3;
            // dL_j = G_p.R^T dG_j
            s_gL[j][lane] = s_G[p][r] * s_gG[j][c] + s_G[p][4 + r] * s_gG[j][4 + c] + s_G[p][8 + r] * s_gG[j][8 + c];
            // dG_p.R += dG_j.R L_j.R^T + dG_j.t L_j.t^T ; dG_p.t += dG_j.t
            if (c < 3) upd = s_gG[j][r * 4] * s_loc[j][c * 4] + s_gG[j][r * 4 + 1] * s_loc[j][c * 4 + 1] + s_gG[j][r * 4 + 2] * s_loc[j][c * 4 + 2]
                             + s_gG[j][r * 4 + 3] * s_loc[j][c * 4 + 3];
            else upd = s_gG[j][r * 4 + 3];
        }
        __syncwarp();
        if (lane < 12) s_gG[p][lane] += upd;
        __syncwarp();
    }
    if (lane < 12) s_gL[0][lane] = s_gG[0][lane];
    __syncwarp();
    // ---- rest joints: L_j.t = J_j - J_parent
    if (lane < NJ) {
        const int j = lane;
        float g0 = s_gJ[j][0] + s_gL[j][3], g1 = s_gJ[j][1] + s_gL[j][7], g2 = s_gJ[j][2] + s_gL[j][11];
        for (int c = j + 1; c < NJ; ++c)
            if (parents[c] == j) { g0 -= s_gL[c][3]; g1 -= s_gL[c][7]; g2 -= s_gL[c][11]; }
        s_gJ[j][0] = g0; s_gJ[j][1] = g1; s_gJ[j][2] = g2;
        // ---- Rodrigues backward: dR = dL_j.R (+ d pose feature for j >= 1)
        float gR[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) gR[r * 3 + c] = s_gL[j][r * 4 + c] + (j > 0 ? ac[ACC_FEAT + (j - 1) * 9 + r * 3 + c] : 0.f);
        const float oc = 1.0f - cs;
        const float g_sn = -z * gR[1] + y * gR[2] + z * gR[3] - x * gR[5] - y * gR[6] + x * gR[7];
        const float g_oc = gR[0] * (-(z * z) - y * y) + (gR[1] + gR[3]) * (x * y) + (gR[2] + gR[6]) * (x * z)
                           + gR[4] * (-(z * z) - x * x) + (gR[5] + gR[7]) * (y * z) + gR[8] * (-(y * y) - x * x);
        const float g_x = oc * (y * (gR[1] + gR[3]) + z * (gR[2] + gR[6]) - 2.f * x * (gR[4] + gR[8])) + sn * (gR[7] - gR[5]);
        const float g_y = oc * (x * (gR[1] + gR[3]) + z * (gR[5] + gR[7]) - 2.f * y * (gR[0] + gR[8])) + sn * (gR[2] - gR[6]);
        const float g_z = oc * (x * (gR[2] + gR[6]) + y * (gR[5] + gR[7]) - 2.f * z * (gR[0] + gR[4])) + sn * (gR[3] - gR[1]);
        float g_ang = g_sn * cs + g_oc * sn;
        g_ang -= (g_x * rx + g_y * ry + g_z * rz) / (ang * ang);
        g_pose[(b * NJ + j) * 3] = g_x / ang + g_ang * (rx + 1e-8f) / ang;
        g_pose[(b * NJ + j) * 3 + 1] = g_y / ang + g_ang * (ry + 1e-8f) / ang;
        g_pose[(b * NJ + j) * 3 + 2] = g_z / ang + g_ang * (rz + 1e-8f) / ang;
    }
    __syncwarp();
    if (lane < NBETA) {
        float s = ac[ACC_BETA + lane];
        for (int e = 0; e < NJ * 3; ++e) s += J_shapedirs[e * NBETA + lane] * s_gJ[e / 3][e % 3];
        g_betas[b * NBETA + lane] = s;
    }
    if (lane < 3 && g_transl) g_transl[b * 3 + lane] = gt[lane] + ac[ACC_TRANSL + lane];
}

extern "C" int64_t an_body_tables_ws_bytes(int B) { return B > 0 ? (int64_t)B * 2 * JWS_FLOATS * 4 : 0; }

extern "C" int64_t an_body_tables_bwd_ws_bytes(int B) { return B > 0 ? (int64_t)B * ACC_FLOATS * 4 : 0; }

extern "C" int an_body_tables_bwd(const float* g_ober2cano, const float* g_ginv,
                                  const float* betas, const float* pose, const float* transl, int B,
                                  const float* shapedirs, const float* posedirs,
                                  const float* J_template, const float* J_shapedirs, const float* lbs_weights,
                                  const int32_t* parents, int V, int J, int n_betas, const void* ws, const float* ginv,
                                  void* bwd_ws, float* g_betas, float* g_pose, float* g_transl, void* stream)
{
    if (!g_ober2cano || !betas || !pose || !shapedirs || !posedirs || !J_template || !J_shapedirs || !lbs_weights ||
        !parents || !ws || !ginv || !bwd_ws || !g_betas || !g_pose || B <= 0 || V <= 0) return AN_ERR_ARG;
    if (J != NJ || n_betas != NBETA) return AN_ERR_UNSUPPORTED;
    if (((uintptr_t)g_ober2cano) & 15) return AN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(bwd_ws, 0, (size_t)B * ACC_FLOATS * 4, st);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)((V + 127) / 128), (unsigned)B);
    body_tables_bwd_kernel<<<grid, 128, 0, st>>>(transl, shapedirs, posedirs, lbs_weights, V, (const float*)ws,
                                                  ginv, g_ober2cano, (float*)bwd_ws);
    AN_CHECK_LAUNCH();
    body_joints_bwd_kernel<<<B, 32, 0, st>>>(betas, pose, transl, J_template, J_shapedirs, parents, (const float*)bwd_ws, g_ginv,
                                              g_betas, g_pose, g_transl);
    AN_CHECK_LAUNCH();
    return AN_OK;
}

extern "C" int an_body_tables_fwd(const float* betas, const float* pose, const float* transl,
                                  const float* betas_t, const float* pose_t, const float* transl_t, int B, int Bt,
                                  const float* v_template, const float* shapedirs, const float* posedirs,
                                  const float* J_template, const float* J_shapedirs, const float* lbs_weights,
                                  const int32_t* parents, int V, int J, int n_betas, void* ws,
                                  float* verts, float* ober2cano, float* ginv, float* verts_template, void* stream)
{
    if (!betas || !pose || !betas_t || !pose_t || !v_template || !shapedirs || !posedirs || !J_template || !J_shapedirs ||
        !lbs_weights || !parents || !ws || !verts || !ober2cano || !ginv || B <= 0 || V <= 0) return AN_ERR_ARG;
    if (J != NJ || n_betas != NBETA) return AN_ERR_UNSUPPORTED;        // SMPL: 24 joints, 10 shape coefficients
    if (Bt != 1 && Bt != B) return AN_ERR_ARG;
    if (((uintptr_t)ober2cano) & 15) return AN_ERR_ALIGN;
    body_joints_kernel<<<2 * B, 32, 0, (cudaStream_t)stream>>>(betas, pose, transl, betas_t, pose_t, transl_t, Bt,
                                                              J_template, J_shapedirs, parents, (float*)ws, ginv);
    AN_CHECK_LAUNCH();
    dim3 grid((unsigned)((V + 127) / 128), (unsigned)B);
    body_tables_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(betas, betas_t, Bt, transl, transl_t, v_template, shapedirs,
                                                              posedirs, lbs_weights, V, (const float*)ws, ginv,
                                                              verts, ober2cano, verts_template);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
