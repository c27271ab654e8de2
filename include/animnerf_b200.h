/* animnerf_b200.h -- C ABI of the B200-native (sm_100a) Anim-NeRF rendering hot path.
 *
 * One shared library (libanimnerf_b200.so), plain pointers and sizes, no torch types.
 * The reference (JanaldoChen/Anim-NeRF) is pure Python: its only native seam is the
 * external KNN_CUDA wheel, every other boundary is a Python call signature.  Each entry
 * point below cites the reference call it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - the caller owns every buffer including scratch; nothing is allocated inside;
 *  - no global mutable state: calls are re-entrant per stream;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - return 0 on success, <0 for an argument error (AN_ERR_*), >0 = cudaError_t;
 *  - nothing throws across the ABI;  there is NO CPU fallback.
 *  - B = frames, R = rays per frame, K = samples per ray, N = R*K points per frame,
 *    V = vertices (6890), J = joints (24), k = 4 neighbours.
 */
#ifndef ANIMNERF_B200_H
#define ANIMNERF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AN_OK 0
#define AN_ERR_ARG (-1)       /* null pointer / non-positive extent                    */
#define AN_ERR_UNSUPPORTED (-2) /* shape outside what the kernels were built for       */
#define AN_ERR_ALIGN (-3)     /* pointer not aligned as documented                      */

#define AN_KNN_K 4
#define AN_MLP_W 256          /* trunk width (models/nerf.py:62)                        */
#define AN_MLP_ENC 63         /* 3 + 3*2*10 positional-encoding channels                */
#define AN_GRID_MAX_DIM 32    /* vertex hash grid is at most 32^3 cells                 */

int an_version(void);
const char* an_error_string(int code);

/* ---- A1/A2: ray generation ------------------------------------------------------------
 * replaces datasets/anim_nerf_dataset.py:56-85 gen_ray_directions/gen_rays (and the dead
 * twin utils/ray_utils.py:74-121) fused with the ray part of
 * models/anim_nerf.py:128-137 convert_to_body_model_space.
 * c2w (B,3,4), focal (B,2), center (B,2) [cx,cy]; pix (B,R,2) int32 (row,col) or NULL for
 * the full H*W grid in row-major order (then R must equal H*W); ginv (B,4,4) inverse root
 * transform or NULL (world-space rays, near/far unclamped).  rays (B,R,8).               */
int an_raygen_fwd(const float* c2w, const float* focal, const float* center, const int32_t* pix,
                  const float* ginv, int B, int R, int H, int W, float near_, float far_,
                  float* rays, void* stream);

/* ---- A1 for training batches: pixel sampling + gathers + ray generation ----------
 * replaces, per training step, datasets/anim_nerf_dataset.py:235-262 (__getitem__): get_pixelcoords'
 * 'foreground_pixel' draws (:10-54; the first n_fg samples from the eroded-silhouette list, the rest from the
 * outside-band list, with replacement), the rgb/alpha gathers at those pixels (:259-261) with the
 * mask / white-background composite (:200-204, :244-245), and gen_rays (:56-85) at the drawn pixels only,
 * fused with the body-space ray transform (ginv as in an_raygen_fwd).
 * images (F,H,W,3) uint8 RGB and masks (F,H,W) uint8 stay resident in device memory; fg/bg lists are CSR per
 * stored frame: *_off (F+1) int32, *_list linear pixel ids row*W+col of mask_inside>0 / mask_outside>0.
 * frame_ids (B) int32 selects the stored frame of each batch entry.  sel (B,n) int32 explicit positions in
 * the lists (parity with the reference's np.random.choice draws) or NULL for the in-kernel Philox stream.
 * Outputs rays (B,n,8), rgbs (B,n,3), alphas (B,n), pix (B,n,2) int32 (row,col; may be NULL).            */
int an_sample_training_rays_fwd(const uint8_t* images, const uint8_t* masks,
                                const int32_t* fg_list, const int32_t* fg_off,
                                const int32_t* bg_list, const int32_t* bg_off,
                                const int32_t* frame_ids, const float* c2w, const float* focal,
                                const float* center, const float* ginv, int B, int n, int n_fg, int H, int W,
                                float near_, float far_, int white_bkgd, int with_background,
                                const int32_t* sel, uint64_t seed,
                                float* rays, float* rgbs, float* alphas, int32_t* pix, void* stream);

/* ---- A3: stratified sampling ------------------------------------------------------------
 * replaces models/volume_rendering.py:29-56 VolumeRenderer.sample_coarse (lindisp=True,
 * i.e. linear in depth).  rays (n_rays,8), z (n_rays,Kc).  perturb>0: noise_u (n_rays,Kc)
 * U[0,1) draws if non-NULL (parity mode), else an in-kernel Philox stream (seed,offset).   */
int an_sample_coarse_fwd(const float* rays, int64_t n_rays, int Kc, float perturb,
                         const float* noise_u, uint64_t seed, float* z, void* stream);

/* ---- A1 + A2 + A3 in one launch ------------------------------------------------------------------
 * an_rays_sample_fwd: ray generation (camera + pixel list / full grid, as an_raygen_fwd) OR given world-space rays
 * rays_world (B,R,8) (the reference's training batches carry rays, train.py:172; then c2w/focal/center/pix may be
 * NULL and near/far come from the rays), the body-space transform with the near/far clamp (ginv (B,4,4) or NULL)
 * and the stratified depths of an_sample_coarse_fwd, fused: rays_body (B,R,8), z (B,R,Kc).
 * an_rays_sample_bwd: gradient with respect to ginv (models/anim_nerf.py:131: the inverse SMPL root transform, the
 * route from the rays to the body parameters under optim_body_params): g_rays_body (B,R,8) = gradients of
 * [o', d', near', far'], g_z (B,R,Kc) or NULL -> g_ginv (B,4,4) (written; rows 0-2).  No gradient to the world rays.
 * an_ray_point_grad: the ray-side gradients of one render pass from its per-point gradients (x = o + z d): g_xyz
 * (n_rays*K,3) is read at valid samples only; g_z_comp (n_rays,K) / g_far_comp (n_rays) from an_composite_bwd or NULL;
 * writes g_rays (n_rays,8) = [sum g_x, sum z g_x, 0, g_far_comp] and g_z (n_rays,K) = g_z_comp + g_x . d.        */
int an_rays_sample_fwd(const float* c2w, const float* focal, const float* center, const int32_t* pix,
                       const float* rays_world, const float* ginv, int B, int R, int H, int W, int Kc,
                       float near_, float far_, float perturb, const float* noise_u, uint64_t seed,
                       float* rays_body, float* z, void* stream);
int an_rays_sample_bwd(const float* c2w, const float* focal, const float* center, const int32_t* pix,
                       const float* rays_world, const float* rays_body, const float* z,
                       const float* g_rays_body, const float* g_z, int B, int R, int H, int W, int Kc,
                       float near_, float far_, float* g_ginv, void* stream);
int an_ray_point_grad(const float* rays, const float* z, const uint8_t* valid, const float* g_xyz,
                      const float* g_z_comp, const float* g_far_comp, int64_t n_rays, int K,
                      float* g_rays, float* g_z, void* stream);

/* ---- A5-A8 (+A4): K-nearest-vertex search fused with inverse skinning --------------------
 * replaces knn_cuda.KNN(k=4, transpose_mode=True).forward (call site
 * models/anim_nerf.py:82-83,158-159) + get_neighbs (:153-178) + unpose (:180-192) +
 * batch_index_select/batch_transform (:24-39) + the point generation in
 * models/volume_rendering.py:117-120.
 *
 * an_vertex_grid_build: per-frame uniform grid over the posed vertices used by the exact pruned
 * search.  REQUIRED: 3*cell >= dis_threshold of the later queries (the kernel scans a 7^3-cell box);
 * pass cell = 1.25 * dis_threshold/3 * 1.001 (a smaller cell only triggers more exhaustive rescans).  The cell grows automatically if the body would need more
 * than AN_GRID_MAX_DIM cells per axis.  ws: an_vertex_grid_bytes(B,V) bytes, 16-byte aligned.
 * flag_radius: the build also marks, for every cell of the grid dilated by 3 cells, whether a query inside the cell can
 * be valid.  flag_radius > 0 (<= 3*cell; pass dis_threshold): exact test "some vertex lies within flag_radius of the
 * cell's box" -- queries in unmarked cells are farther than flag_radius from every vertex and are finished without a
 * search (a later call with a larger dis_threshold ignores the marks).  flag_radius <= 0: the coarser mark "a vertex in
 * the 7^3-cell neighbourhood".                                                                                       */
int64_t an_vertex_grid_bytes(int B, int V);
int an_vertex_grid_build(const float* verts, int B, int V, float cell, float flag_radius, void* ws, void* stream);

/* Query points are either xyz (B,N,3) (rays,z NULL) or generated as o + z*d from rays
 * (B,R,8), z (B,R,K) with N = R*K (xyz NULL).
 * mode 0 = exhaustive shared-memory-tiled search; 1 = grid-pruned search (exact: falls back
 * to the exhaustive scan per query when the pruning radius cannot prove the 4th neighbour).
 * Outputs (any of idx/dist/qw/cidx may be NULL):
 *   xyz_cano (B*N,3); valid (B*N) u8; idx (B*N,4) int32 ascending (d2,index); dist (B*N,4)
 *   Euclidean; qw (B*N,4) normalised blend weights;  for invalid points sigma (B*N) := -1e5
 *   and rgb (B*N,3) := 0 when those pointers are given;  cidx: compacted list of valid point
 *   ids (global id b*N+n), *count incremented atomically (caller zeroes it).
 * ober2cano (B,V,4,4) row-major (rows 0-2 used), lbs_weights (V,J).
 * query_ws (mode 1 only, else NULL): scratch of an_knn_query_ws_bytes(B,N) bytes, 16-byte
 * aligned, holding the work list of the queries that survive the occupancy test.
 * idx/dist hold the exact 4-NN of every query for which the search established them (always in mode
 * 0; in mode 1 every query that may be valid, plus the pruned ones the walk happened to resolve) and
 * -1 / 0 otherwise -- a query with idx -1 is provably farther than dis_threshold from every vertex.
 * Seeds (mode 1, rays+z queries; all NULL/0 when absent): the fine pass of VolumeRenderer.forward
 * (models/volume_rendering.py:199-207) re-queries the coarse samples of the same rays plus Kf new
 * depths.  seed_idx = the idx table (B*R*seed_Kc,4) a previous call wrote for the same rays with
 * K = seed_Kc; seed_src/seed_nn (B*N) u8 come from an_sample_fine_merge_fwd: seed_src[g] < seed_Kc
 * means query g IS that coarse sample (its neighbours are reused, distances re-evaluated: same bits),
 * seed_nn[g] = the coarse sample nearest in depth, whose four neighbours bound the search ball.
 * seed_xyz_cano (B*R*seed_Kc,3), seed_valid (B*R*seed_Kc) u8 and seed_qw (B*R*seed_Kc,4) (optional; qw is needed
 * only when this call emits qw): that previous call's outputs -- a shared sample then takes them as they are
 * instead of re-evaluating the blend (same point, same arithmetic: same bits; not used when dist is requested).
 * Results are bit-identical with and without seeds.                                          */
int64_t an_knn_query_ws_bytes(int B, int64_t N);
int an_knn_unpose_fwd(const float* xyz, const float* rays, const float* z, int B, int R, int K,
                      int64_t N, const float* verts, int V, const void* grid_ws, void* query_ws,
                      const float* ober2cano, const float* lbs_weights, int J,
                      float dis_threshold, int mode,
                      const uint8_t* seed_src, const uint8_t* seed_nn, const int32_t* seed_idx, int seed_Kc,
                      const float* seed_xyz_cano, const uint8_t* seed_valid, const float* seed_qw,
                      float* xyz_cano, uint8_t* valid, int32_t* idx, float* dist, float* qw,
                      float* sigma, float* rgb, int32_t* cidx, int32_t* count, void* stream);

/* ---- A11: ordered list of the valid points ------------------------------------------------------------------------
 * cidx[0..count) = the ids g with valid[g] != 0 in ascending order (valid: n bytes, 16-byte aligned; n < 2^31).
 * an_knn_unpose_fwd can append valid ids itself (cidx/count arguments) but then in scheduling order; this list is a
 * function of the flags only, which makes everything downstream (MLP tiles, the weight gradient's summation order)
 * reproducible bit for bit.  ws: an_compact_ws_bytes() bytes of scratch.  Two launches.                               */
int64_t an_compact_ws_bytes(void);
int an_compact_valid(const uint8_t* valid, int64_t n, int32_t* cidx, int32_t* count, void* ws, void* stream);

/* ---- A5-A8 over the density lattice (cfg4: extract_mesh.py:27-35,152-160) ---------------------------------------
 * Same search / blend / mask as an_knn_unpose_fwd (mode 1) for one frame, the query points generated in the kernel:
 * point (i,j,k) = (x[j], y[i], z[k]) + centre, flat index (i*nj + j)*nk + k -- numpy's 'xy' meshgrid order, the centre
 * added in fp32 as the reference does.  lattice: device floats [cx, cy, cz, (float)nj, x[nj], z[nk], y[ni]], 16-byte
 * aligned.  Outputs as an_knn_unpose_fwd (xyz_cano (N,3), valid (N), sigma/rgb sentinels at invalid points, compact ids). */
int an_knn_unpose_lattice_fwd(const float* lattice, int ni, int nj, int nk, const float* verts, int V,
                              const void* grid_ws, void* query_ws, const float* ober2cano,
                              const float* lbs_weights, int J, float dis_threshold,
                              float* xyz_cano, uint8_t* valid, float* sigma, float* rgb,
                              int32_t* cidx, int32_t* count, void* stream);

/* backward of the blend + affine apply (autograd of models/anim_nerf.py:173-174,188; no
 * gradient through dist/idx/valid, as under the reference's no_grad KNN).
 * g_xyz_cano (B*N,3) is read at the `*count` compacted ids in cidx.  Accumulates (atomically)
 * into g_ober2cano (B,V,4,4) and writes g_xyz (B*N,3) (zero for invalid points) when non-NULL. */
int an_knn_unpose_bwd(const float* g_xyz_cano, const int32_t* cidx, const int32_t* count,
                      const float* xyz, const float* rays, const float* z, int B, int R, int K,
                      int64_t N, int V, const int32_t* idx, const float* qw, const float* ober2cano,
                      float* g_ober2cano, float* g_xyz, void* stream);

/* ---- A9-A11: positional encoding + 8x256 MLP (tcgen05/TMEM) -------------------------------
 * replaces models/embedding.py:22-39 + models/nerf.py:129-175 (NeRF.forward/get_sigma) +
 * the masking in models/anim_nerf.py:304-305.
 *
 * an_mlp_pack: fp32 nn.Linear (out,in) weights -> bf16 UMMA-ready swizzled images (+ fp32
 * biases/heads).  w_host/b_host: HOST arrays of 12 DEVICE pointers in the order
 * xyz_encoding_1..8, xyz_encoding_final, dir_encoding.0, sigma, rgb.0.
 * packed must hold an_mlp_packed_bytes() bytes, 1024-byte aligned.                          */
int64_t an_mlp_packed_bytes(void);
int an_mlp_pack(const float* const* w_host, const float* const* b_host, void* packed, void* stream);

/* forward over the compacted points: for p < *count: point id = cidx[p] (cidx NULL: id = p and
 * the count is n_max), reads xyz_cano[id], writes sigma[id], rgb[id*3..].
 * stash: NULL for inference; else an_mlp_stash_bytes(n_max) bytes receiving the bf16
 * activations the backward needs.                                                            */
int64_t an_mlp_stash_bytes(int64_t n_max);
int an_mlp_fwd(const void* packed, const float* xyz_cano, const int32_t* cidx, const int32_t* count,
               int64_t n_max, float* sigma, float* rgb, void* stash, void* stream);

/* backward: g_sigma (ids), g_rgb (ids,3), rgb = the forward's output (ids,3) -> g_params (fp32 vector
 * of an_mlp_grad_floats() floats: the first 592 388 are the gradient, per nn.Linear weight then
 * bias, in an_mlp_pack's order; the tail is scratch for the fused head layer; accumulated, caller
 * zeroes the whole vector) and g_xyz_cano (ids,3) when non-NULL (caller zeroes).
 * scratch: an_mlp_bwd_scratch_bytes(n_max).  wgrad_ws: an_mlp_wgrad_ws_bytes() bytes, 16-byte aligned: per-CTA
 * partial weight gradients, summed into g_params in a fixed order (no floating-point atomics: for given inputs the
 * gradient is reproducible bit for bit).                                                       */
int64_t an_mlp_grad_floats(void);
int64_t an_mlp_bwd_scratch_bytes(int64_t n_max);
int64_t an_mlp_wgrad_ws_bytes(void);
int an_mlp_bwd(const void* packed, const void* stash, const float* xyz_cano, const float* rgb,
               const int32_t* cidx, const int32_t* count, int64_t n_max,
               const float* g_sigma, const float* g_rgb,
               float* g_params, float* g_xyz_cano, void* scratch, void* wgrad_ws, void* stream);
/* the two stages of an_mlp_bwd, callable separately:
 *   dgrad  activation-gradient chain (writes the dY images to scratch, g_xyz_cano)
 *   wgrad  dW/db of every layer (both heads included) from stash (X) and scratch (dY) on the tensor
 *          cores, then the chain rule from the fused final+colour layer to the two nn.Linear       */
int an_mlp_bwd_dgrad(const void* packed, const void* stash, const float* xyz_cano, const float* rgb,
                     const int32_t* cidx, const int32_t* count, int64_t n_max,
                     const float* g_sigma, const float* g_rgb, float* g_xyz_cano,
                     void* scratch, void* stream);
int an_mlp_bwd_wgrad(const void* packed, const void* stash, const void* scratch, const int32_t* cidx,
                     const int32_t* count, int64_t n_max, float* g_params, void* wgrad_ws, void* stream);

/* ---- A18 (normal-smoothness regulariser): second-order path of NeRF.get_normal --------------
 * replaces the torch double backward of models/nerf.py:177-190 (get_normal: autograd.grad of
 * alpha = 1 - exp(-delta*relu(sigma)) w.r.t. xyz with create_graph=True) as used by
 * train.py:286-309.  With s = d sigma / d xyz (an_mlp_bwd_dgrad run with g_sigma = 1, g_rgb = 0;
 * its g_xyz_cano output is s and its scratch holds delta_l = d sigma / d a_l) and v = dL/ds per
 * point, the weight gradient of L through s is  dL/dW_l = delta_l tau_{l-1}^T  where
 * tau_l = relu'(a_l) * (W_l tau_{l-1}),  tau_0 = (d enc / d xyz) v  is the forward-mode tangent this
 * entry point computes on the tensor cores (same kernel skeleton as an_mlp_fwd; ReLU pattern from
 * the primal stash `pstash`, no biases) and writes to `tstash` in the activation-stash layout, so
 * that an_mlp_bwd_wgrad(packed, tstash, that scratch, ...) accumulates dL/dW (its bias outputs are
 * then meaningless -- biases do not enter s -- and must be discarded by the caller).
 * tvec (ids,3) = v; tsigma (ids) receives w_sigma . tau_8 when non-NULL.
 * tscale (ids) or NULL: c = dL/d sigma per point.  The sigma-only gradient chain is linear in its per-point seed
 * (delta' = c delta), so the first-order term folds into the same product: with tscale the kernel writes
 * T_l = tau_l + c X_l (accumulators start from c * bias, T_0 = tau_0 + c enc(x)), and
 * an_mlp_bwd_wgrad_scaled(packed, tstash, that scratch, bias_scale = c) returns the complete gradient of a loss
 * L(sigma, s): weights = delta T^T, biases = sum_p c_p delta_p (tsigma then holds w_sigma.T_8 + c b_sigma). */
int an_mlp_fwd_tangent(const void* packed, const float* xyz_cano, const float* tvec, const float* tscale,
                       const void* pstash, const int32_t* cidx, const int32_t* count, int64_t n_max,
                       float* tsigma, void* tstash, void* stream);
int an_mlp_bwd_wgrad_scaled(const void* packed, const void* stash, const void* scratch, const int32_t* cidx,
                            const int32_t* count, int64_t n_max, const float* bias_scale, float* g_params,
                            void* wgrad_ws, void* stream);

/* ---- A18 (optimiser): Adam over a list of tensors in one launch -----------------------------------
 * replaces torch.optim.Adam as train.py:217-226 / utils/__init__.py:33-45 configure it (eps 1e-8, betas
 * (0.9, 0.999), L2 weight decay, no amsgrad) for the step's 48 MLP tensors.  params/grads/exp_avg/exp_avg_sq: HOST
 * arrays of n_tensors (<= 64) DEVICE pointers to fp32 tensors of sizes[i] elements.  step: device float, the
 * number of steps taken so far (incremented by the kernel); lr_dev: device float or NULL (then `lr` is used);
 * done_counter: device uint32, zero before the first call (kernel-internal); one_minus_beta*: 1 - beta evaluated
 * in double by the caller (torch does the same; 1.0f - 0.999f differs by 1.3e-5).  Graph-capturable.       */
int an_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                 const int64_t* sizes, int n_tensors, float* step, const float* lr_dev, float lr,
                 float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float eps,
                 float weight_decay, unsigned int* done_counter, void* stream);

/* ---- 8(f)#4: marching cubes of the density lattice -----------------------------------------------------------
 * replaces `mcubes.marching_cubes(-sigmas, 0.)` (extract_mesh.py:165; PyMCubes, CPU).  volume (nx,ny,nz) fp32, last
 * axis fastest; a lattice point is inside when its value < iso.  tri_table: (256,16) int8, cube-edge ids of each
 * configuration's triangles, -1 terminated (bit i of the configuration = corner i inside; corners / edges numbered as
 * in anim-nerf_b200/mesh.py).  Three launches:
 *   an_mc_count  voff (nx*ny*nz) uint16 = each lattice point's vertex offset inside its 1024-point block,
 *                block_counts (ceil(n/1024), 2) int32 = (vertices, triangles) per block;
 *   an_mc_scan   block_counts -> exclusive prefix sums in place, totals[2] int64 = (n_vertices, n_faces): the caller
 *                reads them back and allocates the outputs;
 *   an_mc_emit   (block_counts as left by an_mc_scan, totals as written by it) vertices (n_vertices,3) fp32 in lattice-index coordinates (linear interpolation along the lattice
 *                edge), faces (n_faces,3) int32, normals from inside to outside.  Vertex and face order depend on the
 *                lattice only (no atomics).                                                                          */
int an_mc_count(const float* volume, int nx, int ny, int nz, float iso, const int8_t* tri_table,
                uint16_t* voff, int32_t* block_counts, void* stream);
int an_mc_scan(int32_t* block_counts, int64_t n_blocks, int64_t* totals, void* stream);
int an_mc_emit(const float* volume, int nx, int ny, int nz, float iso, const int8_t* tri_table,
               const uint16_t* voff, const int32_t* block_offsets, const int64_t* totals, float* vertices, int32_t* faces,
               void* stream);

/* ---- A18 (losses): the four render-loss terms and their gradients in one launch ------------------------------------
 * replaces train.py:228-262 (F.mse_loss on rgbs / rgbs_fine, lambda_alphas * F.l1_loss on alphas / alphas_fine) and their
 * autograd.  rgb_* (n_rays,3), acc_* (n_rays), targets likewise; the fine inputs may both be NULL (n_importance = 0 or
 * share_fine).  terms[5] (device) = mse_c, mse_f, l1_c, l1_f, total = mse_c + mse_f + lambda (l1_c + l1_f); g_* receive
 * d total / d input (sign(0) = 0 as torch).  ws: an_render_loss_ws_bytes() of device scratch, zero before its first
 * use (the kernel leaves it zeroed; one scratch per stream).  Partial sums are combined in CTA order: reproducible.   */
int64_t an_render_loss_ws_bytes(void);
int an_render_loss(const float* rgb_coarse, const float* rgb_fine, const float* acc_coarse, const float* acc_fine,
                   const float* tgt_rgb, const float* tgt_acc, int64_t n_rays, float lambda_alphas, float* terms, void* ws,
                   float* g_rgb_coarse, float* g_rgb_fine, float* g_acc_coarse, float* g_acc_fine, void* stream);

/* ---- A12: alpha compositing ------------------------------------------------------------
 * replaces models/volume_rendering.py:128-160 (composite tail), far=True, white_bkgd flag.
 * sigma (n_rays,K), rgb (n_rays,K,3), z (n_rays,K), rays (n_rays,8) (far = rays[:,7]);
 * sigma_noise (n_rays,K) or NULL.  Outputs weights (n_rays,K) (may be NULL), rgb_out
 * (n_rays,3), depth (n_rays), acc (n_rays).                                                  */
int an_composite_fwd(const float* sigma, const float* rgb, const float* z, const float* rays,
                     const float* sigma_noise, int64_t n_rays, int K, int white_bkgd,
                     float* weights, float* rgb_out, float* depth, float* acc, void* stream);
int an_composite_bwd(const float* sigma, const float* rgb, const float* z, const float* rays,
                     const float* sigma_noise, int64_t n_rays, int K, int white_bkgd,
                     const float* g_rgb_out, const float* g_depth, const float* g_acc,
                     float* g_sigma, float* g_rgb, float* g_z, float* g_far, void* stream);

/* ---- A13/A14: inverse-CDF resampling + sort-merge ---------------------------------------
 * replaces models/volume_rendering.py:59-97 sample_fine and :199-207 (mid-points, cat, sort).
 * an_searchsorted_right: the index stage alone (torch.searchsorted(cdf,u,right=True), :85),
 * exposed for the bit-exact index check: cdf (n_rows,M), u (n_rows,F) -> inds int32.
 * an_sample_fine_merge_fwd: weights (n_rays,Kc) coarse weights, z_coarse (n_rays,Kc);
 * u (n_rays,Kf) explicit draws or NULL (det ? linspace(0,1,Kf) : Philox(seed));
 * outputs z_fine (n_rays,Kf) (may be NULL), z_all (n_rays,Kc+Kf) ascending, src (n_rays,Kc+Kf)
 * u8 = index into cat(z_coarse,z_fine) of each sorted entry (may be NULL), nn_coarse (n_rays,Kc+Kf)
 * u8 = the coarse sample nearest in depth to each sorted entry (itself for a coarse entry; may be
 * NULL) -- the seed tables of an_knn_unpose_fwd.                                               */
int an_searchsorted_right(const float* cdf, const float* u, int64_t n_rows, int M, int F,
                          int32_t* inds, void* stream);
int an_sample_fine_merge_fwd(const float* weights, const float* z_coarse, const float* u,
                             int64_t n_rays, int Kc, int Kf, int det, uint64_t seed,
                             float* z_fine, float* z_all, uint8_t* src, uint8_t* nn_coarse, void* stream);

/* ---- A16 (+ vertex part of A2): per-frame tables ------------------------------------------
 * replaces, for frames whose SMPL parameters are not being optimised, the torch chain
 * set_body_model -> SMPL.forward/lbs (smplx/body_models.py:289-387, smplx/lbs.py:152-251),
 * the vertex part of convert_to_body_model_space (models/anim_nerf.py:139-143) and
 * clac_ober2cano_transform (:147-151).  Inputs: betas (B,10), pose (B,24,3) = [global_orient,
 * body_pose] axis-angle, transl (B,3) or NULL -- posed body -- and the same three for the
 * template body with batch Bt (1 = shared by all frames, or B).  Model constants: v_template
 * (V,3), shapedirs (V,3,10), posedirs (207, V*3), J_template (24,3) = J_regressor . v_template,
 * J_shapedirs (24,3,10) = J_regressor . shapedirs, lbs_weights (V,24), parents int32[24].
 * ws: an_body_tables_ws_bytes(B) scratch.  Outputs: verts (B,V,3) in the root frame, ober2cano
 * (B,V,4,4; 16-byte aligned), ginv (B,4,4) = inverse root transform (feeds an_raygen_fwd),
 * verts_template (B,V,3) or NULL.
 *
 * an_body_tables_bwd: autograd of that chain with respect to the POSED body's parameters -- what the reference
 * optimises under its shipped optim_body_params=True (config.py:34, train.py:141-145,330-331; gradients of
 * smplx/lbs.py:152-251 batch_rodrigues / batch_rigid_transform / blend shapes / skinning blend, and of the two
 * matrix inverses of models/anim_nerf.py:131,148).  Inputs: g_ober2cano (B,V,4,4; rows 0-2 read; 16-byte aligned;
 * from an_knn_unpose_bwd), g_ginv (B,4,4; rows 0-2 read) or NULL (gradient of the body-space rays), the forward's
 * inputs, its workspace `ws` (unchanged since the forward) and its ginv output; bwd_ws: an_body_tables_bwd_ws_bytes(B)
 * scratch.  Outputs (written, not accumulated): g_betas (B,10), g_pose (B,24,3), g_transl (B,3) or NULL.  Posed
 * vertices carry no gradient (the neighbour search is not differentiated) and neither do the template parameters. */
int64_t an_body_tables_ws_bytes(int B);
int64_t an_body_tables_bwd_ws_bytes(int B);
int an_body_tables_bwd(const float* g_ober2cano, const float* g_ginv,
                       const float* betas, const float* pose, const float* transl, int B,
                       const float* shapedirs, const float* posedirs,
                       const float* J_template, const float* J_shapedirs, const float* lbs_weights,
                       const int32_t* parents, int V, int J, int n_betas, const void* ws, const float* ginv,
                       void* bwd_ws, float* g_betas, float* g_pose, float* g_transl, void* stream);
int an_body_tables_fwd(const float* betas, const float* pose, const float* transl,
                       const float* betas_t, const float* pose_t, const float* transl_t, int B, int Bt,
                       const float* v_template, const float* shapedirs, const float* posedirs,
                       const float* J_template, const float* J_shapedirs, const float* lbs_weights,
                       const int32_t* parents, int V, int J, int n_betas, void* ws,
                       float* verts, float* ober2cano, float* ginv, float* verts_template, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANIMNERF_B200_H */
