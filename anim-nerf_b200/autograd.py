"""torch.autograd glue: the CUDA kernels as differentiable functions.

`RenderPass` is one composite pass of the reference (`VolumeRenderer.composite`,
models/volume_rendering.py:113-160) with everything between the ray samples and the per-ray
outputs on kernels: point generation + KNN + unpose -> compaction of valid points -> MLP ->
compositing.  Its backward chains composite_bwd -> mlp_bwd -> knn_unpose_bwd and returns
gradients for the MLP parameters, the per-vertex observation->canonical table, the rays and the
sample depths (the three routes to the SMPL parameters, SURVEY §7 traps).  As in the reference
there is no gradient through KNN distances/indices, `valid`, or z_fine.
"""
import torch

from . import ops


COUNT_LOG = None      # bench.py sets this to a list to collect (K, device count tensor) per pass


def _param_grads(net, g_flat, sink):
    """What a backward returns to autograd for the 24 MLP parameters: per-tensor views of this call's gradient
    vector, or nothing when the kernel accumulated straight into the net's attached flat buffer."""
    if sink is not None:
        return [None] * 24
    return net.split_flat_grad(g_flat)


class RenderPass(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays, z, ober2cano, sigma_noise, cfg, *params):
        """rays (B,R,8) body space, z (B,R,K); cfg = dict(verts, lbs, grid, thr, net, white, knn_mode)."""
        net = cfg["net"]
        B, R, K = z.shape
        dev = z.device
        # Function.forward always runs with grad mode off and needs_input_grad ignores torch.no_grad():
        # the caller records the ambient grad mode in cfg["grad"] (inference must not write the stash)
        need_grad = cfg.get("grad", True) and any(ctx.needs_input_grad)
        rays_c, z_c, o2c_c = rays.contiguous(), z.contiguous(), ober2cano.contiguous()
        sigma = torch.empty(B, R, K, device=dev)
        rgb = torch.empty(B, R, K, 3, device=dev)
        # cfg["seed"]: (src, nn, idx) of the coarse pass over the same rays -- the fine pass reuses / starts
        # from those neighbours; cfg["want_seed"]: this pass's idx table is handed back in cfg["knn_idx"]
        want_idx = need_grad or cfg.get("want_seed", False)
        out = ops.knn_unpose(cfg["verts"], o2c_c, cfg["lbs"], cfg["thr"], rays=rays_c, z=z_c, grid=cfg["grid"],
                             mode=cfg.get("knn_mode", 1), want_idx=want_idx, want_qw=need_grad,
                             sigma=sigma, rgb=rgb, compact=True, seed=cfg.get("seed"))
        cfg["knn_idx"] = out["idx"]
        cfg["knn_out"] = out          # the fine pass takes over this pass's results at the samples they share
        if COUNT_LOG is not None:
            COUNT_LOG.append((K, out["count"]))
        packed = net.packed()
        stash = ops.mlp_stash(B * R * K, dev) if need_grad else None
        ops.mlp_fwd(packed, out["xyz_cano"], sigma, rgb, cidx=out["cidx"], count=out["count"], n_max=B * R * K,
                    stash=stash)
        w, rgb_o, depth, acc = ops.composite(sigma, rgb, z_c, rays_c, cfg["white"], sigma_noise)
        if need_grad:
            ctx.cfg = cfg
            ctx.packed, ctx.stash, ctx.aux = packed, stash, out
            ctx.save_for_backward(rays_c, z_c, o2c_c, sigma, rgb, sigma_noise if sigma_noise is not None else torch.empty(0))
        ctx.mark_non_differentiable(w)
        return rgb_o, depth, acc, w

    @staticmethod
    def backward(ctx, g_rgb_o, g_depth, g_acc, _gw):
        rays, z, o2c, sigma, rgb, noise = ctx.saved_tensors
        noise = noise if noise.numel() else None
        cfg, aux = ctx.cfg, ctx.aux
        net = cfg["net"]
        B, R, K = z.shape
        g_sigma, g_rgb, g_z, g_far = ops.composite_bwd(sigma, rgb, z, rays, g_rgb_o, g_depth, g_acc, cfg["white"], noise)
        need_geo = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        sink = net.grad_sink()
        g_flat, g_xc = ops.mlp_bwd(ctx.packed, ctx.stash, aux["xyz_cano"], rgb, g_sigma, g_rgb, cidx=aux["cidx"],
                                   count=aux["count"], n_max=B * R * K, want_g_xyz=need_geo, g_params=sink)
        g_rays = g_zz = g_o2c = None
        if need_geo:
            g_o2c, g_xyz = ops.knn_unpose_bwd(g_xc, aux["cidx"], aux["count"], aux["idx"], aux["qw"], o2c, rays=rays, z=z,
                                              zero_g_xyz=False)
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:      # x = o + z d: one warp-per-ray kernel
                g_rays, g_zz = ops.ray_point_grad(rays, z, aux["valid"], g_xyz, g_z, g_far)
            if not ctx.needs_input_grad[2]:
                g_o2c = None
        g_params = _param_grads(net, g_flat, sink)
        ctx.stash = ctx.aux = None
        return (g_rays, g_zz, g_o2c, None, None) + tuple(g_params)


class BodyTables(torch.autograd.Function):
    """Per-frame tables (A16 + vertex part of A2 + clac_ober2cano_transform) as a differentiable function of the POSED
    body's parameters: forward `an_body_tables_fwd`, backward `an_body_tables_bwd` (the reference's shipped
    optim_body_params=True path, train.py:141-145,330-331).  Outputs: verts (root frame; no gradient -- the
    neighbour search is not differentiated), ober2cano, ginv, verts_template (no gradient).  The template body's
    parameters come from the batch and receive no gradient."""

    @staticmethod
    def forward(ctx, model, template, betas, global_orient, body_pose, transl):
        posed = dict(betas=betas, global_orient=global_orient, body_pose=body_pose, transl=transl)
        verts, o2c, ginv, vt, c = ops.body_tables(model, posed, template, want_ctx=True)
        ctx.c = c
        ctx.shapes = (betas.shape, global_orient.shape, body_pose.shape, None if transl is None else transl.shape)
        ctx.mark_non_differentiable(verts, vt)
        return verts, o2c, ginv, vt

    @staticmethod
    def backward(ctx, _gv, g_o2c, g_ginv, _gvt):
        c = ctx.c
        if g_o2c is None:
            g_o2c = torch.zeros(c["B"], c["V"], 4, 4, device=c["pose"].device)
        g_betas, g_pose, g_transl = ops.body_tables_bwd(c, g_o2c, g_ginv)
        sb, sg, sp, st = ctx.shapes
        if sb[0] != g_betas.shape[0]:                   # one shared shape row (BodyModelParams.betas): sum over the frames
            g_betas = g_betas.sum(0, keepdim=True)
        g_go = g_pose[:, :3].reshape(sg)
        g_bp = g_pose[:, 3:].reshape(sp)
        ctx.c = None
        return None, None, g_betas.reshape(sb), g_go, g_bp, (None if st is None else g_transl.reshape(st))


class PointQuery(torch.autograd.Function):
    """AnimNeRF.forward / NeRF.forward on explicit points: (optional unpose) -> MLP -> mask."""

    @staticmethod
    def forward(ctx, xyz, ober2cano, cfg, *params):
        net = cfg["net"]
        B, N = xyz.shape[:2]
        dev = xyz.device
        need_grad = cfg.get("grad", True) and any(ctx.needs_input_grad)
        xyz_c = xyz.contiguous()
        sigma = torch.empty(B, N, device=dev)
        rgb = torch.empty(B, N, 3, device=dev)
        packed = net.packed()
        stash = ops.mlp_stash(B * N, dev) if need_grad else None
        if cfg.get("unpose", True):
            o2c_c = ober2cano.contiguous()
            out = ops.knn_unpose(cfg["verts"], o2c_c, cfg["lbs"], cfg["thr"], xyz=xyz_c, grid=cfg["grid"],
                                 mode=cfg.get("knn_mode", 1), want_idx=need_grad, want_qw=need_grad,
                                 sigma=sigma, rgb=rgb, compact=True)
            ops.mlp_fwd(packed, out["xyz_cano"], sigma, rgb, cidx=out["cidx"], count=out["count"], n_max=B * N,
                        stash=stash)
        else:
            o2c_c, out = None, dict(xyz_cano=xyz_c, cidx=None, count=None)
            ops.mlp_fwd(packed, xyz_c, sigma, rgb, n_max=B * N, stash=stash)
        if need_grad:
            ctx.cfg, ctx.packed, ctx.stash, ctx.aux = cfg, packed, stash, out
            ctx.save_for_backward(xyz_c, o2c_c if o2c_c is not None else torch.empty(0), rgb)
        return rgb, sigma.unsqueeze(-1)

    @staticmethod
    def backward(ctx, g_rgb, g_sigma):
        xyz, o2c, rgb = ctx.saved_tensors
        cfg, aux = ctx.cfg, ctx.aux
        B, N = xyz.shape[:2]
        need_geo = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        sink = cfg["net"].grad_sink()
        g_flat, g_xc = ops.mlp_bwd(ctx.packed, ctx.stash, aux["xyz_cano"], rgb, g_sigma.contiguous().view(B, N),
                                   g_rgb.contiguous(), cidx=aux["cidx"], count=aux["count"], n_max=B * N,
                                   want_g_xyz=need_geo, g_params=sink)
        g_xyz = g_o2c = None
        if need_geo:
            if cfg.get("unpose", True):
                g_o2c, g_xyz = ops.knn_unpose_bwd(g_xc, aux["cidx"], aux["count"], aux["idx"], aux["qw"], o2c, xyz=xyz)
                if not ctx.needs_input_grad[1]:
                    g_o2c = None
            else:
                g_xyz = g_xc.view_as(xyz)
        ctx.stash = ctx.aux = None
        return (g_xyz, g_o2c, None) + tuple(_param_grads(cfg["net"], g_flat, sink))


class SigmaWithGradient(torch.autograd.Function):
    """Canonical-space density and its spatial gradient: xyz (...,3) -> sigma (...,1), s = d sigma/d xyz (...,3),
    both differentiable with respect to the MLP parameters.  This is the second-order path behind the
    reference's `NeRF.get_normal` (models/nerf.py:177-190: `autograd.grad(alpha, xyz, create_graph=True)`),
    used by the normal-smoothness regulariser (train.py:286-309), without torch double backward:

      forward   an_mlp_fwd (stash) -> sigma;  an_mlp_bwd_dgrad with g_sigma = 1, g_rgb = 0 -> s and the
                delta_l = d sigma/d a_l images (kept)
      backward  T = an_mlp_fwd_tangent(v = dL/ds, c = dL/dsigma): forward-mode tangent on the tensor cores plus c times
                the primal activations, then ONE an_mlp_bwd_wgrad_scaled(T images, delta images, c): weights
                delta T^T (second-order term delta tau^T + first-order term (c delta) X^T), biases sum c delta.
    There is no gradient to xyz (the regulariser's sample points are detached in the reference)."""

    @staticmethod
    def forward(ctx, xyz, net, *params):
        lead = xyz.shape[:-1]
        x = xyz.detach().reshape(-1, 3).contiguous().float()
        n, dev = x.shape[0], x.device
        sigma = torch.empty(n, device=dev)
        rgb = torch.empty(n, 3, device=dev)
        packed = net.packed()
        stash = ops.mlp_stash(n, dev)
        ops.mlp_fwd(packed, x, sigma, rgb, n_max=n, stash=stash)
        scratch = ops.mlp_bwd_scratch(n, dev)
        s = ops.mlp_bwd_dgrad(packed, stash, x, rgb, torch.ones(n, device=dev), torch.zeros(n, 3, device=dev), scratch,
                              n_max=n, want_g_xyz=True)
        ctx.net, ctx.packed, ctx.stash, ctx.scratch, ctx.n = net, packed, stash, scratch, n
        ctx.save_for_backward(x, rgb)
        return sigma.view(*lead, 1), s.view(*lead, 3)

    @staticmethod
    def backward(ctx, g_sigma, g_s):
        x, rgb = ctx.saved_tensors
        net, packed, stash, scratch, n = ctx.net, ctx.packed, ctx.stash, ctx.scratch, ctx.n
        dev = x.device
        # One tangent pass + one weight-gradient pass give the whole gradient.  With v = dL/ds and c = dL/dsigma:
        #   through s:      dW_l = sum_p delta_l tau_{l-1}^T,            no bias term
        #   through sigma:  dW_l = sum_p (c delta_l) X_{l-1}^T,          db_l = sum_p c delta_l
        # (the sigma-only activation-gradient chain is linear in its per-point seed, so its images are c * delta).
        # The tangent kernel writes T = tau + c X; the wgrad kernel forms delta T^T and the c-weighted bias sums.
        c = g_sigma.reshape(n).contiguous().float()
        tstash, _ = ops.mlp_fwd_tangent(packed, x, g_s.reshape(n, 3).contiguous(), stash, n_max=n, tscale=c)
        sink = net.grad_sink()
        flat = ops.mlp_bwd_wgrad(packed, tstash, scratch, n_max=n, bias_scale=c, g_params=sink)
        ctx.stash = ctx.scratch = None
        return (None, None) + tuple(_param_grads(net, flat, sink))


def sigma_with_gradient(net, xyz):
    return SigmaWithGradient.apply(xyz, net, *net.param_list())


def mlp_query(net, xyz):
    """Canonical-space NeRF query (no unposing): xyz (B,N,3) -> rgb (B,N,3), sigma (B,N,1)."""
    cfg = dict(net=net, unpose=False, grad=torch.is_grad_enabled())
    return PointQuery.apply(xyz, None, cfg, *net.param_list())


class RaysSample(torch.autograd.Function):
    """A1 + A2 + A3 fused (`an_rays_sample_fwd`): rays from a camera or given world-space rays -> body-space rays
    (models/anim_nerf.py:128-137) and stratified depths (models/volume_rendering.py:29-56) in one launch.
    Differentiable with respect to ginv (the inverse SMPL root transform): `an_rays_sample_bwd`."""

    @staticmethod
    def forward(ctx, ginv, src, n_coarse, perturb, noise_u, seed):
        rays_body, z = ops.rays_sample(n_coarse, perturb, noise_u, seed, rays_world=src.get("rays_world"),
                                       camera=src.get("camera"), ginv=ginv)
        ctx.src = src
        ctx.save_for_backward(rays_body, z)
        return rays_body, z

    @staticmethod
    def backward(ctx, g_rays, g_z):
        rays_body, z = ctx.saved_tensors
        if g_rays is None:
            g_rays = torch.zeros_like(rays_body)
        g_ginv = ops.rays_sample_bwd(rays_body, z, g_rays.contiguous(), None if g_z is None else g_z.contiguous(),
                                     rays_world=ctx.src.get("rays_world"), camera=ctx.src.get("camera"))
        return g_ginv, None, None, None, None, None


class SampleCoarse(torch.autograd.Function):
    """models/volume_rendering.py:29-56.  z is affine in (near, far): z = near + s*(far-near)."""

    @staticmethod
    def forward(ctx, rays, n_coarse, perturb, noise_u, seed):
        z = ops.sample_coarse(rays, n_coarse, perturb, noise_u, seed)
        ctx.save_for_backward(rays, z)
        return z

    @staticmethod
    def backward(ctx, g_z):
        rays, z = ctx.saved_tensors
        near, far = rays[..., 6:7], rays[..., 7:8]
        s = (z - near) / (far - near)
        g = torch.zeros_like(rays)
        g[..., 6] = (g_z * (1 - s)).sum(-1)
        g[..., 7] = (g_z * s).sum(-1)
        return g, None, None, None, None


class SampleFineMerge(torch.autograd.Function):
    """models/volume_rendering.py:59-97 + :199-207.  z_fine is detached in the reference; the sorted
    union carries the coarse depths' gradient through the sort permutation."""

    @staticmethod
    def forward(ctx, weights, z_coarse, n_fine, det, u, seed):
        z_fine, z_all, src, nn = ops.sample_fine_merge(weights, z_coarse, n_fine, det, u, seed)
        ctx.save_for_backward(src)
        ctx.kc = z_coarse.shape[-1]
        ctx.mark_non_differentiable(z_fine, src, nn)
        return z_all, z_fine, src, nn

    @staticmethod
    def backward(ctx, g_all, _g_fine, _g_src, _g_nn):
        (src,) = ctx.saved_tensors
        g_cat = torch.zeros_like(g_all).scatter_(-1, src.long(), g_all)
        return None, g_cat[..., :ctx.kc].contiguous(), None, None, None, None


class RenderLoss(torch.autograd.Function):
    """train.py:228-262: mse(rgbs) + mse(rgbs_fine) + lambda_alphas (l1(alphas) + l1(alphas_fine)) and its gradients in
    one launch (ops.render_loss).  -> (total, terms (4,) detached: mse_c, mse_f, l1_c, l1_f)."""

    @staticmethod
    def forward(ctx, rgb_c, rgb_f, acc_c, acc_f, tgt_rgb, tgt_acc, lam):
        terms, g = ops.render_loss(rgb_c, rgb_f, acc_c, acc_f, tgt_rgb, tgt_acc, lam)
        ctx.shapes = [t.shape if t is not None else None for t in (rgb_c, rgb_f, acc_c, acc_f)]
        ctx.save_for_backward(*[t for t in g if t is not None])
        ctx.mark_non_differentiable(terms)
        return terms[4], terms[:4]

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        saved = list(ctx.saved_tensors)
        out = []
        for shp in ctx.shapes:
            out.append(None if shp is None else (saved.pop(0) * g_total).view(shp))
        return tuple(out) + (None, None, None)
