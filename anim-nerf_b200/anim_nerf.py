"""AnimNeRF: host-side mirror of reference `models/anim_nerf.py:AnimNeRF` (42-307).

Same constructor keywords, same per-frame state setters (`set_body_model`,
`convert_to_body_model_space`, `clac_ober2cano_transform`, `set_latent_code`) and the same
`forward(xyz, viewdir, use_fine) -> (rgb, sigma)` / `query_canonical_space` contract; the work
between the query points and (rgb, sigma) runs on the sm_100a kernels (KNN + unpose, MLP) --
there is no torch fallback for it.  The per-frame table builder (SMPL LBS, 6890 4x4 inverses; SURVEY §8
row A16) runs on the fused kernels (`an_body_tables_fwd` / `an_body_tables_bwd`, see `setup_frame`), with gradients to
the posed body's SMPL parameters when they are being optimised (`optim_body_params`, the reference's shipped default);
the differentiable torch builder (`body_model.py`) remains for gradients to the template body's parameters.

Only the shipped configuration is built (every reference yaml): use_unpose=True with k_neigh=4,
use_view=False, use_deformation=False, no latent codes, query_inside=False.
"""
import torch
import torch.nn as nn

from . import ops
from .autograd import BodyTables, PointQuery, RenderPass
from .body_model import BodyModel
from .nerf import NeRF


def batch_transform(P, v, pad_ones=True):
    """(P @ [v;1|0])[:3] for P (...,4,4), v (...,3)  (reference models/anim_nerf.py:31-39)."""
    out = torch.matmul(P[..., :3, :3], v[..., None])[..., 0]
    return out + P[..., :3, 3] if pad_ones else out


def affine_inverse(T):
    """Inverse of affine 4x4 transforms (last row [0,0,0,1]) in closed form: adjugate of the 3x3
    block over its determinant, then -R^-1 t.  Replaces the reference's `torch.inverse`
    (models/anim_nerf.py:131,148): same result to fp32 round-off, but no batched-LU launch chain
    and no host synchronisation (`linalg.inv` reads its `info` back), so the per-frame table
    builder can be captured in a CUDA graph.  Differentiable."""
    R, t = T[..., :3, :3], T[..., :3, 3]
    c0, c1, c2 = R[..., :, 0], R[..., :, 1], R[..., :, 2]
    r0, r1, r2 = torch.cross(c1, c2, dim=-1), torch.cross(c2, c0, dim=-1), torch.cross(c0, c1, dim=-1)
    det = (c0 * r0).sum(-1, keepdim=True)
    Rinv = torch.stack([r0 / det, r1 / det, r2 / det], dim=-2)
    tinv = -torch.matmul(Rinv, t[..., None])
    top = torch.cat([Rinv, tinv], dim=-1)
    bottom = torch.zeros_like(top[..., :1, :])
    bottom[..., 0, 3] = 1.0
    return torch.cat([top, bottom], dim=-2)


class AnimNeRF(nn.Module):
    def __init__(self, model_path="smplx/models", model_type="smpl", gender="male", freqs_xyz=10, freqs_dir=4,
                 use_view=False, use_unpose=False, unpose_view=False, k_neigh=4, use_knn=False,
                 use_deformation=False, deformation_dim=0, apperance_dim=0, use_fine=False, share_fine=False,
                 dis_threshold=0.2, query_inside=False, body_model_data=None, **kwargs):
        super().__init__()
        if use_view or use_deformation or deformation_dim or apperance_dim or query_inside or k_neigh != 4:
            raise NotImplementedError("only the shipped configuration is built: k_neigh=4, use_view=False, "
                                      "no deformation/appearance codes, query_inside=False")
        self.freqs_xyz, self.freqs_dir = freqs_xyz, freqs_dir
        self.use_view, self.use_unpose, self.unpose_view = use_view, use_unpose, unpose_view
        self.k_neigh, self.use_knn = k_neigh, use_knn
        self.use_deformation, self.deformation_dim, self.apperance_dim = use_deformation, deformation_dim, apperance_dim
        self.use_fine, self.share_fine = use_fine, share_fine
        self.dis_threshold, self.query_inside = dis_threshold, query_inside
        self.weight_std = 0.1
        self.knn_mode = 1          # 1 = grid-pruned exact search (default), 0 = exhaustive
        self.fused_tables = True   # per-frame tables on the kernels (forward and backward); False: differentiable torch builder
        if body_model_data is None:
            import os
            path = os.path.join(model_path, model_type, "%s_%s.pkl" % (model_type.upper(), gender.upper()))
            if not os.path.exists(path):      # as smplx.create does (smplx/body_models.py:126-136): no silent stand-in body
                raise FileNotFoundError("SMPL model file %s not found (pass body_model_data=<dict or path> explicitly, "
                                        "e.g. synthetic.make_smpl_dict(0) for tests)" % path)
            body_model_data = path
        self.body_model = BodyModel(body_model_data)
        self.lbs_dim = self.body_model.lbs_weights.shape[1]
        self.nerf = NeRF(freqs_xyz=freqs_xyz, freqs_dir=freqs_dir, use_view=use_view)
        if use_fine:
            self.nerf_fine = self.nerf if share_fine else NeRF(freqs_xyz=freqs_xyz, freqs_dir=freqs_dir, use_view=use_view)
        self._grid = None

    def set_latent_code(self, latent_code):
        pass    # no latent codes in the shipped configuration (deformation_dim = apperance_dim = 0)

    # ------------------------------------------------------------------ per-frame state (A16, A2)
    def set_body_model(self, body_model_params, body_model_params_template=None):
        out = self.body_model(**body_model_params)
        self.verts = out["vertices"]
        self.joints = out["joints"][:, :self.lbs_dim]
        self.verts_transform = out["vertices_transform"]
        self.joints_transform = out["joints_transform"]
        self.shape_offsets, self.pose_offsets = out["shape_offsets"], out["pose_offsets"]
        self.global_transform = out["joints_transform"][:, 0, :, :].clone()
        if body_model_params_template is not None:
            t = self.body_model(**body_model_params_template)
            self.verts_template = t["vertices"]
            self.joints_template = t["joints"][:, :self.lbs_dim]
            self.verts_transform_template = t["vertices_transform"]
            self.joints_transform_template = t["joints_transform"]
            self.shape_offsets_template, self.pose_offsets_template = t["shape_offsets"], t["pose_offsets"]
        self._grid = None
        if torch.is_grad_enabled():         # a training step: the optimiser may have moved the weights
            self.nerf.mark_dirty()
            if self.use_fine:
                self.nerf_fine.mark_dirty()

    # ------------------------------------------------------------------ fused per-frame setup
    def setup_frame(self, body_model_params, body_model_params_template, rays=None):
        """`set_body_model` -> `convert_to_body_model_space(rays)` -> `clac_ober2cano_transform` in one
        call (the sequence of train.py:201-203 / novel_view.py:79-85).  The tables come from the fused kernels
        (`an_body_tables_fwd`: two launches instead of ~250; with gradients to the posed body's parameters through
        `an_body_tables_bwd` when they require them) and only the state the rendering path reads is set (`verts`,
        `ober2cano_transform`, `verts_template`, `global_transform`).  Returns (rays in body space or None, ginv (B,4,4))."""
        posed_grad = torch.is_grad_enabled() and any(torch.is_tensor(v) and v.requires_grad for v in body_model_params.values())
        tmpl_grad = torch.is_grad_enabled() and any(torch.is_tensor(v) and v.requires_grad
                                                    for v in body_model_params_template.values())
        if tmpl_grad or not self.fused_tables:
            # gradients to the TEMPLATE body's parameters (never requested by the reference's training loop) or the
            # torch builder asked for explicitly (tests): the differentiable torch chain
            self.set_body_model(body_model_params, body_model_params_template)
            ginv = affine_inverse(self.global_transform)
            rays = self.convert_to_body_model_space(rays)
            self.clac_ober2cano_transform()
            return rays, ginv
        if posed_grad:      # optim_body_params (the reference's shipped default): kernels forward and backward
            p = body_model_params
            verts, o2c, ginv, vt = BodyTables.apply(self.body_model, body_model_params_template, p["betas"], p["global_orient"],
                                                    p["body_pose"], p.get("transl"))
        else:
            verts, o2c, ginv, vt = ops.body_tables(self.body_model, body_model_params, body_model_params_template)
        self.verts, self.ober2cano_transform, self.verts_template = verts, o2c, vt
        self.global_transform = torch.eye(4, device=verts.device).expand(verts.shape[0], 4, 4)
        self.joints = self.verts_transform = self.joints_transform = None      # not built on the fused path
        self._grid = None
        if torch.is_grad_enabled():
            self.nerf.mark_dirty()
            if self.use_fine:
                self.nerf_fine.mark_dirty()
        if rays is not None:
            rays = self.rays_to_body_space(rays, ginv)
        return rays, ginv

    def clear_frame_state(self):
        """Drop the per-frame tensors (and with them the autograd graph of the step that built them: `ober2cano_transform`
        hangs on to the table builder's node when SMPL parameters are optimised).  Needed before a training step is
        captured on another stream: a live graph keeps the parameters' AccumulateGrad nodes bound to the stream they
        were created on."""
        self.verts = self.ober2cano_transform = self.verts_template = self.global_transform = None
        self.joints = self.verts_transform = self.joints_transform = None
        self._grid = None

    @staticmethod
    def rays_to_body_space(rays, ginv):
        """Ray part of `convert_to_body_model_space` (models/anim_nerf.py:128-137)."""
        g = ginv.unsqueeze(1)
        rays_o = batch_transform(g, rays[:, :, 0:3], True)
        rays_d = batch_transform(g, rays[:, :, 3:6], False)
        cam_dist = torch.norm(rays_o, dim=-1, keepdim=True)
        near = torch.max(rays[:, :, 6:7], cam_dist - 1.0)
        far = torch.min(rays[:, :, 7:8], cam_dist + 1.0)
        return torch.cat((rays_o, rays_d, near, far), dim=-1)

    def convert_to_body_model_space(self, rays):
        """rays=None re-expresses only the per-frame tables (the rays then come from `an_raygen_fwd`,
        which applies the same root-frame transform and near/far clamp while generating them)."""
        ginv = affine_inverse(self.global_transform).unsqueeze(1)            # (bs,1,4,4)
        if rays is not None:
            rays_o = batch_transform(ginv, rays[:, :, 0:3], True)
            rays_d = batch_transform(ginv, rays[:, :, 3:6], False)
            cam_dist = torch.norm(rays_o, dim=-1, keepdim=True)
            near = torch.max(rays[:, :, 6:7], cam_dist - 1.0)
            far = torch.min(rays[:, :, 7:8], cam_dist + 1.0)
        self.verts = batch_transform(ginv, self.verts, True)
        self.joints = batch_transform(ginv, self.joints, True)
        self.global_transform = torch.matmul(ginv.squeeze(1), self.global_transform)
        self.verts_transform = torch.matmul(ginv, self.verts_transform)
        self._grid = None
        if rays is None:
            return None
        return torch.cat((rays_o, rays_d, near, far), dim=-1)

    def clac_ober2cano_transform(self):
        inv = affine_inverse(self.verts_transform)
        shift = (self.shape_offsets_template - self.shape_offsets) + (self.pose_offsets_template - self.pose_offsets)
        inv = torch.cat([inv[..., :3], torch.cat([inv[..., :3, 3:] + shift[..., None], inv[..., 3:, 3:]], dim=-2)], dim=-1)
        self.ober2cano_transform = torch.matmul(self.verts_transform_template, inv)

    # ------------------------------------------------------------------ kernel configuration
    def _cfg(self, use_fine):
        verts = self.verts.detach().contiguous()
        if self._grid is None or self._grid[1] != self.dis_threshold:
            self._grid = (ops.vertex_grid(verts, self.dis_threshold) if self.knn_mode == 1 else None, self.dis_threshold)
        net = self.nerf_fine if use_fine else self.nerf
        return dict(verts=verts, lbs=self.body_model.lbs_weights, grid=self._grid[0], thr=float(self.dis_threshold),
                    net=net, knn_mode=self.knn_mode, unpose=self.use_unpose,
                    grad=torch.is_grad_enabled())

    def render_pass(self, rays, z, use_fine=False, sigma_noise=None, white_bkgd=True, want_seed=False, seed=None):
        """Fused composite pass used by VolumeRenderer: -> (weights, rgb, depth, acc).
        want_seed: keep this pass's neighbour table in `self.last_knn_idx`; seed: dict(src, nn, idx) from
        `sample_fine_merge` + an earlier pass over the same rays (see `ops.knn_unpose`)."""
        if not self.use_unpose:
            raise NotImplementedError("render_pass is built for use_unpose=True (every shipped config)")
        cfg = self._cfg(use_fine)
        cfg["white"] = bool(white_bkgd)
        if self.knn_mode == 1:
            cfg["want_seed"], cfg["seed"] = bool(want_seed), seed
        rgb, depth, acc, w = RenderPass.apply(rays, z, self.ober2cano_transform, sigma_noise, cfg, *cfg["net"].param_list())
        keep = want_seed and self.knn_mode == 1
        self.last_knn_idx = cfg.get("knn_idx") if keep else None
        self.last_knn_out = cfg.get("knn_out") if keep else None
        return w, rgb, depth, acc

    @torch.no_grad()
    def lattice_sigma(self, x_axis, y_rows, z_axis, center, use_fine=False):
        """sigma of `forward` on the lattice points (x[j], y[i], z[k]) + center of the frame set on this model, in the
        order of extract_mesh.py's grid ((i*nj + j)*nk + k), without materialising the points: they are generated in
        the KNN kernel (cfg4).  -> (ni*nj*nk,) fp32, -1e5 at invalid points.  Inference only."""
        if not (self.use_unpose and self.knn_mode == 1):
            raise NotImplementedError("lattice queries run on the grid-pruned KNN kernel with use_unpose=True")
        cfg = self._cfg(use_fine)
        n = y_rows.numel() * x_axis.numel() * z_axis.numel()
        dev = cfg["verts"].device
        sigma, rgb = torch.empty(n, device=dev), torch.empty(n, 3, device=dev)
        out = ops.knn_unpose_lattice(cfg["verts"], self.ober2cano_transform.contiguous(), cfg["lbs"], cfg["thr"], x_axis, y_rows,
                                     z_axis, center, grid=cfg["grid"], sigma=sigma, rgb=rgb)
        ops.mlp_fwd(cfg["net"].packed(), out["xyz_cano"], sigma, rgb, cidx=out["cidx"], count=out["count"], n_max=n)
        return sigma

    # ------------------------------------------------------------------ point queries (B2)
    def unpose(self, xyz, viewdir=None):
        cfg = self._cfg(False)
        out = ops.knn_unpose(cfg["verts"], self.ober2cano_transform.detach(), cfg["lbs"], cfg["thr"], xyz=xyz.detach(),
                             grid=cfg["grid"], mode=self.knn_mode)
        return out["xyz_cano"], viewdir, out["valid"].float().unsqueeze(-1)

    def query_canonical_space(self, xyz, viewdir=None, use_fine=False, only_sigma=False, only_normal=False):
        net = self.nerf_fine if use_fine else self.nerf
        if only_sigma:
            return net.get_sigma(xyz, only_sigma=True)         # regulariser queries, on the kernels
        if only_normal:
            return net.get_normal(xyz)
        return net(xyz)

    def forward(self, xyz, viewdir=None, use_fine=False):
        cfg = self._cfg(use_fine)
        o2c = self.ober2cano_transform if self.use_unpose else None
        return PointQuery.apply(xyz, o2c, cfg, *cfg["net"].param_list())
