"""AnimNeRFSystem: host-side mirror of reference `train.py:AnimNeRFSystem` (103-424), duck-type
compatible with a LightningModule (`forward`, `training_step`, `configure_optimizers`,
`compute_loss`, `decode_batch`) but importable without pytorch-lightning (absent here).

`forward(rays (B,h,w,8), body_model_params, body_model_params_template, latent_code, perturb)`
follows train.py:189-215: per-frame tables -> rays to body space -> render -> dict of (B,h,w,.).
The reference's `chunk` loop (train.py:205-210) exists to bound the memory of its materialised
gathers; the fused kernels need no chunking: all rays of the call go through in one pass
(`hparams.render_chunk`, when set, caps the rays per launch; the reference's own `chunk` = 2048 is
accepted in the config and ignored, as it changes no value).

Losses (train.py:228-322): rgb MSE + 0.1 * alpha L1 on coarse and fine run on the render outputs;
the foreground/background density and the normal-smoothness regularisers query the MLP through
`NeRF.get_sigma/get_normal`, which run on the same kernels: the density queries
through `PointQuery`, the normals through `SigmaWithGradient` (forward + dgrad; its backward is the
tensor-core tangent pass + the wgrad kernel instead of the reference's torch double backward).
"""
from collections import defaultdict
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .anim_nerf import AnimNeRF
from .body_model_params import BodyModelParams
from .volume_rendering import VolumeRenderer


def default_hparams(**over):
    """The hot-path-relevant keys with the values of reference config.py:7-78 merged with
    configs/people_snapshot/male-3-casual.yaml (freqs_dir 0, n_importance 32), in the reference's own schema:
    `train.optimizer` / `train.scheduler` are nested nodes (`.type`, `.weight_decay`, `.poly_exp`)."""
    hp = dict(model_path="./smplx/models", model_type="smpl", gender="male", freqs_xyz=10, freqs_dir=0,
              use_view=False, k_neigh=4, use_knn=True, use_unpose=True, unpose_view=False, use_deformation=False,
              deformation_dim=0, apperance_dim=0, latent_dim=0, use_fine=True, share_fine=False, dis_threshold=0.2,
              query_inside=False, n_samples=64, n_importance=32, n_depth=0, chunk=2048, white_bkgd=True,
              optim_body_params=True, num_frames=1,
              train=SimpleNamespace(lr=5e-4, lambda_alphas=0.1, lambda_foreground=0.01, lambda_background=0.01,
                                    lambda_normals=0.01, epsilon=0.01, max_epochs=30,
                                    optimizer=SimpleNamespace(type="adam", momentum=0.9, weight_decay=0),
                                    scheduler=SimpleNamespace(type="poly", poly_exp=0.9)))
    hp.update(over)
    return SimpleNamespace(**hp)


def _node(cfg, name, key, default=None):
    """`cfg.<name>.<key>` of the reference's nested schema (config.py:67-68; yacs CfgNode, dict or namespace), also
    accepting the flat spelling `cfg.<name>` = value / `cfg.<key>`."""
    node = getattr(cfg, name, None)
    if node is not None and not isinstance(node, (str, int, float)):
        if isinstance(node, dict):
            return node.get(key, default)
        return getattr(node, key, default)
    if key == "type" and node is not None:
        return node
    return getattr(cfg, key, default)


class AnimNeRFSystem(nn.Module):
    def __init__(self, hparams=None, body_model_data=None, **over):
        super().__init__()
        self.hparams = hparams if hparams is not None else default_hparams(**over)
        hp = self.hparams
        self.anim_nerf = AnimNeRF(model_path=hp.model_path, model_type=hp.model_type, gender=hp.gender,
                                  freqs_xyz=hp.freqs_xyz, freqs_dir=hp.freqs_dir, use_view=hp.use_view,
                                  k_neigh=hp.k_neigh, use_knn=hp.use_knn, use_unpose=hp.use_unpose,
                                  unpose_view=hp.unpose_view, use_deformation=hp.use_deformation,
                                  deformation_dim=hp.deformation_dim, apperance_dim=hp.apperance_dim,
                                  use_fine=hp.n_importance > 0 or hp.n_depth > 0, share_fine=hp.share_fine,
                                  dis_threshold=hp.dis_threshold, query_inside=hp.query_inside,
                                  body_model_data=body_model_data)
        # train.py:139-146: per-frame SMPL parameter table (filled by `init_body_model_params`; the reference reads
        # <root_dir>/smpls/*.pkl here -- file IO is outside the path, the tensors are handed in instead)
        self.body_model_params = BodyModelParams(hp.num_frames, model_type=hp.model_type)
        if hp.optim_body_params:
            for name in self.body_model_params.param_names:
                self.body_model_params.set_requires_grad(name, True)
        self.volume_renderer = VolumeRenderer(n_coarse=hp.n_samples, n_fine=hp.n_importance, n_fine_depth=hp.n_depth,
                                              share_fine=hp.share_fine, white_bkgd=hp.white_bkgd)

    def init_body_model_params(self, params):
        """train.py:155-165 (`load_body_model_params`) from tensors: params[name] (num_frames, dim) per SMPL parameter."""
        for name in self.body_model_params.param_names:
            self.body_model_params.init_parameters(name, params[name].float(), requires_grad=self.hparams.optim_body_params)

    def load_reference_state_dict(self, state_dict, strict=True):
        """Load the `state_dict` of a reference checkpoint (pytorch-lightning `.ckpt['state_dict']` of
        train.py:AnimNeRFSystem, README.md:110).  Parameter names and shapes are the reference's own
        (`anim_nerf.nerf[_fine].*`, `anim_nerf.body_model.*`, `body_model_params.*.weight`); the smplx module's
        unused default-pose parameters / bookkeeping buffers (`body_model.betas`, `.global_orient`, `.body_pose`,
        `.transl`, `.faces_tensor`, `.vertex_joint_selector.*`) have no counterpart here and are dropped; buffers
        this package derives from the model file (joint template, parents) are kept.  Returns (missing, dropped)."""
        drop = ("anim_nerf.body_model.betas", "anim_nerf.body_model.global_orient", "anim_nerf.body_model.body_pose",
                "anim_nerf.body_model.transl", "anim_nerf.body_model.faces_tensor", "anim_nerf.body_model.vertex_joint_selector.")
        own = self.state_dict()
        sd, dropped = {}, []
        for k, v in state_dict.items():
            if k.startswith(drop) or k.startswith("evaluator."):
                dropped.append(k)
            elif k in own and own[k].shape != v.shape and k.startswith("body_model_params."):
                emb = self.body_model_params
                emb.init_parameters(k.split(".")[1], v.clone(), requires_grad=self.hparams.optim_body_params)   # other frame count
            else:
                sd[k] = v
        res = self.load_state_dict(sd, strict=False)
        derived = ("anim_nerf.body_model.parent_idx", "anim_nerf.body_model.J_template", "anim_nerf.body_model.J_shapedirs",
                   "anim_nerf.body_model.parents_i32")
        missing = [k for k in res.missing_keys if not k.startswith(derived) and not k.startswith("body_model_params.")]
        if strict and (missing or res.unexpected_keys):
            raise RuntimeError("reference checkpoint does not match: missing %s unexpected %s" % (missing, res.unexpected_keys))
        self.anim_nerf.body_model.refresh_derived()       # joint template / tree follow the loaded model buffers
        self.mark_weights_dirty()
        return missing, dropped

    def forward(self, rays, body_model_params, body_model_params_template, latent_code=None, perturb=1.0, noise=None):
        bs, h, w = rays.shape[:3]
        n_rays = h * w
        rays = rays.view(bs, n_rays, rays.shape[-1])
        # per-frame tables; the rays stay in world space: their transform to the body's root frame is fused with the
        # stratified sampling in the renderer's front-end kernel (ginv carries the gradient to the SMPL root)
        _, ginv = self.anim_nerf.setup_frame(body_model_params, body_model_params_template, None)
        chunk = getattr(self.hparams, "render_chunk", None) or n_rays      # the reference's `chunk` (2048) bounds ITS gathers' memory
        results = defaultdict(list)
        for i in range(0, max(n_rays, 1), max(chunk, 1)):
            out = self.volume_renderer(self.anim_nerf, rays[:, i:i + chunk, :], perturb=perturb, ginv=ginv,
                                       noise=None if noise is None else {k: (v[:, i:i + chunk] if v is not None else None) for k, v in noise.items()})
            for k, v in out.items():
                results[k].append(v)
        return {k: torch.cat(v, 1).view(bs, h, w, v[0].shape[-1]) for k, v in results.items()}

    def nets(self):
        a = self.anim_nerf
        return [a.nerf] + ([a.nerf_fine] if a.use_fine and not a.share_fine else [])

    def configure_optimizers(self, flat_grads=None):
        """train.py:217-226 + utils/__init__.py:33-58: Adam(eps 1e-8) on the MLPs at lr, on the SMPL table at lr/2 when
        it is optimised; per-epoch poly decay (1 - epoch/max_epochs)**poly_exp.  Reads the reference's own config
        schema (train.optimizer.{type,weight_decay}, train.scheduler.{type,poly_exp}; config.py:67-68).
        On CUDA the update is `FusedAdam` (one an_adam_step launch per group) and, unless flat_grads=False, the
        gradients live in one `FlatGradBuffer` (`self.flat_grads`): the weight-gradient kernels accumulate into it,
        `optimizer.zero_grad()` is one memset, a data-parallel step all-reduces it in one call."""
        hp = self.hparams
        tr = hp.train
        opt_type = _node(tr, "optimizer", "type", "adam")
        wd = float(_node(tr, "optimizer", "weight_decay", 0.0) or 0.0)
        if opt_type != "adam":
            raise NotImplementedError("train.optimizer.type = 'adam' (every shipped config); got %r" % (opt_type,))
        groups = [{"params": list(self.anim_nerf.parameters()), "lr": tr.lr}]
        body = [p for p in self.body_model_params.parameters() if p.requires_grad] if hp.optim_body_params else []
        if body:
            groups.append({"params": body, "lr": tr.lr * 0.5})
        on_gpu = all(p.is_cuda for g in groups for p in g["params"])
        nets = self.nets()
        if on_gpu and getattr(tr, "fused_adam", True):
            from .optim import FlatGradBuffer, FusedAdam
            self.optimizer = FusedAdam(groups, lr=tr.lr, eps=1e-8, weight_decay=wd)
            if flat_grads is None or flat_grads:
                self.flat_grads = FlatGradBuffer(nets, body)
                self.optimizer.flat = self.flat_grads
            self.optimizer.on_step.append(self.mark_weights_dirty)
        else:               # host-logic tests on CPU tensors
            self.optimizer = torch.optim.Adam(groups, lr=tr.lr, eps=1e-8, weight_decay=wd)
            self.optimizer.register_step_post_hook(lambda *_: self.mark_weights_dirty())
        sched = []
        sched_type = _node(tr, "scheduler", "type", None)
        if sched_type == "poly":
            poly_exp, max_epochs = float(_node(tr, "scheduler", "poly_exp", 0.9)), tr.max_epochs
            self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lambda epoch: (1 - epoch / max_epochs) ** poly_exp)
            sched = [self.scheduler]
        elif sched_type not in (None, "none"):
            raise NotImplementedError("train.scheduler.type = 'poly' (every shipped config); got %r" % (sched_type,))
        return [self.optimizer], sched

    def mark_weights_dirty(self):
        """The optimiser moved the weights: the bf16 images the kernels read are repacked on their next use, in any
        grad mode (FusedAdam / graph replays write the parameters through raw pointers, invisible to `Tensor._version`)."""
        for net in self.nets():
            net.mark_dirty()

    def compute_loss(self, rgbs, alphas, results, frame_idx=None, latent_code=None, fg_points=None, bg_points=None,
                     with_regularizers=True):
        hp = self.hparams
        fine = hp.n_importance > 0 and not hp.share_fine
        loss, det = 0, {}

        def add(name, value, weight=1.0):
            nonlocal loss
            loss = loss + weight * value
            det[name] = value
        if results["rgbs"].is_cuda:      # the four render terms and their gradients in one launch (an_render_loss)
            from .autograd import RenderLoss
            total, terms = RenderLoss.apply(results["rgbs"], results["rgbs_fine"] if fine else None, results["alphas"],
                                            results["alphas_fine"] if fine else None, rgbs, alphas, float(hp.train.lambda_alphas))
            loss = total
            det.update(loss_rgb=terms[0], loss_alphas=terms[2])
            if fine:
                det.update(loss_rgb_fine=terms[1], loss_alphas_fine=terms[3])
        else:                            # host tensors (CPU unit tests of the loss arithmetic)
            add("loss_rgb", F.mse_loss(results["rgbs"], rgbs))
            if fine:
                add("loss_rgb_fine", F.mse_loss(results["rgbs_fine"], rgbs))
            add("loss_alphas", F.l1_loss(results["alphas"], alphas), hp.train.lambda_alphas)
            if fine:
                add("loss_alphas_fine", F.l1_loss(results["alphas_fine"], alphas), hp.train.lambda_alphas)
        if not with_regularizers:
            return loss, det
        k = -2.0 / hp.n_samples
        q = self.anim_nerf.query_canonical_space
        if hp.use_unpose and (fg_points is not None or bg_points is not None):
            # train.py:264-284; foreground and background points of a net go through one MLP launch
            n_fg = fg_points.shape[1] if fg_points is not None else 0
            pts = torch.cat([p for p in (fg_points, bg_points) if p is not None], 1)
            for use_fine, sfx in (((False, ""), (True, "_fine")) if fine else ((False, ""),)):
                e = torch.exp(k * torch.relu(q(pts, use_fine=use_fine, only_sigma=True)))
                if fg_points is not None:
                    add("loss_foreground" + sfx, torch.mean(e[:, :n_fg]), hp.train.lambda_foreground)
                if bg_points is not None:
                    add("loss_background" + sfx, torch.mean(1 - e[:, n_fg:]), hp.train.lambda_background)
        points = self.anim_nerf.verts_template.detach().clone()
        points = points + torch.randn_like(points) * hp.dis_threshold * 0.5
        neighbs = points + torch.randn_like(points) * hp.train.epsilon

        def unit(v):
            return v / (torch.norm(v, p=2, dim=-1, keepdim=True) + 1e-5)

        def normals_loss(use_fine):
            # both point sets in one query (one forward/dgrad launch pair per net instead of two)
            n = q(torch.cat([points, neighbs], 0), use_fine=use_fine, only_normal=True)
            return F.mse_loss(unit(n[:points.shape[0]]), unit(n[points.shape[0]:]))
        add("loss_normals", normals_loss(False), hp.train.lambda_normals)
        if fine:
            add("loss_normals_fine", normals_loss(True), hp.train.lambda_normals)
        return loss, det

    def decode_batch(self, batch):
        """train.py:167-187.  Accepts the reference dataset's flat keys (`betas`, ..., `betas_template`, ...) or the two
        dicts already assembled (`body_model_params`, `body_model_params_template`)."""
        g = batch.get
        names = ("betas", "global_orient", "body_pose", "transl")
        params = batch["body_model_params"] if "body_model_params" in batch else {k: batch[k] for k in names}
        params_t = batch["body_model_params_template"] if "body_model_params_template" in batch \
            else {k: batch[k + "_template"] for k in names}
        return (g("frame_id"), g("cam_id"), g("frame_idx"), batch["rays"], batch["rgbs"], batch["alphas"],
                params, params_t, g("fg_points"), g("bg_points"))

    def training_step(self, batch, batch_idx=0, with_regularizers=True):
        (_, _, frame_idx, rays, rgbs, alphas, params, params_t, fg, bg) = self.decode_batch(batch)
        if self.hparams.optim_body_params:                       # train.py:330-331: the table's rows, not the batch's copies
            if frame_idx is None:
                raise KeyError("training_step with optim_body_params=True needs batch['frame_idx'] (rows of the SMPL table)")
            params = self.body_model_params(frame_idx)
        results = self(rays, params, params_t)
        loss, details = self.compute_loss(rgbs, alphas, results, frame_idx=frame_idx, fg_points=fg, bg_points=bg,
                                          with_regularizers=with_regularizers)
        self.last_details = details
        return loss
