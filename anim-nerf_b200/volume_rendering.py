"""VolumeRenderer: host-side mirror of reference `models/volume_rendering.py` (8-232).

Same constructor, same `forward(model, rays, perturb=0., **kw) -> dict` keys and shapes
('rgbs','alphas','depths'[,'rgbs_fine','alphas_fine','depths_fine']).  The stages run on the
sm_100a kernels: stratified sampling, (model.render_pass =) point generation + KNN/unpose + MLP +
alpha compositing, inverse-CDF resampling + sort-merge.  Randomness: `perturb>0` draws come from
an in-kernel Philox stream seeded per call from torch's generator; pass `noise=dict(coarse_u,
fine_u, sigma_c, sigma_f)` to use explicit draws instead (parity tests).

A `model` without `render_pass` (any callable with the reference's
`model(xyz, viewdir, use_fine=...) -> (rgb, sigma)` contract) goes through the generic path:
torch point generation, the callable, then the compositing kernel -- forward only.
"""
import torch
import torch.nn as nn

from . import ops
from .autograd import RaysSample, SampleCoarse, SampleFineMerge


class VolumeRenderer(nn.Module):
    def __init__(self, n_coarse=64, n_fine=0, n_fine_depth=0, share_fine=False, noise_std=1.0, depth_std=0.02,
                 white_bkgd=True, lindisp=True):
        super().__init__()
        if not lindisp:
            raise NotImplementedError("lindisp=False (sampling linear in disparity) is unused by every reference config")
        if n_fine_depth > 0:
            raise NotImplementedError("n_fine_depth > 0 is unused by every reference config (n_depth=0)")
        self.n_coarse, self.n_fine, self.n_fine_depth = n_coarse, n_fine, n_fine_depth
        self.share_fine, self.noise_std, self.depth_std = share_fine, noise_std, depth_std
        self.lindisp, self.white_bkgd = lindisp, white_bkgd
        # False: in-kernel Philox streams seeded per call from torch's CPU generator (a host value, so it
        # would be frozen into a captured CUDA graph).  True: draws come from torch's device generator
        # (graph-safe: its Philox offset advances on every replay) and reach the kernels as explicit tensors.
        self.device_rng = False

    @staticmethod
    def _seed():
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def sample_coarse(self, rays, perturb=0., noise_u=None):
        rays = rays[..., :8].contiguous()
        if self.device_rng and perturb > 0 and noise_u is None:
            noise_u = torch.rand(*rays.shape[:-1], self.n_coarse, device=rays.device)
        return SampleCoarse.apply(rays, self.n_coarse, float(perturb), noise_u,
                                  self._seed() if (perturb > 0 and noise_u is None) else 0)

    def rays_and_coarse_samples(self, ginv, rays_world=None, camera=None, perturb=0., noise_u=None):
        """Fused front end (one launch): rays from `camera` (dict, see `ops.rays_sample`) or the given world-space
        rays, taken to the body's root frame by ginv (None: left as they are) and sampled -> (rays_body, z_coarse)."""
        src = {"rays_world": rays_world[..., :8].contiguous()} if rays_world is not None else {"camera": camera}
        if self.device_rng and perturb > 0 and noise_u is None:
            ref = rays_world if rays_world is not None else camera["c2w"]
            B = ref.shape[0]
            R = rays_world.shape[1] if rays_world is not None else (camera["pix"].shape[1] if camera.get("pix") is not None
                                                                    else camera["H"] * camera["W"])
            noise_u = torch.rand(B, R, self.n_coarse, device=ref.device)
        seed = self._seed() if (perturb > 0 and noise_u is None) else 0
        if ginv is not None and ginv.requires_grad and torch.is_grad_enabled():
            return RaysSample.apply(ginv, src, self.n_coarse, float(perturb), noise_u, seed)
        return ops.rays_sample(self.n_coarse, float(perturb), noise_u, seed, rays_world=src.get("rays_world"),
                               camera=src.get("camera"), ginv=None if ginv is None else ginv.detach())

    def sample_fine_merge(self, z_coarse, weights, det=False, u=None):
        """Fused `sample_fine` + cat + sort (reference :199-207): takes the coarse depths and the full
        coarse weights (the kernel forms the mid-point bins and the w[1:-1] slice itself); returns
        (z_combine sorted, z_fine)."""
        if self.device_rng and not det and u is None:
            u = torch.rand(*z_coarse.shape[:-1], self.n_fine, device=z_coarse.device)
        return SampleFineMerge.apply(weights.detach(), z_coarse, self.n_fine, bool(det), u,
                                     self._seed() if (not det and u is None) else 0)

    def composite(self, model, rays, z_samp, coarse=True, far=True, perturb=0., sigma_noise=None, **kwargs):
        if not far:
            raise NotImplementedError("far=False is never used by the reference's callers")
        bs, n_rays, K = z_samp.shape
        if self.noise_std > 0.0 and perturb > 0 and sigma_noise is None:
            sigma_noise = torch.randn(bs, n_rays, K, device=z_samp.device)
            if self.noise_std != 1.0:       # (x * 1.0 is x: the reference's default needs no second launch)
                sigma_noise = sigma_noise * self.noise_std
        if hasattr(model, "render_pass"):
            return model.render_pass(rays[..., :8].contiguous(), z_samp, use_fine=not coarse,
                                     sigma_noise=sigma_noise, white_bkgd=self.white_bkgd,
                                     want_seed=kwargs.get("want_seed", False), seed=kwargs.get("seed"))
        # generic callable: reference semantics, compositing on the kernel (forward only)
        xyz = (rays[..., None, :3] + z_samp.unsqueeze(-1) * rays[..., None, 3:6]).reshape(bs, -1, 3)
        viewdir = rays[..., None, 3:6].expand(-1, -1, K, -1).reshape(bs, -1, 3)
        rgbs, sigmas = model(xyz, viewdir, use_fine=not coarse, **kwargs)
        w, rgb, depth, acc = ops.composite(sigmas.reshape(bs, n_rays, K).contiguous().float(),
                                           rgbs.reshape(bs, n_rays, K, 3).contiguous().float(),
                                           z_samp.contiguous(), rays[..., :8].contiguous(), self.white_bkgd, sigma_noise)
        return w, rgb, depth, acc

    def forward(self, model, rays, perturb=0., noise=None, ginv=None, camera=None, **kwargs):
        """reference `VolumeRenderer.forward(model, rays, perturb, **kwargs)`.  Extensions: `noise` (explicit draws),
        and the fused front end -- `ginv` (B,4,4): `rays` are WORLD-space rays (or None with `camera`: rays are
        generated from the camera) and the body-space transform + stratified sampling run in one launch."""
        noise = noise or {}
        fused_front = ginv is not None or camera is not None
        if rays is not None:
            rays = rays[..., :8].contiguous()
        if rays is not None and rays.shape[0] * rays.shape[1] == 0:
            # no rays: the reference's torch chain returns empty tensors of the right shapes; the kernels are not launched
            bs, n = rays.shape[:2]
            keys = ["rgbs", "alphas", "depths"]
            if self.n_fine > 0 and not self.share_fine:
                keys += ["rgbs_fine", "alphas_fine", "depths_fine"]
            return {k: rays.new_zeros(bs, n, 3 if k.startswith("rgbs") else 1) for k in keys}
        if fused_front:
            rays, z_coarse = self.rays_and_coarse_samples(ginv, rays_world=rays, camera=camera, perturb=perturb,
                                                          noise_u=noise.get("coarse_u"))
        else:
            z_coarse = self.sample_coarse(rays, perturb=perturb, noise_u=noise.get("coarse_u"))
        no_grad_coarse = self.n_fine > 0 and self.share_fine
        # the fine pass re-queries the coarse samples of the same rays plus n_fine new depths: a fused model
        # hands its coarse-pass neighbour table over as seeds for the fine pass's search (bit-identical results)
        fused = hasattr(model, "render_pass") and self.n_fine > 0
        if fused:
            kwargs = dict(kwargs, want_seed=True)
        with torch.set_grad_enabled(torch.is_grad_enabled() and not no_grad_coarse):
            weights, rgbs, depths, alphas = self.composite(model, rays, z_coarse, coarse=True, far=True, perturb=perturb,
                                                           sigma_noise=noise.get("sigma_c"), **kwargs)
        output = {"rgbs": rgbs, "alphas": alphas, "depths": depths}
        if self.n_fine > 0:
            z_combine, _, src, nn = self.sample_fine_merge(z_coarse, weights, det=(perturb == 0), u=noise.get("fine_u"))
            if fused:
                idx_c, out_c = getattr(model, "last_knn_idx", None), getattr(model, "last_knn_out", None) or {}
                seed = dict(src=src, nn=nn, idx=idx_c, xyz_cano=out_c.get("xyz_cano"), valid=out_c.get("valid"), qw=out_c.get("qw"))
                kwargs = dict(kwargs, want_seed=False, seed=seed if idx_c is not None else None)
            _, rgbs_f, depths_f, alphas_f = self.composite(model, rays, z_combine, coarse=False, far=True, perturb=perturb,
                                                           sigma_noise=noise.get("sigma_f"), **kwargs)
            if fused:
                model.last_knn_idx = model.last_knn_out = None
            if self.share_fine:
                output = {"rgbs": rgbs_f, "alphas": alphas_f, "depths": depths_f}
            else:
                output.update({"rgbs_fine": rgbs_f, "alphas_fine": alphas_f, "depths_fine": depths_f})
        return output
