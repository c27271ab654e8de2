"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle."""
import os
import sys

import numpy as np
import torch


def run(n_rays=256, n_coarse=64, n_fine=64):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import animnerf_oracle as oracle       # checker only
    from . import synthetic
    from .anim_nerf import AnimNeRF
    from .volume_rendering import VolumeRenderer
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device (B200); there is no CPU fallback")
    dev = torch.device("cuda:0")
    data = synthetic.make_smpl_dict(0)
    net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=True, freqs_dir=0, body_model_data=data).to(dev)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):
        sd = {k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}
        getattr(net, name).load_state_dict(sd, strict=True)
    posed_np, tmpl_np = synthetic.make_body_params(1, seed=1)
    posed = {k: torch.from_numpy(v).to(dev) for k, v in posed_np.items()}
    tmpl = {k: torch.from_numpy(v).to(dev) for k, v in tmpl_np.items()}
    renderer = VolumeRenderer(n_coarse=n_coarse, n_fine=n_fine, white_bkgd=True)
    net.set_body_model(posed, tmpl)
    rays_w = torch.from_numpy(synthetic.rays_at_bbox(net.verts.detach().cpu().numpy(), n_rays, seed=2)).to(dev)
    rays = net.convert_to_body_model_space(rays_w)
    net.clac_ober2cano_transform()
    # forward + backward through the kernels
    out = renderer(net, rays, perturb=0.0)
    loss = ((out["rgbs_fine"] - 0.5) ** 2).mean() + ((out["rgbs"] - 0.5) ** 2).mean()
    loss.backward()
    gnorm = float(torch.sqrt(sum((p.grad ** 2).sum() for p in net.nerf_fine.parameters())))
    torch.cuda.synchronize()
    # oracle on the same tables
    pc = {n: (getattr_path(net.nerf, n).weight.detach().cpu(), getattr_path(net.nerf, n).bias.detach().cpu()) for n in synthetic.NERF_LAYER_NAMES}
    pf = {n: (getattr_path(net.nerf_fine, n).weight.detach().cpu(), getattr_path(net.nerf_fine, n).bias.detach().cpu()) for n in synthetic.NERF_LAYER_NAMES}
    tables = (net.verts.detach().cpu(), net.ober2cano_transform.detach().cpu(), net.body_model.lbs_weights.cpu())
    ref = oracle.render_rays(pc, pf, rays.detach().cpu(), tables, n_coarse=n_coarse, n_fine=n_fine, perturb=0.0)
    errs = {k: float((out[k].detach().cpu() - ref[k]).abs().max()) for k in ("rgbs", "alphas", "rgbs_fine", "alphas_fine")}
    print("smoke: max abs err vs oracle", errs, "| grad norm (fine net) %.4g" % gnorm)
    assert errs["rgbs"] < 1e-2 and errs["alphas"] < 1e-2, errs
    assert (out["rgbs_fine"].detach().cpu() - ref["rgbs_fine"]).abs().mean() < 2e-3
    assert np.isfinite(gnorm) and gnorm > 0
    return errs


def getattr_path(mod, path):
    for part in path.split("."):
        mod = mod[int(part)] if part.isdigit() else getattr(mod, part)
    return mod
