"""Run-to-run determinism probe: the same render + backward twice from identical weights; which outputs / gradients differ
bitwise.  python tools/probe_determinism.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from util import oracle, body_model, synthetic
from anim_nerf_b200.anim_nerf import AnimNeRF
from anim_nerf_b200.volume_rendering import VolumeRenderer
from anim_nerf_b200.optim import FlatGradBuffer, FusedAdam

DEV = "cuda"
bm = body_model()
posed_np, tmpl_np = synthetic.make_body_params(1, seed=5)
posed = {k: torch.from_numpy(v) for k, v in posed_np.items()}
tmpl = {k: torch.from_numpy(v) for k, v in tmpl_np.items()}
with torch.no_grad():
    po = bm(**posed)
    rays_w = torch.from_numpy(synthetic.rays_at_bbox(po["vertices"].numpy(), 64, seed=4, margin=0.02))
    rays = oracle.rays_to_body_space(rays_w, po["joints_transform"][:, 0])
posed_d, tmpl_d = {k: v.to(DEV) for k, v in posed.items()}, {k: v.to(DEV) for k, v in tmpl.items()}
rays_d = rays.to(DEV)
tgt = torch.rand(1, 64, 3, device=DEV)


def run(steps):
    net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=True, freqs_dir=0, body_model_data=synthetic.make_smpl_dict(0)).to(DEV)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):
        getattr(net, name).load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}, strict=True)
    vr = VolumeRenderer(n_coarse=64, n_fine=64, white_bkgd=True)
    params = [p for n in ("nerf", "nerf_fine") for p in getattr(net, n).parameters()]
    opt = FusedAdam(params, lr=5e-4, eps=1e-8)
    flat = FlatGradBuffer([net.nerf, net.nerf_fine])
    opt.flat = flat
    opt.on_step.append(lambda: (net.nerf.mark_dirty(), net.nerf_fine.mark_dirty()))
    rec = []
    for s in range(steps):
        net.setup_frame(posed_d, tmpl_d, None)
        out = vr(net, rays_d, perturb=0.0)
        loss = ((out["rgbs"] - tgt) ** 2).mean() + ((out["rgbs_fine"] - tgt) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        rec.append({"rgbs": out["rgbs"].detach().clone(), "rgbs_fine": out["rgbs_fine"].detach().clone(),
                    "depths_fine": out["depths_fine"].detach().clone(), "grad": flat.buf.detach().clone() if hasattr(flat, "buf") else torch.cat([p.grad.flatten() for p in params])})
        opt.step()
        rec[-1]["w"] = torch.cat([p.detach().flatten() for p in params])
    return rec

a = run(6); b = run(6)
for s, (ra, rb) in enumerate(zip(a, b)):
    line = "step %d:" % s
    for k in ra:
        d = (ra[k] - rb[k]).abs()
        line += "  %s max|d| %.3g (n!= %d, ref %.3g)" % (k, float(d.max()), int((d != 0).sum()), float(ra[k].abs().max()))
    print(line)
