"""Generate golden fixtures by importing and running the UNMODIFIED reference
(/root/reference, JanaldoChen/Anim-NeRF) on CPU in the build container.

    python tests/golden/make_golden.py          # writes tests/golden/*.npz

The reference cannot travel to the GPU box, the fixtures can.  Recipe = SURVEY
Appendix B: synthetic SMPL pickle, `knn_cuda` shim (torch.cdist without the matmul
shortcut + topk, under no_grad), `AnimNeRF(use_knn=True, use_unpose=True, ...)`,
`set_body_model -> convert_to_body_model_space -> clac_ober2cano_transform ->
VolumeRenderer.forward`.  Intermediates are captured by wrapping (never editing)
reference callables.  Nothing from the reference is copied into this repository.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import anim_nerf_b200 as pkg  # noqa: E402
from anim_nerf_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class ShimKNN:
    def __init__(self, k, transpose_mode=True):
        self.k = k

    def __call__(self, ref, query):
        with torch.no_grad():
            d = torch.cdist(query, ref, compute_mode="donot_use_mm_for_euclid_dist")
            return d.topk(self.k, largest=False, dim=-1)


def install_shim():
    m = types.ModuleType("knn_cuda")
    m.KNN = ShimKNN
    sys.modules["knn_cuda"] = m


def to_t(d):
    return {k: torch.from_numpy(np.asarray(v)).float() for k, v in d.items()}


def build_reference(tmp, trained_scale=False):
    install_shim()
    from models.anim_nerf import AnimNeRF
    from models.volume_rendering import VolumeRenderer
    model_path = synthetic.write_smpl_pickle(tmp)
    net = AnimNeRF(model_path=model_path, model_type="smpl", gender="male", freqs_xyz=10, freqs_dir=0,
                   use_view=False, use_unpose=True, k_neigh=4, use_knn=True, use_fine=True,
                   share_fine=False, dis_threshold=0.2)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):
        sd = {k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed, trained_scale=trained_scale).items()}
        getattr(net, name).load_state_dict(sd, strict=True)
    return net, VolumeRenderer


def run_case(net, VolumeRenderer, B, R, Kc, Kf, perturb, tag, with_grad=True, compact=False):
    """compact=True (the cfg1-sized cases, R = 1024): only what the tests assert is kept, in narrow dtypes
    (indices int16, validity uint8, per-point MLP outputs dropped), so that a fixture stays at a few MB."""
    posed_np, tmpl_np = synthetic.make_body_params(B, seed=1)
    posed = to_t(posed_np)
    tmpl = to_t(tmpl_np)
    if with_grad:
        for v in posed.values():
            v.requires_grad_(True)
    renderer = VolumeRenderer(n_coarse=Kc, n_fine=Kf, white_bkgd=True)
    cap = {}

    net.set_body_model(posed, tmpl)
    body_out = dict(vertices=net.verts.detach().clone(), joints_transform=net.joints_transform.detach().clone(),
                    vertices_transform=net.verts_transform.detach().clone(),
                    shape_offsets=net.shape_offsets.detach().clone(), pose_offsets=net.pose_offsets.detach().clone(),
                    vertices_template=net.verts_template.detach().clone())
    rays_world = torch.from_numpy(synthetic.rays_at_bbox(net.verts.detach().numpy(), R, seed=2))
    rays = net.convert_to_body_model_space(rays_world)
    net.clac_ober2cano_transform()
    if with_grad:
        net.ober2cano_transform.retain_grad()
        rays.retain_grad()

    # ---- capture wrappers (no edits to the reference) ----
    knn_calls, unpose_calls, query_calls, ss_calls, rand_calls, randn_calls, comp_calls = [], [], [], [], [], [], []
    orig_knn = net.knn
    net.knn = lambda ref, q: (lambda r: (knn_calls.append((r[0].clone(), r[1].clone())), r)[1])(orig_knn(ref, q))
    orig_unpose = net.unpose
    net.unpose = lambda xyz, vd=None: (lambda r: (unpose_calls.append((r[0].detach().clone(), r[2].detach().clone())), r)[1])(orig_unpose(xyz, vd))
    orig_query = net.query_canonical_space

    def q_wrap(xyz, viewdir=None, use_fine=False, **kw):
        r = orig_query(xyz, viewdir, use_fine, **kw)
        if isinstance(r, tuple):
            query_calls.append((r[0].detach().clone(), r[1].detach().clone()))
        return r
    net.query_canonical_space = q_wrap
    orig_ss, orig_rand, orig_randn_like = torch.searchsorted, torch.rand, torch.randn_like

    def ss_wrap(cdf, u, **kw):
        r = orig_ss(cdf, u, **kw)
        ss_calls.append((cdf.detach().clone(), u.detach().clone(), r.clone()))
        return r

    def rand_wrap(*a, **kw):
        r = orig_rand(*a, **kw)
        rand_calls.append(r.clone())
        return r

    def randn_like_wrap(x, **kw):
        r = orig_randn_like(x, **kw)
        randn_calls.append(r.clone())
        return r
    orig_comp = renderer.composite

    def comp_wrap(model, rays_, z, **kw):
        r = orig_comp(model, rays_, z, **kw)
        comp_calls.append((z.detach().clone(), r[0].detach().clone()))
        return r
    renderer.composite = comp_wrap
    torch.searchsorted, torch.rand, torch.randn_like = ss_wrap, rand_wrap, randn_like_wrap
    try:
        torch.manual_seed(1234)
        out = renderer(net, rays, perturb=perturb)
    finally:
        torch.searchsorted, torch.rand, torch.randn_like = orig_ss, orig_rand, orig_randn_like
        net.knn, net.unpose, net.query_canonical_space = orig_knn, orig_unpose, orig_query

    fx = dict(B=B, R=R, Kc=Kc, Kf=Kf, perturb=np.float32(perturb))
    for k, v in posed_np.items():
        fx["posed_" + k] = v
    for k, v in tmpl_np.items():
        fx["tmpl_" + k] = v
    fx["rays_world"] = rays_world.numpy()
    fx["rays_body"] = rays.detach().numpy()
    fx["verts_body"] = net.verts.detach().numpy()
    fx["ober2cano"] = net.ober2cano_transform.detach().numpy()[:, :, :3, :].copy()
    fx["ober2cano_row3_maxdev"] = np.float32(
        (net.ober2cano_transform.detach()[:, :, 3, :] - torch.tensor([0, 0, 0, 1.0])).abs().max())
    sub = slice(0, None, 53)
    if not compact:
        for k in ("vertices", "vertices_transform", "shape_offsets", "pose_offsets", "vertices_template"):
            fx["body_" + k] = body_out[k].numpy()[:, sub]
        fx["body_joints_transform"] = body_out["joints_transform"].numpy()
    for k, v in out.items():
        fx["out_" + k] = v.detach().numpy()
    # coarse pass intermediates
    fx["z_coarse"] = comp_calls[0][0].numpy()
    fx["weights_coarse"] = comp_calls[0][1].numpy()
    fx["z_combine"] = comp_calls[1][0].numpy()
    fx["weights_fine"] = comp_calls[1][1].numpy()
    fx["knn_idx_coarse"] = knn_calls[0][1].numpy().astype(np.int16)
    fx["knn_idx_fine"] = knn_calls[1][1].numpy().astype(np.int16)
    fx["valid_coarse"] = unpose_calls[0][1].numpy().astype(np.uint8)
    fx["valid_fine"] = unpose_calls[1][1].numpy().astype(np.uint8)
    if not compact:
        fx["knn_dist_coarse"] = knn_calls[0][0].numpy()
        fx["xyz_cano_coarse"] = unpose_calls[0][0].numpy()
        fx["rgb_pts_coarse"] = query_calls[0][0].numpy().astype(np.float16)   # raw MLP output, pre-mask
        fx["sigma_pts_coarse"] = query_calls[0][1].numpy()
    else:       # range of the raw density over the valid coarse points (documents the fixture's dynamic range)
        sg = query_calls[0][1].numpy()[fx["valid_coarse"] > 0]
        fx["sigma_valid_percentiles"] = np.percentile(sg, [0, 1, 25, 50, 75, 99, 100]).astype(np.float32)
    fx["cdf"], fx["u"], fx["inds"] = ss_calls[0][0].numpy(), ss_calls[0][1].numpy(), ss_calls[0][2].numpy().astype(np.int16)
    if perturb > 0:
        fx["noise_coarse_u"] = (rand_calls[0] ).numpy()      # perturb * rand -> the oracle multiplies again
        fx["noise_fine_u"] = rand_calls[1].numpy()
        fx["noise_sigma_c"] = randn_calls[0].numpy()
        fx["noise_sigma_f"] = randn_calls[1].numpy()

    if with_grad:
        rs = np.random.RandomState(7)
        loss = 0
        for k in sorted(out.keys()):
            coef = torch.from_numpy(rs.normal(size=out[k].shape).astype(np.float32))
            fx["coef_" + k] = coef.numpy()
            loss = loss + (out[k] * coef).sum()
        params = {}
        for net_name in ("nerf", "nerf_fine"):
            for n, p in getattr(net, net_name).named_parameters():
                params[net_name + "." + n] = p
                p.grad = None
        loss.backward()
        fx["loss"] = np.float32(loss.item())
        for n, p in params.items():
            g = p.grad.numpy()
            fx["gnorm_" + n] = np.float32(np.linalg.norm(g))
            if g.size <= 1024:
                fx["grad_" + n] = g
            else:
                fx["grad_" + n + "_blk"] = g[:32, :32].copy()
        fx["grad_ober2cano_sumabs"] = np.float32(net.ober2cano_transform.grad.abs().sum())
        go = net.ober2cano_transform.grad.numpy()[:, :, :3, :]
        nz = np.abs(go).reshape(B, -1, 12).sum(-1)
        top = np.argsort(-nz, axis=1)[:, :64]
        fx["grad_ober2cano_top_idx"] = top.astype(np.int16)
        fx["grad_ober2cano_top"] = np.take_along_axis(go.reshape(B, -1, 12), top[..., None], 1)
        fx["grad_rays_body"] = rays.grad.numpy()
        for k, v in posed.items():
            fx["grad_posed_" + k] = v.grad.numpy()
    path = os.path.join(OUT, "render_%s.npz" % tag)
    np.savez_compressed(path, **fx)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024),
          "valid_coarse=%.3f" % fx["valid_coarse"].mean(), "acc_fine=%.3f" % fx["out_alphas_fine"].mean())


def gen_rays_case():
    from datasets.anim_nerf_dataset import gen_rays
    rs = np.random.RandomState(3)
    R = np.linalg.qr(rs.normal(size=(3, 3)))[0]
    c2w = np.concatenate([R, rs.normal(size=(3, 1))], 1).astype(np.float32)
    H, W = 6, 8
    focal = np.array([9.5, 8.25], np.float32)
    c = np.array([3.7, 2.9], np.float32)
    rays = gen_rays(torch.from_numpy(c2w), H, W, focal, 0.1, 10.0, c)
    rays_c = gen_rays(torch.from_numpy(c2w), H, W, focal, 0.1, 10.0, None)
    np.savez_compressed(os.path.join(OUT, "gen_rays.npz"), c2w=c2w, H=H, W=W, focal=focal, c=c,
                        rays=rays.numpy(), rays_default_c=rays_c.numpy())
    print("wrote gen_rays.npz")


def pixel_sampling_case():
    """Training-ray sampling: the reference's own `get_pixelcoords` ('foreground_pixel', seeded numpy global RNG)
    and `gen_rays` (datasets/anim_nerf_dataset.py:10-85) on a synthetic soft-edged silhouette, combined by the
    lines of `__getitem__` (:242-261; the dataset class itself needs People-Snapshot files, so those ten lines
    are restated here around the reference's functions)."""
    from datasets.anim_nerf_dataset import get_pixelcoords, gen_rays
    import cv2
    rs = np.random.RandomState(41)
    H, W, n_side = 120, 96, 8
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    body = (((xx - 50) / 14.0) ** 2 + ((yy - 58) / 36.0) ** 2 < 1) | (((xx - 30) / 5.0) ** 2 + ((yy - 40) / 22.0) ** 2 < 1)
    mask_u8 = cv2.GaussianBlur((body * 255).astype(np.uint8), (5, 5), 1.2)        # soft edge: values strictly between 0 and 255
    img_u8 = rs.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    R = np.linalg.qr(rs.normal(size=(3, 3)))[0]
    c2w = np.concatenate([R, rs.normal(size=(3, 1))], 1).astype(np.float32)
    focal = np.array([105.5, 99.25], np.float32)
    c = np.array([47.3, 61.1], np.float32)
    fx = dict(img_u8=img_u8, mask_u8=mask_u8, c2w=c2w, focal=focal, c=c, n_side=n_side)
    # :200-204, :244-245
    img = torch.from_numpy(img_u8 / 255.).float().permute(2, 0, 1)
    mask = torch.from_numpy(mask_u8 / 255.).float().unsqueeze(0)
    img = img * mask                       # with_background = False
    img = img * mask + (1 - mask)          # white_bkgd = True
    rgbs, alphas = img.permute(1, 2, 0), mask.permute(1, 2, 0)
    rays = gen_rays(torch.from_numpy(c2w), H, W, focal, 0.1, 10.0, c)
    for fore_erode in (3, 5):
        np.random.seed(5)
        coords = get_pixelcoords(H, W, alphas.numpy(), subsampletype="foreground_pixel", subsamplesize=n_side,
                                 fore_rate=0.9, fore_erode=fore_erode)
        fx["coords_e%d" % fore_erode] = coords.astype(np.int32)
    coords = fx["coords_e3"]
    fx["rays"] = rays[coords[:, 0], coords[:, 1]].numpy()
    fx["rgbs"] = rgbs[coords[:, 0], coords[:, 1]].numpy()
    fx["alphas"] = alphas[coords[:, 0], coords[:, 1]].numpy()
    path = os.path.join(OUT, "pixel_sampling.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "soft pixels:", int(((mask_u8 > 0) & (mask_u8 < 255)).sum()))


def state_dict_case(tmp):
    """Names and shapes of the parameters/buffers a reference checkpoint carries for the hot path's modules:
    `AnimNeRF` (under `anim_nerf.`) and `BodyModelParams` (under `body_model_params.`), as train.py:110-146 nests them."""
    import json
    from models.body_model_params import BodyModelParams
    net, _ = build_reference(tmp)
    keys = {"anim_nerf." + k: list(v.shape) for k, v in net.state_dict().items()}
    keys.update({"body_model_params." + k: list(v.shape) for k, v in BodyModelParams(7, model_type="smpl").state_dict().items()})
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("wrote state_dict_keys.json", len(keys), "entries")


def api_signatures_case():
    """Constructor / method parameter names of the reference classes on the path's boundary (B2, B3, and the
    modules B4 is built from), taken with `inspect.signature` from the imported reference."""
    import inspect
    import json
    install_shim()
    from models.anim_nerf import AnimNeRF
    from models.volume_rendering import VolumeRenderer
    from models.nerf import NeRF
    from models.body_model_params import BodyModelParams
    want = {
        "AnimNeRF": (AnimNeRF, ["__init__", "set_latent_code", "set_body_model", "convert_to_body_model_space",
                                "clac_ober2cano_transform", "unpose", "query_canonical_space", "forward"]),
        "VolumeRenderer": (VolumeRenderer, ["__init__", "sample_coarse", "sample_fine", "composite", "forward"]),
        "NeRF": (NeRF, ["__init__", "get_sigma", "get_normal", "forward"]),
        "BodyModelParams": (BodyModelParams, ["__init__", "init_parameters", "set_requires_grad", "forward"]),
    }
    out = {}
    for cname, (cls, methods) in want.items():
        out[cname] = {}
        for m in methods:
            sig = inspect.signature(getattr(cls, m))
            out[cname][m] = [[n, (None if p.default is inspect._empty else repr(p.default)), p.kind.name]
                             for n, p in sig.parameters.items() if n != "self"]
    # train.py cannot be imported here (pytorch-lightning, yacs, ... are absent): its class is read with `ast`
    import ast
    tree = ast.parse(open(os.path.join(REF, "train.py")).read())
    out["AnimNeRFSystem"] = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "AnimNeRFSystem":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in ("__init__", "forward", "compute_loss", "decode_batch",
                                                                   "training_step", "configure_optimizers"):
                    args = [a.arg for a in fn.args.args if a.arg != "self"]
                    defaults = [None] * (len(args) - len(fn.args.defaults)) + [ast.unparse(d) for d in fn.args.defaults]
                    out["AnimNeRFSystem"][fn.name] = [[a, d, "POSITIONAL_OR_KEYWORD"] for a, d in zip(args, defaults)]
    with open(os.path.join(OUT, "api_signatures.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote api_signatures.json", {k: len(v) for k, v in out.items()})


def regularizers_case():
    """The MLP queries of the training regularisers through the reference's own `NeRF` (models/nerf.py:155-190):
    `get_sigma(only_sigma=True)` on foreground/background points and `get_normal` (autograd.grad with
    create_graph=True) on points / neighbours, combined exactly as train.py:264-297 does, then backward
    (torch double backward) -> per-parameter gradients."""
    from models.nerf import NeRF
    import torch.nn.functional as F
    net = NeRF(freqs_xyz=10, freqs_dir=0, use_view=False)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(10).items()}, strict=True)
    rs = np.random.RandomState(31)
    B, n, n_samples, dis_threshold, epsilon = 2, 192, 64, 0.2, 0.02
    base = rs.uniform(-0.8, 0.8, size=(B, n, 3)).astype(np.float32)
    points = base + rs.normal(size=base.shape).astype(np.float32) * dis_threshold * 0.5
    neighbs = points + rs.normal(size=base.shape).astype(np.float32) * epsilon
    fg = rs.normal(0, 0.1, size=(B, 128, 3)).astype(np.float32)
    bg = rs.normal(0, 1.0, size=(B, 128, 3)).astype(np.float32)
    t = torch.from_numpy
    sig_fg = net.get_sigma(t(fg), only_sigma=True)
    sig_bg = net.get_sigma(t(bg), only_sigma=True)
    loss_fg = torch.mean(torch.exp(-2.0 / n_samples * torch.relu(sig_fg)))
    loss_bg = torch.mean(1 - torch.exp(-2.0 / n_samples * torch.relu(sig_bg)))
    n_p = net.get_normal(t(points.copy()))
    n_q = net.get_normal(t(neighbs.copy()))
    u_p = n_p / (torch.norm(n_p, p=2, dim=-1, keepdim=True) + 1e-5)
    u_q = n_q / (torch.norm(n_q, p=2, dim=-1, keepdim=True) + 1e-5)
    loss_n = F.mse_loss(u_p, u_q)
    loss = 0.01 * loss_fg + 0.01 * loss_bg + 0.01 * loss_n
    loss.backward()
    fx = dict(points=points, neighbs=neighbs, fg=fg, bg=bg, sigma_fg=sig_fg.detach().numpy(), sigma_bg=sig_bg.detach().numpy(),
              normal_points=n_p.detach().numpy(), normal_neighbs=n_q.detach().numpy(),
              loss_fg=np.float32(loss_fg.item()), loss_bg=np.float32(loss_bg.item()), loss_normals=np.float32(loss_n.item()))
    for name, prm in net.named_parameters():
        g = prm.grad.numpy() if prm.grad is not None else np.zeros(tuple(prm.shape), np.float32)
        fx["gnorm_" + name] = np.float32(np.linalg.norm(g))
        fx["grad_" + name + ("" if g.size <= 1024 else "_blk")] = g if g.size <= 1024 else g[:32, :32].copy()
    path = os.path.join(OUT, "regularizers.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "loss_normals=%.5f" % fx["loss_normals"])


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "--regularizers-only" in sys.argv:
        regularizers_case()
        sys.exit(0)
    if "--api-only" in sys.argv:
        api_signatures_case()
        sys.exit(0)
    if "--state-dict-only" in sys.argv:
        with tempfile.TemporaryDirectory() as tmp:
            state_dict_case(tmp)
        sys.exit(0)
    if "--pixel-sampling-only" in sys.argv:
        pixel_sampling_case()
        sys.exit(0)
    if "--cfg1-only" not in sys.argv:
        regularizers_case()
        pixel_sampling_case()
        api_signatures_case()
        with tempfile.TemporaryDirectory() as tmp:
            state_dict_case(tmp)
    with tempfile.TemporaryDirectory() as tmp:
        net, VR = build_reference(tmp)
        if "--cfg1-only" not in sys.argv:
            run_case(net, VR, B=2, R=96, Kc=64, Kf=64, perturb=0.0, tag="det")
            run_case(net, VR, B=1, R=64, Kc=64, Kf=32, perturb=1.0, tag="perturb")
        # BASELINE configs[0] at full size: 1024 rays, 64 + 64 samples, deterministic and perturbed
        run_case(net, VR, B=1, R=1024, Kc=64, Kf=64, perturb=0.0, tag="cfg1_det", with_grad=False, compact=True)
        run_case(net, VR, B=1, R=1024, Kc=64, Kf=64, perturb=1.0, tag="cfg1_perturb", with_grad=False, compact=True)
    with tempfile.TemporaryDirectory() as tmp:
        # trained-scale weights (synthetic.make_nerf_weights(trained_scale=True)): sigma spans ~0..100, saturated colours
        net, VR = build_reference(tmp, trained_scale=True)
        run_case(net, VR, B=1, R=256, Kc=64, Kf=64, perturb=0.0, tag="trained_det", with_grad=False, compact=True)
    if "--cfg1-only" in sys.argv:
        sys.exit(0)
    try:
        gen_rays_case()
    except Exception as e:  # cv2/torchvision import problems should not lose the main fixtures
        print("gen_rays fixture skipped:", repr(e))
