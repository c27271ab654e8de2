"""GPU parity of the training regularisers' MLP queries: the density
query (`NeRF.get_sigma`, train.py:264-284) and the second-order normal query (`NeRF.get_normal`,
nerf.py:177-190 -> train.py:286-309) on the kernels, against torch autograd / double backward of
the oracle MLP on the same points and weights."""
import numpy as np
import pytest
import torch

from util import oracle, nerf_params, synthetic, load_golden, regulariser_losses

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _st_bf16(t):
    """round to bf16, straight-through gradient (the kernels treat operand rounding as identity)."""
    return t + (t.bfloat16().float() - t).detach()


def _sigma_oracle(p, x, emulate_bf16):
    """sigma of the trunk (models/nerf.py:155-175); emulate_bf16 rounds the MMA operands where the kernel does."""
    rd = _st_bf16 if emulate_bf16 else (lambda t: t)
    e = rd(oracle.embed(x))
    h = e
    for i in range(8):
        if i == 4:
            h = torch.cat([e, h], -1)
        w, b = p["xyz_encoding_%d.0" % (i + 1)]
        h = rd(torch.relu(h @ rd(w).T + rd(b)))       # trunk biases enter through a bf16 tensor-core step (bias slab)
    return (h @ rd(p["sigma"][0]).T + p["sigma"][1])[:, 0]


def _net(seed):
    from anim_nerf_b200.nerf import NeRF
    net = NeRF(freqs_dir=0, use_view=False).to(DEV)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}, strict=True)
    return net


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def _cos(a, b):
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-20))


def _named_grads(net):
    out = {}
    for name, lin in zip(synthetic.NERF_LAYER_NAMES, net.linears()):
        out[name + ".weight"] = None if lin.weight.grad is None else lin.weight.grad.detach().cpu()
        out[name + ".bias"] = None if lin.bias.grad is None else lin.bias.grad.detach().cpu()
    return out


TRUNK = ["xyz_encoding_%d.0" % (i + 1) for i in range(8)] + ["sigma"]


@pytest.mark.parametrize("n", [300, 2500])
@pytest.mark.parametrize("emu", [True, False])
def test_sigma_gradient_and_linear_functionals(n, emu):
    """SigmaWithGradient on its own: s = d sigma/d xyz, and the parameter gradients of
    L = sum(s . r) + sum(sigma * q) for random r, q -- the first term exercises only the tangent + wgrad
    route (second order), the second only the ordinary backward.
    emu=True (oracle at the kernel's operand precision): rel. L2 per tensor <= 6e-2;
    emu=False (plain fp32 oracle): <= 0.3 (bf16 ReLU-branch flips at random init, see test_mlp_backward)."""
    from anim_nerf_b200.autograd import sigma_with_gradient
    seed = 10
    net = _net(seed)
    p = nerf_params(seed, requires_grad=True)
    rs = np.random.RandomState(21)
    x = torch.from_numpy(rs.uniform(-1, 1, size=(n, 3)).astype(np.float32))
    r = torch.from_numpy(rs.normal(size=(n, 3)).astype(np.float32))
    q = torch.from_numpy(rs.normal(size=(n,)).astype(np.float32))
    # oracle: double backward
    xo = x.clone().requires_grad_(True)
    sig_o = _sigma_oracle(p, xo, emu)
    s_o = torch.autograd.grad(sig_o.sum(), xo, create_graph=True)[0]
    ((s_o * r).sum() + (sig_o * q).sum()).backward()
    # kernels (the allocator's free blocks hold NaN patterns: no result may depend on unwritten scratch)
    junk = torch.full((96 << 20,), float("nan"), device=DEV)
    del junk
    sig, s = sigma_with_gradient(net, x.to(DEV))
    tol = 6e-2 if emu else 0.3
    e_sig, e_s = _rel(sig[:, 0].detach().cpu(), sig_o.detach()), _rel(s.detach().cpu(), s_o.detach())
    print("sigma rel err %.4g, d sigma/d xyz rel err %.4g" % (e_sig, e_s))
    assert e_sig < (2e-2 if emu else 5e-2) and e_s < tol
    ((s * r.to(DEV)).sum() + (sig[:, 0] * q.to(DEV)).sum()).backward()
    g = _named_grads(net)
    errs = {}
    for name in TRUNK:
        W, b = p[name]
        errs[name + ".weight"] = _rel(g[name + ".weight"], W.grad)
        errs[name + ".bias"] = _rel(g[name + ".bias"], b.grad)
    print({k: round(v, 4) for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, bad
    # the colour branch takes no part in sigma or its gradient
    for name in ("xyz_encoding_final", "dir_encoding.0", "rgb.0"):
        assert float(g[name + ".weight"].abs().max()) == 0.0 and float(g[name + ".bias"].abs().max()) == 0.0, name


def test_second_order_term_alone_has_no_bias_gradient():
    """L = sum(s . r) depends on no bias (biases do not enter d sigma/d xyz): bias gradients are exactly zero
    and the weight gradients match the oracle's double backward."""
    from anim_nerf_b200.autograd import sigma_with_gradient
    net = _net(11)
    p = nerf_params(11, requires_grad=True)
    rs = np.random.RandomState(22)
    n = 700
    x = torch.from_numpy(rs.uniform(-1, 1, size=(n, 3)).astype(np.float32))
    r = torch.from_numpy(rs.normal(size=(n, 3)).astype(np.float32))
    xo = x.clone().requires_grad_(True)
    s_o = torch.autograd.grad(_sigma_oracle(p, xo, True).sum(), xo, create_graph=True)[0]
    (s_o * r).sum().backward()
    _, s = sigma_with_gradient(net, x.to(DEV))
    (s * r.to(DEV)).sum().backward()
    g = _named_grads(net)
    for name in TRUNK:
        assert float(g[name + ".bias"].abs().max()) == 0.0, name
        assert p[name][1].grad is None or float(p[name][1].grad.abs().max()) == 0.0
        e = _rel(g[name + ".weight"], p[name][0].grad)
        assert e < 6e-2, (name, e)


def test_normal_and_density_regularisers_match_torch_double_backward():
    """The regulariser terms of `compute_loss` (train.py:264-309) through the kernels vs the reference's torch
    formulation (`oracle.nerf_sigma` / `oracle.nerf_normal`: fp32, autograd.grad(create_graph=True)) on the same
    points: loss values within 2 %, every trunk gradient's direction (cosine) >= 0.9 and norm within 10 %
    (measured on B200: cosine 0.943-0.9994, norm ratio 0.96-1.04; the deficit is bf16 ReLU-branch flips, which the
    unit-normalisation of near-zero normals amplifies -- against the oracle at the kernel's operand precision the
    per-tensor error is 1-4 %, test_sigma_gradient_and_linear_functionals)."""
    B, n_pts = 2, 1500
    rs = np.random.RandomState(23)
    vt = torch.from_numpy(rs.uniform(-0.8, 0.8, size=(B, n_pts, 3)).astype(np.float32)).to(DEV)
    pts = vt + torch.from_numpy(rs.normal(size=(B, n_pts, 3)).astype(np.float32)).to(DEV) * 0.1
    nb = pts + torch.from_numpy(rs.normal(size=(B, n_pts, 3)).astype(np.float32)).to(DEV) * 0.02
    fg = torch.from_numpy(rs.uniform(-0.5, 0.5, size=(B, 128, 3)).astype(np.float32)).to(DEV)
    k = -2.0 / 64

    def unit(v):
        return v / (torch.norm(v, p=2, dim=-1, keepdim=True) + 1e-5)

    def loss_with(net, get_sigma, get_normal):
        l_fg = torch.mean(torch.exp(k * torch.relu(get_sigma(fg))))
        l_bg = torch.mean(1 - torch.exp(k * torch.relu(get_sigma(fg + 0.3))))
        l_n = torch.nn.functional.mse_loss(unit(get_normal(pts)), unit(get_normal(nb)))
        return l_fg, l_bg, l_n

    pr = nerf_params(10, device=DEV, requires_grad=True)
    lr = loss_with(None, lambda x: oracle.nerf_sigma(pr, x), lambda x: oracle.nerf_normal(pr, x))
    (lr[0] + lr[1] + lr[2]).backward()
    g_ref = {}
    for name in TRUNK:
        g_ref[name + ".weight"], g_ref[name + ".bias"] = pr[name][0].grad.cpu(), pr[name][1].grad.cpu()
    net = _net(10)
    lk = loss_with(net, lambda x: net.get_sigma(x, only_sigma=True), lambda x: net.get_normal(x))
    (lk[0] + lk[1] + lk[2]).backward()
    g = _named_grads(net)
    print("losses ref", [float(v) for v in lr], "kernels", [float(v) for v in lk])
    for a, b in zip(lk, lr):
        assert abs(float(a) - float(b)) <= 2e-2 * abs(float(b)) + 1e-6, (float(a), float(b))
    stats = {}
    for name in TRUNK:
        for kind in (".weight", ".bias"):
            a, b = g[name + kind], g_ref[name + kind]
            stats[name + kind] = (round(_cos(a, b), 4), round(float(a.norm() / (b.norm() + 1e-20)), 4))
    print(stats)
    bad = {k2: v for k2, v in stats.items() if not (v[0] >= 0.9 and 0.9 <= v[1] <= 1.1)}
    assert not bad, bad


def test_regularisers_vs_reference_fixture():
    """Kernels vs the values captured from the reference's own NeRF.get_sigma / get_normal + torch double
    backward (tests/golden/regularizers.npz): densities within 2e-2 absolute (bf16 operands, sigma ~ 5), loss terms
    within 2 %, every trunk gradient norm within 10 % and the captured gradient blocks' direction >= 0.95."""
    fx = load_golden("regularizers")
    net = _net(10)
    l_fg, l_bg, l_n, _ = regulariser_losses(fx, lambda x: net.get_sigma(x, only_sigma=True), lambda x: net.get_normal(x), dev=DEV)
    sig = net.get_sigma(torch.from_numpy(fx["fg"]).to(DEV), only_sigma=True).detach().cpu().numpy()
    assert float(np.abs(sig - fx["sigma_fg"]).max()) < 2e-2
    for got, key in ((l_fg, "loss_fg"), (l_bg, "loss_bg"), (l_n, "loss_normals")):
        assert abs(float(got.detach()) - float(fx[key])) <= 2e-2 * abs(float(fx[key])) + 1e-6, (key, float(got.detach()), float(fx[key]))
    (0.01 * (l_fg + l_bg + l_n)).backward()
    g = _named_grads(net)
    stats = {}
    for name in TRUNK:
        for kind in (".weight", ".bias"):
            a = g[name + kind].numpy()
            ref_n = float(fx["gnorm_" + name + kind])
            blk = fx["grad_" + name + kind] if "grad_" + name + kind in fx else fx["grad_" + name + kind + "_blk"]
            mine = a if blk.shape == a.shape else a[:32, :32]
            cos = float((mine * blk).sum() / (np.linalg.norm(mine) * np.linalg.norm(blk) + 1e-30))
            stats[name + kind] = (round(float(np.linalg.norm(a)) / (ref_n + 1e-30), 4), round(cos, 4))
    print(stats)
    bad = {k2: v for k2, v in stats.items() if not (0.9 <= v[0] <= 1.1 and v[1] >= 0.95)}
    assert not bad, bad


def test_compute_loss_with_regularisers_runs_on_kernels():
    """`AnimNeRFSystem.training_step` with fg/bg points: every loss term is produced, finite, and the step's
    MLP gradients are finite and non-zero for both networks (the regularisers add to the render gradients)."""
    from anim_nerf_b200.system import AnimNeRFSystem
    from anim_nerf_b200.body_model import BodyModel
    data = synthetic.make_smpl_dict(0)
    B = 2
    sysm = AnimNeRFSystem(body_model_data=data, n_samples=64, n_importance=64, num_frames=B).to(DEV)    # optim_body_params=True (default)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):
        getattr(sysm.anim_nerf, name).load_state_dict(
            {k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}, strict=True)
    posed_np, tmpl_np = synthetic.make_body_params(B, seed=1)
    sysm.init_body_model_params({k: torch.from_numpy(v) for k, v in posed_np.items()})
    with torch.no_grad():
        verts = BodyModel(data)(**{k: torch.from_numpy(v) for k, v in posed_np.items()})["vertices"].numpy()
    batch_np = synthetic.make_training_batch(verts, n_side=8, seed=3)
    rs = np.random.RandomState(5)
    batch = dict(rays=torch.from_numpy(batch_np["rays"]).to(DEV), rgbs=torch.from_numpy(batch_np["rgbs"]).to(DEV),
                 alphas=torch.from_numpy(batch_np["alphas"]).to(DEV), frame_idx=torch.arange(B, device=DEV),
                 body_model_params={k: torch.from_numpy(v).to(DEV) for k, v in posed_np.items()},
                 body_model_params_template={k: torch.from_numpy(v).to(DEV) for k, v in tmpl_np.items()},
                 fg_points=torch.from_numpy(rs.normal(0, 0.1, size=(B, 128, 3)).astype(np.float32)).to(DEV),
                 bg_points=torch.from_numpy(rs.normal(0, 1.0, size=(B, 128, 3)).astype(np.float32)).to(DEV))
    loss = sysm.training_step(batch, 0)
    want = {"loss_rgb", "loss_rgb_fine", "loss_alphas", "loss_alphas_fine", "loss_foreground", "loss_foreground_fine",
            "loss_background", "loss_background_fine", "loss_normals", "loss_normals_fine"}
    assert set(sysm.last_details) == want
    for k2, v in sysm.last_details.items():
        assert torch.isfinite(v).all(), k2
    loss.backward()
    for name in ("nerf", "nerf_fine"):
        for pname, prm in getattr(sysm.anim_nerf, name).named_parameters():
            assert prm.grad is not None and torch.isfinite(prm.grad).all(), (name, pname)
        assert float(getattr(sysm.anim_nerf, name).sigma.weight.grad.abs().max()) > 0
    # the shipped configuration optimises the SMPL table: its rows received gradients through the kernels
    for pname, prm in sysm.body_model_params.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad).all() and float(prm.grad.abs().max()) > 0, pname
