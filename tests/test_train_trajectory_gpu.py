"""GPU: training-trajectory parity.  200 Adam steps of the render loss on the kernels (bf16 tensor-core MLP forward /
backward, FusedAdam, gradients accumulated in the flat buffer) against 200 steps of the fp32 oracle (torch autograd +
torch.optim.Adam on the host) from the same initial weights on the same mini-batches -- the convergence evidence the
single-step gradient tolerances cannot give ("no training-trajectory parity")."""
import numpy as np
import pytest
import torch

from util import oracle, nerf_params, body_model, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
N_STEPS, N_RAYS, BATCH, KC, KF = 200, 64, 64, 64, 64


def _psnr(mse):
    return -10.0 * np.log10(max(mse, 1e-12))


def _teacher_params(seed):
    """A field that differs visibly from the student's initialisation: thinner (sigma bias 1.5 instead of 5) and tinted
    (rgb bias +1.2 / 0 / -1.2), so that the fit has something to learn (two random-init networks render the same grey)."""
    w = synthetic.make_nerf_weights(seed, sigma_bias=1.5)
    w["rgb.0.bias"] = w["rgb.0.bias"] + np.array([1.2, 0.0, -1.2], np.float32)
    return {n: (torch.from_numpy(w[n + ".weight"]), torch.from_numpy(w[n + ".bias"])) for n in synthetic.NERF_LAYER_NAMES}


def test_200_step_trajectory_matches_fp32_oracle():
    """Student (seeds 10/11) is fitted to the renders of a teacher (a thinner, tinted field; fp32 oracle) of one posed
    frame: loss = mse(rgb) + mse(rgb_fine) + 0.1 (l1(alpha) + l1(alpha_fine)) (train.py:228-262 without the
    regularisers), Adam lr 5e-4 eps 1e-8 with the reference's poly decay (1 - t/T)^0.9 (utils/__init__.py:52) run per
    step over the T = 200 steps, 64 rays x (64 + 64) samples per step, perturb = 0.
    What is comparable and what is not: the first ~20 steps agree step by step (4-5 digits); then Adam at lr ~4e-4
    bounces around the minimum (the fp32 oracle's own loss swings by several x from step to step) and any two runs
    decorrelate -- shown by the oracle's TWIN, the same fp32 run from initial weights perturbed by 1e-6 (relative),
    which drifts from the oracle as far as the kernels do; as the learning rate decays every run settles.
    Asserted (values measured on B200 are printed):
      steps 0..19     per-step loss within 1 % of the oracle's (measured 0.05-0.4 %)
      steps 180..199  mean loss not above 3 x the worst of the fp32 runs (the oracle and three twins); PSNR of the fine
                      render against the teacher above 48 dB and not more than 6 dB below the worst fp32 run
      and every run's loss fell by more than 100 x.
    The kernels' run is reproducible bit for bit (ordered compaction of the valid points, fixed-order weight-gradient
    reduction: tests/test_kernels_gpu.py::test_training_steps_are_bit_reproducible_with_frozen_body_params); measured:
    end loss 3.2e-4 (oracle 3.6e-4, twins 3.8e-4 / 2.1e-4 / 2.5e-4), PSNR 56.8 dB (oracle 58.3, twins 59.8 / 59.0 / ...).
    Before the compaction was ordered the valid points got their slots by atomic allocation, the slot order permuted the
    fp32 summation order of the weight gradient, and Adam's normalisation turned those last-bit differences into +-lr
    steps on the weights whose gradient is at noise level: over 12 such runs the end loss spanned 2.1e-4 .. 6.5e-4 and the
    PSNR 53.1 .. 61.4 dB -- the spread the bands below were sized for, and the same sensitivity the fp32 twins show.  At
    this level (rgb rms error ~1e-3) the fit is at the bf16 MLP's own error floor (mean |d rgb| 7e-4 on the trained-scale
    fixture), three orders of magnitude below what a fit to real images reaches (~30 dB)."""
    from anim_nerf_b200.anim_nerf import AnimNeRF
    from anim_nerf_b200.volume_rendering import VolumeRenderer
    from anim_nerf_b200.optim import FlatGradBuffer, FusedAdam
    torch.manual_seed(0)
    bm = body_model()
    posed_np, tmpl_np = synthetic.make_body_params(1, seed=5)
    posed = {k: torch.from_numpy(v) for k, v in posed_np.items()}
    tmpl = {k: torch.from_numpy(v) for k, v in tmpl_np.items()}
    with torch.no_grad():
        po, pt = bm(**posed), bm(**tmpl)
        verts, o2c = oracle.ober2cano_tables(po, pt)
        rays_w = torch.from_numpy(synthetic.rays_at_bbox(po["vertices"].numpy(), N_RAYS, seed=4, margin=0.02))
        rays = oracle.rays_to_body_space(rays_w, po["joints_transform"][:, 0])
        tables = (verts, o2c, bm.lbs_weights)
        teacher = oracle.render_rays(_teacher_params(20), _teacher_params(21), rays, tables, n_coarse=KC, n_fine=KF)
    tgt_rgb, tgt_a = teacher["rgbs_fine"], teacher["alphas_fine"]
    order = np.stack([np.sort(np.random.RandomState(100 + s).permutation(N_RAYS)[:BATCH]) for s in range(N_STEPS)])
    mse, l1 = torch.nn.functional.mse_loss, torch.nn.functional.l1_loss

    def loss_of(out, sel, dev):
        r, a = tgt_rgb[:, sel].to(dev), tgt_a[:, sel].to(dev)
        mf = mse(out["rgbs_fine"], r)
        return mse(out["rgbs"], r) + mf + 0.1 * (l1(out["alphas"], a) + l1(out["alphas_fine"], a)), float(mf.detach())

    # ---- fp32 oracle trajectories (host): the reference run, and a twin whose initial weights are perturbed by 1e-6
    # (relative) -- the spread between the two is the optimiser's own sensitivity, the yardstick for the kernels' run
    def oracle_run(eps, seed=7):
        pc, pf = nerf_params(10, requires_grad=True), nerf_params(11, requires_grad=True)
        if eps:
            g = torch.Generator().manual_seed(seed)
            with torch.no_grad():
                for p in (pc, pf):
                    for wb in p.values():
                        for t in wb:
                            t.mul_(1 + eps * torch.randn(t.shape, generator=g))
        opt_o = torch.optim.Adam([t for p in (pc, pf) for wb in p.values() for t in wb], lr=5e-4, eps=1e-8)
        sched_o = torch.optim.lr_scheduler.LambdaLR(opt_o, lambda s: (1 - s / N_STEPS) ** 0.9)      # utils/__init__.py:52, per step
        ls, mfs = [], []
        for s in range(N_STEPS):
            sel = torch.from_numpy(order[s])
            out = oracle.render_rays(pc, pf, rays[:, sel], tables, n_coarse=KC, n_fine=KF)
            loss, mf = loss_of(out, sel, "cpu")
            opt_o.zero_grad(set_to_none=True)
            loss.backward()
            opt_o.step()
            sched_o.step()
            ls.append(float(loss.detach())); mfs.append(mf)
        return np.asarray(ls), np.asarray(mfs)
    lo, mf_o = oracle_run(0.0)
    lo2, mf_o2 = oracle_run(1e-6)
    more_twins = [oracle_run(1e-6, seed) for seed in (8, 9)]

    # ---- the kernels' trajectory
    net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=True, freqs_dir=0, body_model_data=synthetic.make_smpl_dict(0)).to(DEV)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):
        getattr(net, name).load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}, strict=True)
    vr = VolumeRenderer(n_coarse=KC, n_fine=KF, white_bkgd=True)
    params = [p for n in ("nerf", "nerf_fine") for p in getattr(net, n).parameters()]
    opt = FusedAdam(params, lr=5e-4, eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: (1 - s / N_STEPS) ** 0.9)
    flat = FlatGradBuffer([net.nerf, net.nerf_fine])
    opt.flat = flat
    opt.on_step.append(lambda: (net.nerf.mark_dirty(), net.nerf_fine.mark_dirty()))
    posed_d, tmpl_d = {k: v.to(DEV) for k, v in posed.items()}, {k: v.to(DEV) for k, v in tmpl.items()}
    rays_d = rays.to(DEV)
    loss_k, mf_k = [], []
    for s in range(N_STEPS):
        sel = torch.from_numpy(order[s]).to(DEV)
        net.setup_frame(posed_d, tmpl_d, None)
        out = vr(net, rays_d[:, sel], perturb=0.0)
        loss, mf = loss_of(out, sel.cpu(), DEV)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        loss_k.append(float(loss.detach())); mf_k.append(mf)
    lk, mf_k = np.asarray(loss_k), np.asarray(mf_k)
    rel, rel2 = np.abs(lk - lo) / lo, np.abs(lo2 - lo) / lo
    end = {"oracle": lo[180:].mean(), "oracle_twin": lo2[180:].mean(), "kernels": lk[180:].mean()}
    psnr = {"oracle": _psnr(float(mf_o[180:].mean())), "oracle_twin": _psnr(float(mf_o2[180:].mean())), "kernels": _psnr(float(mf_k[180:].mean()))}
    for i, (l3, m3) in enumerate(more_twins):
        end["oracle_twin%d" % (i + 2)] = l3[180:].mean(); psnr["oracle_twin%d" % (i + 2)] = _psnr(float(m3[180:].mean()))
    fp32_end = [v for k, v in end.items() if k != "kernels"]; fp32_psnr = [v for k, v in psnr.items() if k != "kernels"]
    print("every 10th step (oracle, oracle twin, kernels):",
          [(i, round(float(lo[i]), 5), round(float(lo2[i]), 5), round(float(lk[i]), 5)) for i in range(0, N_STEPS, 10)])
    print("loss %.5f -> mean of steps 180..199: %s | PSNR vs teacher: %s | first 20 steps max rel diff to the oracle: kernels %.4f, twin %.4f | "
          "steps 20..199 max rel diff: kernels %.2f, twin %.2f" % (lo[0], {k: round(float(v), 6) for k, v in end.items()},
                                                                     {k: round(v, 2) for k, v in psnr.items()}, rel[:20].max(), rel2[:20].max(),
                                                                     rel[20:].max(), rel2[20:].max()))
    assert all(v < 0.01 * lo[0] for v in end.values()), "no convergence: the comparison would be vacuous"
    assert rel[:20].max() < 0.01, rel[:20].max()
    assert end["kernels"] < 3.0 * max(fp32_end), end
    assert psnr["kernels"] > 48.0 and psnr["kernels"] > min(fp32_psnr) - 6.0, psnr
