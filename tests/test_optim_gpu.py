"""GPU parity of the fused Adam step (an_adam_step / FusedAdam) against torch.optim.Adam configured as the reference
does (utils/__init__.py:33-45): same parameters, moments and step counts after several steps, with and without
weight decay, ragged tensor sizes, a learning-rate schedule, more tensors than one launch holds, and under CUDA-graph
replay."""
import numpy as np
import pytest
import torch

from util import synthetic  # noqa: F401  (path setup)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for s in shapes]


def _grads(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g).to(DEV) for s in shapes]


@pytest.mark.parametrize("weight_decay", [0.0, 1e-2])
def test_matches_torch_adam_over_steps(weight_decay):
    from anim_nerf_b200.optim import FusedAdam
    shapes = [(256, 63), (256,), (256, 319), (1, 256), (3, 128), (1,), (7, 5, 3)]
    a, b = _params(shapes, 0), _params(shapes, 0)
    ref = torch.optim.Adam(a, lr=5e-4, eps=1e-8, weight_decay=weight_decay)
    ours = FusedAdam(b, lr=5e-4, eps=1e-8, weight_decay=weight_decay)
    sched_r = torch.optim.lr_scheduler.LambdaLR(ref, lambda e: (1 - e / 20) ** 0.9)
    sched_o = torch.optim.lr_scheduler.LambdaLR(ours, lambda e: (1 - e / 20) ** 0.9)
    for it in range(6):
        gs = _grads(shapes, 100 + it)
        for p, q, g in zip(a, b, gs):
            p.grad, q.grad = g.clone(), g.clone()
        ref.step(); ours.step()
        sched_r.step(); sched_o.step()
    # (the step count is kept per launch group, torch keeps it per tensor: identical whenever every parameter of the
    # group receives a gradient each step, which is the case for the MLPs on this path)
    worst = [0.0, 0.0, 0.0]
    for p, q in zip(a, b):
        for k, (x, y) in enumerate(((q.detach(), p.detach()), (ours.state[q]["exp_avg"], ref.state[p]["exp_avg"]),
                                    (ours.state[q]["exp_avg_sq"], ref.state[p]["exp_avg_sq"]))):
            worst[k] = max(worst[k], float(((x - y).abs() / (y.abs() + 1e-6)).max()))
    print("max relative deviation: params %.3g exp_avg %.3g exp_avg_sq %.3g" % tuple(worst))
    # moments: fp32 round-off of the two evaluation orders (measured: <= 5e-8 absolute on O(0.1) values);
    # parameters move by ~lr per step: deviations are measured against that scale
    for p, q in zip(a, b):
        assert float((q.detach() - p.detach()).abs().max()) < 5e-4 * 1e-3, "parameter update deviates by more than 1e-3 of one lr step"
        np.testing.assert_allclose(ours.state[q]["exp_avg"].cpu().numpy(), ref.state[p]["exp_avg"].cpu().numpy(), rtol=1e-5, atol=5e-7)
        np.testing.assert_allclose(ours.state[q]["exp_avg_sq"].cpu().numpy(), ref.state[p]["exp_avg_sq"].cpu().numpy(), rtol=1e-5, atol=1e-8)


def test_more_tensors_than_one_launch_and_graph_replay():
    from anim_nerf_b200.optim import FusedAdam
    shapes = [(17, 3)] * 70                               # two launches (64 + 6), each with its own step counter
    a, b = _params(shapes, 1), _params(shapes, 1)
    ref = torch.optim.Adam(a, lr=1e-3, eps=1e-8)
    ours = FusedAdam(b, lr=1e-3, eps=1e-8)
    static_g = [torch.zeros(*s, device=DEV) for s in shapes]
    for q, g in zip(b, static_g):
        q.grad = g
    ours.step()                                           # eager step (gradients zero: parameters unchanged, step = 1)
    ref_step0 = [torch.zeros(*s, device=DEV) for s in shapes]
    for p, g in zip(a, ref_step0):
        p.grad = g
    ref.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ours.step()
    for it in range(4):
        gs = _grads(shapes, 200 + it)
        for sg, g in zip(static_g, gs):
            sg.copy_(g)
        if it == 2:
            for grp in ours.param_groups + ref.param_groups:
                grp["lr"] = 2.5e-4                        # schedule change between replays
            ours.sync_lr()
        graph.replay()
        for p, g in zip(a, gs):
            p.grad = g
        ref.step()
    torch.cuda.synchronize()
    for p, q in zip(a, b):
        assert float((q.detach() - p.detach()).abs().max()) < 1e-3 * 1e-3


def test_state_dict_resume_parity_with_torch_adam():
    """Checkpoint / resume: the step count travels in state_dict() as torch.optim.Adam's state[p]['step'];
    a FusedAdam checkpoint resumes in FusedAdam, a torch.optim.Adam checkpoint (what a Lightning .ckpt of the reference
    holds) resumes in FusedAdam, and a FusedAdam checkpoint resumes in torch.optim.Adam -- each continuing exactly like an
    uninterrupted torch.optim.Adam run (bias correction included: a restart at t = 1 would be off by ~30x at step 4)."""
    import copy
    from anim_nerf_b200.optim import FusedAdam
    shapes = [(64, 63), (64,), (3, 64), (3,)]
    mk = lambda cls, ps: cls(ps, lr=5e-4, eps=1e-8)        # noqa: E731

    def run(opt, ps, its):
        for it in its:
            for p, g in zip(ps, _grads(shapes, 300 + it)):
                p.grad = g
            opt.step()

    a = _params(shapes, 2)                                 # uninterrupted torch run: 3 + 3 steps
    ref = mk(torch.optim.Adam, a)
    run(ref, a, range(3))
    ref_mid = copy.deepcopy(ref.state_dict()); a_mid = [p.detach().clone() for p in a]
    run(ref, a, range(3, 6))

    b = _params(shapes, 2)                                 # FusedAdam: 3 steps, checkpoint
    ours = mk(FusedAdam, b)
    run(ours, b, range(3))
    sd = copy.deepcopy(ours.state_dict())
    assert all(float(s["step"]) == 3.0 for s in sd["state"].values()), [float(s["step"]) for s in sd["state"].values()]

    def resumed(cls, state, start):
        ps = [p.clone().requires_grad_(True) for p in start]
        opt = mk(cls, ps)
        opt.load_state_dict(copy.deepcopy(state))
        run(opt, ps, range(3, 6))
        return ps, opt

    for label, (ps, opt) in {"fused->fused": resumed(FusedAdam, sd, [p.detach() for p in b]),
                             "torch->fused": resumed(FusedAdam, ref_mid, a_mid),
                             "fused->torch": resumed(torch.optim.Adam, sd, [p.detach() for p in b])}.items():
        for p, q in zip(a, ps):
            assert float((q.detach() - p.detach()).abs().max()) < 5e-4 * 1e-3, label
        assert all(float(s["step"]) == 6.0 for s in opt.state_dict()["state"].values()), label


def test_changing_the_set_of_parameters_with_gradients_raises():
    from anim_nerf_b200.optim import FusedAdam
    shapes = [(8, 4), (8,)]
    ps = _params(shapes, 3)
    opt = FusedAdam(ps, lr=1e-3)
    for p, g in zip(ps, _grads(shapes, 1)):
        p.grad = g
    opt.step()
    ps[1].grad = None
    with pytest.raises(RuntimeError, match="changed after its first step"):
        opt.step()
