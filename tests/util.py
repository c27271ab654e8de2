"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

import anim_nerf_b200  # noqa: E402,F401
from anim_nerf_b200 import synthetic  # noqa: E402
from anim_nerf_b200.body_model import BodyModel  # noqa: E402
from oracle import animnerf_oracle as oracle  # noqa: E402


def load_golden(tag):
    return dict(np.load(os.path.join(GOLDEN, tag + ".npz")))


def nerf_params(seed, device="cpu", requires_grad=False):
    """dict name -> (weight, bias) torch tensors (oracle layout)."""
    w = synthetic.make_nerf_weights(seed)
    p = {}
    for name in synthetic.NERF_LAYER_NAMES:
        W = torch.from_numpy(w[name + ".weight"]).to(device).requires_grad_(requires_grad)
        b = torch.from_numpy(w[name + ".bias"]).to(device).requires_grad_(requires_grad)
        p[name] = (W, b)
    return p


_body = None


def body_model():
    global _body
    if _body is None:
        _body = BodyModel(synthetic.make_smpl_dict(0))
    return _body


def body_params_from_fixture(fx, prefix, requires_grad=False):
    return {k: torch.from_numpy(fx[prefix + k]).clone().requires_grad_(requires_grad)
            for k in ("betas", "global_orient", "body_pose", "transl")}


def golden_tables(fx):
    """(verts (B,V,3), ober2cano (B,V,4,4), lbs_weights) from the fixture's exact tables."""
    verts = torch.from_numpy(fx["verts_body"])
    o2c3 = torch.from_numpy(fx["ober2cano"])
    B, V = o2c3.shape[:2]
    last = torch.tensor([0, 0, 0, 1.0]).expand(B, V, 1, 4)
    return verts, torch.cat([o2c3, last], 2), body_model().lbs_weights


def regulariser_losses(fx, get_sigma, get_normal, dev="cpu"):
    """train.py:264-297 on the fixture's points; returns (loss_fg, loss_bg, loss_normals, normals)."""
    t = lambda a: torch.from_numpy(a).to(dev)                                   # noqa: E731
    k = -2.0 / 64
    l_fg = torch.mean(torch.exp(k * torch.relu(get_sigma(t(fx["fg"])))))
    l_bg = torch.mean(1 - torch.exp(k * torch.relu(get_sigma(t(fx["bg"])))))
    n_p, n_q = get_normal(t(fx["points"])), get_normal(t(fx["neighbs"]))
    unit = lambda v: v / (torch.norm(v, p=2, dim=-1, keepdim=True) + 1e-5)      # noqa: E731
    return l_fg, l_bg, torch.nn.functional.mse_loss(unit(n_p), unit(n_q)), (n_p, n_q)
