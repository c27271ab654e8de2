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


def nerf_params(seed, device="cpu", requires_grad=False, trained_scale=False):
    """dict name -> (weight, bias) torch tensors (oracle layout)."""
    w = synthetic.make_nerf_weights(seed, trained_scale=trained_scale)
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


def fixture_noise(fx):
    """The reference's own torch.rand / randn draws of a perturb > 0 fixture (explicit-noise mode of oracle and kernels)."""
    if float(fx["perturb"]) == 0:
        return None
    return {k: torch.from_numpy(fx["noise_" + k]) for k in ("coarse_u", "fine_u", "sigma_c", "sigma_f")}


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


_ref_lib = None


def ref_mlp_fwd(packed, xyz_cano, sigma, rgb, cidx=None, count=None, n_max=None):
    """fp32 SIMT reference MLP forward on the device (tests/csrc/mlp_ref.cu in tests/libanimnerf_b200_ref.so)."""
    import ctypes
    from anim_nerf_b200._lib import ptr, stream, check
    global _ref_lib
    if _ref_lib is None:
        _ref_lib = ctypes.CDLL(os.path.join(ROOT, "tests", "libanimnerf_b200_ref.so"))
        _ref_lib.an_test_mlp_fwd_ref.restype = ctypes.c_int
        _ref_lib.an_test_mlp_fwd_ref.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] + [ctypes.c_void_p] * 3
    if n_max is None:
        n_max = xyz_cano.numel() // 3
    check(_ref_lib.an_test_mlp_fwd_ref(ptr(packed), ptr(xyz_cano), ptr(cidx), ptr(count), int(n_max), ptr(sigma), ptr(rgb), stream()),
          "an_test_mlp_fwd_ref")


class ref_mlp:
    """Context manager: route the package's MLP forward (inference calls, no stash) through the fp32 reference kernel,
    to separate the bf16 tensor-core error from everything else in an end-to-end comparison."""

    def __enter__(self):
        from anim_nerf_b200 import ops
        self._ops, self._orig = ops, ops.mlp_fwd

        def patched(packed, xyz_cano, sigma, rgb, cidx=None, count=None, n_max=None, stash=None):
            assert stash is None, "the reference kernel is forward-only"
            ref_mlp_fwd(packed, xyz_cano, sigma, rgb, cidx=cidx, count=count, n_max=n_max)
        ops.mlp_fwd = patched
        return self

    def __exit__(self, *a):
        self._ops.mlp_fwd = self._orig
