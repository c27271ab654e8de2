"""CPU-side tests: the C-ABI library loads and exports every symbol include/animnerf_b200.h
declares (no compute without a GPU), host logic, and the product path's refusal to run on CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from util import ROOT, synthetic, body_model


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "animnerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(an_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from anim_nerf_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "binding missing for " + n
    assert set(_lib.SIGNATURES) == set(names)
    # host-only entry points are callable without a GPU
    L = _lib.load()
    assert L.an_version() >= 100
    assert L.an_mlp_packed_bytes() % 1024 == 0
    assert L.an_mlp_grad_floats() == 592388 + 128 * 256 + 128      # flat gradient + fused head-layer scratch
    assert L.an_mlp_stash_bytes(256) == 2 * 608256
    assert L.an_knn_query_ws_bytes(2, 1000) == 48 + 2 * 1000 * 16      # header (counters + statistics) + work list
    assert L.an_vertex_grid_bytes(2, 6890) > 2 * 6890 * 16
    assert b"argument" in L.an_error_string(-1)


def test_argument_errors_do_not_touch_the_gpu():
    from anim_nerf_b200 import _lib
    L = _lib.load()
    assert L.an_composite_fwd(None, None, None, None, None, 10, 64, 1, None, None, None, None, None) == -1
    assert L.an_sample_coarse_fwd(None, 0, 64, 0.0, None, 0, None, None) == -1
    assert L.an_mlp_fwd(None, None, None, None, 10, None, None, None, 0, None) == -1


def test_product_path_has_no_cpu_fallback():
    from anim_nerf_b200.anim_nerf import AnimNeRF
    from anim_nerf_b200.volume_rendering import VolumeRenderer
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=True, freqs_dir=0, body_model_data=synthetic.make_smpl_dict(0))
    p, t = synthetic.make_body_params(1)
    net.set_body_model({k: torch.from_numpy(v) for k, v in p.items()}, {k: torch.from_numpy(v) for k, v in t.items()})
    rays = net.convert_to_body_model_space(torch.from_numpy(synthetic.rays_at_bbox(net.verts.detach().numpy(), 8)))
    net.clac_ober2cano_transform()
    with pytest.raises(Exception):
        VolumeRenderer(n_coarse=8, n_fine=8)(net, rays)


def test_nerf_state_dict_layout_matches_reference_keys():
    from anim_nerf_b200.nerf import NeRF
    net = NeRF(freqs_dir=0)
    keys = set(net.state_dict().keys())
    want = set(synthetic.make_nerf_weights(0).keys())
    assert keys == want
    assert sum(p.numel() for p in net.parameters()) == 592388
    flat = torch.arange(592388, dtype=torch.float32)
    parts = net.split_flat_grad(flat)
    assert [tuple(x.shape) for x in parts] == [tuple(p.shape) for p in net.param_list()]
    assert parts[0][0, 0] == 0 and parts[12][0] == 256 * 63          # weight then bias per linear


def test_synthetic_body_is_deterministic_and_schema_complete():
    d = synthetic.make_smpl_dict(0)
    assert d["v_template"].shape == (6890, 3) and d["weights"].shape == (6890, 24)
    assert d["posedirs"].shape == (6890, 3, 207) and d["J_regressor"].shape == (24, 6890)
    np.testing.assert_allclose(d["weights"].sum(1), 1.0, atol=1e-5)
    d2 = synthetic.make_smpl_dict(0)
    assert all((d[k] == d2[k]).all() for k in d)


def test_body_model_gradients_reach_smpl_params():
    bm = body_model()
    p, _ = synthetic.make_body_params(1)
    t = {k: torch.from_numpy(v).requires_grad_(True) for k, v in p.items()}
    out = bm(**t)
    (out["vertices"].sum() + out["vertices_transform"].sum()).backward()
    assert all(v.grad is not None and torch.isfinite(v.grad).all() for v in t.values())
