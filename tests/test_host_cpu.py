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
    assert L.an_mlp_stash_bytes(256) == 4 * 608256          # whole CTA-pair iterations: 512 points = 4 tiles of 128
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
    net = NeRF(freqs_dir=0, use_view=False)
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


# ------------------------------------------------------------------ B4: system state, checkpoints, optimiser groups
def _system(**over):
    from anim_nerf_b200.system import AnimNeRFSystem
    return AnimNeRFSystem(body_model_data=synthetic.make_smpl_dict(0), n_samples=64, n_importance=64, **over)


def test_reference_checkpoint_state_dict_loads():
    """A state dict with exactly the names/shapes a reference checkpoint carries (captured from the reference's
    modules, tests/golden/state_dict_keys.json) loads strictly; the MLP weights and the SMPL table arrive intact."""
    import json
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    g = torch.Generator().manual_seed(0)
    sd = {k: torch.randn(*shape, generator=g) if shape else torch.zeros(()) for k, shape in keys.items()}
    for k in list(sd):
        if k.endswith("parents") or k.endswith("extra_joints_idxs") or k.endswith("faces_tensor"):
            sd[k] = torch.zeros(keys[k], dtype=torch.long)
    sysm = _system(num_frames=3, optim_body_params=True)          # table of another length: rows come from the checkpoint
    parents_before = sysm.anim_nerf.body_model.parents.clone()
    sd["anim_nerf.body_model.parents"] = parents_before.clone()
    missing, dropped = sysm.load_reference_state_dict(sd, strict=True)
    assert not missing and len(dropped) == 6
    assert torch.equal(sysm.anim_nerf.nerf.xyz_encoding_5[0].weight, sd["anim_nerf.nerf.xyz_encoding_5.0.weight"])
    assert torch.equal(sysm.anim_nerf.nerf_fine.rgb[0].bias, sd["anim_nerf.nerf_fine.rgb.0.bias"])
    assert torch.equal(sysm.body_model_params.body_pose.weight, sd["body_model_params.body_pose.weight"])
    assert sysm.body_model_params.body_pose.weight.shape == (7, 69) and sysm.body_model_params.body_pose.weight.requires_grad
    # every reference key is either consumed or one of the six smplx leftovers
    own = set(sysm.state_dict())
    assert all(k in own or k in dropped for k in keys)
    with pytest.raises(RuntimeError):
        sysm.load_reference_state_dict({k: v for k, v in sd.items() if "xyz_encoding_3" not in k}, strict=True)


def test_body_model_params_table_and_optimiser_groups():
    """models/body_model_params.py:5-66 semantics + train.py:217-226 optimiser groups and the poly schedule."""
    F = 5
    sysm = _system(num_frames=F, optim_body_params=True)
    posed, _ = synthetic.make_body_params(F, seed=4)
    sysm.init_body_model_params({k: torch.from_numpy(v) for k, v in posed.items()})
    tab = sysm.body_model_params
    assert tab.betas.weight.shape == (1, 10) and tab.body_pose.weight.shape == (F, 69)
    np.testing.assert_allclose(tab.betas.weight.detach().numpy()[0], posed["betas"].mean(0), atol=1e-7)   # betas: mean over frames
    out = tab(torch.tensor([3, 1]))
    assert torch.equal(out["body_pose"], torch.from_numpy(posed["body_pose"][[3, 1]]))
    assert torch.equal(out["betas"][0], out["betas"][1]) and out["betas"].shape == (2, 10)
    assert all(p.requires_grad for p in tab.parameters())
    (opt,), (sched,) = sysm.configure_optimizers()
    assert [g["lr"] for g in opt.param_groups] == [5e-4, 2.5e-4]
    assert len(opt.param_groups[0]["params"]) == 48 and len(opt.param_groups[1]["params"]) == 4
    opt.step(); sched.step()
    assert abs(opt.param_groups[0]["lr"] - 5e-4 * (1 - 1 / 30) ** 0.9) < 1e-12      # config.py:64 max_epochs = 30
    frozen = _system(num_frames=F, optim_body_params=False)      # table frozen, one group
    assert not any(p.requires_grad for p in frozen.body_model_params.parameters())
    assert len(frozen.configure_optimizers()[0][0].param_groups) == 1


def test_system_builds_and_configures_from_the_reference_config_schema():
    """The drop-in claim of B4: the system is constructed from a namespace shaped like the reference's
    `get_default_config()` merged with male-3-casual.yaml (config.py:7-78: nested train.optimizer / train.scheduler
    nodes, optim_body_params True, chunk 2048, epsilon 0.01, max_epochs 30) and `configure_optimizers` honours it;
    the package's own defaults agree with those values."""
    from types import SimpleNamespace as NS
    from anim_nerf_b200.system import AnimNeRFSystem, default_hparams
    cfg = NS(num_gpus=-1, exp_name="male-3-casual", dataset_name="anim_nerf", root_dir="./data/people_snapshot/male-3-casual",
             model_type="smpl", gender="male", model_path="./smplx/models", img_wh=(512, 512), freqs_xyz=10, freqs_dir=0,
             use_view=False, use_knn=True, k_neigh=4, use_unpose=True, unpose_view=False, use_deformation=False,
             deformation_dim=0, apperance_dim=0, latent_dim=0, pose_dim=69, optim_body_params=True, dis_threshold=0.2,
             n_samples=64, n_importance=32, n_depth=0, share_fine=False, chunk=2048, query_inside=False, white_bkgd=True,
             num_frames=4, frame_IDs=[1, 5, 9, 13],
             train=NS(lambda_alphas=0.1, lambda_foreground=0.01, lambda_background=0.01, lambda_normals=0.01, lambda_cycle=0.1,
                      epsilon=0.01, batch_size=16, max_epochs=30, max_steps=200000, lr=5e-4,
                      optimizer={"type": "adam", "momentum": 0.9, "weight_decay": 0},          # a CfgNode is a dict subclass
                      scheduler=NS(type="poly", poly_exp=0.9), num_workers=8))
    sysm = AnimNeRFSystem(cfg, body_model_data=synthetic.make_smpl_dict(0))
    (opt,), (sched,) = sysm.configure_optimizers()
    assert [g["lr"] for g in opt.param_groups] == [5e-4, 2.5e-4] and opt.defaults["weight_decay"] == 0 and opt.defaults["eps"] == 1e-8
    opt.step(); sched.step(); sched.step()
    assert abs(opt.param_groups[1]["lr"] - 2.5e-4 * (1 - 2 / 30) ** 0.9) < 1e-12
    cfg.train.optimizer = {"type": "sgd", "momentum": 0.9, "weight_decay": 0}
    with pytest.raises(NotImplementedError):
        sysm.configure_optimizers()
    d = default_hparams()
    for k in ("freqs_xyz", "freqs_dir", "k_neigh", "use_knn", "use_unpose", "optim_body_params", "dis_threshold", "n_samples",
              "n_importance", "n_depth", "share_fine", "chunk", "query_inside", "white_bkgd"):
        assert getattr(d, k) == getattr(cfg, k), k
    for k in ("lambda_alphas", "lambda_foreground", "lambda_background", "lambda_normals", "epsilon", "max_epochs", "lr"):
        assert getattr(d.train, k) == getattr(cfg.train, k), k
    # a missing model file is an error, not a silently substituted synthetic body
    with pytest.raises(FileNotFoundError):
        AnimNeRFSystem(cfg)


def test_decode_batch_accepts_the_reference_datasets_flat_keys():
    sysm = _system()
    posed, tmpl = synthetic.make_body_params(2, seed=4)
    flat = {k: torch.from_numpy(v) for k, v in posed.items()}
    flat.update({k + "_template": torch.from_numpy(v) for k, v in tmpl.items()})
    flat.update(rays=torch.zeros(2, 4, 4, 8), rgbs=torch.zeros(2, 4, 4, 3), alphas=torch.zeros(2, 4, 4, 1), frame_idx=torch.tensor([0, 1]))
    out = sysm.decode_batch(flat)
    assert set(out[6]) == set(out[7]) == {"betas", "global_orient", "body_pose", "transl"}
    assert torch.equal(out[7]["body_pose"], flat["body_pose_template"]) and out[8] is None


def test_fused_adam_has_no_cpu_path_and_validates_arguments():
    """FusedAdam is the CUDA library's optimiser: on CPU tensors it refuses to step (no silent fallback); bad
    hyper-parameters are rejected like torch.optim.Adam rejects them; param_groups / schedulers work on the host side."""
    from anim_nerf_b200.optim import FusedAdam
    p = torch.zeros(4, 3, requires_grad=True)
    opt = FusedAdam([p], lr=5e-4, eps=1e-8)
    assert opt.param_groups[0]["lr"] == 5e-4 and opt.param_groups[0]["betas"] == (0.9, 0.999)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 0.5 ** e)
    p.grad = torch.ones_like(p)
    with pytest.raises(RuntimeError):
        opt.step()
    opt.zero_grad(set_to_none=True)
    opt.step()                                      # nothing to do without gradients: no launch, no error
    sched.step()
    assert abs(opt.param_groups[0]["lr"] - 2.5e-4) < 1e-12
    for bad in (dict(lr=-1.0), dict(eps=-1e-8), dict(betas=(1.0, 0.999)), dict(weight_decay=-0.1)):
        with pytest.raises(ValueError):
            FusedAdam([p], **bad)


def test_host_mirrors_keep_the_reference_signatures():
    """Drop-in surface: every constructor / method of the reference classes on the boundary exists on the
    mirror with the same parameter names, order and defaults (tests/golden/api_signatures.json is taken with
    inspect.signature from the imported reference; train.py's AnimNeRFSystem, which needs pytorch-lightning to import,
    is read with `ast`).  Extra trailing keywords on our side are allowed."""
    import inspect
    import json
    from anim_nerf_b200.anim_nerf import AnimNeRF
    from anim_nerf_b200.body_model_params import BodyModelParams
    from anim_nerf_b200.nerf import NeRF
    from anim_nerf_b200.volume_rendering import VolumeRenderer
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "api_signatures.json")))
    from anim_nerf_b200.system import AnimNeRFSystem
    ours = {"AnimNeRF": AnimNeRF, "VolumeRenderer": VolumeRenderer, "NeRF": NeRF, "BodyModelParams": BodyModelParams,
            "AnimNeRFSystem": AnimNeRFSystem}
    # documented differences: the stand-alone inverse-CDF sampler is fused with the sort-merge
    known = {("VolumeRenderer", "sample_fine"): "VolumeRenderer.sample_fine_merge (an_sample_fine_merge_fwd)"}
    problems = []
    for cname, methods in ref.items():
        for m, params in methods.items():
            if (cname, m) in known:
                assert hasattr(ours[cname], "sample_fine_merge")
                continue
            if not hasattr(ours[cname], m):
                problems.append(("missing", cname, m)); continue
            mine = {n: p for n, p in inspect.signature(getattr(ours[cname], m)).parameters.items() if n != "self"}
            has_kw = any(p.kind.name == "VAR_KEYWORD" for p in mine.values())
            ref_pos = [n for n, _, k in params if k == "POSITIONAL_OR_KEYWORD"]
            my_pos = [n for n, p in mine.items() if p.kind.name == "POSITIONAL_OR_KEYWORD"]
            if my_pos[:len(ref_pos)] != ref_pos and not (has_kw and all(n in mine or has_kw for n in ref_pos)):
                problems.append(("order", cname, m, ref_pos, my_pos))
            for n, default, kind in params:
                if kind != "POSITIONAL_OR_KEYWORD" or n not in mine:
                    continue
                d = None if mine[n].default is inspect._empty else repr(mine[n].default)
                if default is not None and d != default:          # (a default where the reference has none is a superset)
                    problems.append(("default", cname, m, n, default, d))
    assert not problems, problems


def test_graph_step_tree_helpers_mirror_nested_batches():
    """GraphedTrainStep's static buffers follow the nesting of the example batch (dicts / lists of tensors, other leaves
    kept as they are); a copy reaches every tensor leaf and nothing else."""
    from anim_nerf_b200.graph_step import _tree_map, _tree_copy
    batch = {"rays": torch.arange(6.0).view(2, 3), "smpl": {"posed": {"transl": torch.zeros(2, 3)}, "list": [torch.ones(2), 5]}, "tag": "x"}
    static = _tree_map(lambda t: t.clone(), batch)
    assert static["tag"] == "x" and static["smpl"]["list"][1] == 5
    assert static["rays"].data_ptr() != batch["rays"].data_ptr()
    new = {"rays": batch["rays"] + 1, "smpl": {"posed": {"transl": torch.full((2, 3), 7.0)}, "list": [torch.full((2,), 3.0), 5]}, "tag": "y"}
    ptrs = (static["rays"].data_ptr(), static["smpl"]["posed"]["transl"].data_ptr())
    _tree_copy(static, new)
    assert torch.equal(static["rays"], new["rays"]) and float(static["smpl"]["posed"]["transl"].min()) == 7.0
    assert torch.equal(static["smpl"]["list"][0], new["smpl"]["list"][0])
    assert ptrs == (static["rays"].data_ptr(), static["smpl"]["posed"]["transl"].data_ptr())      # in place: the graph's addresses


def test_body_model_params_lookup_equals_the_reference_embedding_lookup():
    """`BodyModelParams.forward` (reference models/body_model_params.py:60-66: `emb(ids)` per table, betas looked up at
    id 0) reads the rows with gathers / an expand: same values, same shapes (1-D and 2-D id tensors, repeated ids) and
    the same gradients as the nn.Embedding lookups."""
    import torch
    from anim_nerf_b200.body_model_params import BodyModelParams
    m = BodyModelParams(20)
    g = torch.Generator().manual_seed(0)
    for n, d in m.params_dim.items():
        m.init_parameters(n, torch.randn(20, d, generator=g), requires_grad=True)
    for ids in (torch.tensor([3, 7, 7, 0, 19]), torch.tensor([[5], [5], [11]])):
        out = m(ids)
        ref = {n: getattr(m, n)(torch.zeros_like(ids) if n == "betas" else ids) for n in m.param_names}
        grads = []
        for res in (out, ref):
            for n in m.param_names:
                getattr(m, n).weight.grad = None
            sum((o * torch.arange(o.numel(), dtype=torch.float32).view_as(o)).sum() for o in res.values()).backward()
            grads.append({n: getattr(m, n).weight.grad.clone() for n in m.param_names})
        for n in m.param_names:
            assert out[n].shape == ref[n].shape and torch.equal(out[n], ref[n]), n
            assert torch.allclose(grads[0][n], grads[1][n], rtol=1e-6, atol=0), n


def test_packed_neighbour_keys_order_like_the_distance_index_contract():
    """The grid search keeps its best list as 64-bit keys (bits(d2) << 32 | index) and the fine-depth merge sorts
    (ordered bits(z) << 32 | draw index) (csrc/knn_unpose.cu `key_pack`, csrc/sample_fine.cu `ord_bits`): unsigned order of
    the packed keys must be the (value, index) order of the contract -- for squared distances (>= +0, incl. +inf and
    exact ties) and, with the sign-flip map, for depths of either sign."""
    import numpy as np
    rs = np.random.RandomState(0)
    d2 = np.concatenate([rs.uniform(0, 4, 4000).astype(np.float32) ** 2, np.zeros(50, np.float32),
                         np.full(50, np.inf, np.float32), np.float32([1e-38, 1e-45, 3.4e38])])
    d2[:2000] = rs.choice(d2[2000:2100], 2000)                       # many exact ties
    idx = rs.randint(0, 6890, d2.size).astype(np.int64)
    key = (d2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | idx.astype(np.uint64)
    assert np.array_equal(np.argsort(key, kind="stable"), np.lexsort((idx, d2)))
    z = np.concatenate([rs.normal(0, 3, 5000).astype(np.float32), np.float32([0.0, -0.0, 1e-45, -1e-45])])
    b = z.view(np.uint32)
    ob = np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)
    back = np.where(ob & np.uint32(0x80000000), ob & np.uint32(0x7fffffff), ~ob).astype(np.uint32).view(np.float32)
    assert np.array_equal(back.view(np.uint32), b)                   # the map round-trips every bit pattern
    order = np.argsort(ob, kind="stable")
    assert np.all(np.diff(z[order]) >= 0)                            # ... and is monotone (-0.0 sorts before +0.0)
