"""GPU parity of the inference entry points (BASELINE cfg3 / cfg4 shapes at test size):
full-frame rendering from camera parameters and the density-grid query of mesh extraction,
against the CPU oracle on the same tables, plus the size-independent sharding property
(row / lattice slabs rendered separately are bit-identical to the unsharded result)."""
import numpy as np
import pytest
import torch

from util import oracle, nerf_params, body_model, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _system(n_fine=64):
    from anim_nerf_b200.system import AnimNeRFSystem
    sysm = AnimNeRFSystem(body_model_data=synthetic.make_smpl_dict(0), n_samples=64, n_importance=n_fine).to(DEV)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):
        sd = {k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}
        getattr(sysm.anim_nerf, name).load_state_dict(sd, strict=True)
    return sysm


def _frame(B=1, seed=1):
    posed_np, tmpl_np = synthetic.make_body_params(B, seed=seed)
    posed = {k: torch.from_numpy(v) for k, v in posed_np.items()}
    tmpl = {k: torch.from_numpy(v) for k, v in tmpl_np.items()}
    return posed, tmpl


def _cam(W, H, B=1):
    cam = synthetic.make_camera(W, H, focal_scale=0.55)      # body fills the small test frame
    c2w = torch.from_numpy(cam["c2w"])[None].repeat(B, 1, 1)
    focal = torch.from_numpy(cam["focal"])[None].repeat(B, 1)
    c = torch.from_numpy(cam["c"])[None].repeat(B, 1)
    return c2w, focal, c


def _oracle_tables(posed, tmpl):
    bm = body_model()
    with torch.no_grad():
        po, to = bm(**posed), bm(**tmpl)
    verts, o2c = oracle.ober2cano_tables(po, to)
    return po, verts, o2c, bm.lbs_weights


def test_render_frame_matches_oracle_and_row_slabs_are_bit_identical():
    from anim_nerf_b200 import inference
    H = W = 40
    sysm = _system()
    posed, tmpl = _frame()
    c2w, focal, c = _cam(W, H)
    args = (sysm.volume_renderer, sysm.anim_nerf, c2w.to(DEV), focal.to(DEV), c.to(DEV), H, W,
            {k: v.to(DEV) for k, v in posed.items()}, {k: v.to(DEV) for k, v in tmpl.items()})
    full = inference.render_frame(*args)
    assert full["rgbs_fine"].shape == (1, H, W, 3) and full["alphas_fine"].shape == (1, H, W, 1)
    # oracle: reference ray generation -> body space -> render on CPU
    po, verts, o2c, lbs = _oracle_tables(posed, tmpl)
    rays_w = oracle.gen_rays(c2w[0], H, W, focal[0], 0.1, 10.0, c[0]).view(1, H * W, 8)
    rays_b = oracle.rays_to_body_space(rays_w, po["joints_transform"][:, 0])
    ref = oracle.render_rays(nerf_params(10), nerf_params(11), rays_b, (verts, o2c, lbs), n_coarse=64, n_fine=64)
    hit = float((ref["alphas_fine"] > 0.5).float().mean())
    assert 0.05 < hit < 0.95, hit                                  # the frame sees body and background
    for k in ("rgbs", "alphas", "rgbs_fine", "alphas_fine"):
        got = full[k].reshape(1, H * W, -1).cpu()
        # tables are rebuilt on the GPU in fp32: a handful of samples at the validity threshold may flip
        bad = ((got - ref[k]).abs() > 1e-2).float().mean()
        assert float(bad) < 0.01, (k, float(bad))
    mse = float(((full["rgbs_fine"].reshape(1, H * W, 3).cpu() - ref["rgbs_fine"]) ** 2).mean())
    assert mse < 1e-5, mse
    # sharding property: row sets rendered separately (rank r: rows r, r+world, ...) == the unsharded frame, bit for bit
    for world in (2, 3):
        parts = [inference.render_frame_sharded(*args, rank=r, world=world, gather=False) for r in range(world)]
        for k in full:
            for r in range(world):
                assert torch.equal(parts[r][k], full[k][:, r::world]), (world, r, k)
    # chunked launches change no value
    chunked = inference.render_frame(*args, chunk=333)
    for k in full:
        assert torch.equal(chunked[k], full[k]), k


def test_batched_inference_turntable_matches_system_forward():
    """novel_view.py:75-116 mirror: P = identity reproduces AnimNeRFSystem.forward; a rotation about the
    body's y axis keeps the alpha mass (the body stays in view) and changes the image."""
    from anim_nerf_b200 import inference
    H = W = 32
    sysm = _system()
    posed, tmpl = _frame()
    c2w, focal, c = _cam(W, H)
    posed_d = {k: v.to(DEV) for k, v in posed.items()}
    tmpl_d = {k: v.to(DEV) for k, v in tmpl.items()}
    rays_w = oracle.gen_rays(c2w[0], H, W, focal[0], 0.1, 10.0, c[0]).view(1, H * W, 8).to(DEV)
    with torch.no_grad():
        ref = sysm(rays_w.view(1, H, W, 8), posed_d, tmpl_d, perturb=0.0)
    eye = torch.eye(4, device=DEV).view(1, 1, 4, 4)
    out = inference.batched_inference(sysm.volume_renderer, sysm.anim_nerf, rays_w, posed_d, tmpl_d, P=eye)
    for k in ref:
        assert torch.equal(out[k].view_as(ref[k]), ref[k]), k
    ang = np.pi / 2
    Ry = torch.tensor([[np.cos(ang), 0, np.sin(ang), 0], [0, 1, 0, 0], [-np.sin(ang), 0, np.cos(ang), 0], [0, 0, 0, 1]],
                      dtype=torch.float32, device=DEV).view(1, 1, 4, 4)
    rot = inference.batched_inference(sysm.volume_renderer, sysm.anim_nerf, rays_w, posed_d, tmpl_d, P=Ry)
    assert float(rot["alphas_fine"].mean()) > 0.02
    assert float((rot["rgbs_fine"] - out["rgbs_fine"]).abs().max()) > 1e-2


def test_density_grid_matches_oracle_and_slabs_are_bit_identical():
    from anim_nerf_b200 import inference
    N = 20
    sysm = _system()
    posed, tmpl = _frame()
    an = sysm.anim_nerf
    an.set_body_model({k: v.to(DEV) for k, v in posed.items()}, {k: v.to(DEV) for k, v in tmpl.items()})
    an.convert_to_body_model_space(None)
    an.clac_ober2cano_transform()
    with torch.no_grad():
        sig = inference.query_density_grid(an, N)
    assert sig.shape == (N, N, N) and float(sig.min()) >= 0.0
    # lattice construction equals the reference's numpy recipe bit for bit
    center = (an.verts.max(dim=1)[0] + an.verts.min(dim=1)[0]) / 2.0
    grid = inference.create_grid(N, (-1.2, 1.2), (-1.2, 1.2), (-1.2, 1.2))
    pts_ref = torch.from_numpy(grid.reshape(-1, 3)).unsqueeze(0).float().to(DEV)
    pts_ref += center
    pts = inference.grid_slab_points(N, (-1.2, 1.2), (-1.2, 1.2), (-1.2, 1.2), center[0], 0, N, DEV)
    assert torch.equal(pts, pts_ref)
    # oracle on the kernel's own tables (isolates the query path from the fp32 table rebuild)
    tables = (an.verts.cpu(), an.ober2cano_transform.cpu(), an.body_model.lbs_weights.cpu())
    _, s_ref, aux = oracle.field(nerf_params(11), pts_ref.cpu(), tables, 0.2)
    s_ref = torch.relu(s_ref).view(N, N, N)
    inside = float((s_ref > 0).float().mean())
    assert inside > 0.005, inside
    bad = ((sig.cpu() - s_ref).abs() > 5e-2).float().mean()
    assert float(bad) < 2e-3, float(bad)
    for world in (2, 3):
        parts = [inference.query_density_grid_sharded(an, N, rank=r, world=world, gather=False) for r in range(world)]
        for r in range(world):                       # rank r owns lattice rows r, r + world, ...
            assert torch.equal(parts[r], sig[r::world]), (world, r)
    small = inference.query_density_grid(an, N, slab_rows=3)
    assert torch.equal(small, sig)
    # the lattice mode of the KNN kernel (points generated in the kernel) equals the explicit-points path bit for bit,
    # canonical points and validity included, on a ragged lattice slab too
    explicit = inference.batched_point_inference(an, pts_ref).view(N, N, N)
    assert torch.equal(explicit, sig)
    ax = [torch.from_numpy(np.linspace(-1.2, 1.2, N)).float().to(DEV) for _ in range(3)]
    rows = ax[1][3:17:4]
    from anim_nerf_b200 import ops
    cfg = an._cfg(True)
    lat = ops.knn_unpose_lattice(cfg["verts"], an.ober2cano_transform.contiguous(), cfg["lbs"], cfg["thr"], ax[0][:13], rows, ax[2],
                                 center[0], grid=cfg["grid"])
    p4 = pts_ref.view(N, N, N, 3)[3:17:4, :13].reshape(1, -1, 3).contiguous()
    ref = ops.knn_unpose(cfg["verts"], an.ober2cano_transform.contiguous(), cfg["lbs"], cfg["thr"], xyz=p4, grid=cfg["grid"], compact=True)
    assert torch.equal(lat["valid"], ref["valid"]) and torch.equal(lat["xyz_cano"], ref["xyz_cano"])
    assert int(lat["count"]) == int(ref["count"]) > 0
    assert torch.equal(torch.sort(lat["cidx"][:int(lat["count"])])[0], torch.sort(ref["cidx"][:int(ref["count"])])[0])
