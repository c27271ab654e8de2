"""GPU: marching cubes of the density lattice (reference extract_mesh.py:152-169) through the C ABI
(an_mc_count / an_mc_scan / an_mc_emit) against the CPU oracle -- exact on small lattices (faces equal, vertices bit for
bit) -- and through size-independent properties at the reference's default N_grid = 256."""
import numpy as np
import pytest
import torch

from util import ROOT, synthetic  # noqa: F401
from oracle import mcubes_oracle as mo
from anim_nerf_b200 import mesh

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _grid(shape):
    return np.stack(np.meshgrid(*[np.arange(n) for n in shape], indexing="ij"), -1).astype(np.float32)


def _fields():
    rng = np.random.RandomState(3)
    g = _grid((20, 20, 20))
    yield "sphere", (np.linalg.norm(g - np.array([9.3, 9.6, 10.1], np.float32), axis=-1) - 6.5).astype(np.float32), 0.0
    yield "noise_ragged", np.pad(rng.randn(9, 12, 15).astype(np.float32), 1, constant_values=5.0), 0.0      # nx != ny != nz, every configuration
    yield "noise_open", rng.randn(8, 7, 33).astype(np.float32), 0.25                                          # surface runs into the lattice boundary
    g = _grid((24, 18, 40))
    blobs = np.full(g.shape[:3], 1.0, np.float32)
    for _ in range(10):
        c = rng.uniform(4, 14, 3) * np.array([1.0, 1.0, 2.2])
        blobs -= (2.0 * np.exp(-((g - c) ** 2).sum(-1) / (2 * rng.uniform(1.5, 3.0) ** 2))).astype(np.float32)
    yield "blobs", blobs, 0.0
    yield "empty", np.ones((5, 6, 7), np.float32), 0.0


@pytest.mark.parametrize("name,vol,iso", list(_fields()), ids=[f[0] for f in _fields()])
def test_marching_cubes_equals_oracle(name, vol, iso):
    v, f = mesh.marching_cubes(torch.from_numpy(vol).to(DEV), iso)
    vo, fo = mo.marching_cubes(vol, iso)
    assert v.shape == vo.shape and f.shape == fo.shape, (v.shape, vo.shape, f.shape, fo.shape)
    assert np.array_equal(f.cpu().numpy(), fo)                       # index work: exact
    assert np.array_equal(v.cpu().numpy().view(np.uint32), vo.view(np.uint32)), np.abs(v.cpu().numpy() - vo).max()     # same fp32 formula: bit for bit
    if name not in ("noise_open", "empty"):
        assert mo.mesh_report(v.cpu().numpy(), f.cpu().numpy())["closed"]


def test_marching_cubes_256_sphere_properties_and_reproducibility():
    """The reference's default lattice (N_grid = 256, extract_mesh.py:78): 16.8 M lattice points, 16 Ki blocks through the
    scan.  A sphere's mesh is a closed, outward-oriented genus-0 surface whose vertices lie on the sphere and whose
    volume is that of the ball; two runs give identical bytes (no atomics)."""
    N, r = 256, 90.3
    ax = torch.arange(N, device=DEV, dtype=torch.float32)
    c = torch.tensor([127.2, 126.9, 128.4], device=DEV)
    vol = torch.sqrt((ax.view(N, 1, 1) - c[0]) ** 2 + (ax.view(1, N, 1) - c[1]) ** 2 + (ax.view(1, 1, N) - c[2]) ** 2) - r
    v, f = mesh.marching_cubes(vol, 0.0)
    v2, f2 = mesh.marching_cubes(vol, 0.0)
    assert torch.equal(v, v2) and torch.equal(f, f2)
    rep = mo.mesh_report(v.cpu().numpy(), f.cpu().numpy())
    assert rep["closed"] and rep["manifold"] and rep["euler"] == 2, rep
    ball = 4.0 / 3.0 * np.pi * r ** 3
    assert 0.9995 * ball < rep["volume"] < ball, (rep["volume"], ball)
    rad = (v - c).norm(dim=1)
    assert float(rad.max()) <= r + 1e-3 and float(rad.min()) > r - 2e-3, (float(rad.min()), float(rad.max()))
    assert int(f.max()) == v.shape[0] - 1 and int(f.min()) == 0
    print("N=256 sphere: %d vertices, %d faces, volume %.1f (ball %.1f)" % (v.shape[0], f.shape[0], rep["volume"], ball))


def test_extract_mesh_end_to_end_matches_oracle_on_the_same_lattice():
    """extract_mesh.py:152-169 on a synthetic frame: the density lattice comes from the kernels (query_density_grid), the
    mesh from the device marching cubes; the oracle runs marching cubes on the same lattice values."""
    from anim_nerf_b200.anim_nerf import AnimNeRF
    from anim_nerf_b200.inference import query_density_grid
    net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=True, freqs_dir=0, body_model_data=synthetic.make_smpl_dict(0)).to(DEV)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):
        getattr(net, name).load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}, strict=True)
    posed_np, tmpl_np = synthetic.make_body_params(1, seed=5)
    posed = {k: torch.from_numpy(v).to(DEV) for k, v in posed_np.items()}
    tmpl = {k: torch.from_numpy(v).to(DEV) for k, v in tmpl_np.items()}
    N, thr = 40, 2.0                                    # random-init sigma ~ 5 inside the body: the surface is the valid region's boundary
    with torch.no_grad():
        net.setup_frame(posed, tmpl, None)
        verts, faces = mesh.extract_mesh(net, N=N, sigma_threshold=thr)
        center = (net.verts.max(dim=1)[0] + net.verts.min(dim=1)[0]) / 2.0
        sig = query_density_grid(net, N, center=center).reshape(N, N, N)
    assert faces.shape[0] > 1000
    vo, fo = mo.marching_cubes((thr - sig).cpu().numpy(), 0.0)
    assert np.array_equal(faces.cpu().numpy(), fo)
    want = mesh.mcubes_to_world(torch.from_numpy(vo), N, (-1.2, 1.2), (-1.2, 1.2), (-1.2, 1.2)) + center.cpu().reshape(1, 3)
    np.testing.assert_allclose(verts.cpu().numpy(), want.numpy(), atol=1e-6)
    rep = mo.mesh_report(verts.cpu().numpy(), faces.cpu().numpy())
    assert rep["closed"], rep
    # the mesh hugs the posed body: every vertex within the KNN validity radius of some SMPL vertex (+ one lattice cell)
    d = torch.cdist(verts, net.verts[0]).min(dim=1)[0]
    assert float(d.max()) < 0.2 + 2.4 / N * 2, float(d.max())
