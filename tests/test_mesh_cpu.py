"""CPU: the marching-cubes oracle (oracle/mcubes_oracle.py) and the host side of anim_nerf_b200.mesh -- the generated
triangle table against the oracle's run-time polygon construction, mesh properties on analytic fields, the lattice ->
world map and the .obj writer (PyMCubes itself is absent, see the oracle's header)."""
import numpy as np
import torch

from util import ROOT  # noqa: F401  (puts the repo on sys.path)
from oracle import mcubes_oracle as mo
from anim_nerf_b200 import mesh


def _sphere(N, c, r):
    g = np.stack(np.meshgrid(*[np.arange(N)] * 3, indexing="ij"), -1).astype(np.float32)
    return (np.linalg.norm(g - np.asarray(c, np.float32), axis=-1) - r).astype(np.float32)


def test_generated_table_matches_the_runtime_construction():
    t = mesh.tri_table()
    assert t.shape == (256, 16) and t.dtype == np.int8
    n_tri = (t >= 0).sum(1) // 3
    assert n_tri.max() == 5 and n_tri[0] == 0 and n_tri[255] == 0
    for c in range(256):
        inside = [bool((c >> i) & 1) for i in range(8)]
        want = [e for tri in mo.cell_triangles(inside) for e in tri]
        assert list(t[c][t[c] >= 0]) == want, c
        assert (t[c][len(want):] == -1).all()
        # a configuration and its complement cut the same edges
        assert set(t[c][t[c] >= 0]) == set(t[255 - c][t[255 - c] >= 0])


def test_oracle_sphere_is_a_closed_outward_surface():
    c, r = (9.3, 9.6, 10.1), 6.5
    v, f = mo.marching_cubes(_sphere(20, c, r), 0.0)
    rep = mo.mesh_report(v, f)
    assert rep["closed"] and rep["manifold"] and rep["euler"] == 2
    assert 0.97 * 4 / 3 * np.pi * r ** 3 < rep["volume"] < 4 / 3 * np.pi * r ** 3       # inscribed polyhedron, normals outward
    rad = np.linalg.norm(v - np.asarray(c, np.float32), axis=1)
    assert rad.max() <= r + 1e-4 and rad.min() > r - 0.03      # vertices on lattice edges: linear interpolation of a convex field


def test_oracle_is_closed_on_noise_and_on_several_components():
    rng = np.random.RandomState(0)
    vol = np.pad(rng.randn(10, 11, 12).astype(np.float32), 1, constant_values=5.0)      # every ambiguous configuration occurs
    rep = mo.mesh_report(*mo.marching_cubes(vol, 0.0))
    assert rep["closed"]
    two = np.minimum(_sphere(22, (6, 6, 6), 4.2), _sphere(22, (15, 15, 14), 4.7))
    rep = mo.mesh_report(*mo.marching_cubes(two, 0.0))
    assert rep["closed"] and rep["manifold"] and rep["euler"] == 4


def test_mcubes_to_world_and_obj_writer(tmp_path):
    v = torch.tensor([[0.0, 0.0, 0.0], [256.0, 128.0, 64.0], [10.0, 20.0, 30.0]])
    w = mesh.mcubes_to_world(v, 256, (-1.2, 1.2), (-1.0, 1.4), (-0.5, 0.5))
    # extract_mesh.py:37-47: x' = (ymax-ymin) v1/N + ymin ; y' = (xmax-xmin) v0/N + xmin ; z' = (zmax-zmin) v2/N + zmin
    want = np.array([[-1.0, -1.2, -0.5], [2.4 * 0.5 - 1.0, 2.4 * 1.0 - 1.2, 0.25 - 0.5],
                     [2.4 * 20 / 256 - 1.0, 2.4 * 10 / 256 - 1.2, 30 / 256 - 0.5]])
    np.testing.assert_allclose(w.numpy(), want, atol=1e-6)
    f = torch.tensor([[0, 1, 2]], dtype=torch.int32)
    path = tmp_path / "m.obj"
    mesh.export_obj(w, f, str(path))
    lines = path.read_text().splitlines()
    assert len(lines) == 4 and lines[0].startswith("v ") and lines[3] == "f 1 2 3"
    np.testing.assert_allclose([float(x) for x in lines[1].split()[1:]], want[1], atol=1e-6)


def test_marching_cubes_refuses_cpu_tensors():
    import pytest
    with pytest.raises(RuntimeError):
        mesh.marching_cubes(torch.zeros(4, 4, 4), 0.0)
