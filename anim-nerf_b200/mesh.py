"""Mesh extraction downstream of the density-grid query (reference extract_mesh.py:152-173):
sigma lattice -> marching cubes on the device -> lattice-to-world map -> .obj.

The reference calls PyMCubes (`mcubes.marching_cubes(-sigmas, 0.)`, extract_mesh.py:165), a CPU library that is neither
under /root/reference nor installed here, so its triangle table cannot be pinned: the table below is GENERATED (closed
intersection loops per cube configuration, fan-triangulated) with the same cube / edge numbering and the same
vertex rule (linear interpolation along the lattice edge, vertices shared between cells, coordinates in lattice-index
units).  The surface is the same piecewise-linear iso-surface; how a cell's polygon is cut into triangles may differ
from PyMCubes' table.  tests/test_mesh_*.py check it against the CPU restatement (oracle/mcubes_oracle.py) and through
size-independent properties (closed 2-manifold, Euler characteristic, vertices on the iso-surface, enclosed volume)."""
import functools

import numpy as np
import torch

from . import _lib

# cube corners and edges (the numbering PyMCubes and most marching-cubes tables share)
CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
EDGES = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]])
# faces as corner cycles
FACES = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 5, 4], [3, 2, 6, 7], [0, 3, 7, 4], [1, 2, 6, 5]])
TABLE_WIDTH = 16      # up to 5 triangles + terminator (the kernels index rows of 16)


def _edge_id(a, b):
    for e, (p, q) in enumerate(EDGES):
        if (p == a and q == b) or (p == b and q == a):
            return e
    raise KeyError((a, b))


@functools.lru_cache(maxsize=None)
def tri_table():
    """(256, TABLE_WIDTH) int8: for configuration c (bit i set <=> corner i is INSIDE, value < iso) the cube-edge ids of
    its triangles, three per triangle, -1 terminated.  Construction: on every face the crossed edges are joined by
    segments (two crossed edges: one segment; four -- an ambiguous face -- : each inside corner is cut off on its own, a
    rule that depends on the face's four signs only, so the two cells sharing a face agree and the mesh is closed);
    segments chain into closed loops; every loop is oriented so that its normal points from inside to outside and cut
    into a fan from its lowest edge."""
    table = -np.ones((256, TABLE_WIDTH), np.int8)
    for c in range(256):
        inside = [(c >> i) & 1 for i in range(8)]
        nbr = {}                                    # crossed edge -> the (two) crossed edges it is joined to
        for f in FACES:
            crossed = [k for k in range(4) if inside[f[k]] != inside[f[(k + 1) % 4]]]       # edge k = (f[k], f[k+1])
            eid = [_edge_id(f[k], f[(k + 1) % 4]) for k in range(4)]
            if len(crossed) == 2:
                pairs = [(eid[crossed[0]], eid[crossed[1]])]
            elif len(crossed) == 4:                 # alternating signs: cut off every inside corner f[k] (edges k-1 and k)
                pairs = [(eid[(k - 1) % 4], eid[k]) for k in range(4) if inside[f[k]]]
            else:
                pairs = []
            for a, b in pairs:
                nbr.setdefault(a, []).append(b)
                nbr.setdefault(b, []).append(a)
        assert all(len(v) == 2 for v in nbr.values()), c
        todo, tris = sorted(nbr), []
        while todo:
            start = todo[0]
            loop, prev, cur = [start], start, nbr[start][0]
            while cur != start:
                loop.append(cur)
                a, b = nbr[cur]
                prev, cur = cur, (b if a == prev else a)
            for e in loop:
                todo.remove(e)
            # orientation: area vector of the loop (edge midpoints) against the inside -> outside directions of its edges
            mid = np.array([(CORNERS[EDGES[e][0]] + CORNERS[EDGES[e][1]]) / 2.0 for e in loop])
            area = sum(np.cross(mid[i] - mid[0], mid[i + 1] - mid[0]) for i in range(1, len(loop) - 1))
            out = np.zeros(3)
            for e in loop:
                p, q = EDGES[e]
                out += (CORNERS[q] - CORNERS[p]) * (1 if inside[p] else -1)
            assert abs(float(area @ out)) > 1e-9, (c, loop)
            if area @ out < 0:
                loop = [loop[0]] + loop[:0:-1]
            tris += [(loop[0], loop[i], loop[i + 1]) for i in range(1, len(loop) - 1)]
        flat = [e for t in tris for e in t]
        assert len(flat) < TABLE_WIDTH, (c, len(flat))
        table[c, :len(flat)] = flat
    return table


@functools.lru_cache(maxsize=None)
def _device_table(device):
    return torch.from_numpy(tri_table()).to(device)


def marching_cubes(volume, isovalue=0.0):
    """`mcubes.marching_cubes(volume, isovalue)` on the device: volume (nx,ny,nz) fp32 CUDA tensor ->
    (vertices (V,3) fp32 in lattice-index coordinates, faces (F,3) int32).  A lattice point is inside when its value is
    below the isovalue; triangle normals point from inside to outside.  Vertices are shared between cells and ordered
    by the lattice point that owns their edge (x, then y, then z edge), faces by cell -- the output does not depend on
    how the kernels were scheduled."""
    if not volume.is_cuda:
        raise RuntimeError("marching_cubes runs on the device (no CPU fallback)")
    vol = volume.contiguous().float()
    nx, ny, nz = vol.shape
    dev = vol.device
    table = _device_table(dev)
    n_sites = nx * ny * nz
    n_blocks = (n_sites + 1023) // 1024
    voff = torch.empty(n_sites, dtype=torch.int16, device=dev)
    counts = torch.empty(n_blocks, 2, dtype=torch.int32, device=dev)
    totals = torch.empty(2, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.call("an_mc_count", vol.data_ptr(), nx, ny, nz, float(isovalue), table.data_ptr(), voff.data_ptr(), counts.data_ptr(), stream)
    _lib.call("an_mc_scan", counts.data_ptr(), n_blocks, totals.data_ptr(), stream)
    n_v, n_f = (int(t) for t in totals.tolist())
    verts = torch.empty(n_v, 3, device=dev)
    faces = torch.empty(n_f, 3, dtype=torch.int32, device=dev)
    if n_v:
        _lib.call("an_mc_emit", vol.data_ptr(), nx, ny, nz, float(isovalue), table.data_ptr(), voff.data_ptr(), counts.data_ptr(),
                  totals.data_ptr(), verts.data_ptr(), faces.data_ptr(), stream)
    return verts, faces


def mcubes_to_world(vertices, N, x_range, y_range, z_range):
    """extract_mesh.py:37-47, including its axis swap (the lattice comes from numpy's 'xy' meshgrid) and its division
    by N rather than N - 1."""
    v = vertices / N
    out = torch.empty_like(v)
    out[:, 0] = (y_range[1] - y_range[0]) * v[:, 1] + y_range[0]
    out[:, 1] = (x_range[1] - x_range[0]) * v[:, 0] + x_range[0]
    out[:, 2] = (z_range[1] - z_range[0]) * v[:, 2] + z_range[0]
    return out


def extract_mesh(anim_nerf, N=256, x_range=(-1.2, 1.2), y_range=(-1.2, 1.2), z_range=(-1.2, 1.2), sigma_threshold=20.0,
                 sigmas=None):
    """extract_mesh.py:152-169 for the frame set on `anim_nerf`: relu(sigma) on the N^3 lattice around the posed body,
    minus the threshold, marching cubes of the negated field at 0, lattice -> world, plus the bounding-box centre.
    `sigmas` may carry the lattice when it has been queried already (e.g. sharded over ranks).
    Returns (vertices (V,3) fp32, faces (F,3) int32) on the device."""
    from .inference import query_density_grid
    verts = anim_nerf.verts
    center = (verts.max(dim=1)[0] + verts.min(dim=1)[0]) / 2.0
    if sigmas is None:
        sigmas = query_density_grid(anim_nerf, N, x_range, y_range, z_range, center=center)
    field = sigma_threshold - sigmas.reshape(N, N, N)          # = -(max(sigma, 0) - threshold)
    v, f = marching_cubes(field, 0.0)
    return mcubes_to_world(v, N, x_range, y_range, z_range) + center.reshape(1, 3), f


def export_obj(vertices, faces, path):
    """`mcubes.export_obj`: 'v x y z' lines, then 1-based 'f a b c' lines."""
    v = vertices.detach().cpu().numpy() if torch.is_tensor(vertices) else np.asarray(vertices)
    f = faces.detach().cpu().numpy() if torch.is_tensor(faces) else np.asarray(faces)
    with open(path, "w") as fh:
        fh.write("".join("v %s %s %s\n" % (repr(float(a)), repr(float(b)), repr(float(c))) for a, b, c in v))
        fh.write("".join("f %d %d %d\n" % (a + 1, b + 1, c + 1) for a, b, c in f))
