"""TEST INFRASTRUCTURE ONLY (imported by tests/ alone; the product never does).

CPU restatement of marching cubes as the reference uses it (extract_mesh.py:165, `mcubes.marching_cubes(-sigmas, 0.)`),
for small lattices, in plain Python loops.  PARITY UNPINNED: PyMCubes (requirements.txt:9, version not pinned) is neither
under /root/reference nor installed in this image and the reference ships no mesh fixture, so the triangulation of a
cell cannot be compared with PyMCubes' table.  What is restated is the published algorithm (Lorensen & Cline 1987):
  * lattice point inside <=> value < isovalue;
  * one vertex per lattice edge whose ends straddle the isovalue, at linear interpolation, shared by the cells around it,
    in lattice-index coordinates (what mcubes returns and extract_mesh.py:37-47 maps to the world);
  * per cell the closed polygons through its crossed edges (ambiguous faces: every inside corner is cut off on its own),
    oriented from inside to outside and fan-triangulated from their lowest edge.
It builds every cell's polygons at run time (no table), which makes it an independent check of the generated table in
anim-nerf_b200/mesh.py.  Output order: vertices by owning lattice point (last axis fastest), x / y / z edge; faces by cell."""
import numpy as np

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]
_EDGE_ID = {frozenset(e): i for i, e in enumerate(EDGES)}


def cell_triangles(inside):
    """Triangles (as triples of cube-edge ids) of a cell whose corners have the given inside flags."""
    nbr = {}
    for f in FACES:
        eid = [_EDGE_ID[frozenset((f[k], f[(k + 1) % 4]))] for k in range(4)]
        crossed = [k for k in range(4) if inside[f[k]] != inside[f[(k + 1) % 4]]]
        if len(crossed) == 2:
            pairs = [(eid[crossed[0]], eid[crossed[1]])]
        elif len(crossed) == 4:
            pairs = [(eid[(k - 1) % 4], eid[k]) for k in range(4) if inside[f[k]]]
        else:
            pairs = []
        for a, b in pairs:
            nbr.setdefault(a, []).append(b)
            nbr.setdefault(b, []).append(a)
    todo, tris = sorted(nbr), []
    while todo:
        start = todo[0]
        loop, prev, cur = [start], start, nbr[start][0]
        while cur != start:
            loop.append(cur)
            a, b = nbr[cur]
            prev, cur = cur, (b if a == prev else a)
        todo = [e for e in todo if e not in loop]
        mid = np.array([(np.array(CORNERS[EDGES[e][0]]) + np.array(CORNERS[EDGES[e][1]])) / 2.0 for e in loop])
        area = sum(np.cross(mid[i] - mid[0], mid[i + 1] - mid[0]) for i in range(1, len(loop) - 1))
        out = np.zeros(3)
        for e in loop:
            p, q = EDGES[e]
            out += (np.array(CORNERS[q]) - np.array(CORNERS[p])) * (1 if inside[p] else -1)
        if area @ out < 0:
            loop = [loop[0]] + loop[:0:-1]
        tris += [(loop[0], loop[i], loop[i + 1]) for i in range(1, len(loop) - 1)]
    return tris


def marching_cubes(volume, isovalue=0.0):
    """(vertices (V,3) float32 in lattice-index coordinates, faces (F,3) int32)."""
    vol = np.asarray(volume, np.float32)
    iso = np.float32(isovalue)
    nx, ny, nz = vol.shape
    vid, verts = {}, []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                v0 = vol[i, j, k]
                for axis, (di, dj, dk) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
                    a, b, c = i + di, j + dj, k + dk
                    if a < nx and b < ny and c < nz and ((vol[a, b, c] < iso) != (v0 < iso)):
                        t = (iso - v0) / (vol[a, b, c] - v0)              # float32 arithmetic, as the kernel
                        p = [np.float32(i), np.float32(j), np.float32(k)]
                        p[axis] = np.float32(p[axis] + t)
                        vid[(i, j, k, axis)] = len(verts)
                        verts.append(p)
    faces = []
    for i in range(nx - 1):
        for j in range(ny - 1):
            for k in range(nz - 1):
                inside = [bool(vol[i + c[0], j + c[1], k + c[2]] < iso) for c in CORNERS]
                if all(inside) or not any(inside):
                    continue
                for tri in cell_triangles(inside):
                    f = []
                    for e in tri:
                        p, q = EDGES[e]
                        cp, cq = CORNERS[p], CORNERS[q]
                        o = (i + min(cp[0], cq[0]), j + min(cp[1], cq[1]), k + min(cp[2], cq[2]))
                        axis = [x != y for x, y in zip(cp, cq)].index(True)
                        f.append(vid[o + (axis,)])
                    faces.append(f)
    return (np.asarray(verts, np.float32).reshape(-1, 3), np.asarray(faces, np.int32).reshape(-1, 3))


def mesh_report(vertices, faces):
    """Closedness / orientation facts of a triangle mesh.  closed: every directed edge occurs as often as its reverse (the
    surface has no boundary and is consistently oriented); manifold: and each occurs exactly once (on noise a lattice
    edge can carry four triangles -- a fan diagonal lying in a cell face -- without opening the surface);
    euler: V - E + F; volume: signed, positive when the normals point outward."""
    f = np.asarray(faces, np.int64)
    de = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    m = len(vertices) + 1
    key, rev = de[:, 0] * m + de[:, 1], de[:, 1] * m + de[:, 0]
    uk, ck = np.unique(key, return_counts=True)
    ur, cr = np.unique(rev, return_counts=True)
    closed = bool(len(uk) == len(ur) and (uk == ur).all() and (ck == cr).all())
    n_edges = len(np.unique(np.sort(de, axis=1), axis=0))
    v = np.asarray(vertices, np.float64)
    vol = float(np.einsum("ij,ij->i", v[f[:, 0]], np.cross(v[f[:, 1]], v[f[:, 2]])).sum() / 6.0)
    return {"closed": closed, "manifold": bool(closed and (ck == 1).all()), "euler": int(len(v) - n_edges + len(f)), "volume": vol}
