#!/usr/bin/env python
"""KNN + unpose microbenchmark on the cfg2 training batch (16 frames x 1024 rays, 64 coarse + 128
sorted fine depths): times `an_knn_unpose_fwd` per search-kernel variant, with and without the
coarse-pass seeds, prints candidate statistics and checks every variant against the exhaustive mode.

    python tools/bench_knn.py [--variants 1,2,3,4] [--reps 10] [--frame512]
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import anim_nerf_b200  # noqa: E402,F401
from anim_nerf_b200 import _lib, ops  # noqa: E402
from anim_nerf_b200.anim_nerf import AnimNeRF  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]


def stats_of(qws):
    """QueryWs header: n_work, next_chunk, pad[2], stats[4] (u64)."""
    h = qws[:48].cpu().numpy()
    n_work = int(h[:4].view(np.uint32)[0])
    st = h[16:48].view(np.uint64)
    return n_work, int(st[0]), int(st[1]), int(st[2])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants-unused", default="1,2,3,4")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--check", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    data, host, params, tmpl = bench.build_batch(0)
    model = AnimNeRF(body_model_data=data, use_unpose=True, use_fine=True).to(dev)
    rays_w = host["rays"].to(dev).reshape(bench.N_FRAMES, -1, 8)
    with torch.no_grad():
        rays, _ = model.setup_frame({k: v.to(dev) for k, v in params.items()}, {k: v.to(dev) for k, v in tmpl.items()}, rays_w)
    rays = rays.contiguous()
    verts, o2c, lbs = model.verts.contiguous(), model.ober2cano_transform.contiguous(), model.body_model.lbs_weights
    thr = float(model.dis_threshold)
    grid = ops.vertex_grid(verts, thr)
    B, R = rays.shape[:2]
    torch.manual_seed(0)
    zc = ops.sample_coarse(rays, 64, perturb=1.0, noise_u=torch.rand(B, R, 64, device=dev))
    # a plausible coarse weight profile: weight concentrated where the coarse samples are valid
    c0 = ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=zc, grid=grid, mode=1, want_idx=True)
    w = c0["valid"].float().view(B, R, 64) * torch.rand(B, R, 64, device=dev)
    _, z_all, src, nn = ops.sample_fine_merge(w, zc, 64, det=False, u=torch.rand(B, R, 64, device=dev))
    print("valid fraction coarse %.3f" % c0["valid"].float().mean().item())
    kw = dict(grid=grid, want_idx=True, want_qw=True, compact=True)
    ref = {}
    if args.check:
        ref["c"] = ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=zc, mode=0, want_idx=True, want_dist=True)
        ref["f"] = ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=z_all, mode=0, want_idx=True, want_dist=True)
        print("valid fraction fine %.3f" % ref["f"]["valid"].float().mean().item())
        t0 = timed(lambda: ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=zc, mode=0, want_idx=True, compact=True), 3)
        print("mode 0 (exhaustive) coarse: %.3f ms" % t0)
    # candidate statistics need a variant build: AN_LIB_PATH=... AN_NVCC_EXTRA=-DAN_KNN_STATS python -m anim_nerf_b200._build
    qc = torch.empty(lib.an_knn_query_ws_bytes(B, R * 64), device=dev, dtype=torch.uint8)
    qf = torch.empty(lib.an_knn_query_ws_bytes(B, R * 128), device=dev, dtype=torch.uint8)
    out_c = ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=zc, mode=1, qws=qc, want_dist=True, **kw)
    sc = stats_of(qc)
    seed = dict(src=src, nn=nn, idx=out_c["idx"])
    out_f = ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=z_all, mode=1, qws=qf, want_dist=True, **kw)
    sf = stats_of(qf)
    out_s = ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=z_all, mode=1, qws=qf, want_dist=True, seed=seed, **kw)
    ss = stats_of(qf)
    ok = "unchecked"
    if args.check:
        ok = True
        for out, r in ((out_c, ref["c"]), (out_f, ref["f"]), (out_s, ref["f"])):
            f = out["idx"][..., 0] >= 0
            ok = ok and torch.equal(out["valid"], r["valid"]) and torch.equal(out["idx"][f], r["idx"][f]) \
                and torch.equal(out["dist"][f], r["dist"][f]) and bool((f | ~r["valid"].bool()).all())
            v = r["valid"].bool()
            ok = ok and torch.equal(out["xyz_cano"][v], r["xyz_cano"][v])
    tc = timed(lambda: ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=zc, mode=1, qws=qc, **kw), args.reps)
    tf = timed(lambda: ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=z_all, mode=1, qws=qf, **kw), args.reps)
    ts = timed(lambda: ops.knn_unpose(verts, o2c, lbs, thr, rays=rays, z=z_all, mode=1, qws=qf, seed=seed, **kw), args.reps)
    for name, t, s, nq in (("coarse", tc, sc, B * R * 64), ("fine", tf, sf, B * R * 128), ("fine+seed", ts, ss, B * R * 128)):
        n_work, n_cand, n_iter, n_redo = s
        print("%-9s %.3f ms  queries %d searched %d (%.3f)  cand/searched %.1f  lane-eff %.2f  redo %d  exact=%s"
              % (name, t, nq, n_work, n_work / nq, n_cand / max(n_work, 1), n_cand / max(n_iter, 1), n_redo, ok))


if __name__ == "__main__":
    main()
