#!/usr/bin/env python
"""Inference configurations of BASELINE.json at full size, one process per GPU (torchrun) or one GPU:
  cfg3  512x512 novel-view frame (rows sharded)
  cfg4  512^3 canonical density-grid query of mesh extraction (lattice slabs sharded)
  cfg5b 1080x1080 novel-pose frames of a synthetic pose sequence (rows sharded; --frames of the 120)
No data-path collective; device-timed with CUDA events, max over ranks; rank 0 prints one JSON line.
    python tools/bench_inference.py [--grid 512] [--frames 12]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_inference.py
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch.distributed as dist
    import anim_nerf_b200  # noqa: F401
    from anim_nerf_b200 import _lib, synthetic, inference, dist_utils
    from anim_nerf_b200.system import AnimNeRFSystem
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    sysm = AnimNeRFSystem(body_model_data=synthetic.make_smpl_dict(0), n_samples=64, n_importance=64).to(dev)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):
        sd = {k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}
        getattr(sysm.anim_nerf, name).load_state_dict(sd, strict=True)
    vr, an = sysm.volume_renderer, sysm.anim_nerf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # pose sequence: interpolation between two seeded poses
    pa, tmpl_np = synthetic.make_body_params(1, seed=1)
    pb, _ = synthetic.make_body_params(1, seed=2)
    n_seq = 120
    seq = []
    for f in range(n_seq):
        a = f / (n_seq - 1)
        seq.append({k: torch.from_numpy((1 - a) * pa[k] + a * pb[k]).float().to(dev) for k in pa})
    tmpl = {k: torch.from_numpy(v).to(dev) for k, v in tmpl_np.items()}
    out = {"n_gpus": world}

    def frame_fn(H, W):
        cam = synthetic.make_camera(W, H)
        cam_d = [torch.from_numpy(cam[k])[None].to(dev) for k in ("c2w", "focal", "c")]
        n_rows = len(dist_utils.shard_rows(H, rank, world))
        host = {k: torch.empty(1, n_rows, W, c, pin_memory=True)
                for k, c in (("rgbs_fine", 3), ("alphas_fine", 1), ("depths_fine", 1))}

        def fn(i=0):
            o = inference.render_frame_sharded(vr, an, cam_d[0], cam_d[1], cam_d[2], H, W, seq[(i * 7) % n_seq], tmpl,
                                               rank=rank, world=world, gather=False)
            for k, h in host.items():
                h.copy_(o[k], non_blocking=True)
            fn.cov = o["alphas_fine"]
        return fn

    f512 = frame_fn(512, 512)
    ms = timed(f512, max(args.reps, 5))
    out["cfg3_frame_512"] = {"ms": ms, "rays_per_s": 512 * 512 / ms * 1e3}
    f1080 = frame_fn(1080, 1080)
    ms = timed(f1080, args.frames)
    out["cfg5_frame_1080"] = {"ms_per_frame": ms, "rays_per_s": 1080 * 1080 / ms * 1e3, "frames_timed": args.frames,
                              "sequence_120_frames_s": ms * 120 / 1e3,
                              "foreground_fraction_local_slab": float((f1080.cov > 0.5).float().mean())}

    # cfg4: density grid around the posed body
    an.setup_frame(seq[0], tmpl, None)
    N = args.grid
    slab = dist_utils.shard_range(N, rank, world)
    buf = torch.empty(slab[1] - slab[0], N, N, device=dev)

    def grid(i=0):
        inference.query_density_grid(an, N, slab=slab, out=buf)
    ms = timed(grid, args.reps)
    occ = float((buf > 0).float().mean())
    out["cfg4_grid"] = {"N": N, "ms": ms, "points_per_s": N ** 3 / ms * 1e3, "occupied_fraction_local_slab": occ,
                        "note": "AnimNeRF.forward on every lattice point: KNN + unpose, MLP on the valid points "
                                "(exact culling: points farther than dis_threshold from the body have sigma = -1e5 "
                                "in the reference too), relu(sigma); lattice slabs sharded over the ranks"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
