#!/usr/bin/env python
"""bench.py -- rays/s of the Anim-NeRF render_rays forward+backward (+Adam) training step.

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores

Workload (BASELINE.json configs[1], "cfg2"): one training step at the People-Snapshot
male-3-casual shape -- 16 frames x 1024 rays (32x32 pixels each), 64 coarse + 64 fine samples
(192 point queries per ray), forward + backward + Adam, synthetic frames: seeded synthetic
SMPL-topology body, random poses, random-init MLP weights (sigma bias +5), 90 % of the pixels on
the body silhouette.  A "step" = per-frame tables (torch SMPL LBS) -> rays to body space ->
coarse pass -> inverse-CDF resampling -> fine pass -> rgb MSE + 0.1*alpha L1 (coarse+fine) ->
backward -> Adam -> weight repack.  The training regularisers (fg/bg density, normal smoothness;
torch double-backward in the reference) are outside the path the metric names and
are NOT included; stated in config.

N>1: one process per GPU (torchrun), each rank owns its own 16-frame batch (weak scaling, the
global batch is N x 16 384 rays), one NCCL all-reduce of the flat MLP gradient bucket per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rays/sec render_rays fwd+bwd (64+64 samples)"
FLOP_PER_POINT_FWD = 1179904
WGRAD_BYTES_PER_POINT = 4480 + 4608   # bf16 images the weight-gradient kernel reads once: X (enc 128 + 8 x 512 + c 256) + dY (128 + 384 + 8 x 512)
ALGO_BYTES_PER_POINT = 130            # unavoidable HBM traffic per sample of the fused path (ray share + xyz + outputs)
N_FRAMES, N_SIDE, KC, KF = 16, 32, 64, 64


# ----------------------------------------------------------------------------- scene
def build_batch(rank, device=None):
    import anim_nerf_b200  # noqa: F401
    from anim_nerf_b200 import synthetic
    from anim_nerf_b200.body_model import BodyModel
    data = synthetic.make_smpl_dict(0)
    bm = BodyModel(data)
    posed_np, tmpl_np = synthetic.make_body_params(N_FRAMES, seed=1 + 100 * rank)
    with torch.no_grad():
        verts = bm(**{k: torch.from_numpy(v) for k, v in posed_np.items()})["vertices"].numpy()
    batch = synthetic.make_training_batch(verts, n_side=N_SIDE, seed=3 + rank)
    host = dict(rays=torch.from_numpy(batch["rays"]), rgbs=torch.from_numpy(batch["rgbs"]),
                alphas=torch.from_numpy(batch["alphas"]))
    params = {k: torch.from_numpy(v) for k, v in posed_np.items()}
    tmpl = {k: torch.from_numpy(v) for k, v in tmpl_np.items()}
    return data, host, params, tmpl


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Median SM clock / throttle reasons over the samples taken inside [t0, t1] (the timed region); when the
        region is shorter than the sampling period, over every sample since start() (all taken under the same load:
        the warm-up replays, the timed region and the end-to-end loop that follows it)."""
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 is not None and t0 <= ts <= t1]
        window = "timed region"
        if len(rows) < 2:
            rows, window = [r for _, r in self.rows], "warm-up + timed region + e2e loop + untimed replays of the same step (same load)"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


# ----------------------------------------------------------------------------- ours
def _load_nerfs(sysm):
    from anim_nerf_b200 import synthetic
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):      # same init on every rank
        sd = {k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}
        getattr(sysm.anim_nerf, name).load_state_dict(sd, strict=True)


def _timed_replays(step, n, barrier, dev, world):
    """n replays between two CUDA events on the current stream, barrier + synchronize on both sides; max over ranks (ms)."""
    import torch.distributed as dist
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms, t0, t1


def run_ours(args):
    import torch.distributed as dist
    import anim_nerf_b200  # noqa: F401
    from anim_nerf_b200 import _lib, synthetic, dist_utils, inference
    from anim_nerf_b200.system import AnimNeRFSystem
    from anim_nerf_b200.graph_step import GraphedTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rendering path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    data, host, params, tmpl = build_batch(rank)
    n_rays = N_FRAMES * N_SIDE * N_SIDE
    mse, l1 = torch.nn.functional.mse_loss, torch.nn.functional.l1_loss
    pin = {k: v.pin_memory() for k, v in host.items()}
    tmpl_d = {k: v.to(dev) for k, v in tmpl.items()}
    resident = {k: v.to(dev) for k, v in host.items()}

    def make_system(optim_body_params):
        """The training system in the reference's shipped configuration (config.py:34 / male-3-casual.yaml:
        optim_body_params=True: the per-frame SMPL table is optimised next to the two MLPs) or with the table frozen."""
        sysm = AnimNeRFSystem(body_model_data=data, n_samples=KC, n_importance=KF, num_frames=N_FRAMES * world,
                              optim_body_params=optim_body_params).to(dev)
        _load_nerfs(sysm)
        # the table holds the frames of every rank (rank r trains on rows [16 r, 16 r + 16)); identical on all ranks
        allp = [synthetic.make_body_params(N_FRAMES, seed=1 + 100 * r)[0] for r in range(world)]
        sysm.init_body_model_params({k: torch.from_numpy(np.concatenate([q[k] for q in allp], 0)) for k in allp[0]})
        sysm.volume_renderer.device_rng = True            # graph-safe randomness (torch device generator)
        (opt,), _ = sysm.configure_optimizers()           # FusedAdam (lr 5e-4 MLPs, 2.5e-4 SMPL table) + FlatGradBuffer
        frame_idx = torch.arange(N_FRAMES * rank, N_FRAMES * (rank + 1), device=dev)
        params_d = {k: v.to(dev) for k, v in params.items()}
        lam = sysm.hparams.train.lambda_alphas

        def loss_fn(batch_dev, regularizers=False):
            p = sysm.body_model_params(frame_idx) if optim_body_params else params_d      # train.py:330-331
            out = sysm(batch_dev["rays"], p, tmpl_d, perturb=1.0)
            if regularizers:
                return sysm.compute_loss(batch_dev["rgbs"], batch_dev["alphas"], out, fg_points=batch_dev["fg_points"],
                                         bg_points=batch_dev["bg_points"], with_regularizers=True)[0]
            return sysm.compute_loss(batch_dev["rgbs"], batch_dev["alphas"], out, with_regularizers=False)[0]     # train.py:228-262
        return sysm, opt, loss_fn

    sysm, opt, loss_fn = make_system(True)
    flat = sysm.flat_grads

    def step_eager(batch_dev):
        flat.zero()
        loss = loss_fn(batch_dev)
        loss.backward()
        flat.all_reduce(world)                            # one NCCL all-reduce of the flat gradient buffer (MLPs + SMPL table)
        opt.step()
        return loss

    # ---- eager pass: per-kernel CUDA-event timing on the launching stream + launch / valid-point counts
    from anim_nerf_b200 import autograd as _ag
    for _ in range(args.warmup):
        step_eager(resident)
    barrier()
    timing = _lib.enable_timing(True)
    _ag.COUNT_LOG = []
    launches0 = _lib.launch_count
    n_eager = max(2, min(args.steps, 5))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n_eager):
        step_eager(resident)
    ev1.record()
    barrier()
    eager_ms = ev0.elapsed_time(ev1) / n_eager
    launches_per_step = (_lib.launch_count - launches0) // n_eager
    _lib.enable_timing(False)
    per_kernel = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timing.items()}
    calls_per_step = {k: len(v) / n_eager for k, v in timing.items()}
    share = {k: per_kernel[k] * calls_per_step[k] for k in per_kernel}
    counts = _ag.COUNT_LOG
    _ag.COUNT_LOG = None
    pts_coarse = float(np.mean([c.item() for k, c in counts if k == KC]))
    pts_fine = float(np.mean([c.item() for k, c in counts if k == KC + KF]))
    valid_frac_coarse = pts_coarse / (n_rays * KC)
    valid_frac_fine = pts_fine / (n_rays * (KC + KF))

    # ---- headline: the same step replayed from ONE CUDA graph (inputs resident in HBM; the NCCL all-reduce is inside it)
    gstep = GraphedTrainStep(loss_fn, opt, None, resident, world=world, warmup=args.warmup, flat=flat,
                             renderer=sysm.volume_renderer, model=sysm.anim_nerf)
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        gstep()
    ms, t_region0, t_region1 = _timed_replays(gstep, args.steps, barrier, dev, world)
    launches = launches_per_step * args.steps
    step_ms = ms / args.steps

    # ---- end-to-end arm (pinned host buffers; H2D of the batch + D2H of the loss inside the timed region)
    # The loop a training script runs on this API: every step's batch is copied from pinned host memory (on a copy stream,
    # while the previous step computes) and every step's loss is read back on the host (one step late, so the launch of
    # the next step is not held up); both inside the timed region.
    e2e_state = {"pending": None, "losses": []}

    def step_e2e():
        res = gstep.run_staged()                 # this step: staged batch -> static buffers, replay, loss D2H started
        gstep.stage(pin)                         # H2D of the next step's batch
        if e2e_state["pending"] is not None:
            e2e_state["losses"].append(e2e_state["pending"].value())
        e2e_state["pending"] = res

    def e2e_loop(n):
        gstep.stage(pin)
        for _ in range(n):
            step_e2e()
        e2e_state["losses"].append(e2e_state["pending"].value())       # the last step's loss: the loop ends with the GPU idle
        e2e_state["pending"] = None

    e2e_loop(max(1, args.warmup // 2))
    ms_e2e, _, _ = _timed_replays(lambda: e2e_loop(args.steps), 1, barrier, dev, world)
    assert len(e2e_state["losses"]) == max(1, args.warmup // 2) + args.steps and all(np.isfinite(e2e_state["losses"]))
    # nvidia-smi needs up to a second to deliver its first sample (longer with eight of them starting at once): keep the
    # same load running (untimed replays) until a few samples exist, so that the line always carries clocks under load
    for _ in range(100):                       # <= 100 x 5 replays (~4 s); every rank runs the same count (NCCL inside the step)
        need = torch.tensor([1.0 if (clocks.proc is not None and len(clocks.rows) < 3) else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(need, op=dist.ReduceOp.MAX)
        if float(need[0]) == 0.0:
            break
        for _ in range(5):
            gstep()
        torch.cuda.synchronize()
    clk = clocks.stop(t_region0, t_region1)

    # every rank must hold the same weights after the same number of averaged updates
    in_sync = None
    if world > 1:
        h = torch.stack([torch.cat([p.detach().reshape(-1) for p in g["params"]]).double().sum() for g in opt.param_groups])
        hs = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        in_sync = bool(all(torch.equal(hs[0], x) for x in hs))

    # ---- the reference's whole training_step (train.py:324-348): the same step plus the regularisers of
    # compute_loss (fg/bg density on 2 x 16 x 128 points, normal smoothness on 2 x 16 x 6890 points, both nets;
    # torch double backward in the reference, tangent + wgrad kernels here).  Reported beside the headline, which
    # stays the render_rays fwd+bwd step the metric names.
    full = None
    del gstep
    torch.cuda.empty_cache()
    if not args.no_full_step:
        g = torch.Generator().manual_seed(17 + rank)
        with torch.no_grad():
            sysm.anim_nerf.setup_frame({k: v.to(dev) for k, v in params.items()}, tmpl_d, None)
        vt = sysm.anim_nerf.verts_template.detach().cpu()
        pick = torch.randint(0, vt.shape[1], (N_FRAMES, 128), generator=g)
        ctr = vt.mean(1, keepdim=True)
        surf = torch.gather(vt, 1, pick[..., None].expand(-1, -1, 3))
        reg = {"fg_points": (ctr + 0.8 * (surf - ctr)).to(dev), "bg_points": (ctr + 1.5 * (surf - ctr)).to(dev)}
        gfull = GraphedTrainStep(lambda b: loss_fn(b, regularizers=True), opt, None, dict(resident, **reg), world=world,
                                 warmup=args.warmup, flat=flat, renderer=sysm.volume_renderer, model=sysm.anim_nerf)
        for _ in range(args.warmup):
            gfull()
        ms_full, _, _ = _timed_replays(gfull, args.steps, barrier, dev, world)
        full = {"ms_per_step": ms_full / args.steps, "rays_per_s": n_rays * world * args.steps / (ms_full * 1e-3),
                "includes": "render fwd+bwd+Adam (MLPs + SMPL table) plus compute_loss regularisers: fg/bg density (2 nets x 16 x 256 "
                            "points) and normal smoothness (2 nets x 2 x 16 x 6890 points, second order) on the kernels",
                "loss": float(gfull.loss.item())}
        del gfull
        torch.cuda.empty_cache()

    # ---- the same render step with the SMPL table frozen (optim_body_params=False: MLP gradients only) -- what round 1
    # measured; beside the headline it prices the body-parameter path (table-builder backward, blend scatter, ray grads)
    frozen = None
    if not args.no_frozen_step:
        sys2, opt2, loss_fn2 = make_system(False)
        g2 = GraphedTrainStep(loss_fn2, opt2, None, resident, world=world, warmup=args.warmup, flat=sys2.flat_grads,
                              renderer=sys2.volume_renderer, model=sys2.anim_nerf)
        for _ in range(args.warmup):
            g2()
        ms2, _, _ = _timed_replays(g2, args.steps, barrier, dev, world)
        frozen = {"ms_per_step": ms2 / args.steps, "rays_per_s": n_rays * world * args.steps / (ms2 * 1e-3),
                  "body_params_cost": (ms / args.steps) / (ms2 / args.steps),
                  "what": "optim_body_params=False: the same step with MLP gradients only"}
        del g2, sys2, opt2, loss_fn2
        torch.cuda.empty_cache()

    # ---- BASELINE metric, second half and the inference configurations: ms per 512x512 frame (cfg3), the 512^3
    # density grid of mesh extraction (cfg4), 1080x1080 frames of the 120-pose sequence (cfg5).  Rows / lattice slabs
    # are sharded over the ranks (no data-path collective); timed through the public calls (camera parameters in,
    # per-frame tables + fused ray generation + render), with the D2H read of the finished slabs into pinned host
    # memory inside the timed region.
    vr, an = sysm.volume_renderer, sysm.anim_nerf
    pa, _ = synthetic.make_body_params(1, seed=1)
    pb, _ = synthetic.make_body_params(1, seed=2)
    n_seq = 120
    seq = [{k: torch.from_numpy((1 - f / (n_seq - 1)) * pa[k] + f / (n_seq - 1) * pb[k]).float().to(dev) for k in pa}
           for f in range(n_seq)]
    t1 = {k: v[:1] for k, v in tmpl_d.items()}

    def frame_bench(H, W, n_timed, pose_of):
        cam = synthetic.make_camera(W, H)
        cam_d = [torch.from_numpy(cam[k])[None].to(dev) for k in ("c2w", "focal", "c")]
        n_rows = len(dist_utils.shard_rows(H, rank, world))
        host_img = {k: torch.empty(1, n_rows, W, c, pin_memory=True)
                    for k, c in (("rgbs_fine", 3), ("alphas_fine", 1), ("depths_fine", 1))}
        state = {"i": 0}

        def frame():
            out = inference.render_frame_sharded(vr, an, cam_d[0], cam_d[1], cam_d[2], H, W, pose_of(state["i"]), t1,
                                                 rank=rank, world=world, gather=False)
            state["i"] += 1
            for k, h in host_img.items():
                h.copy_(out[k], non_blocking=True)
            state["out"] = out
        for _ in range(args.warmup):
            frame()
        ms_f, _, _ = _timed_replays(frame, n_timed, barrier, dev, world)
        cov = float((state["out"]["alphas_fine"] > 0.5).float().mean())
        return ms_f / n_timed, cov, int(sum(h.numel() * 4 for h in host_img.values())) * world

    p1 = {k: v[:1].to(dev) for k, v in params.items()}
    n_frames_timed = max(3, min(args.steps, 10))
    ms_frame, frame_cov, frame_bytes = frame_bench(512, 512, n_frames_timed, lambda i: p1)
    n_seq_timed = max(3, min(args.steps, 6))
    ms_1080, cov_1080, bytes_1080 = frame_bench(1080, 1080, n_seq_timed, lambda i: seq[(i * 17) % n_seq])
    an.setup_frame(seq[0], t1, None)
    NG = 512
    slab = (rank, NG, world)                         # lattice rows rank, rank + world, ... (interleaved)
    gbuf = torch.empty(len(range(*slab)), NG, NG, device=dev)

    def grid():
        inference.query_density_grid(an, NG, slab=slab, out=gbuf)
    for _ in range(2):
        grid()
    ms_grid, _, _ = _timed_replays(grid, 3, barrier, dev, world)
    ms_grid /= 3
    mesh_512 = None
    if world == 1:      # the lattice is whole on this GPU: marching cubes of it (extract_mesh.py:165), device-side
        from anim_nerf_b200 import mesh as mesh_mod
        field = 2.0 - gbuf            # random-init sigma is ~5 inside the body: threshold 2 puts the surface at the body's boundary
        mv, mf = mesh_mod.marching_cubes(field, 0.0)
        ms_mc, _, _ = _timed_replays(lambda: mesh_mod.marching_cubes(field, 0.0), 3, barrier, dev, world)
        mesh_512 = {"ms": ms_mc / 3, "vertices": int(mv.shape[0]), "faces": int(mf.shape[0]),
                    "workload": "marching cubes of the 512^3 lattice (count + scan + emit kernels, one D2H of the two totals)"}
        del field, mv, mf

    total_rays = n_rays * world * args.steps
    value = total_rays / (ms * 1e-3)
    e2e_value = total_rays / (ms_e2e * 1e-3)

    # ---- rooflines of the MLP kernels (the path is tensor-bound; useful FLOP = valid points x 1 179 904
    # per pass, fwd + dgrad + wgrad = 3 x), against the measured sustained bf16 peak; the weight-gradient kernel is
    # additionally shown against the HBM peak, which is what actually limits it in this stash-based design.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    peak_hbm = float(peaks.get("hbm_gbs", 6500.0))
    dom = max(share, key=share.get) if share else None
    roofline, rooflines = None, {}
    pts_per_launch = (pts_coarse + pts_fine) / 2.0
    try:        # DRAM bytes per launch from the committed ncu --set full capture (tools/ncu_traffic.py)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        tj = {}
    note = ("per launch = mean over the coarse (64/ray) and fine (128/ray) launch; valid (non-culled) points only: "
            "fraction coarse %.3f, fine %.3f; culling is exact (invalid samples have alpha = 0)" % (valid_frac_coarse, valid_frac_fine))
    for k in ("an_mlp_fwd", "an_mlp_bwd_dgrad", "an_mlp_bwd_wgrad"):
        if k not in per_kernel:
            continue
        t_s = per_kernel[k] * 1e-3
        ach = pts_per_launch * FLOP_PER_POINT_FWD / t_s / 1e12
        traffic = tj.get(k, {}).get("bytes_per_launch")
        rooflines[k] = {"kernel": k, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                        "avg_launch_ms": per_kernel[k], "share_of_step": share[k] / step_ms,
                        "algorithmic_flop_per_point": FLOP_PER_POINT_FWD, "peak_source": peak_src,
                        "traffic": traffic, "traffic_unit": "B/launch, dram read+write (ncu)", "traffic_source": tj.get("_source"),
                        "traffic_per_point": (traffic / pts_per_launch) if traffic else None,
                        "algorithmic_bytes_per_point": ALGO_BYTES_PER_POINT,
                        "traffic_ratio": (traffic / pts_per_launch / ALGO_BYTES_PER_POINT) if traffic else None,
                        "dense_equivalent_tflops": (n_rays * (2 * KC + KF) / 2.0) * FLOP_PER_POINT_FWD / t_s / 1e12,
                        "valid_points_per_launch": pts_per_launch, "note": note}
        if k == "an_mlp_bwd_wgrad":
            hb = pts_per_launch * WGRAD_BYTES_PER_POINT / t_s / 1e9
            rooflines[k]["hbm"] = {"achieved_gbs": hb, "peak_gbs": peak_hbm, "frac": hb / peak_hbm,
                                   "stash_bytes_per_point": WGRAD_BYTES_PER_POINT,
                                   "why": "reads every bf16 activation image X (4480 B/point) and dY image (4608 B/point) once: "
                                          "HBM-bound by the stash-based design, which is why its tensor fraction is low"}
    if dom is not None:
        roofline = rooflines.get(dom if dom in rooflines else "an_mlp_fwd")
    mlp_ms = sum(share.get(k, 0.0) for k in ("an_mlp_fwd", "an_mlp_bwd_dgrad", "an_mlp_bwd_wgrad"))
    step_tflops = 3.0 * (pts_coarse + pts_fine) * FLOP_PER_POINT_FWD / (step_ms * 1e-3) / 1e12

    line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg2: training step, 16 frames x 1024 rays, 64+64 samples, fwd+bwd+Adam, per GPU",
                       "optim_body_params": True,
                       "optim_body_params_note": "the reference's shipped setting (config.py:34, male-3-casual.yaml:26): the per-frame SMPL "
                                                 "table (betas 1x10, global_orient/body_pose/transl per frame) is optimised with the two MLPs "
                                                 "(Adam at lr/2, train.py:221-224); its gradients run through an_knn_unpose_bwd + "
                                                 "an_body_tables_bwd and ride in the same all-reduce; frozen_body_params_step times the "
                                                 "step with the table frozen",
                       "launch": "whole step (NCCL all-reduce included) replayed from one CUDA graph (GraphedTrainStep); eager launch: %.3f ms/step" % eager_ms,
                       "optimizer": "Adam lr 5e-4 (MLPs) / 2.5e-4 (SMPL table), eps 1e-8: an_adam_step (FusedAdam), one launch per group",
                       "rays_per_step_per_gpu": n_rays, "points_per_ray": KC + KC + KF, "perturb": 1.0,
                       "regularizers": "not in the headline step (the metric names render_rays fwd+bwd); the whole training_step with them is timed separately in full_training_step",
                       "parallelism": "dp%d (rays sharded by frame; one NCCL all-reduce of the flat gradient buffer: 2 MLPs + SMPL table, "
                                      "%.2f MB)" % (world, flat.buf.numel() * 4 / 1e6),
                       "reference_arm": "--impl reference times the oracle port of the reference algorithm on the host cores "
                                        "(the imported reference lives in /root/reference, which does not travel to the GPU box)",
                       "valid_point_fraction_coarse": valid_frac_coarse, "valid_point_fraction_fine": valid_frac_fine,
                       "l2": "per-step working set (bf16 activation stash + dY scratch, > 5 GB) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in pin.values())),
                    "d2h_bytes_per_step": 4,
                    "loop": "GraphedTrainStep.stage / run_staged: the next batch's H2D runs on a copy stream during the step, "
                            "each step's loss is read on the host one step late (all of them inside the timed region)"},
            "frame_512": {"ms": ms_frame, "rays_per_s": 512 * 512 / (ms_frame * 1e-3), "frames_timed": n_frames_timed,
                          "workload": "cfg3: 512x512 novel-view frame, inference, 64+64 samples, perturb=0, rows interleaved over "
                                      "%d GPU(s); per-frame tables + ray generation + render + D2H of rgb/alpha/depth" % world,
                          "d2h_bytes_per_frame": frame_bytes, "foreground_pixel_fraction": frame_cov},
            "frame_1080_seq": {"ms_per_frame": ms_1080, "rays_per_s": 1080 * 1080 / (ms_1080 * 1e-3), "frames_timed": n_seq_timed,
                               "sequence_120_frames_s": ms_1080 * 120 / 1e3,
                               "workload": "cfg5: 1080x1080 novel-pose frames of the 120-pose synthetic sequence (a different pose "
                                           "every frame: tables rebuilt per frame), rows interleaved over %d GPU(s), D2H included" % world,
                               "d2h_bytes_per_frame": bytes_1080, "foreground_pixel_fraction": cov_1080},
            "grid_512": {"ms": ms_grid, "points_per_s": NG ** 3 / (ms_grid * 1e-3),
                         "workload": "cfg4: extract_mesh density query, 512^3 lattice around the posed body, AnimNeRF.forward on "
                                     "every point (KNN + unpose, MLP on the valid ones, relu(sigma)), lattice rows interleaved over %d GPU(s)" % world},
            "mesh_512": mesh_512,
            "full_training_step": full,
            "frozen_body_params_step": frozen,
            "weights_in_sync": in_sync,
            "gpu_launches": int(launches),
            "kernel_ms_per_step": {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
            "step_useful_tflops": {"achieved": step_tflops, "frac_of_peak": step_tflops / peak_tf, "mlp_kernel_ms": mlp_ms,
                                   "what": "3 x valid points x 1 179 904 FLOP / whole step time"},
            "clocks": clk, "roofline": roofline,
            "rooflines_mlp": {k: {kk: v[kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "traffic", "traffic_ratio")}
                              for k, v in rooflines.items()}}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sample_rays=args.cpu_rays)
        if world == 1 and not args.no_gpu_eager:
            del sysm, opt
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- reference / CPU
def _oracle_step(n_rays_sample, with_grad=True, seed_offset=0, frame0=0):
    """One bounded sample of the workload through the oracle port (torch CPU + C KNN)."""
    import anim_nerf_b200  # noqa: F401
    from anim_nerf_b200 import synthetic
    from anim_nerf_b200.body_model import BodyModel
    from oracle import animnerf_oracle as oracle
    st = _oracle_step.__dict__.setdefault("state", {})
    if not st:
        data = synthetic.make_smpl_dict(0)
        st["bm"] = BodyModel(data)
        posed_np, tmpl_np = synthetic.make_body_params(N_FRAMES, seed=1)
        st["posed"] = {k: torch.from_numpy(v) for k, v in posed_np.items()}
        st["tmpl"] = {k: torch.from_numpy(v) for k, v in tmpl_np.items()}
        with torch.no_grad():
            verts = st["bm"](**st["posed"])["vertices"].numpy()
        st["batch"] = synthetic.make_training_batch(verts, n_side=N_SIDE, seed=3)
        st["p"] = []
        for seed in (10, 11):
            w = synthetic.make_nerf_weights(seed)
            st["p"].append({n: (torch.from_numpy(w[n + ".weight"]).requires_grad_(True),
                                torch.from_numpy(w[n + ".bias"]).requires_grad_(True)) for n in synthetic.NERF_LAYER_NAMES})
    bm = st["bm"]
    # sample: whole frames first (1024 rays each), then a slice of one frame
    nf = max(1, min(N_FRAMES, n_rays_sample // (N_SIDE * N_SIDE)))
    per = min(N_SIDE * N_SIDE, n_rays_sample)
    frame0 = frame0 % (N_FRAMES - nf + 1)
    sl = slice(frame0, frame0 + nf)
    posed = bm(**{k: v[sl] for k, v in st["posed"].items()})
    tmpl = bm(**{k: v[sl] for k, v in st["tmpl"].items()})
    rays = torch.from_numpy(st["batch"]["rays"][sl]).reshape(nf, -1, 8)[:, :per]
    tgt = torch.from_numpy(st["batch"]["rgbs"][sl]).reshape(nf, -1, 3)[:, :per]
    tga = torch.from_numpy(st["batch"]["alphas"][sl]).reshape(nf, -1, 1)[:, :per]
    g = torch.Generator().manual_seed(5)
    noise = dict(coarse_u=torch.rand(nf, per, KC, generator=g), fine_u=torch.rand(nf, per, KF, generator=g),
                 sigma_c=torch.randn(nf, per, KC, generator=g), sigma_f=torch.randn(nf, per, KC + KF, generator=g))
    with torch.set_grad_enabled(with_grad):
        out, _, _ = oracle.system_forward(st["p"][0], st["p"][1], rays, posed, tmpl, bm.lbs_weights,
                                          n_coarse=KC, n_fine=KF, perturb=1.0, noise=noise)
        mse, l1 = torch.nn.functional.mse_loss, torch.nn.functional.l1_loss
        loss = mse(out["rgbs"], tgt) + mse(out["rgbs_fine"], tgt) + 0.1 * (l1(out["alphas"], tga) + l1(out["alphas_fine"], tga))
        if with_grad:
            loss.backward()
    return nf * per


def cpu_baseline(sample_rays=12288):
    """Oracle port (reference algorithm, torch CPU + OpenMP C KNN) on the box's host cores: whole frames of
    the cfg2 batch (1024 rays each, one frame per call so the saved activations stay ~6 GB), 10-30 s in all."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    _oracle_step(64)                                    # warm-up (builds state)
    t0 = time.time()
    n, f = 0, 0
    while n < sample_rays and time.time() - t0 < 40.0:
        n += _oracle_step(min(1024, sample_rays - n), frame0=f)
        f += 1
    dt = time.time() - t0
    return {"value": n / dt, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": "%d rays (%d frames of the cfg2 batch, 64+64 samples, fwd+bwd, no Adam) after warm-up, %.1f s" % (n, f, dt)}


def gpu_eager_baseline(dev, n_frames=4, steps=3):
    """The "before" a user of the reference has on this GPU: the reference's algorithm as torch eager
    ops on CUDA tensors -- fp32 MLP through cuBLAS, `torch.cdist(...).topk(4)` standing in for KNN_CUDA, autograd
    backward -- through the oracle port (the reference itself lives in /root/reference and does not travel).  A bounded
    sample: `n_frames` whole frames of the cfg2 batch per step (the full 16-frame batch keeps ~40 GB of saved fp32
    activations alive), fwd+bwd, no Adam, no table-builder gradients."""
    import anim_nerf_b200  # noqa: F401
    from anim_nerf_b200 import synthetic
    from anim_nerf_b200.body_model import BodyModel
    from oracle import animnerf_oracle as oracle
    bm = BodyModel(synthetic.make_smpl_dict(0)).to(dev)
    posed_np, tmpl_np = synthetic.make_body_params(N_FRAMES, seed=1)
    with torch.no_grad():
        verts = BodyModel(synthetic.make_smpl_dict(0))(**{k: torch.from_numpy(v) for k, v in posed_np.items()})["vertices"].numpy()
    batch = synthetic.make_training_batch(verts, n_side=N_SIDE, seed=3)
    ps = []
    for seed in (10, 11):
        w = synthetic.make_nerf_weights(seed)
        ps.append({n: (torch.from_numpy(w[n + ".weight"]).to(dev).requires_grad_(True),
                       torch.from_numpy(w[n + ".bias"]).to(dev).requires_grad_(True)) for n in synthetic.NERF_LAYER_NAMES})
    sl = slice(0, n_frames)
    per = N_SIDE * N_SIDE
    rays = torch.from_numpy(batch["rays"][sl]).reshape(n_frames, -1, 8).to(dev)
    tgt = torch.from_numpy(batch["rgbs"][sl]).reshape(n_frames, -1, 3).to(dev)
    tga = torch.from_numpy(batch["alphas"][sl]).reshape(n_frames, -1, 1).to(dev)
    mse, l1 = torch.nn.functional.mse_loss, torch.nn.functional.l1_loss

    def step():
        with torch.no_grad():
            posed = bm(**{k: torch.from_numpy(v[sl]).to(dev) for k, v in posed_np.items()})
            tmpl = bm(**{k: torch.from_numpy(v[sl]).to(dev) for k, v in tmpl_np.items()})
        noise = dict(coarse_u=torch.rand(n_frames, per, KC, device=dev), fine_u=torch.rand(n_frames, per, KF, device=dev),
                     sigma_c=torch.randn(n_frames, per, KC, device=dev), sigma_f=torch.randn(n_frames, per, KC + KF, device=dev))
        out, _, _ = oracle.system_forward(ps[0], ps[1], rays, posed, tmpl, bm.lbs_weights, n_coarse=KC, n_fine=KF, perturb=1.0,
                                          noise=noise)
        loss = mse(out["rgbs"], tgt) + mse(out["rgbs_fine"], tgt) + 0.1 * (l1(out["alphas"], tga) + l1(out["alphas_fine"], tga))
        loss.backward()
        for p in ps:
            for w, b in p.values():
                w.grad = None; b.grad = None
    def timed():
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def exhaustive_knn(verts, xyz, k):
        # stand-in for the KNN_CUDA wheel (an exhaustive shared-memory-tiled search, as that package is): this library's
        # mode-0 kernel, dist / idx only -- everything else of the step stays torch eager
        from anim_nerf_b200 import ops
        V = verts.shape[0]
        eye = torch.eye(4, device=verts.device).expand(1, V, 4, 4).contiguous()
        o = ops.knn_unpose(verts[None].contiguous(), eye, bm.lbs_weights, 1e9, xyz=xyz[None].contiguous(), mode=0,
                           want_idx=True, want_dist=True)
        return o["dist"][0], o["idx"][0].long()
    n = n_frames * per
    try:
        ms = timed()
        out = {"value": n / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "kind": "port on CUDA tensors (torch eager fp32, cdist+topk KNN)",
               "sample": "%d rays/step (%d whole frames of the cfg2 batch, 64+64 samples, fwd+bwd, no Adam), %d steps after warm-up" % (n, n_frames, steps),
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    except Exception as ex:                      # noqa: BLE001  (a baseline leg must not take the bench line down)
        return {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    oracle.KNN_OVERRIDE = exhaustive_knn
    try:
        ms2 = timed()
        out["with_native_knn"] = {"value": n / (ms2 * 1e-3), "unit": "rays/s", "ms_per_step": ms2,
                                  "what": "the same torch-eager step with the k-NN done by an exhaustive CUDA kernel (this library's "
                                          "mode 0) where the reference calls the KNN_CUDA wheel: the closest stand-in for 'the "
                                          "reference on this GPU' that can be run offline"}
    except Exception as ex:                      # noqa: BLE001
        out["with_native_knn"] = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    finally:
        oracle.KNN_OVERRIDE = None
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t0 = time.time()
    n0 = _oracle_step(64)
    rate = n0 / (time.time() - t0)                      # first estimate incl. setup: conservative
    budget_s = 150.0
    per_step = int(min(1024, max(32, rate * budget_s / max(1, args.steps + args.warmup))))
    per_step = max(32, per_step // 32 * 32)
    for i in range(args.warmup):
        _oracle_step(per_step, frame0=i)
    t0 = time.time()
    n = 0
    for i in range(args.steps):
        n += _oracle_step(per_step, frame0=args.warmup + i)
    dt = time.time() - t0
    value = n / dt
    sample = "%d rays/step of cfg2 (64+64 samples, fwd+bwd, no Adam) through the oracle port of the reference algorithm" % per_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: training step, 16 frames x 1024 rays, 64+64 samples, fwd+bwd+Adam, per GPU",
                       "reference_arm": "bounded sample of that workload per step (%d rays of the cfg2 batch, fwd+bwd, no Adam) "
                                        "through the oracle port of the reference algorithm on %d host threads" % (per_step, threads)},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_OUT_FD = None


def emit(line):
    """The one JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _OUT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_OUT_FD, data)


def main():
    # stdout carries exactly one JSON line: anything else written to fd 1 by libraries (NCCL prints its version
    # banner there) is sent to stderr instead
    global _OUT_FD
    sys.stdout.flush()
    _OUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rays", type=int, default=12288)
    ap.add_argument("--no-full-step", action="store_true", help="skip the step-with-regularisers timing")
    ap.add_argument("--no-frozen-step", action="store_true", help="skip the optim_body_params=False comparison step")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-algorithm-on-CUDA-tensors baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
