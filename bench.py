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
            rows, window = [r for _, r in self.rows], "warm-up + timed region + e2e loop (same load)"
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
def run_ours(args):
    import torch.distributed as dist
    import anim_nerf_b200  # noqa: F401
    from anim_nerf_b200 import _lib, synthetic, dist_utils
    from anim_nerf_b200.system import AnimNeRFSystem

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

    data, host, params, tmpl = build_batch(rank)
    sysm = AnimNeRFSystem(body_model_data=data, n_samples=KC, n_importance=KF).to(dev)
    for name, seed in (("nerf", 10), ("nerf_fine", 11)):      # same init on every rank
        sd = {k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}
        getattr(sysm.anim_nerf, name).load_state_dict(sd, strict=True)
    mlp_params = [p for n in ("nerf", "nerf_fine") for p in getattr(sysm.anim_nerf, n).parameters()]
    # FusedAdam was validated on one GPU (tests/test_optim_gpu.py, single-graph replay); the two-graph N>1 step keeps
    # torch's fused capturable Adam until it has had its own multi-GPU run (AN_FUSED_ADAM=1 forces it)
    fused_adam = os.environ.get("AN_FUSED_ADAM", "1" if world == 1 else "0") == "1"
    if fused_adam:
        from anim_nerf_b200.optim import FusedAdam
        opt = FusedAdam(mlp_params, lr=5e-4, eps=1e-8)       # torch.optim.Adam's update in one an_adam_step launch
    else:
        opt = torch.optim.Adam(mlp_params, lr=5e-4, eps=1e-8, fused=True, capturable=True)
    sysm.volume_renderer.device_rng = True            # graph-safe randomness (torch device generator)
    pin = {k: v.pin_memory() for k, v in host.items()}
    params_d = {k: v.to(dev) for k, v in params.items()}
    tmpl_d = {k: v.to(dev) for k, v in tmpl.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    lam = sysm.hparams.train.lambda_alphas
    n_rays = N_FRAMES * N_SIDE * N_SIDE
    mse, l1 = torch.nn.functional.mse_loss, torch.nn.functional.l1_loss

    def loss_fn(batch_dev):
        out = sysm(batch_dev["rays"], params_d, tmpl_d, perturb=1.0)
        return (mse(out["rgbs"], batch_dev["rgbs"]) + mse(out["rgbs_fine"], batch_dev["rgbs"])
                + lam * (l1(out["alphas"], batch_dev["alphas"]) + l1(out["alphas_fine"], batch_dev["alphas"])))

    def step_eager(batch_dev):
        loss = loss_fn(batch_dev)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            dist_utils.allreduce_grads(mlp_params, world)      # one flat 4.74 MB NCCL all-reduce
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    # ---- headline: the same step replayed from CUDA graphs (inputs resident in HBM)
    from anim_nerf_b200.graph_step import GraphedTrainStep
    gstep = GraphedTrainStep(loss_fn, opt, mlp_params, resident, world=world, warmup=args.warmup)
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        gstep()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        gstep()
    ev1.record()
    barrier()
    t_region1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = launches_per_step * args.steps
    step_ms = ms / args.steps

    # ---- end-to-end arm (pinned host buffers; H2D of the batch + D2H of the loss inside the timed region)
    def step_e2e():
        return float(gstep(pin).item())

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clk = clocks.stop(t_region0, t_region1)

    # ---- the reference's whole training_step (train.py:324-348): the same step plus the regularisers of
    # compute_loss (fg/bg density on 2 x 16 x 128 points, normal smoothness on 2 x 16 x 6890 points, both nets;
    # torch double backward in the reference, tangent + wgrad kernels here).  Reported beside the headline, which
    # stays the render_rays fwd+bwd step the metric names.
    full = None
    if not args.no_full_step:
        del gstep
        torch.cuda.empty_cache()
        g = torch.Generator().manual_seed(17 + rank)
        vt = sysm.anim_nerf.verts_template.detach().cpu()
        pick = torch.randint(0, vt.shape[1], (N_FRAMES, 128), generator=g)
        ctr = vt.mean(1, keepdim=True)
        surf = torch.gather(vt, 1, pick[..., None].expand(-1, -1, 3))
        reg = {"fg_points": (ctr + 0.8 * (surf - ctr)).to(dev), "bg_points": (ctr + 1.5 * (surf - ctr)).to(dev)}

        def loss_fn_full(batch_dev):
            out = sysm(batch_dev["rays"], params_d, tmpl_d, perturb=1.0)
            loss, _ = sysm.compute_loss(batch_dev["rgbs"], batch_dev["alphas"], out, fg_points=batch_dev["fg_points"],
                                        bg_points=batch_dev["bg_points"], with_regularizers=True)
            return loss
        resident_full = dict(resident, **reg)
        gfull = GraphedTrainStep(loss_fn_full, opt, mlp_params, resident_full, world=world, warmup=args.warmup)
        for _ in range(args.warmup):
            gfull()
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            gfull()
        q1.record()
        barrier()
        ms_full = q0.elapsed_time(q1)
        if world > 1:
            tf = torch.tensor([ms_full], device=dev, dtype=torch.float64)
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            ms_full = float(tf[0])
        full = {"ms_per_step": ms_full / args.steps, "rays_per_s": n_rays * world * args.steps / (ms_full * 1e-3),
                "includes": "render fwd+bwd+Adam plus compute_loss regularisers: fg/bg density (2 nets x 16 x 256 points) and "
                            "normal smoothness (2 nets x 2 x 16 x 6890 points, second order) on the kernels",
                "loss": float(gfull.loss.item())}
        del gfull
    else:
        del gstep

    # ---- BASELINE metric, second half: ms per 512x512 frame (cfg3: inference, coarse+fine, perturb=0).  The
    # frame's rows are sharded over the ranks (no data-path collective); timed through the public call
    # (camera parameters in, per-frame tables + fused ray generation + render), with the D2H read of the
    # finished rgb/alpha/depth slabs into pinned host memory inside the timed region.
    from anim_nerf_b200 import inference
    torch.cuda.empty_cache()
    FH = FW = 512
    cam = synthetic.make_camera(FW, FH)
    cam_d = [torch.from_numpy(cam[k])[None].to(dev) for k in ("c2w", "focal", "c")]
    p1 = {k: v[:1] for k, v in params_d.items()}
    t1 = {k: v[:1] for k, v in tmpl_d.items()}
    rows = dist_utils.shard_range(FH, rank, world)
    host_img = {k: torch.empty(1, rows[1] - rows[0], FW, c, pin_memory=True)
                for k, c in (("rgbs_fine", 3), ("alphas_fine", 1), ("depths_fine", 1))}

    def frame():
        out = inference.render_frame_sharded(sysm.volume_renderer, sysm.anim_nerf, cam_d[0], cam_d[1], cam_d[2], FH, FW,
                                             p1, t1, rank=rank, world=world, gather=False)
        for k, h in host_img.items():
            h.copy_(out[k], non_blocking=True)
        return out

    for _ in range(args.warmup):
        out_f = frame()
    barrier()
    n_frames_timed = max(3, min(args.steps, 10))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(n_frames_timed):
        out_f = frame()
    f1.record()
    barrier()
    ms_frame = f0.elapsed_time(f1) / n_frames_timed
    frame_cov = float((out_f["alphas_fine"] > 0.5).float().mean())

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_frame], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_frame = float(t[0]), float(t[1]), float(t[2])
        step_ms = ms / args.steps
    total_rays = n_rays * world * args.steps
    value = total_rays / (ms * 1e-3)
    e2e_value = total_rays / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (tensor-bound MLP kernels; measured peaks)
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
    if dom is not None:
        try:        # DRAM bytes per launch from the committed ncu --set full capture (tools/ncu_traffic.py)
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            tj = {}
        # mean valid points per launch over the coarse (64/ray) and fine (128/ray) pass
        pts_per_launch = (pts_coarse + pts_fine) / 2.0
        note = ("per launch = mean over the coarse (64/ray) and fine (128/ray) launch; valid (non-culled) points only: "
                "fraction coarse %.3f, fine %.3f; culling is exact (invalid samples have alpha = 0)" % (valid_frac_coarse, valid_frac_fine))
        for k in ("an_mlp_fwd", "an_mlp_bwd_dgrad", "an_mlp_bwd_wgrad"):
            if k not in per_kernel:
                continue
            t_s = per_kernel[k] * 1e-3
            common = {"kernel": k, "avg_launch_ms": per_kernel[k], "share_of_step": share[k] / step_ms,
                      "traffic": tj.get(k, {}).get("bytes_per_launch"), "traffic_unit": "B/launch, dram read+write (ncu)",
                      "traffic_source": tj.get("_source"), "valid_points_per_launch": pts_per_launch, "note": note}
            if k == "an_mlp_bwd_wgrad":
                # HBM-bound by construction: every X image (4480 B/point) and dY image (4608 B/point) is read once
                ach = pts_per_launch * WGRAD_BYTES_PER_POINT / t_s / 1e9
                rooflines[k] = dict(common, bound="hbm", achieved=ach, peak=peak_hbm, unit="GB/s", frac=ach / peak_hbm,
                                    algorithmic_bytes_per_point=WGRAD_BYTES_PER_POINT,
                                    peak_source="measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6500 GB/s",
                                    tensor_tflops=pts_per_launch * FLOP_PER_POINT_FWD / t_s / 1e12)
            else:
                ach = pts_per_launch * FLOP_PER_POINT_FWD / t_s / 1e12
                rooflines[k] = dict(common, bound="tensor", achieved=ach, peak=peak_tf, unit="TFLOP/s", frac=ach / peak_tf,
                                    algorithmic_flop_per_point=FLOP_PER_POINT_FWD, peak_source=peak_src,
                                    dense_equivalent_tflops=(n_rays * (2 * KC + KF) / 2.0) * FLOP_PER_POINT_FWD / t_s / 1e12)
        roofline = rooflines.get(dom if dom in rooflines else "an_mlp_fwd")

    line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg2: training step, 16 frames x 1024 rays, 64+64 samples, fwd+bwd+Adam, per GPU",
                       "launch": "whole step replayed from CUDA graphs (GraphedTrainStep); eager launch: %.3f ms/step" % eager_ms,
                       "optimizer": "Adam lr 5e-4 eps 1e-8: " + ("an_adam_step (FusedAdam), one launch" if fused_adam
                                                                  else "torch.optim.Adam(fused, capturable)"),
                       "rays_per_step_per_gpu": n_rays, "points_per_ray": KC + KC + KF, "perturb": 1.0,
                       "regularizers": "not in the headline step (the metric names render_rays fwd+bwd); the whole training_step with them is timed separately in full_training_step",
                       "parallelism": "dp%d (rays sharded by frame, NCCL all-reduce of MLP grads)" % world,
                       "valid_point_fraction_coarse": valid_frac_coarse, "valid_point_fraction_fine": valid_frac_fine,
                       "l2": "per-step working set (bf16 activation stash + dY scratch, > 5 GB) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in pin.values())),
                    "d2h_bytes_per_step": 4},
            "frame_512": {"ms": ms_frame, "rays_per_s": FH * FW / (ms_frame * 1e-3), "frames_timed": n_frames_timed,
                          "workload": "cfg3: 512x512 novel-view frame, inference, 64+64 samples, perturb=0, rows sharded over "
                                      "%d GPU(s); per-frame tables + ray generation + render + D2H of rgb/alpha/depth" % world,
                          "d2h_bytes_per_frame": int(sum(h.numel() * 4 for h in host_img.values())) * world,
                          "foreground_pixel_fraction": frame_cov},
            "full_training_step": full,
            "gpu_launches": int(launches),
            "kernel_ms_per_step": {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
            "clocks": clk, "roofline": roofline,
            "rooflines_mlp": {k: {kk: v[kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "traffic")}
                              for k, v in rooflines.items()}}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sample_rays=args.cpu_rays)
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
