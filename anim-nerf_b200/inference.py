"""Inference entry points of the rendering path: full-frame novel-view / novel-pose rendering
(BASELINE cfg3 / cfg5) and the canonical-space density-grid query of mesh extraction (cfg4).

Mirrors the reference's helper functions -- `novel_view.py:75-116` / `novel_pose.py`
(`batched_inference(volume_renderer, anim_nerf, rays, params, params_template, latent_code, P, chunk)`),
`extract_mesh.py:27-35` (`create_grid`) and `:49-61` (`batched_inference(anim_nerf, points, chunk)`)
-- on the sm_100a kernels.  The reference chunks rays (2048) and points (65 536) to bound the
memory of its materialised gathers; the fused kernels keep 45 B per point, so `chunk` here only
bounds the per-launch working set (default: a whole 512x512 frame / 16 Mi grid points at once).

Multi-GPU: rays and grid points are independent given the per-frame tables, so a
frame's rows are dealt out round-robin (rank r: rows r, r+world, ...: equal foreground share per rank)
and a grid is cut into slabs along its first axis, one per rank; every rank rebuilds the (tiny) per-frame tables itself; there is no data-path collective --
only an optional all_gather of the finished slabs.
"""
from collections import defaultdict

import numpy as np
import torch

from . import ops
from .anim_nerf import batch_transform
from .dist_utils import gather_rows, gather_slabs, shard_range, shard_rows


def _rank_world(rank, world):
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return rank, world


_PIX_CACHE = {}


# ------------------------------------------------------------------ full-frame rendering
@torch.no_grad()
def batched_inference(volume_renderer, anim_nerf, rays, body_model_params, body_model_params_template,
                      latent_code=None, P=None, chunk=None):
    """reference novel_view.py:75-116.  rays (B,R,8) world space -> dict of (B,R,.) outputs.
    `P` (B,1|R,4,4): extra rigid transform of the rays in body space (the novel-view turntable)."""
    _, ginv = anim_nerf.setup_frame(body_model_params, body_model_params_template, None)
    if latent_code is not None:
        anim_nerf.set_latent_code(latent_code)
    if P is None and chunk is None:      # body-space transform + stratified sampling fused into the renderer's front end
        return volume_renderer(anim_nerf, rays, perturb=0.0, ginv=ginv)
    rays_body, _ = ops.rays_sample(volume_renderer.n_coarse, 0.0, rays_world=rays, ginv=ginv)     # same kernel, same bits
    return _render_body_space(volume_renderer, anim_nerf, rays_body, P, chunk)


def _render_body_space(volume_renderer, anim_nerf, rays, P=None, chunk=None):
    if P is not None:
        rays = rays.clone()
        rays[:, :, 0:3] = batch_transform(P, rays[:, :, 0:3], True)
        rays[:, :, 3:6] = batch_transform(P, rays[:, :, 3:6], False)
    n_rays = rays.shape[1]
    chunk = chunk or n_rays
    results = defaultdict(list)
    for i in range(0, n_rays, chunk):
        out = volume_renderer(anim_nerf, rays[:, i:i + chunk, :], perturb=0.0)
        for k, v in out.items():
            results[k].append(v)
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 1)) for k, v in results.items()}


@torch.no_grad()
def render_frame(volume_renderer, anim_nerf, c2w, focal, center, H, W, body_model_params,
                 body_model_params_template, near=0.1, far=10.0, P=None, rows=None, chunk=None):
    """One full frame per batch entry from camera parameters: the ray generation of
    `datasets/anim_nerf_dataset.py:56-85` and the ray part of `convert_to_body_model_space`
    (`models/anim_nerf.py:128-137`) run fused in `an_raygen_fwd`, then the render path.
    c2w (B,3,4), focal (B,2), center (B,2).  rows: (r0,r1) renders only that slab of image rows, a list / 1-D tensor
    of row indices only those rows (in that order).  Returns dict of (B, rows, W, .)."""
    _, ginv = anim_nerf.setup_frame(body_model_params, body_model_params_template, None)
    B = c2w.shape[0]
    if rows is None or (isinstance(rows, tuple) and tuple(rows) == (0, H)):
        pix, n_rows = None, H
    else:       # explicit pixel list (row, col); cached per (rows, W, B, device): a sequence renders the same rows every frame
        key = (rows if isinstance(rows, tuple) else tuple(int(r) for r in rows), W, B, str(c2w.device))
        pix = _PIX_CACHE.get(key)
        if pix is None:
            if isinstance(rows, tuple):
                rr = torch.arange(rows[0], rows[1], device=c2w.device, dtype=torch.int32)
            else:
                rr = torch.as_tensor(rows, device=c2w.device, dtype=torch.int32)
            cc = torch.arange(W, device=c2w.device, dtype=torch.int32)
            pix = torch.stack(torch.meshgrid(rr, cc, indexing="ij"), -1).view(1, -1, 2).expand(B, -1, -1).contiguous()
            if len(_PIX_CACHE) > 16:
                _PIX_CACHE.clear()
            _PIX_CACHE[key] = pix
        n_rows = pix.shape[1] // W
    if P is None and chunk is None:
        # ray generation + body-space transform + stratified sampling in ONE launch (an_rays_sample_fwd)
        cam = dict(c2w=c2w, focal=focal, center=center, H=H, W=W, near=near, far=far, pix=pix)
        out = volume_renderer(anim_nerf, None, perturb=0.0, ginv=ginv, camera=cam)
    else:
        rays = ops.raygen(c2w, focal, center, H, W, near, far, pix=pix, ginv=ginv)
        out = _render_body_space(volume_renderer, anim_nerf, rays, P, chunk)
    return {k: v.view(B, n_rows, W, -1) for k, v in out.items()}


@torch.no_grad()
def render_frame_sharded(volume_renderer, anim_nerf, c2w, focal, center, H, W, body_model_params,
                         body_model_params_template, rank=None, world=None, gather=True, **kw):
    """Row-sharded frame: rank r renders rows r, r + world, ... (`shard_rows`: interleaved, so every rank gets the same
    share of the body's rows); no collective on the data path.  gather=True all-gathers the finished rows, back in
    image order, so every rank returns the full frame; gather=False returns this rank's rows (B, len(rows), W, .)."""
    rank, world = _rank_world(rank, world)
    rows = shard_rows(H, rank, world) if world > 1 else None
    out = render_frame(volume_renderer, anim_nerf, c2w, focal, center, H, W, body_model_params,
                       body_model_params_template, rows=rows, **kw)
    if gather and world > 1:
        out = {k: gather_rows(v.transpose(0, 1).contiguous(), H, dim=0).transpose(0, 1) for k, v in out.items()}
    return out


# ------------------------------------------------------------------ density grid (mesh extraction)
def create_grid(N, x_range, y_range, z_range):
    """extract_mesh.py:27-35: (N,N,N,3) float64 lattice, numpy `meshgrid` 'xy' indexing:
    grid[i,j,k] = (x[j], y[i], z[k])."""
    x = np.linspace(x_range[0], x_range[1], N)
    y = np.linspace(y_range[0], y_range[1], N)
    z = np.linspace(z_range[0], z_range[1], N)
    return np.stack(np.meshgrid(x, y, z), -1)


def grid_slab_points(N, x_range, y_range, z_range, center, i0, i1, device, step=1):
    """Points of lattice rows grid[i0:i1:step] as (1, n*N*N, 3) fp32 on `device`, equal bit for bit
    to `torch.from_numpy(create_grid(...)[i0:i1:step].reshape(-1,3)).float() + center` (extract_mesh.py:152-156)
    without materialising the 3.2 GB float64 lattice on the host: the three 1-D axes come from numpy
    (same linspace rounding), the outer product is formed on the device."""
    x = torch.from_numpy(np.linspace(x_range[0], x_range[1], N)).float().to(device)
    y = torch.from_numpy(np.linspace(y_range[0], y_range[1], N)[i0:i1:step]).float().to(device)
    z = torch.from_numpy(np.linspace(z_range[0], z_range[1], N)).float().to(device)
    n = y.numel()
    pts = torch.empty(n, N, N, 3, device=device)
    pts[..., 0] = x.view(1, N, 1)
    pts[..., 1] = y.view(n, 1, 1)
    pts[..., 2] = z.view(1, 1, N)
    pts = pts.view(1, -1, 3)
    if center is not None:
        pts += center.view(1, 1, 3).to(device)
    return pts


@torch.no_grad()
def batched_point_inference(anim_nerf, points, chunk=None):
    """extract_mesh.py:49-61: relu(sigma) of `AnimNeRF.forward` on explicit points (B,N,3) -> (B,N,1)."""
    n = points.shape[1]
    chunk = chunk or n
    sig = []
    for i in range(0, n, chunk):
        _, s = anim_nerf(points[:, i:i + chunk, :], None, use_fine=anim_nerf.use_fine)
        sig.append(torch.relu_(s))
    return sig[0] if len(sig) == 1 else torch.cat(sig, 1)


@torch.no_grad()
def query_density_grid(anim_nerf, N, x_range=(-1.2, 1.2), y_range=(-1.2, 1.2), z_range=(-1.2, 1.2),
                       center=None, slab=None, slab_rows=None, out=None):
    """relu(sigma) on the N^3 lattice around the posed body (extract_mesh.py:152-160) -> (n_slab,N,N) fp32.
    Uses the per-frame state already set on `anim_nerf` (batch of one frame).  `center` defaults to the
    posed bounding-box centre (:155).  slab=(i0,i1[,step]): only lattice rows i0, i0+step, ... < i1 of the first axis;
    slab_rows bounds the rows per launch (default: 16 Mi points)."""
    verts = anim_nerf.verts
    dev = verts.device
    if center is None:
        center = (verts.max(dim=1)[0] + verts.min(dim=1)[0]) / 2.0
    i0, i1, step = (tuple(slab) + (1,))[:3] if slab is not None else (0, N, 1)
    n_out = len(range(i0, i1, step))
    slab_rows = slab_rows or max(1, (1 << 24) // (N * N))
    if out is None:
        out = torch.empty(n_out, N, N, device=dev)
    fused = anim_nerf.use_unpose and getattr(anim_nerf, "knn_mode", 1) == 1
    if fused:       # lattice points generated inside the KNN kernel: same fp32 values as grid_slab_points, nothing materialised
        ax = [torch.from_numpy(np.linspace(r[0], r[1], N)).float().to(dev) for r in (x_range, y_range, z_range)]
    for k in range(0, n_out, slab_rows):
        m = min(slab_rows, n_out - k)
        a = i0 + k * step
        if fused:
            sig = anim_nerf.lattice_sigma(ax[0], ax[1][a:a + (m - 1) * step + 1:step], ax[2], center[0], use_fine=anim_nerf.use_fine)
            torch.clamp(sig.view(m, N, N), min=0.0, out=out[k:k + m])
        else:
            pts = grid_slab_points(N, x_range, y_range, z_range, center[0], a, a + (m - 1) * step + 1, dev, step)
            out[k:k + m] = batched_point_inference(anim_nerf, pts).view(m, N, N)
    return out


@torch.no_grad()
def query_density_grid_sharded(anim_nerf, N, rank=None, world=None, gather=True, **kw):
    """Sharded grid query: rank r owns lattice rows r, r + world, ... of the first axis (interleaved: the body fills the
    middle of the lattice, contiguous slabs would leave the outer ranks idle); no data-path collective, optional
    all_gather of the finished rows back in lattice order."""
    rank, world = _rank_world(rank, world)
    out = query_density_grid(anim_nerf, N, slab=(rank, N, world), **kw)
    if gather and world > 1:
        out = gather_rows(out, N, dim=0)
    return out
