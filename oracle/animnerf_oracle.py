"""CPU oracle for the Anim-NeRF per-ray rendering hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module; the product path (`anim-nerf_b200/`) never
does and has no CPU fallback.

It is a from-scratch restatement (torch-CPU fp32 tensors so that autograd yields the
reference gradients; integer stages in numpy / `knn_oracle.c`) of the algorithm in
the reference repo JanaldoChen/Anim-NeRF, each function citing the file:line it
follows (paths relative to the reference root):

  gen_rays                 datasets/anim_nerf_dataset.py:56-85
  rays_to_body_space       models/anim_nerf.py:128-137 (ray part), batch_transform :31-39
  ober2cano_tables         models/anim_nerf.py:139-151 (vertex part + clac_ober2cano_transform)
  sample_coarse            models/volume_rendering.py:29-56
  knn                      external KNN_CUDA v0.2 (call site models/anim_nerf.py:82-83,158-159)
  unpose                   models/anim_nerf.py:153-192
  embed                    models/embedding.py:22-39
  nerf_forward             models/nerf.py:129-175
  field                    models/anim_nerf.py:290-307
  composite                models/volume_rendering.py:113-160
  searchsorted_right, sample_fine   models/volume_rendering.py:59-97
  render_rays              models/volume_rendering.py:163-232
  system_forward           train.py:189-215

Parity pinning: the reference ships no tests / golden vectors for this path.
This oracle is pinned against outputs of the reference itself, imported and run in the
build container by `tests/golden/make_golden.py` (fixtures committed under
`tests/golden/`).  The one third-party stage, KNN_CUDA (not vendored, not installable
offline), is **parity-unpinned**: the stand-in contract is fp32
d2 = ((qx-vx)^2 + (qy-vy)^2) + (qz-vz)^2 with every operation rounded (no FMA), ordered
by (d2, index) ascending, dist = sqrt(d2) -- checked against `torch.cdist(...,
compute_mode='donot_use_mm_for_euclid_dist').topk` which the golden generator injects
as the `knn_cuda` shim.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- rays
def gen_rays(c2w, H, W, focal, near, far, c=None):
    """Full-frame world-space rays (H,W,8) = [o, d, near, far].  Pixel centres are the
    integer coordinates (no +0.5); camera looks down -z, y flipped."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    if c is None:
        c = (W * 0.5, H * 0.5)
    col = torch.arange(W, dtype=torch.float32)[None, :].expand(H, W)
    row = torch.arange(H, dtype=torch.float32)[:, None].expand(H, W)
    dirs = torch.stack([(col - float(c[0])) / float(focal[0]),
                        -(row - float(c[1])) / float(focal[1]),
                        -torch.ones_like(col)], dim=-1)
    dirs = dirs / torch.linalg.vector_norm(dirs, dim=-1, keepdim=True)
    d = dirs @ c2w[:, :3].T
    o = c2w[:, 3].expand_as(d)
    one = torch.ones_like(d[..., :1])
    return torch.cat([o, d, near * one, far * one], dim=-1)


# --------------------------------------------------------------------- training-ray sampling (8(f)#3)
def _morph(mask, k, op):
    """cv2.erode / cv2.dilate with a k x k all-ones kernel, default anchor (k//2, k//2) and default border
    (out-of-image pixels are ignored), as called at datasets/anim_nerf_dataset.py:32-37:
    dst(r,c) = op over src(r+i-k//2, c+j-k//2), 0 <= i,j < k."""
    H, W = mask.shape
    a = k // 2
    fill = np.inf if op == "erode" else -np.inf
    pad = np.full((H + k - 1, W + k - 1), fill, np.float64)
    pad[a:a + H, a:a + W] = mask
    red = np.minimum if op == "erode" else np.maximum
    # separable: rows then columns
    tmp = pad[:, 0:W].copy()
    for j in range(1, k):
        tmp = red(tmp, pad[:, j:j + W])
    out = tmp[0:H].copy()
    for i in range(1, k):
        out = red(out, tmp[i:i + H])
    return out.astype(mask.dtype)


def pixel_candidate_masks(mask, fore_erode=3):
    """get_pixelcoords('foreground_pixel'), datasets/anim_nerf_dataset.py:30-37: mask (H,W) float ->
    (mask_inside > 0, mask_outside > 0) boolean maps: the silhouette eroded by fore_erode, and the band between
    its 64-px and fore_erode-px dilations."""
    inside = _morph(mask, fore_erode, "erode")
    d1 = _morph(mask, fore_erode, "dilate")
    d2 = _morph(mask, 64, "dilate")
    return inside > 0, (d2 - d1) > 0


def get_pixelcoords(mask, subsamplesize=32, fore_rate=0.9, fore_erode=3, rng=np.random):
    """datasets/anim_nerf_dataset.py:10-54, subsampletype='foreground_pixel'.  mask (H,W,1) or (H,W) float.
    Draws come from `rng.choice` in the reference's order (foreground, then background) -> coords (n,2)
    (row, col) int, plus the draw positions (for the kernel's parity mode)."""
    m = np.asarray(mask, np.float32).reshape(mask.shape[0], mask.shape[1])
    inside, outside = pixel_candidate_masks(m, fore_erode)
    n = subsamplesize * subsamplesize
    n_fg = int(n * fore_rate)
    fx, fy = np.where(inside)
    sf = rng.choice(fx.shape[0], n_fg, replace=True)
    bx, by = np.where(outside)
    sb = rng.choice(bx.shape[0], n - n_fg, replace=True)
    px = np.concatenate((fx[sf], bx[sb]), 0)
    py = np.concatenate((fy[sf], by[sb]), 0)
    return np.stack((px, py), -1).reshape(-1, 2), np.concatenate((sf, sb), 0)


def training_sample(img_u8, mask_u8, coords, c2w, focal, c, near=0.1, far=10.0, white_bkgd=True, with_background=False):
    """datasets/anim_nerf_dataset.py:200-204 (img/255, mask/255, img*mask), :244-245 (white background),
    :246-261 (full-frame gen_rays then the gathers at coords).  img_u8 (H,W,3), mask_u8 (H,W) ->
    rays (n,8), rgbs (n,3), alphas (n,1) torch fp32."""
    img = torch.from_numpy(img_u8 / 255.).float().permute(2, 0, 1)
    mask = torch.from_numpy(mask_u8 / 255.).float().unsqueeze(0)
    if not with_background:
        img = img * mask
    if white_bkgd:
        img = img * mask + (1 - mask)
    rgbs, alphas = img.permute(1, 2, 0), mask.permute(1, 2, 0)
    H, W = mask_u8.shape
    rays = gen_rays(c2w, H, W, focal, near, far, c)
    r, cc = coords[:, 0], coords[:, 1]
    return rays[r, cc], rgbs[r, cc], alphas[r, cc]


def _affine(M, v, w):
    """(M @ [v; w])[:3] for batched 4x4 M broadcast against points v (...,3)."""
    out = torch.matmul(M[..., :3, :3], v[..., None])[..., 0]
    if w:
        out = out + M[..., :3, 3]
    return out


def rays_to_body_space(rays, global_transform):
    """rays (B,R,8) world -> root-joint ("body model") frame; near/far clamped to
    [|o|-1, |o|+1].  global_transform (B,4,4) = joints_transform[:,0]."""
    Ginv = torch.inverse(global_transform)[:, None]
    o = _affine(Ginv, rays[..., 0:3], True)
    d = _affine(Ginv, rays[..., 3:6], False)
    dist = torch.linalg.vector_norm(o, dim=-1, keepdim=True)
    near = torch.maximum(rays[..., 6:7], dist - 1.0)
    far = torch.minimum(rays[..., 7:8], dist + 1.0)
    return torch.cat([o, d, near, far], dim=-1)


def ober2cano_tables(posed, template):
    """Per-frame tables in the root frame: verts (B,V,3) and the observation->canonical
    per-vertex transforms (B,V,4,4) = T_tmpl @ (T^-1 with the translation shifted by the
    shape/pose offset differences).  `posed`/`template` are body-model output dicts."""
    Ginv = torch.inverse(posed["joints_transform"][:, 0])[:, None]
    verts = _affine(Ginv, posed["vertices"], True)
    T = torch.matmul(Ginv, posed["vertices_transform"])
    Tinv = torch.inverse(T).clone()
    Tinv[..., :3, 3] = Tinv[..., :3, 3] + (template["shape_offsets"] - posed["shape_offsets"])
    Tinv[..., :3, 3] = Tinv[..., :3, 3] + (template["pose_offsets"] - posed["pose_offsets"])
    return verts, torch.matmul(template["vertices_transform"], Tinv)


# ----------------------------------------------------------------------- sampling
def sample_coarse(rays, n_coarse, perturb=0.0, noise=None):
    """Stratified depths (B,R,Kc), linear in depth between near and far*(1-1/Kc).
    `noise` = the U[0,1) tensor the reference draws with torch.rand (explicit for parity)."""
    near, far = rays[..., 6:7], rays[..., 7:8]
    t = torch.linspace(0, 1 - 1.0 / n_coarse, n_coarse).to(rays.device)
    z = near * (1 - t) + far * t
    if perturb > 0:
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        hi = torch.cat([mid, z[..., -1:]], -1)
        lo = torch.cat([z[..., :1], mid], -1)
        z = lo + (hi - lo) * (perturb * noise)
    return z


def searchsorted_right(cdf, u):
    """#{m : cdf[m] <= u} per row (numpy int64); cdf (...,M) ascending, u (...,F)."""
    cdf = np.asarray(cdf)
    u = np.asarray(u)
    return (cdf[..., None, :] <= u[..., :, None]).sum(-1).astype(np.int64)


def sample_fine(bins, weights, n_fine, det, u=None, eps=1e-5, inds=None):
    """Inverse-CDF resampling.  bins (B,R,Kc-1), weights (B,R,Kc-2) -> z_fine (B,R,Kf)
    plus the (cdf, u, inds) triple for the bit-exact index check."""
    n_bins = bins.shape[-1]
    w = weights.detach() + eps
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    if det:
        u = torch.linspace(0.0, 1.0, n_fine).to(bins.device).expand(*bins.shape[:-1], n_fine)
    u = u.contiguous()
    if inds is None:
        if cdf.is_cuda:      # CUDA tensors (bench.py's gpu_eager_baseline): the reference's own call, volume_rendering.py:85
            inds = torch.searchsorted(cdf, u, right=True)
        else:
            inds = torch.from_numpy(searchsorted_right(cdf.numpy(), u.numpy()))
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=n_bins - 1)
    c0, c1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    b0, b1 = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    den = c1 - c0
    den = torch.where(den < eps, torch.ones_like(den), den)
    z = b0 + (u - c0) / den * (b1 - b0)
    return z, cdf, u, inds


# ---------------------------------------------------------------------------- KNN
_knn_lib = None


def _load_knn_lib():
    global _knn_lib
    if _knn_lib is None:
        path = os.path.join(_HERE, "libknn_oracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.knn_oracle.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
            lib.knn_oracle.restype = None
            _knn_lib = lib
        else:
            _knn_lib = False
    return _knn_lib


def knn_numpy(verts, xyz, k=4, chunk=2048):
    """Reference contract in numpy (small inputs): returns (dist f32 (N,k), idx i32 (N,k))."""
    verts = np.ascontiguousarray(verts, np.float32)
    xyz = np.ascontiguousarray(xyz, np.float32)
    N = xyz.shape[0]
    dist = np.empty((N, k), np.float32)
    idx = np.empty((N, k), np.int32)
    for s in range(0, N, chunk):
        q = xyz[s:s + chunk]
        dx = q[:, None, 0] - verts[None, :, 0]
        dy = q[:, None, 1] - verts[None, :, 1]
        dz = q[:, None, 2] - verts[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz                      # fp32, each op rounded
        order = np.argsort(d2, axis=1, kind="stable")[:, :k]    # stable => lowest index on ties
        idx[s:s + chunk] = order
        dist[s:s + chunk] = np.sqrt(np.take_along_axis(d2, order, 1))
    return dist, idx


def knn(verts, xyz, k=4):
    """k nearest of verts (V,3) for each xyz (N,3).  Uses the C restatement
    (`oracle/knn_oracle.c`, built by `__graft_entry__.build()`) when present, numpy else."""
    lib = _load_knn_lib()
    verts = np.ascontiguousarray(verts, np.float32)
    xyz = np.ascontiguousarray(xyz, np.float32)
    if not lib:
        return knn_numpy(verts, xyz, k)
    N = xyz.shape[0]
    dist = np.empty((N, k), np.float32)
    idx = np.empty((N, k), np.int32)
    lib.knn_oracle(verts.ctypes.data, verts.shape[0], xyz.ctypes.data, N, k,
                   dist.ctypes.data, idx.ctypes.data)
    return dist, idx


# ------------------------------------------------------------------------- unpose
def knn_cdist(verts, xyz, k=4, chunk=16384):
    """The `knn_cuda` stand-in of the reference run on torch tensors of any device:
    torch.cdist(query, ref, compute_mode='donot_use_mm_for_euclid_dist').topk(k, largest=False), chunked over the
    queries.  verts (V,3), xyz (N,3) -> dist (N,k), idx (N,k) int64.  Used for CUDA tensors (bench.py's
    gpu_eager_baseline); CPU tensors go through the C / numpy restatement of the contract (`knn`)."""
    ds, is_ = [], []
    with torch.no_grad():
        for s in range(0, xyz.shape[0], chunk):
            d = torch.cdist(xyz[s:s + chunk][None], verts[None], compute_mode="donot_use_mm_for_euclid_dist")[0]
            dd, ii = d.topk(k, dim=-1, largest=False)
            ds.append(dd); is_.append(ii)
    return torch.cat(ds, 0), torch.cat(is_, 0)


# bench.py's gpu_eager_baseline can plug another k-NN in here for CUDA tensors (callable(verts (V,3), xyz (N,3), k) ->
# (dist (N,k), idx (N,k) int64)): the reference uses the external KNN_CUDA wheel at this point, not cdist + topk
KNN_OVERRIDE = None


def unpose(xyz, verts, ober2cano, lbs_weights, dis_threshold=0.2, k=4, weight_std=0.1):
    """xyz (B,N,3) body space -> (xyz_cano (B,N,3), valid (B,N,1) float, dist (B,N,k), idx (B,N,k)).
    Distances/indices are constants (the reference runs KNN under no_grad)."""
    B, N = xyz.shape[:2]
    if xyz.is_cuda:
        res = [(KNN_OVERRIDE or knn_cdist)(verts[b].detach(), xyz[b].detach(), k) for b in range(B)]
        dist_t = torch.stack([r[0] for r in res], 0)
        idx_t = torch.stack([r[1] for r in res], 0)
    else:
        dist = np.empty((B, N, k), np.float32)
        idx = np.empty((B, N, k), np.int64)
        for b in range(B):
            d, i = knn(verts[b].detach().numpy(), xyz[b].detach().numpy(), k)
            dist[b], idx[b] = d, i
        dist_t = torch.from_numpy(dist)
        idx_t = torch.from_numpy(idx)
    W = lbs_weights[idx_t]                                            # (B,N,k,24)
    l1 = torch.sum(torch.abs(W - W[..., 0:1, :]), dim=-1)
    conf = (torch.exp(-l1 / (2.0 * weight_std ** 2)) > 0.9).float()
    q = torch.exp(-dist_t) * conf
    q = q / q.sum(-1, keepdim=True)
    flat = ober2cano.reshape(B * ober2cano.shape[1], 4, 4)
    M = flat[idx_t + (torch.arange(B, device=idx_t.device) * ober2cano.shape[1])[:, None, None]]   # (B,N,k,4,4)
    That = torch.sum(q[..., None, None] * M, dim=2)
    dbar = torch.sum(q * dist_t, dim=2, keepdim=True)
    valid = (dbar < dis_threshold).float()
    xc = _affine(That, xyz, True)
    return xc, valid, dist_t, idx_t


# ---------------------------------------------------------------------------- MLP
def embed(x, n_freqs=10):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(n-1) x), cos(2^(n-1) x)] -> (...,3+6n)."""
    out = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


NERF_LAYERS = ([("xyz_encoding_%d.0" % (i + 1)) for i in range(8)]
               + ["xyz_encoding_final", "dir_encoding.0", "sigma", "rgb.0"])


def nerf_param_shapes(W=256, in_xyz=63):
    shapes = {}
    for i in range(8):
        fan_in = in_xyz if i == 0 else (W + in_xyz if i == 4 else W)
        shapes["xyz_encoding_%d.0" % (i + 1)] = (W, fan_in)
    shapes["xyz_encoding_final"] = (W, W)
    shapes["dir_encoding.0"] = (W // 2, W)
    shapes["sigma"] = (1, W)
    shapes["rgb.0"] = (3, W // 2)
    return shapes


def nerf_forward(p, xc):
    """8x256 ReLU trunk with the encoding re-injected at layer 5 (encoding first), raw
    sigma head, 256 'final' (no activation), 128 ReLU colour layer, sigmoid rgb.
    p: dict name -> (weight (out,in), bias (out,)) fp32 torch tensors."""
    e = embed(xc)
    h = e
    for i in range(8):
        if i == 4:
            h = torch.cat([e, h], -1)
        w, b = p["xyz_encoding_%d.0" % (i + 1)]
        h = torch.relu(h @ w.T + b)
    sigma = h @ p["sigma"][0].T + p["sigma"][1]
    f = h @ p["xyz_encoding_final"][0].T + p["xyz_encoding_final"][1]
    c = torch.relu(f @ p["dir_encoding.0"][0].T + p["dir_encoding.0"][1])
    rgb = torch.sigmoid(c @ p["rgb.0"][0].T + p["rgb.0"][1])
    return rgb, sigma


def nerf_sigma(p, xc):
    """NeRF.get_sigma(only_sigma=True) (models/nerf.py:155-175): raw density of canonical points (...,3) -> (...,1)."""
    e = embed(xc)
    h = e
    for i in range(8):
        if i == 4:
            h = torch.cat([e, h], -1)
        w, b = p["xyz_encoding_%d.0" % (i + 1)]
        h = torch.relu(h @ w.T + b)
    return h @ p["sigma"][0].T + p["sigma"][1]


def nerf_normal(p, xc, delta=0.02):
    """NeRF.get_normal (models/nerf.py:177-190): d alpha/d xyz with alpha = 1 - exp(-delta * relu(sigma)),
    taken by autograd with create_graph=True so that a loss on it can be differentiated w.r.t. the weights
    (double backward) -- the formulation the normal-smoothness regulariser uses (train.py:286-309)."""
    with torch.set_grad_enabled(True):
        xc = xc.detach().requires_grad_(True)
        sigma = nerf_sigma(p, xc)
        alpha = 1 - torch.exp(-delta * torch.relu(sigma))
        return torch.autograd.grad(alpha, xc, torch.ones_like(alpha), create_graph=True, retain_graph=True,
                                   only_inputs=True)[0]


def field(p, xyz, tables, dis_threshold=0.2):
    """AnimNeRF.forward: unpose -> MLP -> sigma := -1e5 where invalid."""
    verts, ober2cano, lbs_weights = tables
    xc, valid, dist, idx = unpose(xyz, verts, ober2cano, lbs_weights, dis_threshold)
    rgb, sigma = nerf_forward(p, xc)
    sigma = torch.where(valid < 1, torch.full_like(sigma, -1e5), sigma)
    return rgb, sigma, dict(xyz_cano=xc, valid=valid, dist=dist, idx=idx)


# --------------------------------------------------------------------- compositing
def composite(rgb, sigma, z, far, white_bkgd=True, sigma_noise=None):
    """rgb (B,R,K,3), sigma (B,R,K), z (B,R,K), far (B,R,1) -> weights, rgb, depth, acc."""
    if sigma_noise is not None:
        sigma = sigma + sigma_noise
    delta = torch.cat([z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)], -1)
    alpha = 1 - torch.exp(-delta * torch.relu(sigma))
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1 - alpha + 1e-10], -1), -1)
    w = alpha * trans[..., :-1]
    acc = w.sum(-1, keepdim=True)
    rgb_o = torch.sum(w[..., None] * rgb, -2)
    depth = torch.sum(w * z, -1, keepdim=True)
    if white_bkgd:
        depth = depth + (1 - acc) * far
        rgb_o = rgb_o + 1 - acc
    return w, rgb_o, depth, acc


def render_rays(p_coarse, p_fine, rays, tables, n_coarse=64, n_fine=64, perturb=0.0,
                noise=None, dis_threshold=0.2, white_bkgd=True, capture=None):
    """VolumeRenderer.forward for share_fine=False, n_fine_depth=0.
    `noise` (perturb>0): dict(coarse_u (B,R,Kc), fine_u (B,R,Kf), sigma_c (B,R,Kc), sigma_f (B,R,Kc+Kf)).
    `capture`: optional dict that receives intermediates."""
    B, R = rays.shape[:2]
    o, d, far = rays[..., None, 0:3], rays[..., None, 3:6], rays[..., 7:8]

    def pass_(p, z, sn):
        K = z.shape[-1]
        xyz = (o + z[..., None] * d).reshape(B, R * K, 3)
        rgb, sigma, aux = field(p, xyz, tables, dis_threshold)
        out = composite(rgb.reshape(B, R, K, 3), sigma.reshape(B, R, K), z, far, white_bkgd, sn)
        return out, aux, rgb, sigma

    nz = noise or {}
    zc = sample_coarse(rays, n_coarse, perturb, nz.get("coarse_u"))
    (w, rgb_c, dep_c, acc_c), aux_c, rgb_pts, sig_pts = pass_(p_coarse, zc, nz.get("sigma_c") if perturb > 0 else None)
    out = dict(rgbs=rgb_c, alphas=acc_c, depths=dep_c)
    if n_fine > 0:
        mid = 0.5 * (zc[..., :-1] + zc[..., 1:])
        zf, cdf, u, inds = sample_fine(mid, w[..., 1:-1], n_fine, det=(perturb == 0), u=nz.get("fine_u"))
        zf = zf.detach()
        z2, _ = torch.sort(torch.cat([zc, zf], -1), dim=-1)
        (w2, rgb_f, dep_f, acc_f), aux_f, _, _ = pass_(p_fine, z2, nz.get("sigma_f") if perturb > 0 else None)
        out.update(rgbs_fine=rgb_f, alphas_fine=acc_f, depths_fine=dep_f)
        if capture is not None:
            capture.update(z_fine=zf, cdf=cdf, u=u, inds=inds, z_combine=z2, weights_fine=w2, aux_fine=aux_f)
    if capture is not None:
        capture.update(z_coarse=zc, weights=w, aux_coarse=aux_c, rgb_pts=rgb_pts, sigma_pts=sig_pts)
    return out


def system_forward(p_coarse, p_fine, rays_world, posed, template, lbs_weights, **kw):
    """AnimNeRFSystem.forward without the chunk loop (chunking changes no value):
    per-frame tables, rays to body space, render."""
    verts, o2c = ober2cano_tables(posed, template)
    rays = rays_to_body_space(rays_world, posed["joints_transform"][:, 0])
    return render_rays(p_coarse, p_fine, rays, (verts, o2c, lbs_weights), **kw), rays, (verts, o2c)
