"""Thin functional wrappers over the C ABI (one Python function per entry point).

Tensors in, tensors out; allocation and stream selection happen here (torch = plumbing), the
arithmetic happens in the CUDA library.  Every function raises if the library is missing.
"""
import ctypes

import torch

from . import _lib
from ._lib import ptr, stream, call

K_NEIGH = 4


def _f32c(t):
    return t.contiguous().float() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


# ------------------------------------------------------------------------------ rays
def raygen(c2w, focal, center, H, W, near, far, pix=None, ginv=None):
    """A1/A2.  c2w (B,3,4), focal (B,2), center (B,2); pix (B,R,2) int32 (row,col) or None for the
    full frame; ginv (B,4,4) or None.  Returns rays (B,R,8)."""
    c2w, focal, center = _f32c(c2w), _f32c(focal), _f32c(center)
    B = c2w.shape[0]
    R = pix.shape[1] if pix is not None else H * W
    if pix is not None:
        pix = pix.contiguous().to(torch.int32)
    if ginv is not None:
        ginv = _f32c(ginv)
    rays = torch.empty(B, R, 8, device=c2w.device, dtype=torch.float32)
    call("an_raygen_fwd", ptr(c2w), ptr(focal), ptr(center), ptr(pix), ptr(ginv), B, R, H, W,
         float(near), float(far), ptr(rays), stream())
    return rays


def sample_training_rays(store, frame_ids, c2w, focal, center, ginv, n, n_fg, near=0.1, far=10.0, sel=None, seed=0):
    """SURVEY 8(f)#3 (an_sample_training_rays_fwd).  store: DeviceFrameStore.  -> rays (B,n,8), rgbs (B,n,3),
    alphas (B,n), pix (B,n,2) int32."""
    dev = store.images.device
    frame_ids = torch.as_tensor(frame_ids, dtype=torch.int32, device=dev).contiguous()
    B = frame_ids.shape[0]
    c2w, focal, center = _f32c(c2w), _f32c(focal), _f32c(center)
    if ginv is not None:
        ginv = _f32c(ginv)
    if sel is not None:
        sel = sel.to(torch.int32).contiguous()
    rays = torch.empty(B, n, 8, device=dev)
    rgbs = torch.empty(B, n, 3, device=dev)
    alphas = torch.empty(B, n, device=dev)
    pix = torch.empty(B, n, 2, device=dev, dtype=torch.int32)
    call("an_sample_training_rays_fwd", ptr(store.images), ptr(store.masks), ptr(store.fg_list), ptr(store.fg_off),
         ptr(store.bg_list), ptr(store.bg_off), ptr(frame_ids), ptr(c2w), ptr(focal), ptr(center), ptr(ginv),
         B, n, n_fg, store.H, store.W, float(near), float(far), int(store.white_bkgd), int(store.with_background),
         ptr(sel), int(seed), ptr(rays), ptr(rgbs), ptr(alphas), ptr(pix), stream())
    return rays, rgbs, alphas, pix


def sample_coarse(rays, n_coarse, perturb=0.0, noise_u=None, seed=0):
    """A3.  rays (...,8) -> z (...,Kc)."""
    rays = _f32c(rays)
    lead = rays.shape[:-1]
    n = rays.numel() // 8
    z = torch.empty(*lead, n_coarse, device=rays.device, dtype=torch.float32)
    if noise_u is not None:
        noise_u = _f32c(noise_u)
    call("an_sample_coarse_fwd", ptr(rays), n, n_coarse, float(perturb), ptr(noise_u), int(seed), ptr(z), stream())
    return z


def rays_sample(n_coarse, perturb=0.0, noise_u=None, seed=0, rays_world=None, camera=None, ginv=None):
    """A1 + A2 + A3 in one launch (an_rays_sample_fwd).  Rays come from `rays_world` (B,R,8) or from
    camera = dict(c2w (B,3,4), focal (B,2), center (B,2), H, W, near, far[, pix (B,R,2) int32 (row,col)]);
    ginv (B,4,4) takes them to the body's root frame (near/far clamp included).  -> rays_body (B,R,8), z (B,R,Kc)."""
    c2w = focal = center = pix = None
    H = W = 0
    near, far = 0.1, 10.0
    if rays_world is not None:
        rays_world = _f32c(rays_world)
        B, R = rays_world.shape[:2]
        dev = rays_world.device
    else:
        c2w, focal, center = _f32c(camera["c2w"]), _f32c(camera["focal"]), _f32c(camera["center"])
        H, W, near, far = int(camera["H"]), int(camera["W"]), float(camera.get("near", 0.1)), float(camera.get("far", 10.0))
        pix = camera.get("pix")
        if pix is not None:
            pix = pix.contiguous().to(torch.int32)
        B, R = c2w.shape[0], (pix.shape[1] if pix is not None else H * W)
        dev = c2w.device
    if ginv is not None:
        ginv = _f32c(ginv)
    if noise_u is not None:
        noise_u = _f32c(noise_u)
    rays_body = torch.empty(B, R, 8, device=dev)
    z = torch.empty(B, R, n_coarse, device=dev)
    call("an_rays_sample_fwd", ptr(c2w), ptr(focal), ptr(center), ptr(pix), ptr(rays_world), ptr(ginv), B, R, H, W, int(n_coarse),
         near, far, float(perturb), ptr(noise_u), int(seed), ptr(rays_body), ptr(z), stream())
    return rays_body, z


def rays_sample_bwd(rays_body, z, g_rays_body, g_z, rays_world=None, camera=None):
    """Gradient of `rays_sample` with respect to ginv -> (B,4,4)."""
    c2w = focal = center = pix = None
    H = W = 0
    near, far = 0.1, 10.0
    if rays_world is not None:
        rays_world = _f32c(rays_world)
    else:
        c2w, focal, center = _f32c(camera["c2w"]), _f32c(camera["focal"]), _f32c(camera["center"])
        H, W, near, far = int(camera["H"]), int(camera["W"]), float(camera.get("near", 0.1)), float(camera.get("far", 10.0))
        pix = camera.get("pix")
        if pix is not None:
            pix = pix.contiguous().to(torch.int32)
    B, R, Kc = z.shape
    g_ginv = torch.empty(B, 4, 4, device=z.device)
    call("an_rays_sample_bwd", ptr(c2w), ptr(focal), ptr(center), ptr(pix), ptr(rays_world), ptr(rays_body), ptr(z),
         ptr(_f32c(g_rays_body)), ptr(None if g_z is None else _f32c(g_z)), B, R, H, W, Kc, near, far, ptr(g_ginv), stream())
    return g_ginv


def ray_point_grad(rays, z, valid, g_xyz, g_z_comp=None, g_far_comp=None):
    """Ray-side gradients of one render pass (an_ray_point_grad): -> g_rays (B,R,8) = [sum g_x, sum z g_x, 0, g_far],
    g_z (B,R,K) = g_z_comp + g_x . d.  g_xyz is read at valid samples only (it may be uninitialised elsewhere)."""
    B, R, K = z.shape
    g_rays = torch.empty(B, R, 8, device=z.device)
    g_z = torch.empty(B, R, K, device=z.device)
    call("an_ray_point_grad", ptr(rays), ptr(z), ptr(valid), ptr(g_xyz), ptr(g_z_comp), ptr(g_far_comp), B * R, K,
         ptr(g_rays), ptr(g_z), stream())
    return g_rays, g_z


# ------------------------------------------------------------------------ per-frame tables
def _body_params(d):
    """dict(betas (B|1,10), global_orient (B,3), body_pose (B,69), transl (B,3)|None) -> B, betas (B,10), pose (B,72), transl."""
    go, bp = _f32c(d["global_orient"].detach()), _f32c(d["body_pose"].detach())
    B = max(go.shape[0], d["betas"].shape[0])
    pose = torch.cat([go.reshape(go.shape[0], -1), bp.reshape(bp.shape[0], -1)], 1).contiguous()
    betas = _f32c(d["betas"].detach())
    if betas.shape[0] != B:
        betas = betas.expand(B, -1).contiguous()
    tr = d.get("transl")
    return B, betas, pose, (_f32c(tr.detach()) if tr is not None else None)


def body_tables(model, posed, template, want_template_verts=True, want_ctx=False):
    """A16 + A2 (vertex part) + clac_ober2cano_transform in two kernels.  `model`: BodyModel (constant
    buffers); posed / template: dicts betas (B|1,10), global_orient (B,3), body_pose (B,69), transl (B,3).
    Returns verts (B,V,3) root frame, ober2cano (B,V,4,4), ginv (B,4,4), verts_template (B,V,3)
    [, ctx: what `body_tables_bwd` needs]."""
    B, betas, pose, transl = _body_params(posed)
    Bt, betas_t, pose_t, transl_t = _body_params(template)
    V = model.v_template.shape[0]
    dev = pose.device
    ws = torch.empty(_lib.load().an_body_tables_ws_bytes(B), device=dev, dtype=torch.uint8)
    verts = torch.empty(B, V, 3, device=dev)
    o2c = torch.empty(B, V, 4, 4, device=dev)
    ginv = torch.empty(B, 4, 4, device=dev)
    vt = torch.empty(B, V, 3, device=dev) if want_template_verts else None
    call("an_body_tables_fwd", ptr(betas), ptr(pose), ptr(transl), ptr(betas_t), ptr(pose_t), ptr(transl_t), B, Bt,
         ptr(model.v_template), ptr(model.shapedirs), ptr(model.posedirs), ptr(model.J_template), ptr(model.J_shapedirs),
         ptr(model.lbs_weights), ptr(model.parents_i32), V, model.J_regressor.shape[0], model.shapedirs.shape[-1], ptr(ws),
         ptr(verts), ptr(o2c), ptr(ginv), ptr(vt), stream())
    if want_ctx:
        return verts, o2c, ginv, vt, dict(model=model, betas=betas, pose=pose, transl=transl, ws=ws, ginv=ginv, B=B, V=V)
    return verts, o2c, ginv, vt


def body_tables_bwd(ctx, g_o2c, g_ginv=None):
    """Gradients of `body_tables` w.r.t. the posed body's parameters (an_body_tables_bwd): g_o2c (B,V,4,4),
    g_ginv (B,4,4) or None -> g_betas (B,10), g_pose (B,72) = [global_orient, body_pose], g_transl (B,3)."""
    model, B, V = ctx["model"], ctx["B"], ctx["V"]
    dev = ctx["pose"].device
    g_o2c = _f32c(g_o2c)
    if g_ginv is not None:
        g_ginv = _f32c(g_ginv)
    bws = torch.empty(_lib.load().an_body_tables_bwd_ws_bytes(B), device=dev, dtype=torch.uint8)
    g_betas = torch.empty(B, 10, device=dev)
    g_pose = torch.empty(B, 72, device=dev)
    g_transl = torch.empty(B, 3, device=dev) if ctx["transl"] is not None else None
    call("an_body_tables_bwd", ptr(g_o2c), ptr(g_ginv), ptr(ctx["betas"]), ptr(ctx["pose"]), ptr(ctx["transl"]), B,
         ptr(model.shapedirs), ptr(model.posedirs), ptr(model.J_template), ptr(model.J_shapedirs), ptr(model.lbs_weights),
         ptr(model.parents_i32), V, model.J_regressor.shape[0], model.shapedirs.shape[-1], ptr(ctx["ws"]), ptr(ctx["ginv"]),
         ptr(bws), ptr(g_betas), ptr(g_pose), ptr(g_transl), stream())
    return g_betas, g_pose, g_transl


# ------------------------------------------------------------------------ KNN + unpose
def vertex_grid(verts, dis_threshold, exact_flags=True):
    """Per-frame vertex grid for the pruned search.  cell = 1.25*threshold/3 (+0.1 %): the kernel's
    7^3-cell search box then covers radius >= 1.25*threshold, so 'nothing found' proves
    d_min > threshold and a 4th neighbour up to 25 % beyond the threshold needs no exhaustive rescan.
    exact_flags: cells are marked "a query in here can be valid" by the exact box-to-vertex distance (< threshold)
    instead of "a vertex in the 7^3-cell neighbourhood": ~40 % fewer coarse-pass queries reach the search."""
    verts = _f32c(verts)
    B, V = verts.shape[:2]
    nbytes = _lib.load().an_vertex_grid_bytes(B, V)
    ws = torch.empty(nbytes, device=verts.device, dtype=torch.uint8)
    call("an_vertex_grid_build", ptr(verts), B, V, float(dis_threshold) * 1.25 / 3.0 * 1.001, float(dis_threshold) if exact_flags else 0.0,
         ptr(ws), stream())
    return ws


def knn_unpose(verts, ober2cano, lbs_weights, dis_threshold, xyz=None, rays=None, z=None, grid=None,
               mode=1, want_idx=False, want_dist=False, want_qw=False, sigma=None, rgb=None, compact=False,
               seed=None, qws=None):
    """A5-A8.  Query points: xyz (B,N,3) or rays (B,R,8) + z (B,R,K).  Returns a dict with
    xyz_cano (B,N,3), valid (B,N) uint8 and the optional idx/dist/qw/cidx/count.
    seed (mode 1, rays+z): dict(src, nn (B,R,K) uint8 from `sample_fine_merge`, idx (B,R*Kc,4) int32 = the
    `idx` output of the coarse pass over the same rays; optionally that pass's xyz_cano / valid / qw, which the shared
    samples then take over unchanged): same results, far less search."""
    verts, ober2cano, lbs_weights = _f32c(verts), _f32c(ober2cano), _f32c(lbs_weights)
    B, V = verts.shape[:2]
    dev = verts.device
    if xyz is not None:
        xyz = _f32c(xyz)
        N, R, K = xyz.shape[1], 0, 0
    else:
        rays, z = _f32c(rays), _f32c(z)
        R, K = z.shape[1], z.shape[2]
        N = R * K
    out = dict(xyz_cano=torch.empty(B, N, 3, device=dev), valid=torch.empty(B, N, device=dev, dtype=torch.uint8))
    out["idx"] = torch.empty(B, N, 4, device=dev, dtype=torch.int32) if want_idx else None
    out["dist"] = torch.empty(B, N, 4, device=dev) if want_dist else None
    out["qw"] = torch.empty(B, N, 4, device=dev) if want_qw else None
    ordered = compact and ORDERED_COMPACTION
    out["cidx"] = torch.empty(B * N, device=dev, dtype=torch.int32) if compact else None
    out["count"] = (torch.empty if ordered else torch.zeros)(1, device=dev, dtype=torch.int32) if compact else None
    if mode == 1 and grid is None:
        grid = vertex_grid(verts, dis_threshold)
    if mode == 1 and qws is None:   # work list of the queries that survive the occupancy test (scratch, freed on return)
        qws = torch.empty(_lib.load().an_knn_query_ws_bytes(B, N), device=dev, dtype=torch.uint8)
    s_src = s_nn = s_idx = s_xc = s_valid = s_qw = None
    s_kc = 0
    if seed is not None and mode == 1 and xyz is None:
        s_src, s_nn, s_idx = seed["src"], seed["nn"], seed["idx"]
        s_kc = s_idx.numel() // (4 * B * R)
        assert s_src.dtype == torch.uint8 and s_nn.dtype == torch.uint8 and s_idx.dtype == torch.int32
        assert s_src.numel() == B * N and s_nn.numel() == B * N and s_idx.numel() == B * R * s_kc * 4
        # the seeding pass's own results for the shared samples (copied instead of re-blended)
        s_xc, s_valid, s_qw = seed.get("xyz_cano"), seed.get("valid"), seed.get("qw")
        if s_xc is not None:
            assert s_valid is not None and s_xc.numel() == B * R * s_kc * 3 and s_valid.numel() == B * R * s_kc
            assert s_xc.dtype == torch.float32 and s_valid.dtype == torch.uint8 and s_xc.is_contiguous()
            assert s_qw is None or (s_qw.numel() == B * R * s_kc * 4 and s_qw.dtype == torch.float32)
    call("an_knn_unpose_fwd", ptr(xyz), ptr(rays), ptr(z), B, R, K, N, ptr(verts), V, ptr(grid), ptr(qws),
         ptr(ober2cano), ptr(lbs_weights), lbs_weights.shape[1], float(dis_threshold), int(mode),
         ptr(s_src), ptr(s_nn), ptr(s_idx), int(s_kc), ptr(s_xc), ptr(s_valid), ptr(s_qw),
         ptr(out["xyz_cano"]), ptr(out["valid"]), ptr(out["idx"]), ptr(out["dist"]), ptr(out["qw"]),
         ptr(sigma), ptr(rgb), None if ordered else ptr(out["cidx"]), None if ordered else ptr(out["count"]), stream())
    if ordered:
        compact_valid(out["valid"], out["cidx"], out["count"])
    return out


ORDERED_COMPACTION = True      # valid ids in ascending order (reproducible); False: appended by the KNN kernels in scheduling order


def compact_valid(valid, cidx, count):
    """A11: cidx[:count] = ids of the non-zero flags in ascending order (an_compact_valid)."""
    ws = torch.empty(_lib.load().an_compact_ws_bytes(), device=valid.device, dtype=torch.uint8)
    call("an_compact_valid", ptr(valid), valid.numel(), ptr(cidx), ptr(count), ptr(ws), stream())


def knn_unpose_lattice(verts, ober2cano, lbs_weights, dis_threshold, x_axis, y_rows, z_axis, center, grid=None,
                       sigma=None, rgb=None):
    """A5-A8 on the lattice points (x[j], y[i], z[k]) + center (extract_mesh.py:27-35,152-156; flat index (i*nj + j)*nk + k)
    of ONE frame, generated inside the kernel.  x_axis/y_rows/z_axis: 1-D fp32 tensors, center (3,).  Returns xyz_cano
    (1,N,3), valid (1,N), cidx, count like `knn_unpose(compact=True)`."""
    verts, ober2cano, lbs_weights = _f32c(verts), _f32c(ober2cano), _f32c(lbs_weights)
    assert verts.shape[0] == 1, "lattice queries address one frame"
    V, dev = verts.shape[1], verts.device
    ni, nj, nk = y_rows.numel(), x_axis.numel(), z_axis.numel()
    N = ni * nj * nk
    lat = torch.cat([center.reshape(3).float(), torch.tensor([float(nj)], device=dev), x_axis.float(), z_axis.float(), y_rows.float()])
    out = dict(xyz_cano=torch.empty(1, N, 3, device=dev), valid=torch.empty(1, N, device=dev, dtype=torch.uint8),
               cidx=torch.empty(N, device=dev, dtype=torch.int32), count=torch.zeros(1, device=dev, dtype=torch.int32))
    if grid is None:
        grid = vertex_grid(verts, dis_threshold)
    qws = torch.empty(_lib.load().an_knn_query_ws_bytes(1, N), device=dev, dtype=torch.uint8)
    call("an_knn_unpose_lattice_fwd", ptr(lat), ni, nj, nk, ptr(verts), V, ptr(grid), ptr(qws), ptr(ober2cano), ptr(lbs_weights),
         lbs_weights.shape[1], float(dis_threshold), ptr(out["xyz_cano"]), ptr(out["valid"]), ptr(sigma), ptr(rgb),
         None if ORDERED_COMPACTION else ptr(out["cidx"]), None if ORDERED_COMPACTION else ptr(out["count"]), stream())
    if ORDERED_COMPACTION:
        compact_valid(out["valid"], out["cidx"], out["count"])
    return out


def knn_unpose_bwd(g_xyz_cano, cidx, count, idx, qw, ober2cano, xyz=None, rays=None, z=None, want_g_xyz=True, zero_g_xyz=True):
    """zero_g_xyz=False: g_xyz is left uninitialised at invalid points (for consumers that read valid points only)."""
    ober2cano = _f32c(ober2cano)
    B, V = ober2cano.shape[:2]
    if xyz is not None:
        N, R, K = xyz.shape[1], 0, 0
    else:
        R, K = z.shape[1], z.shape[2]
        N = R * K
    g_o2c = torch.zeros_like(ober2cano)
    g_xyz = (torch.zeros if zero_g_xyz else torch.empty)(B, N, 3, device=ober2cano.device) if want_g_xyz else None
    call("an_knn_unpose_bwd", ptr(g_xyz_cano), ptr(cidx), ptr(count), ptr(xyz), ptr(rays), ptr(z), B, R, K, N, V,
         ptr(idx), ptr(qw), ptr(ober2cano), ptr(g_o2c), ptr(g_xyz), stream())
    return g_o2c, g_xyz


# ------------------------------------------------------------------------------ MLP
def mlp_packed_bytes():
    return _lib.load().an_mlp_packed_bytes()


def mlp_grad_floats():
    return _lib.load().an_mlp_grad_floats()


def mlp_pack(weights, biases, packed=None):
    """weights/biases: 12 fp32 CUDA tensors in the order xyz_encoding_1..8, xyz_encoding_final,
    dir_encoding.0, sigma, rgb.0 (nn.Linear (out,in) layout).  Returns the packed uint8 buffer."""
    assert len(weights) == 12 and len(biases) == 12
    dev = weights[0].device
    if packed is None:
        packed = torch.empty(mlp_packed_bytes() + 1024, device=dev, dtype=torch.uint8)
    off = (-packed.data_ptr()) % 1024
    view = packed[off:off + mlp_packed_bytes()]
    keep = [_f32c(w.detach()) for w in weights] + [_f32c(b.detach()) for b in biases]
    wp = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in keep[:12]])
    bp = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in keep[12:]])
    call("an_mlp_pack", wp, bp, ptr(view), stream())
    return view


def mlp_fwd(packed, xyz_cano, sigma, rgb, cidx=None, count=None, n_max=None, stash=None):
    """A9-A11 over compacted ids (or all n_max points when cidx is None); writes sigma/rgb in place."""
    if n_max is None:
        n_max = xyz_cano.numel() // 3
    call("an_mlp_fwd", ptr(packed), ptr(xyz_cano), ptr(cidx), ptr(count), int(n_max), ptr(sigma), ptr(rgb),
         ptr(stash), stream())


def mlp_stash(n_max, device):
    nbytes = _lib.load().an_mlp_stash_bytes(int(n_max))
    buf = torch.empty(nbytes + 128, device=device, dtype=torch.uint8)
    off = (-buf.data_ptr()) % 128
    return buf[off:off + nbytes]


def _grad_target(g_params, dev):
    """Gradient vector the wgrad kernel accumulates into: a fresh zeroed one, or the caller's persistent buffer
    (`NeRF.attach_flat_grad`), whose head-layer scratch tail is cleared first -- the chain rule back to the two
    nn.Linear of the fused head layer adds this call's dW', db' only."""
    if g_params is None:
        return torch.zeros(mlp_grad_floats(), device=dev)
    assert g_params.numel() == mlp_grad_floats() and g_params.is_contiguous() and g_params.dtype == torch.float32
    g_params[FLAT_FLOATS:].zero_()
    return g_params


FLAT_FLOATS = 592388          # mlp_layout.cuh: parameters of one NeRF (the gradient vector's head)
_wgrad_ws = {}


def _wgrad_workspace(dev):
    """Per-device workspace of the weight-gradient kernel's per-CTA partial sums (33 MB; reused by every call on the
    device: calls are stream-ordered and each reduces its partials before it returns the stream to the next)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    ws = _wgrad_ws.get(key)
    if ws is None:
        ws = torch.empty(_lib.load().an_mlp_wgrad_ws_bytes(), device=dev, dtype=torch.uint8)
        _wgrad_ws[key] = ws
    return ws


def mlp_bwd(packed, stash, xyz_cano, rgb, g_sigma, g_rgb, cidx=None, count=None, n_max=None, want_g_xyz=True, g_params=None):
    if n_max is None:
        n_max = xyz_cano.numel() // 3
    dev = xyz_cano.device
    g_params = _grad_target(g_params, dev)
    g_xyz = torch.zeros_like(xyz_cano) if want_g_xyz else None
    nscr = _lib.load().an_mlp_bwd_scratch_bytes(int(n_max))
    scratch = torch.empty(nscr + 128, device=dev, dtype=torch.uint8)
    off = (-scratch.data_ptr()) % 128
    global _last_bwd_scratch
    _last_bwd_scratch = scratch[off:off + nscr]      # kept alive until the next call (tests inspect the dY images)
    scr = ptr(_last_bwd_scratch)
    call("an_mlp_bwd_dgrad", ptr(packed), ptr(stash), ptr(xyz_cano), ptr(rgb), ptr(cidx), ptr(count), int(n_max),
         ptr(g_sigma), ptr(g_rgb), ptr(g_xyz), scr, stream())
    call("an_mlp_bwd_wgrad", ptr(packed), ptr(stash), scr, ptr(cidx), ptr(count), int(n_max), ptr(g_params),
         ptr(_wgrad_workspace(dev)), stream())
    return g_params, g_xyz


_last_bwd_scratch = None


def mlp_bwd_scratch(n_max, device):
    nscr = _lib.load().an_mlp_bwd_scratch_bytes(int(n_max))
    buf = torch.empty(nscr + 128, device=device, dtype=torch.uint8)
    off = (-buf.data_ptr()) % 128
    return buf[off:off + nscr]


def mlp_bwd_dgrad(packed, stash, xyz_cano, rgb, g_sigma, g_rgb, scratch, cidx=None, count=None, n_max=None,
                  want_g_xyz=True):
    """Activation-gradient chain only: fills `scratch` with the dY images, returns g_xyz_cano (or None)."""
    if n_max is None:
        n_max = xyz_cano.numel() // 3
    g_xyz = torch.zeros_like(xyz_cano) if want_g_xyz else None
    call("an_mlp_bwd_dgrad", ptr(packed), ptr(stash), ptr(xyz_cano), ptr(rgb), ptr(cidx), ptr(count), int(n_max),
         ptr(g_sigma), ptr(g_rgb), ptr(g_xyz), ptr(scratch), stream())
    return g_xyz


def mlp_bwd_wgrad(packed, stash, scratch, cidx=None, count=None, n_max=None, g_params=None, bias_scale=None):
    """dW/db of every layer from the images in `stash` (X) and `scratch` (dY); accumulates into g_params.
    bias_scale (n_max, compact order): db = sum_p bias_scale[p] dY_p instead of the plain column sums."""
    g_params = _grad_target(g_params, stash.device)
    if bias_scale is None:
        call("an_mlp_bwd_wgrad", ptr(packed), ptr(stash), ptr(scratch), ptr(cidx), ptr(count), int(n_max), ptr(g_params),
             ptr(_wgrad_workspace(stash.device)), stream())
    else:
        call("an_mlp_bwd_wgrad_scaled", ptr(packed), ptr(stash), ptr(scratch), ptr(cidx), ptr(count), int(n_max),
             ptr(_f32c(bias_scale)), ptr(g_params), ptr(_wgrad_workspace(stash.device)), stream())
    return g_params


def mlp_fwd_tangent(packed, xyz_cano, tvec, pstash, cidx=None, count=None, n_max=None, want_tsigma=False, tscale=None):
    """Forward-mode tangent of the trunk (an_mlp_fwd_tangent): returns (tstash, tsigma or None).
    tscale (ids): the images become tau + tscale * X (see the header)."""
    if n_max is None:
        n_max = xyz_cano.numel() // 3
    tstash = mlp_stash(n_max, xyz_cano.device)
    tsig = torch.zeros(xyz_cano.numel() // 3, device=xyz_cano.device) if want_tsigma else None
    call("an_mlp_fwd_tangent", ptr(packed), ptr(xyz_cano), ptr(_f32c(tvec)), ptr(None if tscale is None else _f32c(tscale)),
         ptr(pstash), ptr(cidx), ptr(count), int(n_max), ptr(tsig), ptr(tstash), stream())
    return tstash, tsig


# ------------------------------------------------------------------------ compositing
def composite(sigma, rgb, z, rays, white_bkgd=True, sigma_noise=None, want_weights=True):
    """A12.  sigma (...,K), rgb (...,K,3), z (...,K), rays (...,8) -> weights, rgb (...,3), depth (...,1), acc (...,1)."""
    lead = z.shape[:-1]
    K = z.shape[-1]
    n = z.numel() // K
    dev = z.device
    w = torch.empty(*lead, K, device=dev) if want_weights else None
    rgb_o = torch.empty(*lead, 3, device=dev)
    depth = torch.empty(*lead, 1, device=dev)
    acc = torch.empty(*lead, 1, device=dev)
    call("an_composite_fwd", ptr(sigma), ptr(rgb), ptr(z), ptr(rays), ptr(sigma_noise), n, K, int(white_bkgd),
         ptr(w), ptr(rgb_o), ptr(depth), ptr(acc), stream())
    return w, rgb_o, depth, acc


def composite_bwd(sigma, rgb, z, rays, g_rgb_out, g_depth, g_acc, white_bkgd=True, sigma_noise=None):
    lead = z.shape[:-1]
    K = z.shape[-1]
    n = z.numel() // K
    dev = z.device
    g_sigma = torch.empty(*lead, K, device=dev)
    g_rgb = torch.empty(*lead, K, 3, device=dev)
    g_z = torch.empty(*lead, K, device=dev)
    g_far = torch.empty(*lead, device=dev)
    call("an_composite_bwd", ptr(sigma), ptr(rgb), ptr(z), ptr(rays), ptr(sigma_noise), n, K, int(white_bkgd),
         ptr(_f32c(g_rgb_out)), ptr(_f32c(g_depth)), ptr(_f32c(g_acc)), ptr(g_sigma), ptr(g_rgb), ptr(g_z), ptr(g_far), stream())
    return g_sigma, g_rgb, g_z, g_far


# ------------------------------------------------------------------------- resampling
def searchsorted_right(cdf, u):
    cdf, u = _f32c(cdf), _f32c(u)
    M, F = cdf.shape[-1], u.shape[-1]
    n = cdf.numel() // M
    inds = torch.empty(u.shape, device=u.device, dtype=torch.int32)
    call("an_searchsorted_right", ptr(cdf), ptr(u), n, M, F, ptr(inds), stream())
    return inds


def sample_fine_merge(weights, z_coarse, n_fine, det, u=None, seed=0, want_src=True):
    """A13/A14.  weights, z_coarse (...,Kc) -> z_fine (...,Kf), z_all (...,Kc+Kf) ascending, src uint8 (index
    into cat(z_coarse, z_fine) of every sorted entry), nn uint8 (nearest coarse sample of every sorted entry)."""
    lead = z_coarse.shape[:-1]
    Kc = z_coarse.shape[-1]
    n = z_coarse.numel() // Kc
    dev = z_coarse.device
    z_fine = torch.empty(*lead, n_fine, device=dev)
    z_all = torch.empty(*lead, Kc + n_fine, device=dev)
    src = torch.empty(*lead, Kc + n_fine, device=dev, dtype=torch.uint8) if want_src else None
    nn = torch.empty(*lead, Kc + n_fine, device=dev, dtype=torch.uint8) if want_src else None
    if u is not None:
        u = _f32c(u)
    call("an_sample_fine_merge_fwd", ptr(_f32c(weights)), ptr(_f32c(z_coarse)), ptr(u), n, Kc, n_fine, int(det), int(seed),
         ptr(z_fine), ptr(z_all), ptr(src), ptr(nn), stream())
    return z_fine, z_all, src, nn


# ------------------------------------------------------------------------------ losses
_LOSS_WS = {}


def render_loss(rgb_c, rgb_f, acc_c, acc_f, tgt_rgb, tgt_acc, lambda_alphas):
    """A18: -> (terms (5,) = mse_c, mse_f, l1_c, l1_f, total; gradients of the total w.r.t. the four inputs).
    rgb_* (...,3), acc_* (...,1) or (...); the fine pair may be None."""
    rgb_c, acc_c, tgt_rgb, tgt_acc = _f32c(rgb_c), _f32c(acc_c), _f32c(tgt_rgb), _f32c(tgt_acc)
    n = acc_c.numel()
    assert rgb_c.numel() == 3 * n and tgt_rgb.numel() == 3 * n and tgt_acc.numel() == n
    fine = rgb_f is not None
    if fine:
        rgb_f, acc_f = _f32c(rgb_f), _f32c(acc_f)
        assert rgb_f.numel() == 3 * n and acc_f.numel() == n
    dev = rgb_c.device
    terms = torch.empty(5, device=dev)
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _LOSS_WS.get(key)
    if ws is None:          # per (device, stream) scratch: the kernel leaves it zeroed for its next launch
        ws = _LOSS_WS[key] = torch.zeros(_lib.load().an_render_loss_ws_bytes(), device=dev, dtype=torch.uint8)
    g = [torch.empty_like(rgb_c), torch.empty_like(rgb_f) if fine else None, torch.empty_like(acc_c),
         torch.empty_like(acc_f) if fine else None]
    call("an_render_loss", ptr(rgb_c), ptr(rgb_f), ptr(acc_c), ptr(acc_f), ptr(tgt_rgb), ptr(tgt_acc), n, float(lambda_alphas),
         ptr(terms), ptr(ws), ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), stream())
    return terms, g
