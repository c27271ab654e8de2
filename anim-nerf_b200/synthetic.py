"""Synthetic SMPL-topology body template and scenes (seeded, no dataset needed).

The SMPL model files are licence-gated and not shipped with the reference
(`.gitignore: /smplx/models`), so every parity fixture and every bench run uses a
synthetic template with the *schema* the reference loader reads
(reference `smplx/body_models.py:126-252`): ``v_template (6890,3)``,
``shapedirs (6890,3,10)``, ``posedirs (6890,3,207)``, ``J_regressor (24,6890)``,
``kintree_table (2,24)``, ``weights (6890,24)``, ``f (F,3)``.

Geometry: 6890 points on capsules around a 24-joint humanoid skeleton so that
vertex spacing (~1-2 cm) and skinning-weight coherence resemble the real mesh:
the K=4 nearest vertices of a query usually share a bone, so the reference's
confidence test (`models/anim_nerf.py:165-168`) keeps more than one neighbour.

Everything is generated from ``numpy.random.RandomState(seed)`` and is
bit-reproducible across machines (float64 maths, cast to float32 at the end).
"""
import math
import os
import pickle

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24

# SMPL kinematic tree (parents); index 0 is the pelvis / root.
SMPL_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
    dtype=np.int64)

# Approximate SMPL rest-pose joint locations (metres, y up, T-pose).
_JOINTS_REST = np.array([
    [0.000, -0.240, 0.030],   # 0 pelvis
    [0.060, -0.330, 0.020],   # 1 l_hip
    [-0.060, -0.330, 0.020],  # 2 r_hip
    [0.000, -0.130, 0.000],   # 3 spine1
    [0.100, -0.710, 0.020],   # 4 l_knee
    [-0.100, -0.710, 0.020],  # 5 r_knee
    [0.000, 0.010, 0.000],    # 6 spine2
    [0.090, -1.110, -0.020],  # 7 l_ankle
    [-0.090, -1.110, -0.020], # 8 r_ankle
    [0.000, 0.060, 0.020],    # 9 spine3
    [0.110, -1.170, 0.100],   # 10 l_foot
    [-0.110, -1.170, 0.100],  # 11 r_foot
    [0.000, 0.270, -0.010],   # 12 neck
    [0.080, 0.180, 0.000],    # 13 l_collar
    [-0.080, 0.180, 0.000],   # 14 r_collar
    [0.000, 0.350, 0.030],    # 15 head
    [0.180, 0.220, -0.010],   # 16 l_shoulder
    [-0.180, 0.220, -0.010],  # 17 r_shoulder
    [0.440, 0.210, -0.030],   # 18 l_elbow
    [-0.440, 0.210, -0.030],  # 19 r_elbow
    [0.690, 0.210, -0.020],   # 20 l_wrist
    [-0.690, 0.210, -0.020],  # 21 r_wrist
    [0.780, 0.200, -0.020],   # 22 l_hand
    [-0.780, 0.200, -0.020],  # 23 r_hand
], dtype=np.float64)

# capsule radius of the bone that ends at joint j (bone = parent(j) -> j)
_BONE_RADIUS = np.array([
    0.13, 0.09, 0.09, 0.14, 0.075, 0.075, 0.15, 0.055, 0.055, 0.15, 0.045, 0.045,
    0.07, 0.08, 0.08, 0.10, 0.06, 0.06, 0.048, 0.048, 0.04, 0.04, 0.035, 0.035],
    dtype=np.float64)


def _segment_closest(p, a, b):
    """closest-point parameter/distance from points p (n,3) to segment a-b."""
    ab = b - a
    den = float(ab @ ab)
    t = np.zeros(len(p)) if den < 1e-12 else np.clip(((p - a) @ ab) / den, 0.0, 1.0)
    c = a + t[:, None] * ab
    return t, np.linalg.norm(p - c, axis=1)


def make_smpl_dict(seed=0):
    """Build the synthetic SMPL-schema dict (see module docstring)."""
    rs = np.random.RandomState(seed)
    J = _JOINTS_REST
    # bones: (parent -> child) for j>=1, plus a head blob and a pelvis blob
    bones = [(J[SMPL_PARENTS[j]], J[j], _BONE_RADIUS[j], j) for j in range(1, NUM_JOINTS)]
    bones.append((J[15], J[15] + np.array([0.0, 0.12, 0.0]), 0.10, 15))   # skull
    bones.append((J[0] + np.array([-0.08, 0.0, 0.0]), J[0] + np.array([0.08, 0.0, 0.0]), 0.13, 0))
    # area-proportional vertex budget per capsule
    areas = np.array([2 * math.pi * r * (np.linalg.norm(b - a) + 2 * r) for a, b, r, _ in bones])
    counts = np.floor(areas / areas.sum() * NUM_VERTS).astype(int)
    counts[0] += NUM_VERTS - counts.sum()
    verts = []
    for (a, b, r, _), n in zip(bones, counts):
        axis = b - a
        L = np.linalg.norm(axis)
        axis = axis / max(L, 1e-9)
        # orthonormal frame
        tmp = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
        u = np.cross(axis, tmp); u /= np.linalg.norm(u)
        v = np.cross(axis, u)
        s = rs.uniform(-r, L + r, size=n)            # along-axis coordinate incl. caps
        phi = rs.uniform(0, 2 * math.pi, size=n)
        sc = np.clip(s, 0.0, L)
        cap = s - sc                                  # signed overshoot into the cap
        rad = np.sqrt(np.maximum(r * r - cap * cap, 0.0))
        pts = (a[None] + sc[:, None] * axis[None] + cap[:, None] * axis[None]
               + rad[:, None] * (np.cos(phi)[:, None] * u[None] + np.sin(phi)[:, None] * v[None]))
        verts.append(pts)
    v_template = np.concatenate(verts, 0)
    assert v_template.shape == (NUM_VERTS, 3)
    v_template = v_template[rs.permutation(NUM_VERTS)]  # no index/space correlation

    # skinning weights: soft assignment to the 2 bones nearest to the vertex,
    # sharpened so that most vertices are dominated by one joint (rand**8-like sparsity)
    dist = np.full((NUM_VERTS, NUM_JOINTS), 1e3)
    for a, b, r, j in bones:
        jp = SMPL_PARENTS[j] if j > 0 else 0
        _, d = _segment_closest(v_template, a, b)
        # the bone parent->j is driven by the parent joint's rotation
        dist[:, jp] = np.minimum(dist[:, jp], np.maximum(d - r, 0.0) + 0.01)
    w = np.exp(-dist / 0.012)
    w[w < 2e-2 * w.max(axis=1, keepdims=True)] = 0.0
    weights = w / w.sum(axis=1, keepdims=True)
    # quantise so that neighbouring vertices frequently have *identical* rows
    weights = np.round(weights * 64.0) / 64.0
    weights = weights / weights.sum(axis=1, keepdims=True)

    # joint regressor: gaussian of distance to the rest joint, rows sum to 1
    d2 = ((v_template[None, :, :] - J[:, None, :]) ** 2).sum(-1)
    jr = np.exp(-d2 / (2 * 0.06 ** 2)) + 1e-12
    J_regressor = jr / jr.sum(axis=1, keepdims=True)

    shapedirs = rs.normal(0, 0.004, size=(NUM_VERTS, 3, 10))
    posedirs = rs.normal(0, 0.0015, size=(NUM_VERTS, 3, 207))
    kintree = np.stack([SMPL_PARENTS.copy(), np.arange(NUM_JOINTS)], 0)
    kintree[0, 0] = 0  # arbitrary: the loader forces parents[0] = -1 (body_models.py:246-247)
    faces = np.stack([np.arange(0, 300), np.arange(1, 301), np.arange(2, 302)], 1).astype(np.int64)
    return dict(
        v_template=v_template.astype(np.float32),
        shapedirs=shapedirs.astype(np.float32),
        posedirs=posedirs.astype(np.float32),
        J_regressor=J_regressor.astype(np.float32),
        kintree_table=kintree.astype(np.int64),
        weights=weights.astype(np.float32),
        f=faces,
    )


def write_smpl_pickle(root, gender="male", seed=0):
    """Write ``<root>/smpl/SMPL_<GENDER>.pkl`` in the layout `smplx.create` opens
    (reference `smplx/body_models.py:2441-2447`). Returns the model_path to pass."""
    d = os.path.join(root, "smpl")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "SMPL_%s.pkl" % gender.upper()), "wb") as fh:
        pickle.dump(make_smpl_dict(seed), fh, protocol=2)
    return root


def x_pose():
    """The canonical template pose, values of reference `assets/X_pose.pkl`:
    all zero except body_pose[2]=+0.5, body_pose[5]=-0.5 (legs apart)."""
    body_pose = np.zeros(69, np.float32)
    body_pose[2] = 0.5
    body_pose[5] = -0.5
    return dict(betas=np.zeros(10, np.float32), global_orient=np.zeros(3, np.float32),
                body_pose=body_pose, transl=np.zeros(3, np.float32))


def make_body_params(num_frames, seed=1, pose_std=0.2, orient_std=0.1, transl=(0.0, 0.0, 3.0)):
    """Per-frame posed params + X-pose template params, as dicts of float32 arrays
    with the keys `AnimNeRFSystem.decode_batch` builds (reference `train.py:172-183`)."""
    rs = np.random.RandomState(seed)
    posed = dict(
        betas=np.tile(rs.normal(0, 0.5, size=(1, 10)), (num_frames, 1)).astype(np.float32),
        global_orient=rs.normal(0, orient_std, size=(num_frames, 3)).astype(np.float32),
        body_pose=rs.normal(0, pose_std, size=(num_frames, 69)).astype(np.float32),
        transl=np.tile(np.asarray(transl, np.float32)[None], (num_frames, 1)),
    )
    xp = x_pose()
    template = {k: np.tile(v[None], (num_frames, 1)).astype(np.float32) for k, v in xp.items()}
    template["betas"] = posed["betas"].copy()
    return posed, template


def make_camera(W, H, focal_scale=1.1):
    """Pinhole camera at the world origin looking down -z after the reference's
    convention flip (`datasets/anim_nerf_dataset.py:211-226` builds c2w from R,t);
    here the body sits at transl=(0,0,3) in *camera* coordinates (OpenCV: +z forward),
    and c2w = [diag(1,-1,-1)|0]^-1-style flip is folded so that cam dir (x,-y,-1) maps
    to world (x, y, z) pointing at +z."""
    fx = fy = focal_scale * W
    c2w = np.array([[1.0, 0, 0, 0], [0, -1.0, 0, 0], [0, 0, -1.0, 0]], np.float32)
    return dict(c2w=c2w, focal=np.array([fx, fy], np.float32),
                c=np.array([W * 0.5, H * 0.5], np.float32), W=W, H=H)


def rays_at_bbox(verts, n_rays, seed=2, near=0.1, far=10.0, margin=0.15):
    """cfg1-style rays: origin at 0, aimed uniformly at the (padded) bbox of posed verts.
    verts (B,V,3) world space -> rays (B,n_rays,8) float32."""
    rs = np.random.RandomState(seed)
    B = verts.shape[0]
    out = np.zeros((B, n_rays, 8), np.float32)
    for b in range(B):
        lo = verts[b].min(0) - margin
        hi = verts[b].max(0) + margin
        tgt = rs.uniform(lo, hi, size=(n_rays, 3))
        d = tgt / np.linalg.norm(tgt, axis=1, keepdims=True)
        out[b, :, 3:6] = d
        out[b, :, 6] = near
        out[b, :, 7] = far
    return out


# --------------------------------------------------------------------- MLP weights
NERF_LAYER_NAMES = (["xyz_encoding_%d.0" % (i + 1) for i in range(8)]
                    + ["xyz_encoding_final", "dir_encoding.0", "sigma", "rgb.0"])


def nerf_layer_shapes(W=256, in_xyz=63):
    """(out, in) of every nn.Linear of the reference NeRF (models/nerf.py:107-127) with
    freqs_xyz=10, use_view=False, no latent codes: 592 388 parameters."""
    shapes = {}
    for i in range(8):
        fan_in = in_xyz if i == 0 else (W + in_xyz if i == 4 else W)
        shapes["xyz_encoding_%d.0" % (i + 1)] = (W, fan_in)
    shapes["xyz_encoding_final"] = (W, W)
    shapes["dir_encoding.0"] = (W // 2, W)
    shapes["sigma"] = (1, W)
    shapes["rgb.0"] = (3, W // 2)
    return shapes


def make_nerf_weights(seed, sigma_bias=5.0, trained_scale=False):
    """Deterministic nn.Linear-style init (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight
    and bias) from numpy so fixtures do not depend on torch's RNG stream.  `sigma_bias`
    is added to the density head's bias: random-init sigma is ~0.016 and would render
    pure white.
    trained_scale=True: a stand-in for a trained network's dynamic range -- the trunk / final / colour-branch
    weights are scaled by 2.7 (just above the critical ReLU gain, so the output varies strongly with the input),
    the density head by 40 with bias 25: over canonical points in [-1,1]^3 sigma then spans about -13 .. 76
    (1st..99th percentile; max > 100) and rgb 0.15 .. 0.95, instead of 5.016 +- 0.003 and 0.5 +- 0.02."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, (o, i) in nerf_layer_shapes().items():
        bound = 1.0 / math.sqrt(i)
        w = rs.uniform(-bound, bound, size=(o, i)).astype(np.float32)
        b = rs.uniform(-bound, bound, size=(o,)).astype(np.float32)
        if trained_scale:
            if name == "sigma":
                w, b = w * np.float32(40.0), b * 0 + np.float32(25.0)
            elif name != "rgb.0":
                w = w * np.float32(2.7)
        elif name == "sigma":
            b = b + np.float32(sigma_bias)
        out[name + ".weight"] = w
        out[name + ".bias"] = b
    return out


# ----------------------------------------------------------------- synthetic training frames
def project_points(pts, cam):
    """world points (n,3) -> (row, col) float pixel coordinates under `make_camera`'s convention
    (world dir of pixel (row j, col i) is ((i-cx)/fx, (j-cy)/fy, 1), camera at the origin)."""
    fx, fy = cam["focal"]
    cx, cy = cam["c"]
    col = cx + fx * pts[:, 0] / pts[:, 2]
    row = cy + fy * pts[:, 1] / pts[:, 2]
    return row, col


def camera_rays(cam, rows, cols, near=0.1, far=10.0):
    """world-space rays (n,8) of the given pixels (numpy restatement of gen_rays for fixtures;
    the timed path uses the an_raygen_fwd kernel)."""
    fx, fy = cam["focal"]
    cx, cy = cam["c"]
    d = np.stack([(cols - cx) / fx, -(rows - cy) / fy, -np.ones_like(cols, dtype=np.float64)], -1)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    dw = d @ cam["c2w"][:, :3].T.astype(np.float64)
    o = np.broadcast_to(cam["c2w"][:, 3].astype(np.float64), dw.shape)
    return np.concatenate([o, dw, np.full_like(dw[:, :1], near), np.full_like(dw[:, :1], far)], -1).astype(np.float32)


def make_training_batch(verts_world, n_side=32, W=512, H=512, seed=3, fg_frac=0.9, band=64):
    """cfg2-style batch: per frame n_side^2 pixels, 90 % on the body silhouette and
    10 % within a `band`-pixel ring outside it (mimics the reference's foreground_pixel sampling,
    datasets/anim_nerf_dataset.py:30-48), random colour targets, alphas in {0,1}.
    verts_world (B,V,3) numpy.  Returns dict of numpy arrays."""
    from scipy import ndimage
    rs = np.random.RandomState(seed)
    cam = make_camera(W, H)
    B = verts_world.shape[0]
    n = n_side * n_side
    rays = np.zeros((B, n, 8), np.float32)
    pix = np.zeros((B, n, 2), np.int32)
    alphas = np.zeros((B, n, 1), np.float32)
    for b in range(B):
        row, col = project_points(verts_world[b].astype(np.float64), cam)
        mask = np.zeros((H, W), bool)
        r = np.clip(np.round(row).astype(int), 0, H - 1)
        c = np.clip(np.round(col).astype(int), 0, W - 1)
        mask[r, c] = True
        mask = ndimage.binary_closing(ndimage.binary_dilation(mask, iterations=3), iterations=2)
        ring = ndimage.binary_dilation(mask, iterations=band) & ~mask
        fg = np.argwhere(mask)
        bg = np.argwhere(ring)
        n_fg = int(round(fg_frac * n))
        sel = np.concatenate([fg[rs.randint(0, len(fg), n_fg)], bg[rs.randint(0, len(bg), n - n_fg)]], 0)
        perm = rs.permutation(n)
        sel = sel[perm]
        pix[b] = sel
        alphas[b, :, 0] = (np.arange(n) < n_fg)[perm]
        rays[b] = camera_rays(cam, sel[:, 0].astype(np.float64), sel[:, 1].astype(np.float64))
    rgbs = rs.uniform(size=(B, n, 3)).astype(np.float32)
    return dict(rays=rays.reshape(B, n_side, n_side, 8), pix=pix, rgbs=rgbs.reshape(B, n_side, n_side, 3),
                alphas=alphas.reshape(B, n_side, n_side, 1), cam=cam)
