"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against the CPU oracle.
Tolerances are stated per test; index outputs are compared bit for bit."""
import numpy as np
import pytest
import torch

from util import oracle, load_golden, nerf_params, body_model, golden_tables, synthetic, ref_mlp_fwd

pytestmark = pytest.mark.gpu

DEV = "cuda"


def ops():
    from anim_nerf_b200 import ops as o
    return o


@pytest.fixture(scope="module")
def det():
    return load_golden("render_det")


@pytest.fixture(scope="module")
def pert():
    return load_golden("render_perturb")


# ----------------------------------------------------------------------------- rays
def test_raygen_full_frame_matches_reference_fixture():
    fx = load_golden("gen_rays")
    H, W = int(fx["H"]), int(fx["W"])
    t = lambda a: torch.from_numpy(np.asarray(a)).to(DEV)[None]
    rays = ops().raygen(t(fx["c2w"]), t(fx["focal"]), t(fx["c"]), H, W, 0.1, 10.0)
    np.testing.assert_allclose(rays[0].cpu().numpy().reshape(H, W, 8), fx["rays"], atol=2e-6)


def test_raygen_pixels_and_body_space():
    rs = np.random.RandomState(0)
    B, H, W, R = 3, 40, 56, 500
    c2w = torch.from_numpy(rs.normal(size=(B, 3, 4)).astype(np.float32))
    focal = torch.from_numpy(rs.uniform(40, 60, size=(B, 2)).astype(np.float32))
    cen = torch.from_numpy(rs.uniform(15, 30, size=(B, 2)).astype(np.float32))
    G = torch.eye(4).repeat(B, 1, 1)
    G[:, :3, :] = torch.from_numpy(rs.normal(size=(B, 3, 4)).astype(np.float32))
    pix = torch.from_numpy(np.stack([rs.randint(0, H, size=(B, R)), rs.randint(0, W, size=(B, R))], -1).astype(np.int32))
    rays = ops().raygen(c2w.to(DEV), focal.to(DEV), cen.to(DEV), H, W, 0.1, 10.0, pix=pix.to(DEV),
                        ginv=torch.inverse(G).to(DEV)).cpu()
    for b in range(B):
        full = oracle.gen_rays(c2w[b], H, W, focal[b], 0.1, 10.0, cen[b])
        sel = full[pix[b, :, 0].long(), pix[b, :, 1].long()][None]
        ref = oracle.rays_to_body_space(sel, G[b:b + 1])
        np.testing.assert_allclose(rays[b].numpy(), ref[0].numpy(), atol=2e-5, rtol=1e-5)


def test_sample_coarse(det, pert):
    rays = torch.from_numpy(det["rays_body"])
    z = ops().sample_coarse(rays.to(DEV), 64).cpu()
    np.testing.assert_allclose(z.numpy(), det["z_coarse"], atol=1e-6)
    rays = torch.from_numpy(pert["rays_body"])
    z = ops().sample_coarse(rays.to(DEV), 64, perturb=1.0, noise_u=torch.from_numpy(pert["noise_coarse_u"]).to(DEV)).cpu()
    np.testing.assert_allclose(z.numpy(), pert["z_coarse"], atol=1e-6)
    # Philox mode: stratified property (each sample stays inside its stratum, ascending)
    z = ops().sample_coarse(rays.to(DEV), 64, perturb=1.0, seed=123).cpu()
    assert (z[..., 1:] >= z[..., :-1]).all()
    assert (z >= rays[..., 6:7] - 1e-6).all() and (z <= rays[..., 7:8]).all()


# ----------------------------------------------------------------------- KNN + unpose
@pytest.mark.parametrize("mode", [0, 1])
def test_knn_unpose_fixture(det, mode):
    verts, o2c, lbs = golden_tables(det)
    rays = torch.from_numpy(det["rays_body"]).to(DEV)
    z = torch.from_numpy(det["z_coarse"]).to(DEV)
    out = ops().knn_unpose(verts.to(DEV), o2c.to(DEV), lbs.to(DEV), 0.2, rays=rays, z=z, mode=mode,
                           want_idx=True, want_dist=True, want_qw=True, compact=True)
    idx = out["idx"].cpu().numpy()
    ref_idx = det["knn_idx_coarse"].astype(np.int32)
    valid = out["valid"].cpu().numpy()
    ref_valid = det["valid_coarse"][..., 0]
    assert (valid == ref_valid).mean() > 0.9999
    found = (idx[..., 0] >= 0)
    if mode == 0:
        assert found.all()
    assert found[ref_valid > 0].all()
    # KNN indices bit-exact (golden = torch.cdist shim of the reference run; 1e-4 slack = cdist ties)
    mism = (idx[found] != ref_idx[found]).any(-1).mean()
    assert mism < 1e-4, mism
    # ... and bit-exact vs the oracle contract on every found query
    xyz = (rays[..., None, 0:3] + z[..., None] * rays[..., None, 3:6]).reshape(z.shape[0], -1, 3).cpu().numpy()
    for b in range(verts.shape[0]):
        d_o, i_o = oracle.knn(verts[b].numpy(), xyz[b], 4)
        f = found[b]
        assert (idx[b][f] == i_o[f]).all()
        assert (out["dist"][b].cpu().numpy()[f] == d_o[f]).all()
    v = ref_valid > 0
    np.testing.assert_allclose(out["xyz_cano"].cpu().numpy()[v], det["xyz_cano_coarse"][v], atol=1e-5)
    # compaction: the list holds exactly the valid ids
    n = int(out["count"].item())
    ids = np.sort(out["cidx"].cpu().numpy()[:n])
    assert (ids == np.nonzero(valid.reshape(-1))[0]).all()


def test_knn_million_queries_bit_exact():
    """>= 1e6 synthetic queries: both search modes vs the C oracle, indices bit for bit."""
    bm = body_model()
    posed_np, tmpl_np = synthetic.make_body_params(2, seed=5)
    t = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
    posed, tmpl = bm(**t(posed_np)), bm(**t(tmpl_np))
    verts, o2c = oracle.ober2cano_tables(posed, tmpl)
    rs = np.random.RandomState(11)
    N = 524288
    lo, hi = verts.numpy().min((0, 1)) - 0.3, verts.numpy().max((0, 1)) + 0.3
    xyz = torch.from_numpy(rs.uniform(lo, hi, size=(2, N, 3)).astype(np.float32))
    lbs = bm.lbs_weights
    outs = [ops().knn_unpose(verts.to(DEV), o2c.to(DEV), lbs.to(DEV), 0.2, xyz=xyz.to(DEV), mode=m,
                             want_idx=True, want_dist=True) for m in (0, 1)]
    for b in range(2):
        d_o, i_o = oracle.knn(verts[b].numpy(), xyz[b].numpy(), 4)
        i0 = outs[0]["idx"][b].cpu().numpy()
        assert (i0 == i_o).all()
        assert (outs[0]["dist"][b].cpu().numpy() == d_o).all()
        i1 = outs[1]["idx"][b].cpu().numpy()
        f = i1[:, 0] >= 0
        assert (i1[f] == i_o[f]).all()
        assert (d_o[~f, 0] >= 0.2).all()          # pruned queries are provably invalid
    assert (outs[0]["valid"] == outs[1]["valid"]).all()
    v = outs[0]["valid"].bool()
    assert torch.equal(outs[0]["xyz_cano"][v], outs[1]["xyz_cano"][v])


def test_knn_seeded_fine_pass_bit_identical(det):
    """Fine pass of VolumeRenderer.forward: seeds from the coarse pass (neighbour reuse for the shared
    samples, search-ball bound for the new ones) change nothing -- every output bit for bit equal to the
    unseeded search and to the exhaustive one."""
    if True:
        verts, o2c, lbs = (t.to(DEV) for t in golden_tables(det))
        rays = torch.from_numpy(det["rays_body"]).to(DEV)
        zc = torch.from_numpy(det["z_coarse"]).to(DEV)
        w = torch.from_numpy(det["weights_coarse"]).to(DEV)
        kw = dict(want_idx=True, want_dist=True, want_qw=True, compact=True)
        coarse = ops().knn_unpose(verts, o2c, lbs, 0.2, rays=rays, z=zc, mode=1, **kw)
        brute_c = ops().knn_unpose(verts, o2c, lbs, 0.2, rays=rays, z=zc, mode=0, **kw)
        fc = coarse["idx"][..., 0] >= 0
        assert torch.equal(coarse["idx"][fc], brute_c["idx"][fc]) and torch.equal(coarse["dist"][fc], brute_c["dist"][fc])
        assert torch.equal(coarse["valid"], brute_c["valid"])
        for u, nf in ((None, 64), (torch.rand(*zc.shape[:-1], 48, device=DEV, generator=torch.Generator(DEV).manual_seed(3)), 48)):
            _, z_all, src, nn = ops().sample_fine_merge(w, zc, nf, det=u is None, u=u)
            seed = dict(src=src, nn=nn, idx=coarse["idx"])
            a = ops().knn_unpose(verts, o2c, lbs, 0.2, rays=rays, z=z_all, mode=1, seed=seed, **kw)
            b = ops().knn_unpose(verts, o2c, lbs, 0.2, rays=rays, z=z_all, mode=1, **kw)
            c = ops().knn_unpose(verts, o2c, lbs, 0.2, rays=rays, z=z_all, mode=0, **kw)
            assert torch.equal(a["valid"], b["valid"]) and torch.equal(a["valid"], c["valid"])
            assert int(a["valid"].sum()) > 1000
            v = a["valid"].bool()
            for k in ("idx", "dist", "qw"):
                assert torch.equal(a[k][v], b[k][v]) and torch.equal(a[k][v], c[k][v]), k
            assert torch.equal(a["xyz_cano"][v], b["xyz_cano"][v]) and torch.equal(a["xyz_cano"][v], c["xyz_cano"][v])
            # every emitted neighbour set is the exact one
            f = a["idx"][..., 0] >= 0
            assert torch.equal(a["idx"][f], c["idx"][f]) and torch.equal(a["dist"][f], c["dist"][f])
            na, nb = int(a["count"]), int(b["count"])
            assert na == nb == int(v.sum())
            assert torch.equal(torch.sort(a["cidx"][:na])[0], torch.sort(b["cidx"][:nb])[0])
            # ... and when the coarse pass also hands over its xyz_cano / valid / qw, the shared samples copy them
            # (no dist output in that mode): still the same bits everywhere, invalid points included
            kw2 = dict(want_idx=True, want_qw=True, compact=True)
            seed2 = dict(seed, xyz_cano=coarse["xyz_cano"], valid=coarse["valid"], qw=coarse["qw"])
            for kwv, sd in ((kw2, seed2), (dict(compact=True), dict(seed2, qw=None))):        # training / inference outputs
                d = ops().knn_unpose(verts, o2c, lbs, 0.2, rays=rays, z=z_all, mode=1, seed=sd, **kwv)
                assert torch.equal(d["valid"], b["valid"]) and torch.equal(d["xyz_cano"], b["xyz_cano"])
                assert int(d["count"]) == nb and torch.equal(torch.sort(d["cidx"][:nb])[0], torch.sort(b["cidx"][:nb])[0])
                if kwv.get("want_qw"):
                    assert torch.equal(d["qw"], b["qw"])
                    assert torch.equal(d["idx"][v], b["idx"][v])


# ------------------------------------------------------------------------------ MLP
def _packed(seed):
    w = synthetic.make_nerf_weights(seed)
    ws = [torch.from_numpy(w[n + ".weight"]).to(DEV) for n in synthetic.NERF_LAYER_NAMES]
    bs = [torch.from_numpy(w[n + ".bias"]).to(DEV) for n in synthetic.NERF_LAYER_NAMES]
    return ops().mlp_pack(ws, bs)


def _unswizzle(img, rows):
    """(rows*128 B uint8 image of one 64-column chunk) -> (rows,64) float32."""
    raw = img.view(torch.int16).reshape(rows, 8, 8)       # (row, physical unit, 8 bf16)
    r = torch.arange(rows, device=img.device)[:, None]
    u = torch.arange(8, device=img.device)[None, :]
    logical = torch.gather(raw, 1, ((u ^ (r & 7))[..., None]).expand(rows, 8, 8))
    return logical.reshape(rows, 64).view(torch.bfloat16).float()


def test_mlp_pack_images():
    packed = _packed(10)
    w = synthetic.make_nerf_weights(10)
    W2 = torch.from_numpy(w["xyz_encoding_2.0.weight"]).to(DEV)
    # fwd chunk (g=1, kc=2) starts after layer 0's single 32 KB chunk and its 8 KB bias slab
    off = 32768 + 8192 + 2 * 32768
    img = _unswizzle(packed[off:off + 32768], 256)
    assert torch.equal(img, W2[:, 128:192].bfloat16().float())
    # bias slab of layer 1 (after its 4 chunks): 8-row groups of 256 B = core matrices k 0..7 | k 8..15; k = 15 = bias
    off = 32768 + 8192 + 4 * 32768
    slab = packed[off:off + 8192].view(torch.bfloat16).reshape(32, 2, 8, 8)      # (group, k half, row in group, k in half)
    b2 = torch.from_numpy(w["xyz_encoding_2.0.bias"]).to(DEV)
    assert torch.equal(slab[:, 1, :, 7].reshape(256).float(), b2.bfloat16().float())
    assert (slab[:, 0] == 0).all() and (slab[:, 1, :, :7] == 0).all()
    W5 = torch.from_numpy(w["xyz_encoding_5.0.weight"]).to(DEV)
    off = (32768 + 8192) + 3 * (4 * 32768 + 8192)
    img = _unswizzle(packed[off:off + 32768], 256)
    assert torch.equal(img[:, :63], W5[:, :63].bfloat16().float()) and (img[:, 63] == 0).all()
    img = _unswizzle(packed[off + 32768:off + 65536], 256)
    assert torch.equal(img, W5[:, 63:127].bfloat16().float())


@pytest.mark.parametrize("impl,n", [(1, 3000), (0, 3000), (0, 70001)])
def test_mlp_forward(impl, n):
    """impl 1 (fp32 SIMT reference kernel): 2e-4.  impl 0 (tcgen05, bf16 operands, fp32 accumulate):
    sigma within 3e-2 abs + 1e-2 rel, rgb within 1e-2 of the fp32 oracle."""
    packed = _packed(10)
    p = nerf_params(10)
    rs = np.random.RandomState(3)
    xc = torch.from_numpy(rs.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32))
    rgb_ref, sig_ref = oracle.nerf_forward(p, xc)
    sigma = torch.full((n,), -7.0, device=DEV)
    rgb = torch.zeros(n, 3, device=DEV)
    if impl == 1:
        ref_mlp_fwd(packed, xc.to(DEV), sigma, rgb)
    else:
        ops().mlp_fwd(packed, xc.to(DEV), sigma, rgb)
    torch.cuda.synchronize()
    ds = (sigma.cpu() - sig_ref[:, 0]).abs()
    dr = (rgb.cpu() - rgb_ref).abs()
    print("impl", impl, "n", n, "max|dsigma|", ds.max().item(), "max|drgb|", dr.max().item())
    if impl == 1:
        assert ds.max() < 2e-4 and dr.max() < 2e-5
    else:
        assert (ds <= 3e-2 + 1e-2 * sig_ref[:, 0].abs()).all(), ds.max()
        assert dr.max() < 1e-2


def test_mlp_forward_compacted_ids():
    """scatter through cidx / device-side count: untouched ids keep their initial values."""
    packed = _packed(11)
    p = nerf_params(11)
    rs = np.random.RandomState(4)
    n_all, n_val = 5000, 1234
    xc = torch.from_numpy(rs.uniform(-1, 1, size=(n_all, 3)).astype(np.float32))
    ids = torch.from_numpy(rs.permutation(n_all)[:n_val].astype(np.int32))
    cidx = torch.zeros(n_all, dtype=torch.int32)
    cidx[:n_val] = ids
    sigma = torch.full((n_all,), -1e5, device=DEV)
    rgb = torch.zeros(n_all, 3, device=DEV)
    count = torch.tensor([n_val], dtype=torch.int32, device=DEV)
    ops().mlp_fwd(packed, xc.to(DEV), sigma, rgb, cidx=cidx.to(DEV), count=count, n_max=n_all)
    rgb_ref, sig_ref = oracle.nerf_forward(p, xc[ids.long()])
    s = sigma.cpu()
    untouched = torch.ones(n_all, dtype=torch.bool)
    untouched[ids.long()] = False
    assert (s[untouched] == -1e5).all() and (rgb.cpu()[untouched] == 0).all()
    assert ((s[ids.long()] - sig_ref[:, 0]).abs() <= 3e-2 + 1e-2 * sig_ref[:, 0].abs()).all()
    assert (rgb.cpu()[ids.long()] - rgb_ref).abs().max() < 1e-2


def test_mlp_stash_images_match_layerwise_oracle():
    """training mode: every stashed activation image vs the fp32 oracle (bf16-level tolerance)."""
    from anim_nerf_b200 import ops as o
    packed = _packed(10)
    p = nerf_params(10)
    n = 700
    rs = np.random.RandomState(5)
    xc = torch.from_numpy(rs.uniform(-1, 1, size=(n, 3)).astype(np.float32))
    sigma = torch.zeros(n, device=DEV)
    rgb = torch.zeros(n, 3, device=DEV)
    stash = o.mlp_stash(n, DEV)
    stash.zero_()
    o.mlp_fwd(packed, xc.to(DEV), sigma, rgb, stash=stash)
    torch.cuda.synchronize()
    e = oracle.embed(xc)
    hs, h = [], e
    for i in range(8):
        if i == 4:
            h = torch.cat([e, h], -1)
        w, b = p["xyz_encoding_%d.0" % (i + 1)]
        h = torch.relu(h @ w.T + b)
        hs.append(h)
    tile_bytes = 608256                      # mlp_layout.cuh ST_TILE: enc 16K + 8 x 64K (h1..h8) + c 32K + masks
    for tile in range((n + 127) // 128):
        base = tile * tile_bytes
        rows = min(128, n - tile * 128)
        sl = slice(tile * 128, tile * 128 + rows)
        enc = _unswizzle(stash[base:base + 16384], 128)[:rows].cpu()
        assert (enc[:, :63] - e[sl]).abs().max() < 1e-2, "enc tile %d" % tile
        for l in range(8):
            img = torch.cat([_unswizzle(stash[base + 16384 + l * 65536 + c * 16384:][:16384], 128) for c in range(4)], 1)[:rows].cpu()
            err = (img - hs[l][sl]).abs().max().item()
            assert err < 3e-2 * max(1.0, hs[l][sl].abs().max().item()), "h%d tile %d err %g" % (l + 1, tile, err)
            m = stash[base + 16384 + 8 * 65536 + 32768 + l * 4096:][:4096].view(torch.int32).reshape(8, 128).T[:rows].cpu()
            # mask words are stored [block][row]; bit layout per 32-column block: bit (31-c) <-> column c
            col = torch.arange(32)
            bits = ((m[:, :, None] >> (31 - col)) & 1).reshape(rows, 256).bool()
            assert (bits == (img > 0)).all(), "mask h%d" % (l + 1)


# ------------------------------------------------------------------------ compositing
def _comp_inputs(seed, n, K):
    rs = np.random.RandomState(seed)
    sigma = torch.from_numpy(rs.normal(2.0, 6.0, size=(n, K)).astype(np.float32))
    sigma[rs.uniform(size=(n, K)) < 0.4] = -1e5
    rgb = torch.from_numpy(rs.uniform(size=(n, K, 3)).astype(np.float32))
    z = torch.from_numpy(np.sort(rs.uniform(2.0, 4.0, size=(n, K)), -1).astype(np.float32))
    rays = torch.zeros(n, 8)
    rays[:, 7] = 4.0
    return sigma, rgb, z, rays


@pytest.mark.parametrize("K", [64, 96, 128])
def test_composite_forward_backward(K):
    n = 777
    sigma, rgb, z, rays = _comp_inputs(K, n, K)
    noise = torch.from_numpy(np.random.RandomState(1).normal(size=(n, K)).astype(np.float32))
    for nz in (None, noise):
        s, c, zz = sigma.clone().requires_grad_(True), rgb.clone().requires_grad_(True), z.clone().requires_grad_(True)
        far = rays[:, 7:8].clone().requires_grad_(True)
        w_r, rgb_r, dep_r, acc_r = oracle.composite(c[None], s[None], zz[None], far[None], True, None if nz is None else nz[None])
        g = [o.to(DEV) for o in (sigma, rgb, z, rays)]
        w, rgb_o, dep, acc = ops().composite(g[0], g[1], g[2], g[3], True, None if nz is None else nz.to(DEV))
        np.testing.assert_allclose(w.cpu().numpy(), w_r[0].detach().numpy(), atol=2e-6)
        np.testing.assert_allclose(rgb_o.cpu().numpy(), rgb_r[0].detach().numpy(), atol=5e-6)
        np.testing.assert_allclose(dep.cpu().numpy(), dep_r[0].detach().numpy(), atol=2e-5)
        np.testing.assert_allclose(acc.cpu().numpy(), acc_r[0].detach().numpy(), atol=5e-6)
        rs = np.random.RandomState(9)
        gr, gd, ga = [torch.from_numpy(rs.normal(size=s_).astype(np.float32)) for s_ in ((n, 3), (n, 1), (n, 1))]
        ((rgb_r[0] * gr).sum() + (dep_r[0] * gd).sum() + (acc_r[0] * ga).sum()).backward()
        g_sigma, g_rgb, g_z, g_far = ops().composite_bwd(g[0], g[1], g[2], g[3], gr.to(DEV), gd.to(DEV), ga.to(DEV), True,
                                                         None if nz is None else nz.to(DEV))
        np.testing.assert_allclose(g_rgb.cpu().numpy(), c.grad.numpy(), atol=1e-5)
        np.testing.assert_allclose(g_sigma.cpu().numpy(), s.grad.numpy(), atol=1e-4, rtol=1e-3)
        np.testing.assert_allclose(g_z.cpu().numpy(), zz.grad.numpy(), atol=2e-3, rtol=2e-3)
        np.testing.assert_allclose(g_far.cpu().numpy(), far.grad.numpy()[:, 0], atol=1e-5)


# ------------------------------------------------------------------------- resampling
def test_searchsorted_bit_exact(det, pert):
    for fx in (det, pert):
        inds = ops().searchsorted_right(torch.from_numpy(fx["cdf"]).to(DEV), torch.from_numpy(fx["u"]).to(DEV))
        assert (inds.cpu().numpy() == fx["inds"].astype(np.int32)).all()


def test_sample_fine_merge(det, pert):
    # deterministic u (in-kernel linspace) on the reference's captured coarse weights
    w = torch.from_numpy(det["weights_coarse"]).to(DEV)
    zc = torch.from_numpy(det["z_coarse"]).to(DEV)
    z_fine, z_all, src, nn = ops().sample_fine_merge(w, zc, 64, det=True)
    ref_all = det["z_combine"]
    za = z_all.cpu().numpy()
    assert (np.diff(za, axis=-1) >= 0).all()
    close = np.abs(za - ref_all) < 2e-4
    assert close.mean() > 0.97, close.mean()     # re-derived cdf flips a few u==cdf ties
    cat = torch.cat([zc, z_fine], -1)
    assert torch.equal(torch.gather(cat, -1, src.long()), z_all)        # src is the sort permutation
    assert (torch.sort(src.long(), -1)[0] == torch.arange(128, device=DEV)).all()
    # nn = the coarse sample nearest in depth (itself for a coarse entry)
    gap = (z_all[..., :, None] - zc[..., None, :]).abs()
    assert torch.equal(torch.gather(gap, -1, nn.long()[..., None])[..., 0], gap.min(-1)[0])
    assert torch.equal(nn[src < 64], src[src < 64])
    # explicit random u
    w = torch.from_numpy(pert["weights_coarse"]).to(DEV)
    zc = torch.from_numpy(pert["z_coarse"]).to(DEV)
    _, z_all, _, _ = ops().sample_fine_merge(w, zc, 32, det=False, u=torch.from_numpy(pert["noise_fine_u"]).to(DEV))
    np.testing.assert_allclose(z_all.cpu().numpy(), pert["z_combine"], atol=2e-4)


# ------------------------------------------------------------------------ MLP backward
def _st_bf16(t):
    """round to bf16, straight-through gradient (the kernels treat operand rounding as identity)."""
    return t + (t.bfloat16().float() - t).detach()


def _mlp_bwd_case(n, seed=10, with_gx=True, emulate_bf16=True):
    """Kernel gradients + autograd reference.  emulate_bf16: the reference forward rounds the MMA
    operands (activations, weights) to bf16 exactly where the kernel does, so both take the same
    ReLU branches; with False it is the plain fp32 oracle."""
    from anim_nerf_b200 import ops as o
    packed = _packed(seed)
    p = nerf_params(seed, requires_grad=True)
    rd = _st_bf16 if emulate_bf16 else (lambda t: t)
    rs = np.random.RandomState(6)
    xc = torch.from_numpy(rs.uniform(-1, 1, size=(n, 3)).astype(np.float32)).requires_grad_(True)
    gs = torch.from_numpy(rs.normal(size=(n,)).astype(np.float32))
    grgb = torch.from_numpy(rs.normal(size=(n, 3)).astype(np.float32))
    e = rd(oracle.embed(xc))
    pre, h = [], e
    for i in range(8):
        if i == 4:
            h = torch.cat([e, h], -1)
        w, b = p["xyz_encoding_%d.0" % (i + 1)]
        a = h @ rd(w).T + rd(b)              # the bias enters through a bf16 tensor-core step (mlp_layout.cuh: bias slab)
        a.retain_grad(); pre.append(a)
        h32 = torch.relu(a)
        h = rd(h32)
    # the kernels run final+colour as ONE layer W' = W_dir W_final (formed in fp32, rounded once to bf16),
    # with the density head as one more output column of it, and the rgb head as a bf16 GEMM on c
    Wf, bf = p["xyz_encoding_final"]
    Wd, bd = p["dir_encoding.0"]
    sig = (h @ rd(p["sigma"][0]).T + p["sigma"][1])[:, 0]
    cpre = h @ rd(Wd @ Wf).T + rd(Wd @ bf + bd)
    cpre.retain_grad()
    rgbpre = rd(torch.relu(cpre)) @ rd(p["rgb.0"][0]).T + p["rgb.0"][1]
    rgbpre.retain_grad()
    rgb_ref = torch.sigmoid(rgbpre)
    ((sig * gs).sum() + (rgb_ref * grgb).sum()).backward()
    sigma = torch.zeros(n, device=DEV)
    rgb = torch.zeros(n, 3, device=DEV)
    stash = o.mlp_stash(n, DEV)
    o.mlp_fwd(packed, xc.detach().to(DEV), sigma, rgb, stash=stash)
    g_params, g_xyz = o.mlp_bwd(packed, stash, xc.detach().to(DEV), rgb, gs.to(DEV), grgb.to(DEV), want_g_xyz=with_gx)
    torch.cuda.synchronize()
    return dict(p=p, xc=xc, pre=pre, cpre=cpre, rgbpre=rgbpre, gs=gs, g_params=g_params.cpu(), g_xyz=None if g_xyz is None else g_xyz.cpu(),
                scratch=o._last_bwd_scratch)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("n,with_gx,emu", [(700, True, True), (33000, True, True), (5000, False, True), (5000, True, False)])
def test_mlp_backward(n, with_gx, emu):
    """Gradients of the bf16 tensor-core backward vs autograd of the oracle.
    emu=True: oracle evaluated at the kernel's operand precision (bf16 operands, fp32 accumulate):
      relative L2 error per tensor <= 5e-2 (measured 0.1-3.2 %).
    emu=False: plain fp32 oracle: <= 0.25 -- at random init ~0.2 % of the near-zero pre-activations take
      the other ReLU branch in bf16, each flip is a full-size error in that unit's gradient
      (rel. L2 ~ sqrt(flip fraction)); reported, not a kernel defect."""
    r = _mlp_bwd_case(n, with_gx=with_gx, emulate_bf16=emu)
    tol = 5e-2 if emu else 0.25
    # dgrad diagnostics: every pre-activation gradient image vs autograd
    DY_TILE, DY_HEAD, DY_H, dy_err = 589824, 16384, 65536, {}       # mlp_layout.cuh
    for tile in range((n + 127) // 128):
        rows = min(128, n - tile * 128)
        sl = slice(tile * 128, tile * 128 + rows)
        base = tile * DY_TILE
        sc = r["scratch"]
        img = _unswizzle(sc[base:][:16384], 128)[:rows, :3].cpu()                       # d rgb_pre (columns 0..2)
        dy_err.setdefault("rgbpre", []).append(_rel(img, r["rgbpre"].grad[sl]))
        img = torch.cat([_unswizzle(sc[base + DY_HEAD + c * 16384:][:16384], 128) for c in range(2)], 1)[:rows].cpu()
        dy_err.setdefault("cpre", []).append(_rel(img, r["cpre"].grad[sl]))
        img = _unswizzle(sc[base + DY_HEAD + 2 * 16384:][:16384], 128)[:rows, 0].cpu()  # d sigma rides in the head layer's third chunk
        dy_err.setdefault("dsigma", []).append(_rel(img, r["gs"][sl]))
        for g in range(8):
            img = torch.cat([_unswizzle(sc[base + DY_H + g * 65536 + c * 16384:][:16384], 128) for c in range(4)], 1)[:rows].cpu()
            dy_err.setdefault("pre%d" % (g + 1), []).append(_rel(img, r["pre"][g].grad[sl]))
    print("dY rel err (max over tiles):", {k: round(max(v), 4) for k, v in dy_err.items()})
    names = synthetic.NERF_LAYER_NAMES
    off = 0
    errs = {}
    for name in names:
        W, b = r["p"][name]
        gw = r["g_params"][off:off + W.numel()].view_as(W); off += W.numel()
        gb = r["g_params"][off:off + b.numel()].view_as(b); off += b.numel()
        errs[name + ".weight"] = _rel(gw, W.grad)
        errs[name + ".bias"] = _rel(gb, b.grad)
    if with_gx:
        errs["xyz"] = _rel(r["g_xyz"], r["xc"].grad)
    print({k: round(v, 4) for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, bad


def test_mlp_backward_is_bit_reproducible():
    """The weight gradient is reduced in a fixed order (per-CTA partial slices + mlp_wgrad_reduce_kernel, no
    floating-point atomics): two runs over the same inputs give bit-identical gradients, also when accumulating
    into a caller-owned buffer."""
    from anim_nerf_b200 import ops as o
    packed = _packed(10)
    n = 40000
    rs = np.random.RandomState(8)
    xc = torch.from_numpy(rs.uniform(-1, 1, size=(n, 3)).astype(np.float32)).to(DEV)
    gs = torch.from_numpy(rs.normal(size=(n,)).astype(np.float32)).to(DEV)
    grgb = torch.from_numpy(rs.normal(size=(n, 3)).astype(np.float32)).to(DEV)
    sigma, rgb = torch.zeros(n, device=DEV), torch.zeros(n, 3, device=DEV)
    stash = o.mlp_stash(n, DEV)
    o.mlp_fwd(packed, xc, sigma, rgb, stash=stash)
    runs = []
    for _ in range(3):
        g, gx = o.mlp_bwd(packed, stash, xc, rgb, gs, grgb)
        runs.append((g.clone(), gx.clone()))
    for g, gx in runs[1:]:
        assert torch.equal(g, runs[0][0]) and torch.equal(gx, runs[0][1])
    acc = torch.zeros_like(runs[0][0])
    o.mlp_bwd(packed, stash, xc, rgb, gs, grgb, g_params=acc)
    o.mlp_bwd(packed, stash, xc, rgb, gs, grgb, g_params=acc)
    assert torch.equal(acc[:o.FLAT_FLOATS], 2 * runs[0][0][:o.FLAT_FLOATS])


# ------------------------------------------------------------------ per-frame tables (A16, 8(f)#1)
@pytest.mark.parametrize("shared_template", [False, True])
def test_body_tables_match_torch_builder(shared_template):
    """an_body_tables_fwd vs the differentiable torch builder (BodyModel + AnimNeRF.set_body_model /
    convert_to_body_model_space / clac_ober2cano_transform, themselves pinned to the reference's tables by
    the golden fixtures): posed vertices in the root frame, observation->canonical transforms, inverse root
    transform, template vertices.  fp32, different summation order: tolerance 2e-5 absolute."""
    from anim_nerf_b200.anim_nerf import AnimNeRF, affine_inverse
    B = 5
    net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=False, freqs_dir=0, body_model_data=synthetic.make_smpl_dict(0)).to(DEV)
    posed_np, tmpl_np = synthetic.make_body_params(B, seed=7)
    rs = np.random.RandomState(3)
    posed_np["betas"] = rs.normal(0, 0.5, size=posed_np["betas"].shape).astype(np.float32)
    posed = {k: torch.from_numpy(v).to(DEV) for k, v in posed_np.items()}
    tmpl = {k: torch.from_numpy(v[:1] if shared_template else v).to(DEV) for k, v in tmpl_np.items()}
    with torch.no_grad():
        net.set_body_model(posed, {k: v.expand(B, *v.shape[1:]) for k, v in tmpl.items()} if shared_template else tmpl)
        ginv_ref = affine_inverse(net.global_transform)
        net.convert_to_body_model_space(None)
        net.clac_ober2cano_transform()
        verts_ref, o2c_ref, vt_ref = net.verts.clone(), net.ober2cano_transform.clone(), net.verts_template.clone()
    verts, o2c, ginv, vt = ops().body_tables(net.body_model, posed, tmpl)
    assert float((verts - verts_ref).abs().max()) < 2e-5
    assert float((ginv - ginv_ref).abs().max()) < 2e-5
    assert float((vt - vt_ref).abs().max()) < 2e-5
    assert float((o2c - o2c_ref).abs().max()) < 1e-4
    assert torch.equal(o2c[..., 3, :], torch.tensor([0.0, 0.0, 0.0, 1.0], device=DEV).expand(B, o2c.shape[1], 4))
    # the path uses the fused builder with or without gradients to the posed body's parameters
    with torch.no_grad():
        rays_w = torch.from_numpy(synthetic.rays_at_bbox(posed_np["transl"][:, None] + np.zeros((B, 4, 3), np.float32), 8)).to(DEV)
        rays_b, ginv2 = net.setup_frame(posed, tmpl, rays_w)
    assert net.verts_transform is None and torch.equal(ginv2, ginv) and torch.equal(net.verts, verts)
    posed_g = dict(posed, body_pose=posed["body_pose"].clone().requires_grad_(True))
    rays_b2, _ = net.setup_frame(posed_g, tmpl, rays_w)
    assert net.verts_transform is None and net.ober2cano_transform.requires_grad and not net.verts.requires_grad
    assert float((rays_b2.detach() - rays_b).abs().max()) < 2e-5
    net.ober2cano_transform.sum().backward()
    assert posed_g["body_pose"].grad is not None and float(posed_g["body_pose"].grad.abs().sum()) > 0


@pytest.mark.parametrize("shared_template", [False, True])
def test_body_tables_backward_matches_torch_builder(shared_template):
    """an_body_tables_bwd vs torch autograd through the differentiable torch builder (BodyModel +
    set_body_model / convert_to_body_model_space / clac_ober2cano_transform): gradients of a random linear
    functional of (ober2cano, body-space rays) w.r.t. betas, global_orient, body_pose, transl of the posed body.
    The shared shape row (BodyModelParams.betas: one embedding row for all frames) is covered by passing (1,10) betas.
    fp32 both sides, different summation order: relative L2 error <= 2e-4 per parameter."""
    from anim_nerf_b200.anim_nerf import AnimNeRF
    B = 4
    net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=False, freqs_dir=0, body_model_data=synthetic.make_smpl_dict(0)).to(DEV)
    posed_np, tmpl_np = synthetic.make_body_params(B, seed=9)
    rs = np.random.RandomState(5)
    posed_np["betas"] = rs.normal(0, 0.5, size=(1, 10)).astype(np.float32).repeat(B, 0)
    tmpl = {k: torch.from_numpy(v[:1] if shared_template else v).to(DEV) for k, v in tmpl_np.items()}
    tmpl_full = {k: v.expand(B, *v.shape[1:]) for k, v in tmpl.items()} if shared_template else tmpl
    rays_w = torch.from_numpy(synthetic.rays_at_bbox(posed_np["transl"][:, None] + np.zeros((B, 4, 3), np.float32), 16)).to(DEV)
    V = net.body_model.v_template.shape[0]
    c_o2c = torch.from_numpy(rs.normal(size=(B, V, 4, 4)).astype(np.float32)).to(DEV)
    c_o2c[:, :, 3] = 0          # the last row is the constant [0,0,0,1]
    c_rays = torch.from_numpy(rs.normal(size=tuple(rays_w.shape)).astype(np.float32)).to(DEV)

    def run(fused, shared_betas):
        net.fused_tables = fused
        leaf = {k: torch.from_numpy(v).to(DEV).requires_grad_(True) for k, v in posed_np.items()}
        if shared_betas:
            leaf["betas"] = torch.from_numpy(posed_np["betas"][:1]).to(DEV).requires_grad_(True)
        posed = dict(leaf)
        if shared_betas and not fused:
            posed["betas"] = leaf["betas"].expand(B, -1)
        rays_b, _ = net.setup_frame(posed, tmpl if fused else tmpl_full, rays_w)
        ((net.ober2cano_transform * c_o2c).sum() + (rays_b * c_rays).sum()).backward()
        return {k: v.grad.clone() for k, v in leaf.items()}

    for shared_betas in (False, True):
        g_ref = run(False, shared_betas)
        g_ker = run(True, shared_betas)
        for k in g_ref:
            err = float((g_ker[k] - g_ref[k]).norm() / (g_ref[k].norm() + 1e-20))
            print(shared_betas, k, "rel err %.3g" % err, "norm %.3g" % float(g_ref[k].norm()))
            assert g_ker[k].shape == g_ref[k].shape and err < 2e-4, (k, err)
    net.fused_tables = True


# ------------------------------------------------------------------ fused ray front end (A1 + A2 + A3) and its backward
def test_rays_sample_fused_matches_separate_kernels_and_autograd():
    """an_rays_sample_fwd == an_raygen_fwd + an_sample_coarse_fwd bit for bit (camera and world-ray input, with explicit
    noise); an_rays_sample_bwd == torch autograd of the reference arithmetic (models/anim_nerf.py:128-137 +
    models/volume_rendering.py:29-56) with respect to ginv; an_ray_point_grad == the torch reductions it replaces."""
    from anim_nerf_b200 import ops as o
    rs = np.random.RandomState(12)
    B, H, W, Kc = 3, 9, 7, 64
    R = H * W
    c2w = torch.from_numpy(np.stack([np.concatenate([np.linalg.qr(rs.normal(size=(3, 3)))[0], rs.normal(size=(3, 1))], 1) for _ in range(B)]).astype(np.float32)).to(DEV)
    focal = torch.from_numpy(rs.uniform(8, 12, size=(B, 2)).astype(np.float32)).to(DEV)
    center = torch.from_numpy(rs.uniform(3, 5, size=(B, 2)).astype(np.float32)).to(DEV)
    A = rs.normal(size=(B, 3, 3)).astype(np.float32) * 0.3 + np.eye(3, dtype=np.float32)
    g = np.zeros((B, 4, 4), np.float32); g[:, :3, :3] = A; g[:, :3, 3] = rs.normal(size=(B, 3)) * 0.5 + np.array([0, 0, 2.5]); g[:, 3, 3] = 1
    ginv = torch.from_numpy(g).to(DEV)
    noise = torch.rand(B, R, Kc, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    cam = dict(c2w=c2w, focal=focal, center=center, H=H, W=W, near=0.1, far=10.0)
    for perturb in (0.0, 1.0):
        rays_ref = o.raygen(c2w, focal, center, H, W, 0.1, 10.0, ginv=ginv)
        z_ref = o.sample_coarse(rays_ref, Kc, perturb, noise_u=noise if perturb > 0 else None)
        rb, z = o.rays_sample(Kc, perturb, noise if perturb > 0 else None, camera=cam, ginv=ginv)
        assert torch.equal(rb, rays_ref) and torch.equal(z, z_ref)
        rays_w = o.raygen(c2w, focal, center, H, W, 0.1, 10.0)
        rb2, z2 = o.rays_sample(Kc, perturb, noise if perturb > 0 else None, rays_world=rays_w, ginv=ginv)
        assert (rb2 - rays_ref).abs().max() < 1e-5 and (z2 - z_ref).abs().max() < 1e-5
    # backward w.r.t. ginv against torch autograd of the same arithmetic
    from anim_nerf_b200.anim_nerf import AnimNeRF
    c_r = torch.from_numpy(rs.normal(size=(B, R, 8)).astype(np.float32)).to(DEV)
    c_z = torch.from_numpy(rs.normal(size=(B, R, Kc)).astype(np.float32)).to(DEV)
    gi = ginv.clone().requires_grad_(True)
    rb_t = AnimNeRF.rays_to_body_space(rays_w, gi)
    near, far = rb_t[..., 6:7], rb_t[..., 7:8]
    t = torch.linspace(0, 1 - 1.0 / Kc, Kc, device=DEV)
    zt = near * (1 - t) + far * t
    mid = 0.5 * (zt[..., 1:] + zt[..., :-1])
    zt = torch.cat([zt[..., :1], mid], -1) + (torch.cat([mid, zt[..., -1:]], -1) - torch.cat([zt[..., :1], mid], -1)) * noise
    ((rb_t * c_r).sum() + (zt * c_z).sum()).backward()
    rb, z = o.rays_sample(Kc, 1.0, noise, rays_world=rays_w, ginv=ginv)
    g_k = o.rays_sample_bwd(rb, z, c_r, c_z, rays_world=rays_w)
    err = float((g_k[:, :3] - gi.grad[:, :3]).norm() / gi.grad[:, :3].norm())
    assert err < 1e-4, err
    g_c = o.rays_sample_bwd(rb, z, c_r, c_z, camera=cam)          # world rays regenerated from the camera
    assert float((g_c[:, :3] - gi.grad[:, :3]).norm() / gi.grad[:, :3].norm()) < 1e-4
    # ray-side gradients of a pass
    K = 96
    zz = torch.sort(torch.rand(B, R, K, device=DEV) * 3 + 1, -1)[0]
    valid = (torch.rand(B, R, K, device=DEV) < 0.4).to(torch.uint8)
    gx = torch.randn(B, R, K, 3, device=DEV)
    gx_dirty = torch.where(valid[..., None].bool(), gx, torch.full_like(gx, float("nan")))      # invalid entries must not be read
    gzc, gfar = torch.randn(B, R, K, device=DEV), torch.randn(B, R, device=DEV)
    g_rays, g_z = o.ray_point_grad(rb, zz, valid, gx_dirty.view(B, R * K, 3), gzc, gfar)
    gxm = gx * valid[..., None]
    assert (g_rays[..., 0:3] - gxm.sum(2)).abs().max() < 1e-4 and (g_rays[..., 3:6] - (gxm * zz[..., None]).sum(2)).abs().max() < 2e-4
    assert torch.equal(g_rays[..., 7], gfar) and float(g_rays[..., 6].abs().max()) == 0
    assert (g_z - (gzc + (gxm * rb[:, :, None, 3:6]).sum(-1))).abs().max() < 1e-5


@pytest.mark.parametrize("fine", [True, False])
def test_render_loss_matches_torch_losses_and_autograd(fine):
    """A18: an_render_loss against F.mse_loss / F.l1_loss (train.py:228-262) -- the four terms, the total and the
    gradients through autograd (scaled by an upstream factor), sign(0) = 0 at exact hits, ragged ray counts."""
    from anim_nerf_b200.autograd import RenderLoss
    F = torch.nn.functional
    for n, seed in ((16384, 0), (1, 1), (1237, 2)):
        g = torch.Generator().manual_seed(seed)
        mk = lambda *s: torch.rand(*s, generator=g).to(DEV)                   # noqa: E731
        rc, rf, ac, af = mk(n, 3), mk(n, 3), mk(n, 1), mk(n, 1)
        tr, ta = mk(n, 3), (mk(n, 1) > 0.5).float()
        ac[: n // 3] = ta[: n // 3]                                            # exact hits: l1 gradient 0 there
        lam = 0.1
        a = [t.clone().requires_grad_(True) for t in (rc, rf, ac, af)]
        b = [t.clone().requires_grad_(True) for t in (rc, rf, ac, af)]
        total, terms = RenderLoss.apply(a[0], a[1] if fine else None, a[2], a[3] if fine else None, tr, ta, lam)
        want_terms = [F.mse_loss(b[0], tr), F.mse_loss(b[1], tr), F.l1_loss(b[2], ta), F.l1_loss(b[3], ta)]
        want = want_terms[0] + lam * want_terms[2] + ((want_terms[1] + lam * want_terms[3]) if fine else 0.0)
        (3.0 * total).backward()
        (3.0 * want).backward()
        assert abs(float(total) - float(want)) <= 2e-6 * abs(float(want)) + 1e-9, (float(total), float(want))
        for k in range(4):
            if fine or k in (0, 2):
                assert abs(float(terms[k]) - float(want_terms[k])) <= 2e-6 * abs(float(want_terms[k])) + 1e-9
                np.testing.assert_allclose(a[k].grad.cpu().numpy(), b[k].grad.cpu().numpy(), rtol=1e-6, atol=1e-12)
            else:
                assert a[k].grad is None
        assert float(a[2].grad[: n // 3].abs().max()) == 0.0 if n >= 3 else True


def test_compact_valid_is_the_ordered_nonzero_list():
    """A11: an_compact_valid against torch.nonzero on ragged sizes (one flag, partial 16-byte groups, more ranges than
    one block handles, ranges that double to stay within 1024 blocks), dense, sparse and empty flag arrays."""
    g = torch.Generator().manual_seed(5)
    for n, p in ((1, 1.0), (1, 0.0), (15, 0.5), (16, 1.0), (4096, 0.3), (4097, 0.3), (70001, 0.02), (1 << 20, 0.6),
                 ((1 << 23) + 5, 0.25)):
        valid = (torch.rand(n, generator=g) < p).to(torch.uint8).to(DEV)
        cidx = torch.full((n,), -7, dtype=torch.int32, device=DEV)
        count = torch.full((1,), -1, dtype=torch.int32, device=DEV)
        ops().compact_valid(valid, cidx, count)
        want = torch.nonzero(valid).flatten().to(torch.int32)
        assert int(count) == want.numel(), (n, p, int(count), want.numel())
        assert torch.equal(cidx[:want.numel()], want), (n, p)
        assert bool((cidx[want.numel():] == -7).all())          # nothing written past the list


def test_training_steps_are_bit_reproducible_with_frozen_body_params():
    """Two runs of three optimiser steps from the same weights on the same rays give identical bits (losses, gradients,
    weights): ordered compaction + fixed-order weight-gradient reduction + no floating-point atomics on the MLP path."""
    from anim_nerf_b200.anim_nerf import AnimNeRF
    from anim_nerf_b200.volume_rendering import VolumeRenderer
    from anim_nerf_b200.optim import FlatGradBuffer, FusedAdam
    from util import body_model, oracle
    posed_np, tmpl_np = synthetic.make_body_params(1, seed=5)
    posed = {k: torch.from_numpy(v) for k, v in posed_np.items()}
    with torch.no_grad():
        po = body_model()(**posed)
        rays_w = torch.from_numpy(synthetic.rays_at_bbox(po["vertices"].numpy(), 96, seed=4, margin=0.02))
        rays = oracle.rays_to_body_space(rays_w, po["joints_transform"][:, 0]).to(DEV)
    posed_d = {k: v.to(DEV) for k, v in posed.items()}
    tmpl_d = {k: torch.from_numpy(v).to(DEV) for k, v in tmpl_np.items()}
    tgt = torch.rand(1, 96, 3, generator=torch.Generator().manual_seed(1)).to(DEV)

    def run():
        net = AnimNeRF(use_unpose=True, use_knn=True, use_fine=True, freqs_dir=0, body_model_data=synthetic.make_smpl_dict(0)).to(DEV)
        for name, seed in (("nerf", 10), ("nerf_fine", 11)):
            getattr(net, name).load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.make_nerf_weights(seed).items()}, strict=True)
        vr = VolumeRenderer(n_coarse=64, n_fine=64, white_bkgd=True)
        params = [p for n in ("nerf", "nerf_fine") for p in getattr(net, n).parameters()]
        opt = FusedAdam(params, lr=5e-4, eps=1e-8)
        flat = FlatGradBuffer([net.nerf, net.nerf_fine])
        opt.flat = flat
        opt.on_step.append(lambda: (net.nerf.mark_dirty(), net.nerf_fine.mark_dirty()))
        rec = []
        for _ in range(3):
            net.setup_frame(posed_d, tmpl_d, None)
            out = vr(net, rays, perturb=0.0)
            loss = ((out["rgbs"] - tgt) ** 2).mean() + ((out["rgbs_fine"] - tgt) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            rec.append((loss.detach().clone(), flat.buf.clone()))
            opt.step()
        return rec, torch.cat([p.detach().flatten() for p in params])
    (ra, wa), (rb, wb) = run(), run()
    for (la, ga), (lb, gb) in zip(ra, rb):
        assert torch.equal(la, lb) and torch.equal(ga, gb)
    assert torch.equal(wa, wb)
