"""GPU parity of the training-ray sampler (an_sample_training_rays_fwd) against the values
captured from the reference (tests/golden/pixel_sampling.npz), the oracle, and an_raygen_fwd."""
import numpy as np
import pytest
import torch

from util import oracle, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def fx():
    return load_golden("pixel_sampling")


def _store(fx, frames=1, **kw):
    from anim_nerf_b200.train_rays import DeviceFrameStore
    imgs = [fx["img_u8"]] + [np.roll(fx["img_u8"], 7 * f, axis=1) for f in range(1, frames)]
    masks = [fx["mask_u8"]] + [np.roll(fx["mask_u8"], 7 * f, axis=1) for f in range(1, frames)]
    return DeviceFrameStore(torch.from_numpy(np.stack(imgs)), torch.from_numpy(np.stack(masks)), device=DEV, **kw)


def _cam(fx, B):
    t = lambda a: torch.from_numpy(a)[None].repeat(B, *([1] * a.ndim)).to(DEV)      # noqa: E731
    return t(fx["c2w"]), t(fx["focal"]), t(fx["c"])


def test_reference_draws_reproduce_the_reference_batch(fx):
    """Parity mode: fed the reference's own np.random.choice positions, the kernel returns the reference's pixels
    (bit-exact), colours and alphas (bit-exact) and rays (<= 1e-6)."""
    s = int(fx["n_side"])
    np.random.seed(5)
    coords, sel = oracle.get_pixelcoords(np.float32(fx["mask_u8"] / 255.), s, 0.9, 3)
    st = _store(fx)
    c2w, focal, c = _cam(fx, 1)
    out = st.sample([0], c2w, focal, c, subsamplesize=s, fore_rate=0.9, sel=torch.from_numpy(sel)[None].to(DEV))
    assert np.array_equal(out["pix"][0].cpu().numpy(), fx["coords_e3"])
    assert np.array_equal(out["rgbs"].view(-1, 3).cpu().numpy(), fx["rgbs"])
    assert np.array_equal(out["alphas"].view(-1, 1).cpu().numpy(), fx["alphas"])
    np.testing.assert_allclose(out["rays"].view(-1, 8).cpu().numpy(), fx["rays"], rtol=0, atol=1e-6)


def test_body_space_rays_and_raygen_identity(fx):
    """With ginv the rays equal the oracle's convert_to_body_model_space of the world rays, and are bit-identical to
    an_raygen_fwd at the same pixels (shared an_make_ray)."""
    from anim_nerf_b200 import ops
    st = _store(fx, frames=3)
    B, s = 4, 8
    c2w, focal, c = _cam(fx, B)
    rs = np.random.RandomState(3)
    G = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    for b in range(B):
        G[b, :3, :3] = np.linalg.qr(rs.normal(size=(3, 3)))[0]
        G[b, :3, 3] = rs.normal(size=3) * 0.3
    ginv = torch.from_numpy(np.linalg.inv(G).astype(np.float32)).to(DEV)
    out = st.sample([2, 0, 1, 2], c2w, focal, c, ginv=ginv, subsamplesize=s, seed=11)
    H, W = fx["mask_u8"].shape
    rg = ops.raygen(c2w, focal, c, H, W, 0.1, 10.0, pix=out["pix"], ginv=ginv)
    assert torch.equal(rg, out["rays"].view(B, s * s, 8))
    world = ops.raygen(c2w, focal, c, H, W, 0.1, 10.0, pix=out["pix"], ginv=None).cpu()
    ref = oracle.rays_to_body_space(world, torch.from_numpy(G))
    np.testing.assert_allclose(out["rays"].view(B, s * s, 8).cpu().numpy(), ref.numpy(), rtol=0, atol=2e-5)
    # colours / alphas gathered from the right stored frame
    for b, f in enumerate([2, 0, 1, 2]):
        img = np.roll(fx["img_u8"], 7 * f, axis=1) if f else fx["img_u8"]
        msk = np.roll(fx["mask_u8"], 7 * f, axis=1) if f else fx["mask_u8"]
        _, rgbs, alphas = oracle.training_sample(img, msk, out["pix"][b].cpu().numpy(), fx["c2w"], fx["focal"], fx["c"])
        assert np.array_equal(out["rgbs"][b].view(-1, 3).cpu().numpy(), rgbs.numpy())
        assert np.array_equal(out["alphas"][b].view(-1, 1).cpu().numpy(), alphas.numpy())


def test_in_kernel_draws_respect_the_candidate_maps(fx):
    """Philox mode: the first int(n*fore_rate) pixels lie in the eroded silhouette, the rest in the outside band;
    same seed -> same batch, other seed -> other batch; draws cover the lists roughly uniformly."""
    st = _store(fx, fore_erode=5)
    s = 32
    c2w, focal, c = _cam(fx, 1)
    a = st.sample([0], c2w, focal, c, subsamplesize=s, fore_rate=0.9, seed=123)
    b = st.sample([0], c2w, focal, c, subsamplesize=s, fore_rate=0.9, seed=123)
    d = st.sample([0], c2w, focal, c, subsamplesize=s, fore_rate=0.9, seed=124)
    assert torch.equal(a["pix"], b["pix"]) and torch.equal(a["rays"], b["rays"])
    assert not torch.equal(a["pix"], d["pix"])
    ins, outs = oracle.pixel_candidate_masks(np.float32(fx["mask_u8"] / 255.), 5)
    pix = a["pix"][0].cpu().numpy()
    n_fg = int(s * s * 0.9)
    assert ins[pix[:n_fg, 0], pix[:n_fg, 1]].all()
    assert outs[pix[n_fg:, 0], pix[n_fg:, 1]].all()
    # uniformity over the foreground list: 64 equal bins of list positions, chi-square with 63 dof (99.9 % < 104)
    W = fx["mask_u8"].shape[1]
    r, cc = np.where(ins)
    pos = {int(p): i for i, p in enumerate(r * W + cc)}
    big = [st.sample([0], c2w, focal, c, subsamplesize=s, fore_rate=1.0, seed=1000 + k)["pix"][0].cpu().numpy() for k in range(8)]
    k = np.array([pos[int(p[0]) * W + int(p[1])] for q in big for p in q])
    hist = np.bincount(k * 64 // len(pos), minlength=64)
    exp = len(k) / 64.0
    assert float(((hist - exp) ** 2 / exp).sum()) < 104.0
