"""CPU: training-ray sampling -- the oracle's restatement of the reference's
`get_pixelcoords` / `__getitem__` gathers against values captured from the reference itself
(tests/golden/pixel_sampling.npz), and the host-side candidate-list builder against the oracle."""
import numpy as np
import pytest
import torch

from util import oracle, load_golden


@pytest.fixture(scope="module")
def fx():
    return load_golden("pixel_sampling")


@pytest.mark.parametrize("fore_erode", [3, 5])
def test_pixelcoords_match_reference(fx, fore_erode):
    """Same numpy global-RNG seed as the generator -> identical draws: pins the morphology (incl. the even 64x64
    kernel's anchor) and the draw order of the restatement."""
    np.random.seed(5)
    coords, sel = oracle.get_pixelcoords(np.float32(fx["mask_u8"] / 255.), int(fx["n_side"]), 0.9, fore_erode)
    assert np.array_equal(coords, fx["coords_e%d" % fore_erode])
    assert sel.shape == (int(fx["n_side"]) ** 2,)


def test_training_sample_matches_reference(fx):
    rays, rgbs, alphas = oracle.training_sample(fx["img_u8"], fx["mask_u8"], fx["coords_e3"], fx["c2w"], fx["focal"], fx["c"])
    assert np.array_equal(rgbs.numpy(), fx["rgbs"]) and np.array_equal(alphas.numpy(), fx["alphas"])
    np.testing.assert_allclose(rays.numpy(), fx["rays"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("fore_erode", [3, 5])
def test_candidate_lists_match_oracle_masks(fx, fore_erode):
    from anim_nerf_b200.train_rays import candidate_lists
    m8 = torch.from_numpy(np.stack([fx["mask_u8"], fx["mask_u8"][::-1].copy()]))       # two frames
    fl, fo, bl, bo = candidate_lists(m8, fore_erode)
    W = m8.shape[2]
    for f in range(2):
        ins, outs = oracle.pixel_candidate_masks(np.float32(m8[f].numpy() / 255.), fore_erode)
        r, c = np.where(ins)
        assert np.array_equal(fl[fo[f]:fo[f + 1]].numpy(), r * W + c)
        r, c = np.where(outs)
        assert np.array_equal(bl[bo[f]:bo[f + 1]].numpy(), r * W + c)


def test_u8_normalisation_is_bit_identical_in_fp32():
    """The kernel computes float(v)/255.f; the reference float32(float64(v)/255.): same bits for all 256 values."""
    v = np.arange(256)
    assert np.array_equal(v.astype(np.float32) / np.float32(255.0), (v / 255.).astype(np.float32))


def test_empty_candidate_list_is_an_error():
    """np.random.choice(0, ...) raises ValueError in the reference (anim_nerf_dataset.py:13); so does the store."""
    from anim_nerf_b200.train_rays import DeviceFrameStore
    img = torch.zeros(1, 80, 80, 3, dtype=torch.uint8)
    with pytest.raises(ValueError):
        DeviceFrameStore(img, torch.zeros(1, 80, 80, dtype=torch.uint8), device="cpu")
