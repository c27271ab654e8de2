"""Device-resident training frames and per-step ray sampling.

Mirrors the training branch of the reference dataset's `__getitem__`
(`datasets/anim_nerf_dataset.py:235-262`) and `get_pixelcoords(..., 'foreground_pixel')` (`:10-54`):
per frame, `subsamplesize**2` pixels, `fore_rate` of them drawn (with replacement) from the silhouette
eroded by `fore_erode`, the rest from the band between its 64-px and `fore_erode`-px dilations; rays,
colours (`img*mask`, white background) and alphas gathered at those pixels.

The reference does this on CPU dataloader workers, building all H*W rays of a frame to keep 1024 and
re-running three cv2 morphology passes per item.  Here the uint8 frames live in HBM, the two candidate
maps -- which depend only on the frame -- are reduced once at load time to per-frame index lists, and one
kernel launch per step (`an_sample_training_rays_fwd`) draws the pixels, gathers colour/alpha and
generates only the sampled rays, already in body space.  No per-step host work, no H2D copy.
"""
import torch
import torch.nn.functional as F

from . import ops


def _morph(m, k, erode):
    """cv2.erode / cv2.dilate with a k x k ones kernel (anchor k//2, out-of-image pixels ignored) on (F,H,W)
    float maps, as min / max pooling with -inf padding (datasets/anim_nerf_dataset.py:32-37)."""
    a = k // 2
    x = (-m if erode else m).unsqueeze(1)
    x = F.pad(x, (a, k - 1 - a, a, k - 1 - a), value=float("-inf"))
    x = F.max_pool2d(x, kernel_size=k, stride=1).squeeze(1)
    return -x if erode else x


def candidate_lists(masks_u8, fore_erode=3):
    """masks (F,H,W) uint8 -> (fg_list, fg_off, bg_list, bg_off): CSR lists of linear pixel ids (row*W+col, row-major
    = np.where order) with mask_inside > 0 / mask_outside > 0 per frame; int32 tensors on the masks' device."""
    m = masks_u8.float() / 255.0
    inside = _morph(m, fore_erode, True) > 0
    outside = (_morph(m, 64, False) - _morph(m, fore_erode, False)) > 0
    out = []
    for sel in (inside, outside):
        flat = sel.flatten(1)
        cnt = flat.sum(1)
        off = torch.zeros(flat.shape[0] + 1, dtype=torch.int64, device=flat.device)
        off[1:] = torch.cumsum(cnt, 0)
        ids = torch.nonzero(flat)[:, 1]
        out += [ids.to(torch.int32).contiguous(), off.to(torch.int32).contiguous()]
    return tuple(out)


class DeviceFrameStore:
    """Training frames resident on one GPU.  images (F,H,W,3) uint8 RGB, masks (F,H,W) uint8 (the files the
    reference reads with cv2, before its /255)."""

    def __init__(self, images_u8, masks_u8, device=None, fore_erode=3, white_bkgd=True, with_background=False):
        device = device or images_u8.device
        assert images_u8.dtype == torch.uint8 and masks_u8.dtype == torch.uint8
        assert images_u8.shape[:3] == masks_u8.shape and images_u8.shape[3] == 3
        self.images = images_u8.to(device).contiguous()
        self.masks = masks_u8.to(device).contiguous()
        self.F, self.H, self.W = masks_u8.shape
        self.white_bkgd, self.with_background, self.fore_erode = white_bkgd, with_background, fore_erode
        self.fg_list, self.fg_off, self.bg_list, self.bg_off = candidate_lists(self.masks, fore_erode)
        fg_n = (self.fg_off[1:] - self.fg_off[:-1]).cpu()
        bg_n = (self.bg_off[1:] - self.bg_off[:-1]).cpu()
        if int(fg_n.min()) == 0 or int(bg_n.min()) == 0:
            # the reference fails the same way: np.random.choice(0, ...) raises ValueError (anim_nerf_dataset.py:13)
            raise ValueError("a frame has no foreground or no outside-band pixels to sample from")
        self.fg_count, self.bg_count = fg_n, bg_n

    def sample(self, frame_ids, c2w, focal, center, ginv=None, subsamplesize=32, fore_rate=0.9, near=0.1, far=10.0,
               sel=None, seed=0):
        """frame_ids (B) ints; c2w (B,3,4), focal (B,2), center (B,2); ginv (B,4,4) or None.
        Returns dict rays (B,s,s,8), rgbs (B,s,s,3), alphas (B,s,s,1), pix (B,s*s,2)."""
        s = subsamplesize
        n = s * s
        n_fg = int(n * fore_rate)
        rays, rgbs, alphas, pix = ops.sample_training_rays(self, frame_ids, c2w, focal, center, ginv, n, n_fg, near, far,
                                                           sel=sel, seed=seed)
        B = rays.shape[0]
        return dict(rays=rays.view(B, s, s, 8), rgbs=rgbs.view(B, s, s, 3), alphas=alphas.view(B, s, s, 1), pix=pix)
