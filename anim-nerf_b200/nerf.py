"""NeRF parameter container mirroring reference `models/nerf.py:NeRF` (60-190).

The nn.Linear parameters keep the reference's names, shapes and (out,in) fp32 layout so
checkpoints stay interchangeable (state-dict keys `xyz_encoding_{1..8}.0.*`,
`xyz_encoding_final.*`, `dir_encoding.0.*`, `sigma.*`, `rgb.0.*`).  The rendering
path never runs these modules: it consumes `packed()` -- the bf16 UMMA-ready repack made by
the `an_mlp_pack` kernel, refreshed whenever a parameter's version counter moved (i.e. after
each optimiser step).  `get_sigma` / `get_normal` -- the queries of the training regularisers
(train.py:264-309) -- run on the same kernels; their torch formulations live in
the oracle (`oracle.nerf_sigma` / `oracle.nerf_normal`, pinned to the reference by a golden fixture).
"""
import torch
import torch.nn as nn

from . import ops


class Embedding(nn.Module):
    """models/embedding.py:22-39 (torch form, regulariser path only)."""

    def __init__(self, in_channels, N_freqs, logscale=True):
        super().__init__()
        self.N_freqs, self.in_channels = N_freqs, in_channels
        self.out_channels = in_channels * (2 * N_freqs + 1)
        self.freq_bands = 2 ** torch.linspace(0, N_freqs - 1, N_freqs) if logscale \
            else torch.linspace(1, 2 ** (N_freqs - 1), N_freqs)

    def forward(self, x):
        out = [x]
        for f in self.freq_bands:
            out += [torch.sin(f * x), torch.cos(f * x)]
        return torch.cat(out, -1)


class NeRF(nn.Module):
    def __init__(self, D=8, W=256, freqs_xyz=10, freqs_dir=4, use_view=True, use_normal=False,
                 deformation_dim=0, apperance_dim=0, skips=[4], actvn_type="relu"):    # noqa: B006  (the reference's defaults)
        super().__init__()
        if (D, W, freqs_xyz, tuple(skips)) != (8, 256, 10, (4,)) or use_view or use_normal \
                or deformation_dim or apperance_dim or actvn_type != "relu":
            raise NotImplementedError(
                "the sm_100a MLP kernels are built for the shipped configuration: D=8, W=256, freqs_xyz=10, "
                "skips=[4], relu, use_view=False, no latent codes (every configs/**/*.yaml of the reference)")
        self.D, self.W, self.skips = D, W, list(skips)
        self.freqs_xyz, self.freqs_dir = freqs_xyz, freqs_dir
        self.use_view, self.use_normal = use_view, use_normal
        self.deformation_dim, self.apperance_dim = deformation_dim, apperance_dim
        self.encoding_xyz = Embedding(3, freqs_xyz)
        self.in_channels_xyz = 3 + 3 * freqs_xyz * 2
        self.in_channels_dir = 0
        for i in range(D):
            fan_in = self.in_channels_xyz if i == 0 else (W + self.in_channels_xyz if i in self.skips else W)
            setattr(self, "xyz_encoding_%d" % (i + 1), nn.Sequential(nn.Linear(fan_in, W), nn.ReLU(True)))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.dir_encoding = nn.Sequential(nn.Linear(W, W // 2), nn.ReLU(True))
        self.sigma = nn.Linear(W, 1)
        self.rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
        self._packed = None
        self._packed_key = None
        self._packed_buf = None
        self._dirty = False
        self._flat_grad = None

    # ---- kernel-side view of the parameters
    def linears(self):
        return [getattr(self, "xyz_encoding_%d" % (i + 1))[0] for i in range(8)] + \
               [self.xyz_encoding_final, self.dir_encoding[0], self.sigma, self.rgb[0]]

    def param_list(self):
        """24 tensors: 12 weights then 12 biases, kernel order."""
        ls = self.linears()
        return [l.weight for l in ls] + [l.bias for l in ls]

    def mark_dirty(self):
        """Force a repack on the next `packed()` call.  The training path calls this once per step
        (`AnimNeRF.set_body_model` under grad mode): an optimiser step is not reliably visible from the
        host -- fused/capturable Adam does not bump `Tensor._version`, and a replayed CUDA graph never
        re-runs this Python -- so the repack must be an unconditional part of every training step."""
        self._dirty = True

    def packed(self):
        ps = self.param_list()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if (self._packed is None or self._dirty or key != self._packed_key
                or self._packed.device != ps[0].device):
            if self._packed_buf is None or self._packed_buf.device != ps[0].device:
                self._packed_buf = torch.empty(ops.mlp_packed_bytes() + 1024, device=ps[0].device, dtype=torch.uint8)
            # always the same storage: kernels captured in a CUDA graph keep reading the refreshed images
            self._packed = ops.mlp_pack(ps[:12], ps[12:], packed=self._packed_buf)
            self._packed_key = key
            self._dirty = False
        return self._packed

    def attach_flat_grad(self, region):
        """Training plumbing: `region` (an_mlp_grad_floats fp32, zeroed by the caller every step) becomes the
        storage of every parameter's `.grad` -- views in kernel order, which is the order the weight-gradient
        kernel writes -- and the backward of the render / query functions accumulates into it directly instead of
        returning per-tensor copies to autograd.  One memset clears all gradients, one NCCL all-reduce of the
        buffer exchanges them, and the optimiser reads the views.  `region=None` detaches."""
        self._flat_grad = region
        if region is None:
            return
        assert region.numel() == ops.mlp_grad_floats() and region.is_contiguous() and region.dtype == torch.float32
        for p, g in zip(self.param_list(), self.split_flat_grad(region)):
            p.grad = g

    def grad_sink(self):
        """The attached flat gradient buffer, or None (gradients then flow through autograd as usual)."""
        return self._flat_grad

    def split_flat_grad(self, flat):
        """flat gradient (an_mlp_grad_floats) -> 24 tensors matching param_list()."""
        ws, bs, o = [], [], 0
        for l in self.linears():
            nw = l.weight.numel()
            ws.append(flat[o:o + nw].view_as(l.weight)); o += nw
            nb = l.bias.numel()
            bs.append(flat[o:o + nb].view_as(l.bias)); o += nb
        return ws + bs

    # ---- regulariser queries (reference nerf.py:155-190) on the kernels
    def get_sigma(self, xyz, deformation_code=None, only_sigma=False):
        """Raw density of canonical-space points (nerf.py:155-175) through the CUDA MLP."""
        if not only_sigma:
            raise NotImplementedError("get_sigma(only_sigma=False) has no caller on the rendering/training path")
        from .autograd import mlp_query
        lead = xyz.shape[:-1]
        return mlp_query(self, xyz.reshape(1, -1, 3))[1].view(*lead, 1)

    def get_normal(self, xyz, deformation_code=None, delta=0.02):
        """d alpha/d xyz with alpha = 1 - exp(-delta * relu(sigma)) (nerf.py:177-190), differentiable w.r.t. the
        weights: d alpha/d xyz = delta * exp(-delta * relu(sigma)) * [sigma > 0] * d sigma/d xyz, with
        (sigma, d sigma/d xyz) from `SigmaWithGradient` (forward + dgrad kernels; its backward is the
        tangent + wgrad kernels instead of torch double backward)."""
        from .autograd import sigma_with_gradient
        sigma, s = sigma_with_gradient(self, xyz)
        return delta * torch.exp(-delta * torch.relu(sigma)) * (sigma > 0).to(sigma.dtype) * s

    def forward(self, xyz, viewdir=None, deformation_code=None, apperance_code=None):
        """Canonical-space query through the CUDA MLP (no unposing)."""
        from .autograd import mlp_query
        return mlp_query(self, xyz)
