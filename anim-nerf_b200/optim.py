"""Adam on the kernels: `FusedAdam` is `torch.optim.Adam` as the reference configures it (train.py:217-226,
utils/__init__.py:33-45: eps 1e-8, betas (0.9, 0.999), L2 weight decay, no amsgrad) with the update of a whole
parameter group done by one `an_adam_step` launch (<= 64 tensors per launch) instead of torch's multi-tensor chain.

Same constructor arguments, `param_groups` (so LR schedulers work unchanged), `zero_grad`, and torch.optim.Adam's
`state_dict` layout (`state[p]["step"]`, `["exp_avg"]`, `["exp_avg_sq"]`): a checkpoint written by either optimiser
resumes in the other.  Graph-capturable: the step count lives in device memory (one scalar per launch, which
`state[p]["step"]` aliases) and is advanced by the kernel; the learning rate is read from a device scalar that
`sync_lr()` refreshes from `group["lr"]` (called by `step()` outside capture, and by `GraphedTrainStep` before every
replay).  The parameters of a group that take part in the update are fixed by its first step (torch.optim.Adam skips a
parameter whose grad is None and keeps a per-parameter step; here a launch shares one step, so a later change of the
set raises instead of drifting)."""
import ctypes

import torch

from ._lib import call, ptr, stream

_MAX = 64


class FlatGradBuffer:
    """One contiguous fp32 buffer holding every gradient of a training step: per NeRF the gradient vector the
    weight-gradient kernel writes (an_mlp_grad_floats: 592 388 parameters + the fused head layer's scratch), then
    any further parameters (the per-frame SMPL table of `optim_body_params`: <= 114 x 75 + 10 floats).
    `NeRF.attach_flat_grad` makes the MLP parameters' `.grad` views of it and routes the kernels' accumulation
    there; `zero()` is the step's one memset, `all_reduce()` the step's one exchange (NCCL all-reduce of
    the flat bucket, averaged; the body-parameter gradients ride in the same bucket)."""

    def __init__(self, nets, extra_params=()):
        from . import ops
        nets = list(dict.fromkeys(nets))              # share_fine: nerf_fine is nerf
        extra = [p for p in extra_params if p.requires_grad]
        dev = next(nets[0].parameters()).device
        gf = ops.mlp_grad_floats()
        per_net = (gf + 3) // 4 * 4                   # 16-byte aligned regions
        sizes = [(p.numel() + 3) // 4 * 4 for p in extra]
        self.buf = torch.zeros(per_net * len(nets) + sum(sizes), device=dev)
        self.nets, self.extra = nets, extra
        o = 0
        for net in nets:
            net.attach_flat_grad(self.buf[o:o + gf])
            o += per_net
        for p, n in zip(extra, sizes):
            p.grad = self.buf[o:o + p.numel()].view_as(p)
            o += n

    def zero(self):
        self.buf.zero_()

    def all_reduce(self, world=None):
        import torch.distributed as dist
        world = world or (dist.get_world_size() if dist.is_initialized() else 1)
        if world > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.AVG)

    def detach(self):
        for net in self.nets:
            net.attach_flat_grad(None)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._g = {}        # group index -> dict(lr_t, lr_host, members, chunks: list of dict(step, done))
        self.flat = None    # FlatGradBuffer when the gradients live in one buffer (zero_grad is then one memset)
        self.on_step = []   # callbacks after every update (the NeRFs' packed-weight invalidation)

    def zero_grad(self, set_to_none=True):
        if self.flat is not None:
            self.flat.zero()       # .grad tensors are views of the flat buffer: they must stay in place
            return
        super().zero_grad(set_to_none=set_to_none)

    def _gstate(self, gi, device, ps):
        """Device scalars of group gi, created at its first step over the parameters `ps` (or after load_state_dict: the
        step count then continues from the loaded state[p]["step"])."""
        st = self._g.get(gi)
        if st is not None and st["lr_t"].device == device:
            if st["members"] != [id(p) for p in ps]:
                raise RuntimeError("FusedAdam: the set of parameters with gradients in group %d changed after its first step "
                                   "(one device step count is shared per launch)" % gi)
            return st
        chunks = []
        for ci in range((len(ps) + _MAX - 1) // _MAX):
            loaded = [float(self.state[p]["step"]) for p in ps[ci * _MAX:(ci + 1) * _MAX] if "step" in self.state[p]]
            if loaded and min(loaded) != max(loaded):
                raise RuntimeError("FusedAdam: loaded state has different step counts inside one launch group")
            step = torch.full((1,), loaded[0] if loaded else 0.0, device=device)
            for p in ps[ci * _MAX:(ci + 1) * _MAX]:
                self.state[p]["step"] = step[0]          # a view: state_dict() sees the kernel's count
            chunks.append(dict(step=step, done=torch.zeros(1, device=device, dtype=torch.int32)))
        st = dict(lr_t=torch.zeros(1, device=device), lr_host=None, members=[id(p) for p in ps], chunks=chunks)
        self._g[gi] = st
        return st

    def state_dict(self):
        """torch.optim.Adam's layout; the step counts as host scalars (what a non-capturable torch.optim.Adam expects)."""
        sd = super().state_dict()
        sd["state"] = {k: {n: (v.detach().to("cpu", copy=True) if n == "step" and torch.is_tensor(v) else v) for n, v in s.items()}
                       for k, s in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        """torch.optim.Adam / Lightning checkpoints load as they are; the device step scalars are rebuilt from the
        loaded per-parameter steps at the next step()."""
        super().load_state_dict(state_dict)
        self._g = {}

    def sync_lr(self):
        """Copy every group's `lr` to its device scalar when it changed (a host-side write: not inside graph capture)."""
        for gi, group in enumerate(self.param_groups):
            st = self._g.get(gi)
            if st is not None and st["lr_host"] != group["lr"]:
                st["lr_t"].fill_(float(group["lr"]))
                st["lr_host"] = group["lr"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise RuntimeError("FusedAdam runs on the CUDA library only; use torch.optim.Adam for CPU tensors")
            if capturing is None:
                capturing = torch.cuda.is_current_stream_capturing()
            n_chunks = (len(ps) + _MAX - 1) // _MAX
            if gi not in self._g and capturing:
                raise RuntimeError("FusedAdam: take one eager step before capturing a CUDA graph")
            st = self._gstate(gi, dev, ps)
            if not capturing:
                self.sync_lr()
            elif st["lr_host"] is None:
                raise RuntimeError("FusedAdam: call sync_lr() (or take one eager step) before capturing a CUDA graph")
            b1, b2 = group["betas"]
            for ci in range(n_chunks):
                chunk = ps[ci * _MAX:(ci + 1) * _MAX]
                P, G, M, V = [], [], [], []
                for p in chunk:
                    if p.dtype != torch.float32 or not p.is_contiguous():
                        raise RuntimeError("FusedAdam expects contiguous fp32 parameters")
                    s = self.state[p]
                    if "exp_avg" not in s:
                        s["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                        s["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    g = p.grad
                    if g.dtype != torch.float32 or not g.is_contiguous():
                        g = g.float().contiguous()
                    P.append(p); G.append(g); M.append(s["exp_avg"]); V.append(s["exp_avg_sq"])
                n = len(chunk)
                arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])            # noqa: E731
                sizes = (ctypes.c_int64 * n)(*[p.numel() for p in P])
                cs = st["chunks"][ci]
                call("an_adam_step", arr(P), arr(G), arr(M), arr(V), sizes, n, ptr(cs["step"]), ptr(st["lr_t"]),
                     float(group["lr"]), float(b1), float(b2), float(1.0 - b1), float(1.0 - b2), float(group["eps"]),
                     float(group["weight_decay"]),
                     ptr(cs["done"]), stream())
        for fn in self.on_step:
            fn()
        return loss
