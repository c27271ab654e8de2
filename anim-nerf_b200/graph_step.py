"""Training step replayed from CUDA graphs.

One step of the hot path is ~270 kernel launches, most of them microseconds long (per-frame
table builder, autograd glue, loss, Adam): launched one by one from Python the GPU idles between
them.  `GraphedTrainStep` captures the whole step -- tables -> render -> loss -> backward ->
optimiser -> weight repack happens at the start of the next replay -- once, and replays it with
one launch.  Requirements, all met by this package's path: no host synchronisation inside the
step (`affine_inverse` instead of `torch.inverse`), randomness from torch's device generator
(`VolumeRenderer.device_rng = True`: graph-safe Philox offsets), a `capturable=True` optimiser,
static input buffers (`batch` tensors are copied into them before every replay).

N>1 (one process per GPU): the step is captured as two graphs around the single NCCL
all-reduce of the flat MLP-gradient bucket, which stays an ordinary stream-ordered
NCCL call between them.
"""
import torch
import torch.distributed as dist


class GraphedTrainStep:
    def __init__(self, loss_fn, optimizer, params, example_batch, world=1, warmup=3):
        """loss_fn(batch: dict of device tensors) -> scalar loss tensor (forward only);
        params: the tensors whose .grad is all-reduced when world > 1."""
        self.loss_fn, self.opt, self.params, self.world = loss_fn, optimizer, list(params), world
        self.static = {k: v.clone() for k, v in example_batch.items()}
        self.bucket = None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                      # warm-up on a side stream (allocator, lazy inits)
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.g_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_a):
            self.opt.zero_grad(set_to_none=True)
            self.loss = self.loss_fn(self.static)
            self.loss.backward()
            if world > 1:
                self.bucket = torch.cat([p.grad.reshape(-1) for p in self.params])
            else:
                self.opt.step()
        self.g_b = None
        if world > 1:
            self.g_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_b, pool=self.g_a.pool()):
                self.bucket /= world
                o = 0
                for p in self.params:
                    n = p.numel()
                    p.grad.copy_(self.bucket[o:o + n].view_as(p))
                    o += n
                self.opt.step()

    def _eager(self):
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.static)
        loss.backward()
        if self.world > 1:
            bucket = torch.cat([p.grad.reshape(-1) for p in self.params])
            dist.all_reduce(bucket)
            bucket /= self.world
            o = 0
            for p in self.params:
                n = p.numel()
                p.grad.copy_(bucket[o:o + n].view_as(p))
                o += n
        self.opt.step()
        return loss

    def __call__(self, batch=None):
        """Copies `batch` (device or pinned-host tensors) into the static buffers, replays the step and
        returns the (device) loss tensor of this step."""
        if batch is not None:
            for k, v in batch.items():
                self.static[k].copy_(v, non_blocking=True)
        if hasattr(self.opt, "sync_lr"):
            self.opt.sync_lr()              # FusedAdam: the schedule's current lr reaches the captured kernel through a device scalar
        self.g_a.replay()
        if self.world > 1:
            dist.all_reduce(self.bucket)
            self.g_b.replay()
        return self.loss
