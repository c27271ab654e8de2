"""Training step replayed from CUDA graphs.

One step of the hot path is ~270 kernel launches, most of them microseconds long (per-frame
table builder, autograd glue, loss, Adam): launched one by one from Python the GPU idles between
them.  `GraphedTrainStep` captures the whole step -- tables -> render -> loss -> backward ->
optimiser -> weight repack happens at the start of the next replay -- once, and replays it with
one launch.  Requirements, all met by this package's path: no host synchronisation inside the
step (`affine_inverse` instead of `torch.inverse`), randomness from torch's device generator
(`VolumeRenderer.device_rng = True`: graph-safe Philox offsets), a `capturable=True` optimiser,
static input buffers (`batch` tensors are copied into them before every replay).

N>1 (one process per GPU): with a `FlatGradBuffer` the single NCCL all-reduce of the step's flat
gradient bucket is captured inside the one graph of the step; without one the step is
captured as two graphs around an ordinary stream-ordered NCCL call.
"""
import torch
import torch.distributed as dist


def _tree_map(fn, x):
    if isinstance(x, dict):
        return {k: _tree_map(fn, v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_tree_map(fn, v) for v in x)
    return fn(x) if torch.is_tensor(x) else x


def _tree_copy(dst, src):
    if isinstance(src, dict):
        for k, v in src.items():
            _tree_copy(dst[k], v)
    elif isinstance(src, (list, tuple)):
        for d, v in zip(dst, src):
            _tree_copy(d, v)
    elif torch.is_tensor(src):
        dst.copy_(src, non_blocking=True)


class GraphedTrainStep:
    def __init__(self, loss_fn, optimizer, params, example_batch, world=1, warmup=3, flat=None, renderer=None, model=None):
        """loss_fn(batch: (nested) dict of device tensors) -> scalar loss tensor (forward only).
        params: every tensor whose .grad must be exchanged when world > 1 (all optimiser parameters, the SMPL table
        included when it is optimised -- otherwise ranks drift apart); flat: the `FlatGradBuffer` holding those
        gradients -- then the exchange is one all-reduce of that buffer, captured INSIDE the single graph of the step,
        and zero_grad is one memset; without it the gradients are flattened / copied back around an eager NCCL call
        between two graphs.  renderer: the VolumeRenderer used by loss_fn, checked for graph-safe randomness.
        model: the AnimNeRF used by loss_fn; its per-frame state of earlier (eager) steps is dropped first, so that no
        old autograd graph ties parameter AccumulateGrad nodes to the default stream."""
        if renderer is not None and not getattr(renderer, "device_rng", False):
            raise ValueError("GraphedTrainStep: set VolumeRenderer.device_rng = True -- a host-drawn Philox seed would be "
                             "frozen into the captured graph and every replay would reuse the same noise")
        self.loss_fn, self.opt, self.params, self.world, self.flat = loss_fn, optimizer, list(params or ()), world, flat
        self.static = _tree_map(lambda t: t.clone(), example_batch)
        self.bucket = None
        self._staging = None
        if model is not None:
            model.clear_frame_state()
        if flat is not None and hasattr(optimizer, "flat"):
            optimizer.flat = flat
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                      # warm-up on a side stream (allocator, lazy inits, NCCL communicator)
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if model is not None:
            model.clear_frame_state()                   # ... nor to the warm-up stream
        self.g_a = torch.cuda.CUDAGraph()
        self.g_b = None
        if flat is not None:
            with torch.cuda.graph(self.g_a):
                flat.zero()
                loss = self.loss_fn(self.static)
                loss.backward()
                flat.all_reduce(world)                  # NCCL all-reduce captured in the graph (no-op for world == 1)
                self.opt.step()
                self.loss = loss.detach()               # the value only: the step's autograd graph is not kept alive
            del loss
            if model is not None:
                model.clear_frame_state()
            return
        with torch.cuda.graph(self.g_a):
            self.opt.zero_grad(set_to_none=True)
            self.loss = self.loss_fn(self.static)
            self.loss.backward()
            if world > 1:
                self.bucket = torch.cat([p.grad.reshape(-1) for p in self.params])
            else:
                self.opt.step()
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
        if self.flat is not None:
            self.flat.zero()
            loss = self.loss_fn(self.static)
            loss.backward()
            self.flat.all_reduce(self.world)
            self.opt.step()
            return loss
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
        """Copies `batch` (device or pinned-host tensors, same nesting as the example) into the static buffers,
        replays the step and returns the (device) loss tensor of this step."""
        if batch is not None:
            _tree_copy(self.static, batch)
        if hasattr(self.opt, "sync_lr"):
            self.opt.sync_lr()              # FusedAdam: the schedule's current lr reaches the captured kernel through a device scalar
        self.g_a.replay()
        if self.g_b is not None:
            dist.all_reduce(self.bucket)
            self.g_b.replay()
        for fn in getattr(self.opt, "on_step", ()):     # a replay runs no Python: invalidate the packed weights here
            fn()
        return self.loss

    # ---- pipelined input feed: the next batch's host->device copy runs on a copy stream while the current step computes,
    # and a step's loss is read back without stalling the launch of the next one
    def stage(self, batch):
        """Start copying `batch` (pinned host tensors, same nesting as the example) into the staging buffers on the copy
        stream; returns at once.  `run_staged()` consumes it."""
        if self._staging is None:
            self._staging = _tree_map(lambda t: torch.empty_like(t), self.static)
            self._copy_stream = torch.cuda.Stream()
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record()
        self._copy_stream.wait_event(self._consumed)        # the previous staging content has been moved on
        with torch.cuda.stream(self._copy_stream):
            _tree_copy(self._staging, batch)
            self._staged.record()

    def run_staged(self):
        """Replay the step on the batch handed to `stage()`; returns a `StepResult` whose `.value()` is the step's loss
        (a host float; blocks only until THIS step has finished, so it can be called one step late without idling the GPU)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        _tree_copy(self.static, self._staging)              # device -> device, a few microseconds
        self._consumed.record(cur)
        self(None)
        res = StepResult(self.loss)
        return res


class StepResult:
    """Loss of one replayed step on its way to the host: a pinned scalar filled by an async copy + the event after it."""

    def __init__(self, loss_dev):
        self.host = torch.empty((), dtype=loss_dev.dtype, pin_memory=True)
        self.host.copy_(loss_dev, non_blocking=True)
        self.done = torch.cuda.Event()
        self.done.record()

    def value(self):
        self.done.synchronize()
        return float(self.host)
