"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Rays shard naturally: every (frame, ray) is independent given the per-frame tables.
Training needs exactly one exchange per step -- the sum of the MLP gradients -- done as a single
all-reduce over one flat fp32 bucket (2 x 592 388 floats = 4.74 MB for nerf + nerf_fine);
inference and grid queries need no collective (each rank writes its own slab).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous block partition of n_items (frames, image rows, grid slabs) -> (start, stop)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rows(n_rows, rank, world):
    """Interleaved partition of image rows: rank r owns rows r, r + world, r + 2 world, ...  The body covers a band
    of rows in the middle of a frame, so contiguous slabs are load-imbalanced (foreground fraction 0.22-0.28 of a slab
    at 8 ranks, some slabs nearly empty); interleaving gives every rank the same share of every region."""
    return list(range(rank, n_rows, world))


def allreduce_grads(params, world=None, average=True):
    """Sum (and average) the .grad of `params` across ranks through one flat bucket."""
    world = world or (dist.get_world_size() if dist.is_initialized() else 1)
    if world == 1:
        return
    params = [p for p in params if p.grad is not None]
    bucket = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    if average:
        bucket /= world
    o = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(bucket[o:o + n].view_as(p))
        o += n


def gather_rows(local, n_rows, dim=0):
    """all_gather of the per-rank row sets of a `shard_rows` partition along `dim`, back in image order."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    per = (n_rows + world - 1) // world
    pad = per - local.shape[dim]
    if pad:
        shape = list(local.shape); shape[dim] = pad
        local = torch.cat([local, local.new_zeros(shape)], dim)
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local.contiguous())
    full = torch.stack(out, dim + 1)                       # (.., per, world, ..): row = i * world + r
    shape = list(local.shape); shape[dim] = per * world
    return full.reshape(shape).narrow(dim, 0, n_rows)


def gather_slabs(local, n_total, dim=0):
    """Inference: all_gather the per-rank slabs of a `shard_range(n_total, rank, world)` partition along
    `dim` (slabs are padded to the largest one for the collective, the padding is dropped afterwards)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [b - a for a, b in (shard_range(n_total, r, world) for r in range(world))]
    per = max(sizes)
    pad = per - local.shape[dim]
    if pad:
        shape = list(local.shape); shape[dim] = pad
        local = torch.cat([local, local.new_zeros(shape)], dim)
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local.contiguous())
    return torch.cat([o.narrow(dim, 0, n) for o, n in zip(out, sizes)], dim)
