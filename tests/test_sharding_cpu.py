"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame/ray sharding, the single flat
gradient all-reduce of the training step, and slab gathering for sharded inference."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT  # noqa: F401  (puts the repo on sys.path)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import anim_nerf_b200  # noqa: F401
    from anim_nerf_b200 import dist_utils as du
    from anim_nerf_b200.nerf import NeRF
    torch.manual_seed(0)
    net = NeRF(freqs_dir=0, use_view=False)      # same init on both ranks
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    du.allreduce_grads(list(net.parameters()))
    ok_grad = all(torch.allclose(p.grad, torch.full_like(p, (world + 1) / 2.0 * (i + 1))) for i, p in enumerate(net.parameters()))
    a, b = du.shard_range(17, rank, world)
    local = torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 3)
    full = du.gather_slabs(local, 17)
    ok_gather = torch.equal(full[:, 0], torch.arange(17, dtype=torch.float32))
    # unequal slabs with the remainder on the first ranks (16 rows over 3 would be 6,5,5; here 2 ranks: 5 -> 3,2)
    for n in (5, 16):
        a5, b5 = du.shard_range(n, rank, world)
        img = torch.arange(n * 4, dtype=torch.float32).view(n, 4)
        ok_gather = ok_gather and torch.equal(du.gather_slabs(img[a5:b5].clone(), n), img)
        # interleaved rows (rank r: r, r+world, ...): gathered back in image order, also when n % world != 0
        rows = du.shard_rows(n, rank, world)
        ok_gather = ok_gather and rows == list(range(rank, n, world)) and torch.equal(du.gather_rows(img[rows].clone(), n), img)
    # the training exchange: ONE all-reduce (average) of the flat gradient buffer the kernels accumulate into --
    # both MLPs' gradient vectors and the SMPL table's rows; afterwards every parameter's .grad (a view) is the mean
    from anim_nerf_b200.optim import FlatGradBuffer
    net2 = NeRF(freqs_dir=0, use_view=False)
    table = torch.nn.Parameter(torch.zeros(7, 69))
    frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)
    fb = FlatGradBuffer([net, net2], [table, frozen])
    ok_flat = frozen.grad is None and table.grad.data_ptr() >= fb.buf.data_ptr()
    for i, p in enumerate(list(net.parameters()) + list(net2.parameters()) + [table]):
        ok_flat = ok_flat and p.grad.data_ptr() >= fb.buf.data_ptr() and p.grad.shape == p.shape
        p.grad.fill_(float(rank + 1) * (i + 1))
    fb.all_reduce(world)
    for i, p in enumerate(list(net.parameters()) + list(net2.parameters()) + [table]):
        ok_flat = ok_flat and torch.allclose(p.grad, torch.full_like(p, (world + 1) / 2.0 * (i + 1)))
    fb.zero()
    ok_flat = ok_flat and float(table.grad.abs().sum()) == 0.0 and float(net.sigma.weight.grad.abs().sum()) == 0.0
    q.put((rank, ok_grad and ok_flat, (a, b), ok_gather))
    dist.destroy_process_group()


def _spawn(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    return res


def test_two_rank_allreduce_and_sharding():
    res = _spawn(2)
    assert all(r[1] for r in res), res
    assert [r[2] for r in res] == [(0, 9), (9, 17)]
    assert all(r[3] for r in res), res


def test_three_rank_unequal_slabs_gather():
    res = _spawn(3)          # 16 rows -> 6,5,5: the padded slab is not the last one
    assert [r[2] for r in res] == [(0, 6), (6, 12), (12, 17)]
    assert all(r[3] for r in res), res


def test_shard_range_partitions_everything():
    from anim_nerf_b200 import dist_utils as du
    for n in (1, 7, 16, 120, 262144):
        for w in (1, 2, 4, 8):
            spans = [du.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
