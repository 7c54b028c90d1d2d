"""CPU, world_size 2 over gloo: group sharding and the bucket all-reduce reproduce the single-process mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rlt_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_groups, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, L = 4, 10
    torch.manual_seed(0)
    x = torch.randn(n_groups * S, L, 3)
    y = (torch.rand(n_groups * S, L) < 0.3).float()
    xs, ys = parallel.shard_lists(x, y, S, rank, world)
    mine = parallel.shard_groups(n_groups, rank, world)
    assert xs.shape[0] == len(mine) * S
    # a stand-in "per-group gradient": any deterministic function of the group's lists
    def group_grad(g):
        xb, yb = x[g * S:(g + 1) * S], y[g * S:(g + 1) * S]
        return torch.stack([xb.sum(), (xb[..., 0] * yb).sum(), yb.sum(), torch.tensor(float(g))])
    local = torch.stack([group_grad(g) for g in mine]).mean(0) if mine else torch.zeros(4)
    # check the shard holds exactly those groups, in order
    for i, g in enumerate(mine):
        assert torch.equal(xs[i * S:(i + 1) * S], x[g * S:(g + 1) * S])
    parallel.allreduce_mean_(local, len(mine), n_groups)
    expect = torch.stack([group_grad(g) for g in range(n_groups)]).mean(0)
    ok = torch.allclose(local, expect, rtol=1e-5, atol=1e-5)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _run(n_groups):
    world = 2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        out = m.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_groups, out)) for port in [_free_port()] for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert dict(out) == {0: True, 1: True}


def test_even_and_uneven_group_shards_world2():
    _run(6)   # 3 + 3 groups
    _run(5)   # 3 + 2 groups: weighted mean must still equal the global mean


def test_shard_groups_partition():
    for n in (1, 5, 8, 64):
        for w in (1, 2, 4, 8):
            parts = [parallel.shard_groups(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _loader_worker(rank, world, port, n_lists, bs, out):
    from rlt_b200.data import batch_slices, shuffled_order
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(99)                      # every rank seeds alike: the epoch order needs no communication
    ok = True
    for _ in range(2):
        order = shuffled_order(n_lists)
        mine = batch_slices(n_lists, bs, rank, world)
        steps = torch.tensor([len(mine)])
        all_steps = [torch.zeros_like(steps) for _ in range(world)]
        dist.all_gather(all_steps, steps)
        ok &= len({int(s) for s in all_steps}) == 1                      # same number of steps (all-reduces) everywhere
        seen = torch.full((n_lists,), -1, dtype=torch.long)
        for lo, hi in mine:
            seen[order[lo:hi]] = rank
        gathered = [torch.zeros_like(seen) for _ in range(world)]
        dist.all_gather(gathered, seen)
        owners = torch.stack(gathered)                                   # [world, n_lists]
        ok &= bool(((owners >= 0).sum(0) <= 1).all())                    # no list collated twice
        covered = int((owners >= 0).any(0).sum())
        n_b = (n_lists + bs - 1) // bs
        kept = n_b - n_b % world
        ok &= covered == min(n_lists, kept * bs)                         # exactly the complete rounds of W batches
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_device_loader_batches_shard_over_ranks_world2():
    world = 2
    ctx = mp.get_context("spawn")
    for n_lists, bs in ((249, 64), (250, 25), (64, 64)):
        with ctx.Manager() as m:
            out = m.dict()
            port = _free_port()
            procs = [ctx.Process(target=_loader_worker, args=(r, world, port, n_lists, bs, out)) for r in range(world)]
            for p in procs:
                p.start()
            for p in procs:
                p.join(120)
                assert p.exitcode == 0
            assert dict(out) == {0: True, 1: True}, (n_lists, bs)


def test_batch_slices_single_process():
    from rlt_b200.data import batch_slices
    assert batch_slices(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert batch_slices(10, 4, 0, 2) == [(0, 4)] and batch_slices(10, 4, 1, 2) == [(4, 8)]
    assert batch_slices(10, 4, 0, 2, drop_uneven=False) == [(0, 4), (8, 10)]
    assert batch_slices(10, 4, 1, 2, drop_uneven=False) == [(4, 8)]
    assert batch_slices(3, 4, 1, 2) == []
