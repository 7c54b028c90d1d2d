"""2 GPUs over NCCL (skipped on a 1-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_zzzz_nccl_gpu.py`): the
all-reduced gradient bucket of the data-parallel Engine equals (a) the gradient digests of the unmodified reference
(tests/golden/model_choopy_B5.npz) when every group holds the golden's lists, and (b) the mean of the per-group
nn.Module-path gradients for distinct, UNEVENLY sharded groups (3 groups: ranks own 2 + 1); then one fused Adam step leaves
identical parameters on both ranks.  The CPU twin of the host logic is tests/test_parallel_cpu.py (gloo)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    from pathlib import Path
    here = Path(__file__).resolve().parent
    sys.path.insert(0, str(here))
    sys.path.insert(0, str(here.parent / "ranked-list-truncation_b200"))
    sys.path.insert(0, str(here.parent))
    import torch.distributed as dist
    from helpers import build_model, grad_errors, load_golden
    from rlt_b200 import parallel
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    from rlt_b200.optim import FusedAdam
    from utils import losses
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    res = {}
    try:
        S, L = 5, 300
        g = load_golden("model_choopy_B5.npz")
        gx, gy = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
        # ---- (a) every group = the golden's five lists: the reduced bucket must reproduce the reference's gradients
        n_groups = 3
        mine = parallel.shard_groups(n_groups, rank, world)
        model = build_model("choopy").cuda().train()
        eng = Engine(model, n_groups=len(mine), group_size=S, seq_len=L)
        eng.train_step(gx.repeat(len(mine), 1, 1), gy.repeat(len(mine), 1))
        parallel.allreduce_mean_(eng.grad_bucket, len(mine), n_groups)
        rel_l2, rel_max, _ = grad_errors({n: eng.grads[n] for n, _ in model.named_parameters()}, g)
        res["golden"] = (rel_l2, rel_max)
        # ---- (b) distinct groups, uneven shards, against the module path group by group (every rank computes the expectation)
        x, y = synthetic_lists(n_groups * S, L, 1, seed=77, device="cuda")
        crit = losses.ChoopyLoss(metric="f1").cuda()
        expect = {n: torch.zeros_like(p) for n, p in model.named_parameters()}
        for grp in range(n_groups):
            model.zero_grad(set_to_none=True)
            crit(model(x[grp * S:(grp + 1) * S]), y[grp * S:(grp + 1) * S]).backward()
            for n, p in model.named_parameters():
                expect[n] += p.grad / n_groups
        model.zero_grad(set_to_none=True)
        xs, ys = parallel.shard_lists(x, y, S, rank, world)
        eng.train_step(xs, ys)
        local = eng.grad_bucket.clone()
        parallel.allreduce_mean_(eng.grad_bucket, len(mine), n_groups)
        gmax = max(v.abs().max().item() for v in expect.values())
        res["mean"] = max((eng.grads[n] - expect[n]).abs().max().item() for n in expect) / gmax
        res["local_differs"] = (local - eng.grad_bucket).abs().max().item() / gmax     # the reduction did something
        # ---- the bucket is bit-identical on both ranks, and so are the parameters after the fused Adam step
        opt = FusedAdam.for_engine(eng, lr=3e-5, weight_decay=1e-3)
        opt.step()
        flat = torch.cat([eng.grad_bucket.flatten()] + [p.detach().flatten() for p in model.parameters()])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        res["identical"] = bool(torch.equal(both[0], both[1]))
    finally:
        out[rank] = res
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_nccl_reduced_bucket_is_the_mean_of_the_per_group_gradients():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        out = m.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        out = dict(out)
    for r in range(world):
        rel_l2, rel_max = out[r]["golden"]
        assert rel_l2 <= 2e-3 and rel_max <= 1e-3, (r, out[r])
        assert out[r]["mean"] <= 1e-3, (r, out[r])
        assert out[r]["local_differs"] > 1e-3, (r, out[r])
        assert out[r]["identical"], (r, out[r])
