#!/usr/bin/env python
"""bench.py — ranked lists/sec of the truncation-model hot path (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[1]): Choopy cut-transformer, score-only input, 65 536 synthetic
robust04-shaped lists x 300 resident in HBM; attention groups of S = 64 lists (the reference's batch).
A "step" = one pass of the hot path over one batch of G groups per GPU (default 64 -> 4096 lists):
forward + ChoopyLoss + backward into the flat gradient bucket (+ NCCL all-reduce of the bucket when
N > 1).  Inference (forward + fused argmax-cut + F1/DCG) is timed as well and reported in `inference`.

One JSON line on stdout (rank 0).  See DESIGN.md section "Measurement" for every field.
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--model choopy] [--groups G]
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

DATASET_LISTS = 65536
SEQ_LEN = 300
GROUP = 64
METRIC = "ranked lists/sec (train step, L=300)"


N_FEATURES = {"choopy": 1, "mtchoopy": 1, "bicut": 3, "attncut": 3, "mtattncut": 3, "mmoecut": 3}
CRITERION = {"choopy": "ChoopyLoss f1", "mtchoopy": "MtCutLoss f1", "bicut": "BiCutLoss", "attncut": "DivLoss js f1",
             "mtattncut": "MtCutLoss f1", "mmoecut": "MtCutLoss f1"}


def build_model(models, name):
    if name == "choopy":
        return models.Choopy(seq_len=SEQ_LEN, dropout=0.0)
    if name == "mtchoopy":
        return models.MtChoopy(seq_len=SEQ_LEN, num_tasks=3, dropout=0.0)
    if name == "bicut":
        return models.BiCut(input_size=3, dropout=0.0)
    if name == "attncut":
        return models.AttnCut(input_size=3, dropout=0.0)
    if name == "mtattncut":
        return models.MtAttnCut(input_size=3, num_tasks=3, dropout=0.0)
    if name == "mmoecut":
        return models.MMOECut(seq_len=SEQ_LEN, num_tasks=3, input_size=3, dropout=0.0, num_experts=3)
    raise SystemExit(f"unknown model {name}")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="choopy")
    ap.add_argument("--groups", type=int, default=64, help="attention groups (of 64 lists) per GPU per step")
    ap.add_argument("--time-tag", type=int, default=-1,
                    help="call site timed in situ for the roofline (-1 = all tagged sites, the largest is reported)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (other model families, K3 / K4, side baselines)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port: same torch.nn calls + Python reward loop) on the host
    cores, same metric / config.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import torch_port
    from rlt_b200.data import synthetic_lists
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_batches = 2
    x, y = synthetic_lists(GROUP * n_batches, SEQ_LEN, N_FEATURES[args.model], seed=20240229, device="cpu")
    steps, warm = max(1, min(args.steps, 4)), max(1, min(args.warmup, 1))
    lps, times = torch_port.time_lists_per_s(args.model, x, y, GROUP, "train", steps=steps, warmup=warm)
    ilps, _ = torch_port.time_lists_per_s(args.model, x, y, GROUP, "infer", steps=steps, warmup=warm)
    sample = f"{steps} train steps of one batch of {GROUP} lists x {SEQ_LEN} (median), after {warm} warm-up"
    line = {"impl": "reference", "metric": METRIC, "value": lps, "unit": "lists/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * GROUP / lps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} train step (fwd + criterion + bwd + torch Adam step), synthetic robust04-shaped lists x {SEQ_LEN}, groups of {GROUP}",
                       "note": "reference algorithm on host CPU cores (oracle port of the reference's torch calls + Python reward loop)"},
            "inference": {"value": ilps, "unit": "lists/s"},
            "cpu_baseline": {"value": lps, "unit": "lists/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": lps, "unit": "lists/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide handles of one bench run (rank, device, library)."""


def _max_over_ranks(c, ms: float) -> float:
    import torch.distributed as dist
    t = torch.tensor([ms], device=c.dev)
    if c.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(c):
    import torch.distributed as dist
    if c.world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure_model(c, name: str, G: int, steps: int, warmup: int, sites: bool = False, e2e: bool = False):
    """Device-resident train throughput, inference throughput and (optionally) the end-to-end figure of one model
    family at G attention groups of 64 lists per GPU and step.

    A train step is what run.py:121-145 does per batch: forward + criterion + backward (+ gradient all-reduce when
    N > 1) + Adam step + argmax cut + per-list F1 / DCG (K4, on the device)."""
    from rlt_b200 import _lib, ops, parallel
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    from rlt_b200.optim import FusedAdam
    import models
    lib = c.lib
    B = G * GROUP
    shard = max(B, min(DATASET_LISTS // c.world, 4 * B) if not sites else DATASET_LISTS // c.world)
    x_all, y_all = synthetic_lists(shard, SEQ_LEN, N_FEATURES[name], seed=20240229 + c.rank, device=c.dev)
    n_chunks = shard // B
    torch.manual_seed(1234)
    model = build_model(models, name).to(c.dev)
    eng = Engine(model, n_groups=G, group_size=GROUP, seq_len=SEQ_LEN, training=True)
    # run.py:104,129: Adam with L2 decay, stepped once per batch -- one fused launch reading the (all-reduced) bucket
    opt = FusedAdam.for_engine(eng, lr=3e-5, weight_decay=1e-3)

    def cut_metrics():     # run.py:131-145 on the device: argmax cut (BiCut: first "truncate") + per-list F1 / DCG
        if name == "bicut":
            return ops.eval_cut(eng.probs2, eng._y, mode=1)
        return ops.eval_cut(eng.z[eng.H - 1], eng._y, mode=0)

    def step(i):
        ch = i % n_chunks
        eng._y = y_all[ch * B:(ch + 1) * B]
        eng.train_step(x_all[ch * B:(ch + 1) * B], eng._y)
        parallel.allreduce_mean_(eng.grad_bucket, G, G * c.world)
        opt.step()
        cut_metrics()

    for i in range(warmup):
        step(i)
    _barrier(c)
    if sites:
        c.sampler = ClockSampler(c.local)
        c.sampler.start()
        _lib.set_option("time_tag", c.args.time_tag)
        lib.rlt_timing_reset()
    launches0 = lib.rlt_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    _barrier(c)
    launches = int(lib.rlt_launch_count() - launches0)
    ms_local = e0.elapsed_time(e1)
    res = {"model": name, "criterion": CRITERION[name], "lists_per_step_per_gpu": B, "steps": steps, "warmup": warmup}
    if sites:
        _lib.set_option("time_tag", 0)
        res["_site_ms"] = ms_local
    ms = _max_over_ranks(c, ms_local)
    res.update({"train_lists_per_s": c.world * B * steps / (ms * 1e-3), "ms_per_step": ms / steps,
                "gpu_launches": launches, "loss": float(eng.loss.item()),
                "train_tflops": TRAIN_GFLOP_PER_LIST[name] * 1e-3 * c.world * B * steps / (ms * 1e-3) / c.world})

    if sites:      # the optimizer step alone, so that the figure without it can be read off
        _barrier(c)
        e0.record()
        for _ in range(20):
            opt.step()
        e1.record()
        _barrier(c)
        res["opt_us"] = e0.elapsed_time(e1) / 20 * 1e3

    # ---- inference: forward + fused argmax cut + per-list F1 / DCG (no communication)
    for i in range(2):
        eng.infer(x_all[:B], y_all[:B])
    _barrier(c)
    e0.record()
    for i in range(steps):
        ch = i % n_chunks
        eng.infer(x_all[ch * B:(ch + 1) * B], y_all[ch * B:(ch + 1) * B])
    e1.record()
    _barrier(c)
    ms_i = _max_over_ranks(c, e0.elapsed_time(e1))
    res["inference_lists_per_s"] = c.world * B * steps / (ms_i * 1e-3)

    if e2e:
        # ---- end to end: pinned host inputs -> H2D -> the same step -> D2H of the cut probabilities' logits (run.py
        # :131-142 pulls the B x L output to the host every step) and of the loss (run.py:146 loss.item())
        hx = x_all[:B].cpu().pin_memory()
        hy = y_all[:B].cpu().pin_memory()
        dx, dy = torch.empty_like(x_all[:B]), torch.empty_like(y_all[:B])
        out_dev = eng.probs2 if name == "bicut" else eng.z[eng.H - 1]
        hout = torch.empty(out_dev.shape, dtype=torch.float32).pin_memory()

        def e2e_step():
            dx.copy_(hx, non_blocking=True)
            dy.copy_(hy, non_blocking=True)
            eng._y = dy
            eng.train_step(dx, dy)
            parallel.allreduce_mean_(eng.grad_bucket, G, G * c.world)
            opt.step()
            cut_metrics()
            hout.copy_(out_dev, non_blocking=True)
            return eng.loss.item()        # device -> host read of the step's result (synchronises)
        for _ in range(2):
            e2e_step()
        _barrier(c)
        e0.record()
        for _ in range(steps):
            e2e_step()
        e1.record()
        _barrier(c)
        ms_e = _max_over_ranks(c, e0.elapsed_time(e1))
        res["e2e"] = {"value": c.world * B * steps / (ms_e * 1e-3), "unit": "lists/s",
                      "h2d_bytes_per_step": int(hx.numel() * 4 + hy.numel() * 4),
                      "d2h_bytes_per_step": int(hout.numel() * 4 + 4)}
    res["_eng"] = eng if sites else None
    if not sites:
        del eng, opt, model, x_all, y_all
        torch.cuda.empty_cache()
    return res


def measure_heads(c, steps: int):
    """K3 (cut head + reward loss + gradient) and K4 (argmax cut + F1 / DCG) on resident synthetic logits / labels:
    HBM-bound streaming kernels, SURVEY 8(d) bytes per list.  K4 is BASELINE config 5: 10 M lists, L = 300 -> 1000,
    sharded over the ranks without communication (each rank sweeps its resident shard until its share is covered)."""
    from rlt_b200 import ops
    hbm = c.peaks[0]
    out = {}
    total_lists = 10_000_000
    for L, n in ((300, 2_000_000), (500, 1_200_000), (1000, 600_000)):     # 4.8 GB of inputs per rank: far larger than L2
        g = torch.Generator(device=c.dev).manual_seed(L + c.rank)
        z = torch.randn(n, L, device=c.dev, generator=g)
        y = (torch.rand(n, L, device=c.dev, generator=g) < 0.1).float()
        passes = max(1, -(-total_lists // (n * c.world)))
        for _ in range(2):
            ops.eval_cut(z, y)
        _barrier(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(passes):
            ops.eval_cut(z, y)
        e1.record()
        _barrier(c)
        t = _max_over_ranks(c, e0.elapsed_time(e1)) * 1e-3
        nbytes = n * passes * (8 * L + 20)
        out[f"k4_{L}"] = {"what": "argmax cut + per-list F1 / DCG (rlt_eval_cut)", "seq_len": L,
                          "lists_swept": n * passes * c.world, "lists_per_s": n * passes * c.world / t,
                          "gbs_per_gpu": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / hbm,
                          "resident_lists_per_gpu": n, "passes": passes}
        if L == 300:
            grad = torch.empty_like(z)
            lpl = torch.empty(n, device=c.dev)
            bits = ops.pack_labels(y)
            words = bits.shape[1]
            for kind, metric, tau, packed in (("choopy", "f1", 1.0, False), ("js", "f1", 0.85, False), ("js", "dcg", 0.85, False),
                                              ("raml", "f1", 0.95, False), ("raml", "dcg", 0.95, False),
                                              ("choopy", "f1", 1.0, True), ("js", "f1", 0.85, True)):
                if packed:      # labels as the bit masks of rlt_pack_labels: 8 L + 4 ceil(L/32) + 4 bytes per list
                    fn = lambda: ops.cut_loss(z, None, label_bits=bits, loss_kind=kind, metric=metric, tau=tau, grad=grad,  # noqa: E731
                                              loss_per_list=lpl)
                else:
                    fn = lambda: ops.cut_loss(z, y, loss_kind=kind, metric=metric, tau=tau, grad=grad, loss_per_list=lpl)  # noqa: E731
                for _ in range(2):
                    fn()
                _barrier(c)
                e0.record()
                for _ in range(max(3, steps // 2)):
                    fn()
                e1.record()
                _barrier(c)
                k = max(3, steps // 2)
                t = _max_over_ranks(c, e0.elapsed_time(e1)) * 1e-3
                nbytes = n * k * ((8 * L + 4 * words + 4) if packed else (12 * L + 4))
                out[f"k3_{kind}_{metric}" + ("_bits" if packed else "")] = {
                    "what": "softmax + reward + loss + d/dlogits (" + ("rlt_cut_loss_bits, labels as bit masks" if packed else "rlt_cut_loss") + ")",
                    "seq_len": L, "lists_per_s": n * k * c.world / t, "gbs_per_gpu": nbytes / t / 1e9,
                    "frac": nbytes / t / 1e9 / hbm}
            del bits
            del grad, lpl
        del z, y
        torch.cuda.empty_cache()
    return out


def measure_module_api(c, name: str, steps: int):
    """run.py's own call sequence at its batch size on ONE GPU: nn.Module forward -> criterion -> loss.backward() ->
    optimizer.step() -> output.cpu() -> np.argmax -> Metric.f1 / Metric.dcg -> loss.item(), inputs from pinned host
    memory (DataLoader pin_memory=True, attncut_dataloader.py:87)."""
    import numpy as np
    import models
    from utils import losses
    from utils.metrics import Metric
    from rlt_b200.data import synthetic_lists
    from rlt_b200.optim import FusedAdam
    torch.manual_seed(1234)
    model = build_model(models, name).to(c.dev).train()
    crit = {"bicut": lambda: losses.BiCutLoss(metric="f1"), "choopy": lambda: losses.ChoopyLoss(metric="f1"),
            "attncut": lambda: losses.DivLoss(metric="f1", div_type="js", augmented=True)}.get(
        name, lambda: losses.MtCutLoss(metric="f1", num_tasks=3))().to(c.dev)
    opt = FusedAdam(model.parameters(), lr=3e-5, weight_decay=1e-3)
    x, y = synthetic_lists(GROUP * 4, SEQ_LEN, N_FEATURES[name], seed=11, device="cpu")
    x, y = x.pin_memory(), y.pin_memory()

    def one(i):
        b = i % 4
        xb = x[b * GROUP:(b + 1) * GROUP].to(c.dev, non_blocking=True)
        yb = y[b * GROUP:(b + 1) * GROUP].to(c.dev, non_blocking=True)
        opt.zero_grad()
        out = model(xb)
        loss = crit(out, yb)
        loss.backward()
        opt.step()
        last = out[-1] if isinstance(out, list) else out
        p = last.detach().cpu().squeeze().numpy()
        if name == "bicut":
            pred = np.argmax(p, axis=2)
            ks = [SEQ_LEN if r.sum() == SEQ_LEN else int(np.argmin(r)) + 1 for r in pred]
        else:
            ks = np.argmax(p, axis=1) + 1
        yn = yb.cpu().numpy()
        return loss.item(), Metric.f1(yn, ks), Metric.dcg(yn, ks)
    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"what": "nn.Module + criterion + loss.backward() + FusedAdam.step() + output.cpu() + host Metric.f1/dcg at the "
                    "reference's batch (run.py:121-146)", "model": name, "batch": GROUP, "lists_per_s": GROUP * steps / dt,
            "ms_per_step": dt / steps * 1e3}


def measure_graph_step(c, name: str, steps: int):
    """The reference's batch (run.py: 63-64 lists) as ONE CUDA-graph launch per step: Engine forward + criterion + backward +
    cut metrics (rlt_eval_cut) + fused Adam, inputs copied from pinned host memory into the graph's static buffers and the
    loss + per-list F1 / DCG read back every step."""
    import models
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    from rlt_b200.optim import FusedAdam
    torch.manual_seed(1234)
    model = build_model(models, name).to(c.dev).train()
    eng = Engine(model, n_groups=1, group_size=GROUP, seq_len=SEQ_LEN, training=True)
    opt = FusedAdam.for_engine(eng, lr=3e-5, weight_decay=1e-3)
    x, y = synthetic_lists(GROUP * 4, SEQ_LEN, N_FEATURES[name], seed=11, device="cpu")
    x, y = x.pin_memory(), y.pin_memory()
    xs, ys = x[:GROUP].to(c.dev), y[:GROUP].to(c.dev)
    replay = eng.capture_train_step(xs, ys, optimizer=opt, metrics=True)

    def one(i):
        b = i % 4
        xs.copy_(x[b * GROUP:(b + 1) * GROUP], non_blocking=True)
        ys.copy_(y[b * GROUP:(b + 1) * GROUP], non_blocking=True)
        replay()
        _, _, _, f1, dcg = eng.graph_metrics
        return eng.loss.item(), f1.mean().item(), dcg.mean().item()
    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        replay()
    e1.record()
    torch.cuda.synchronize()
    return {"what": "Engine train step + cut metrics + fused Adam as one CUDA graph per batch of the reference's size; "
                    "ms_per_step: wall clock with pinned-host inputs in and loss / F1 / DCG out every step; "
                    "device_ms_per_step: back-to-back replays (CUDA events)", "model": name, "batch": GROUP,
            "lists_per_s": GROUP * steps / dt, "ms_per_step": dt / steps * 1e3,
            "device_ms_per_step": e0.elapsed_time(e1) / steps}


def measure_torch_cuda_eager(c, name: str, steps: int):
    """Side baseline ("the existing Blackwell path"): the reference's module graph in eager torch-CUDA on the
    same B200 (cuDNN LSTM, cuBLASLt, SDPA) with the VECTORISED reward (the reference's Python B x L loop would dominate),
    torch.optim.Adam, batch 64.  BASELINE measurement only; nothing of the product runs here."""
    from oracle import rlt_oracle as O
    from oracle import torch_port
    from rlt_b200.data import synthetic_lists
    torch.manual_seed(1234)
    model = torch_port.PortModel(name, seq_len=SEQ_LEN, n_features=N_FEATURES[name]).to(c.dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=3e-5, weight_decay=1e-3)
    crit = O.criterion_for(name, metric="f1", loop=False)
    x, y = synthetic_lists(GROUP * 4, SEQ_LEN, N_FEATURES[name], seed=11, device=c.dev)

    def one(i):
        b = i % 4
        xb, yb = x[b * GROUP:(b + 1) * GROUP], y[b * GROUP:(b + 1) * GROUP]
        opt.zero_grad(set_to_none=True)
        out = model(xb)
        loss = crit(out, yb)
        loss.backward()
        opt.step()
        last = out[-1] if isinstance(out, list) else out
        return last.detach().argmax(dim=1), loss
    try:
        for i in range(3):
            one(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    except Exception as ex:      # a baseline must never take the bench line down
        return {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
    return {"what": "reference module graph in eager torch-CUDA (cuDNN / cuBLASLt / SDPA), vectorised reward, torch Adam, "
                    "device-resident inputs, no host metrics", "model": name, "batch": GROUP,
            "lists_per_s": GROUP * steps / (ms * 1e-3), "ms_per_step": ms / steps}


# forward GFLOP per list at L = 300, S = 64; a train step is 3x (recomputation not counted)
FWD_GFLOP_PER_LIST = {"bicut": 0.356, "choopy": 1.091, "attncut": 1.123, "mtchoopy": 1.091, "mtattncut": 1.123, "mmoecut": 2.738}
TRAIN_GFLOP_PER_LIST = {k: 3 * v for k, v in FWD_GFLOP_PER_LIST.items()}
CONFIG_GROUPS = {"bicut": 64, "attncut": 64, "mtattncut": 64, "mmoecut": 32}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from rlt_b200 import _lib, ops

    c = Ctx()
    c.args = args
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rlt_b200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=c.dev)
    c.lib = ops.lib()
    c.lib.rlt_launch_count.restype = ctypes.c_ulonglong
    c.peaks = measured_peaks()
    lib, world, rank = c.lib, c.world, c.rank
    hbm, tf_burst, tf_sust, src = c.peaks

    # ---------------- headline: BASELINE configs[1] (or --model), every tagged call site timed in situ
    G = args.groups
    B = G * GROUP
    head = measure_model(c, args.model, G, args.steps, args.warmup, sites=True, e2e=True)
    eng = head.pop("_eng")
    ms = head["ms_per_step"] * args.steps
    site_ms = head.pop("_site_ms")
    tot_ms, cnt = ctypes.c_double(0), ctypes.c_int(0)
    _lib.check(lib.rlt_timing_read(ctypes.byref(tot_ms), ctypes.byref(cnt)), "rlt_timing_read")
    c.sampler.stop_flag = True
    c.sampler.join(timeout=2)
    clocks = c.sampler.summary()

    # ---------------- roofline of the call sites (rank 0's own launches; CUDA events on the launch stream)
    T = B * SEQ_LEN
    d, dff, nh = eng.d, 2048, getattr(eng, "n_head", 8)
    S = GROUP
    # tag -> (label, algorithmic HBM bytes per launch, algorithmic flops per launch)   DESIGN.md section 3
    site = {
        1: ("QKV projection (gemm_tn, TF32)", T * (d + 3 * d) * 4 + 3 * d * d * 4, 2 * T * d * 3 * d),
        2: ("attention out-projection + residual (gemm_tn, TF32)", T * 3 * d * 4 + d * d * 4, 2 * T * d * d),
        3: ("FFN1: h = relu(y W1^T + b1), fp16 operands -> fp16 hidden (gemm_tn, tcgen05 kind::f16)", T * (d + dff) * 2 + dff * d * 2, 2 * T * d * dff),
        4: ("FFN2: u2 = y + h W2^T + b2, fp16 hidden (gemm_tn, kind::f16)", T * dff * 2 + T * 2 * d * 4 + dff * d * 2, 2 * T * d * dff),
        5: ("FFN backward, one pass over h: dH = s (dU2 W2) [h > 0] (fp16), db1 += colsum, dW2 += dU2^T h "
            "(ffn_bwd_kernel, tcgen05 kind::f16, K-major + MN-major views of the same tiles)" if d == 128 else
            "dH = s (dU2 W2) [h > 0], fp16 in / fp16 out + bias-gradient column sums (gemm_tn, kind::f16)",
            T * d * 2 + T * 2 * dff * 2 + dff * d * 2, (4 if d == 128 else 2) * T * d * dff),
        6: ("dY = dU2 + dH W1 / s (gemm_tn, kind::f16)", T * dff * 2 + T * 2 * d * 4 + dff * d * 2, 2 * T * d * dff),
        7: ("dW2 += dU2^T h (gemm_dw, kind::f16, MN-major operands)", T * (d + dff) * 2, 2 * T * d * dff),
        8: ("dW1 += dH^T y (gemm_dw, kind::f16, MN-major operands)", T * (d + dff) * 2, 2 * T * d * dff),
        9: ("cross-list attention forward", T * (3 * d + d + nh) * 4, 4 * T * S * d),
        10: ("cross-list attention backward", T * (3 * d + d + nh + 3 * d) * 4, 10 * T * S * d),
        # fused FFN forward (ffn_fwd_fused.cuh): y fp16 in, LN2 output fp32 (+ fp16 hidden in train mode) out
        11: ("fused FFN forward: relu(y W1^T + b1) W2^T + b2 + y -> LayerNorm2, hidden on chip (ffn_fwd_kernel, tcgen05 "
             "cta_group::2, kind::f16)", T * (d * 2 + d * 4 * 2) + 2 * dff * d * 2, 4 * T * d * dff),
        12: ("BiLSTM recurrence (tcgen05 kind::f16, unit-major; mean of 2 forward + 2 backward launches)", T * 7680, 2 * T * 512 * 128 * 2),
    }
    if getattr(eng, "_ffn_fwd_saves_hidden", True):
        site[11] = (site[11][0], site[11][1] + T * dff * 2, site[11][2])
    kernels = {}
    if args.time_tag == -1:
        for tag, (label, nbytes, flops) in site.items():
            tm, tc = ctypes.c_double(0), ctypes.c_int(0)
            _lib.check(lib.rlt_timing_read_tag(tag, ctypes.byref(tm), ctypes.byref(tc)), "rlt_timing_read_tag")
            if tc.value:
                avg = tm.value / tc.value
                kernels[tag] = {"site": label, "launches": tc.value, "avg_launch_ms": avg, "share_of_step": tm.value / site_ms,
                                "gbs": nbytes / (avg * 1e-3) / 1e9, "hbm_frac": nbytes / (avg * 1e-3) / 1e9 / hbm,
                                "tflops": flops / (avg * 1e-3) / 1e12, "tensor_frac": flops / (avg * 1e-3) / 1e12 / tf_sust}
    elif cnt.value > 0 and args.time_tag in site:
        label, nbytes, flops = site[args.time_tag]
        avg = tot_ms.value / cnt.value
        kernels[args.time_tag] = {"site": label, "launches": cnt.value, "avg_launch_ms": avg, "share_of_step": tot_ms.value / site_ms,
                                  "gbs": nbytes / (avg * 1e-3) / 1e9, "hbm_frac": nbytes / (avg * 1e-3) / 1e9 / hbm,
                                  "tflops": flops / (avg * 1e-3) / 1e12, "tensor_frac": flops / (avg * 1e-3) / 1e12 / tf_sust}
    lib.rlt_timing_reset()
    top = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else None
    roof = None
    if top is not None:
        k = kernels[top]
        label, nbytes, flops = site[top]
        bound = "tensor" if k["tensor_frac"] >= k["hbm_frac"] else "hbm"     # the resource the site is closest to
        traffic = None
        tp = ROOT / "profiles" / "ncu_traffic.json"      # dram bytes per token of one launch, from `ncu --set full`
        if tp.exists():
            per_tok = json.loads(tp.read_text()).get(str(top), {}).get("dram_bytes_per_token")
            traffic = per_tok * T if per_tok else None
        peak_src = (src + " (MEASURED_PEAKS.json: STREAM-style copy / cuBLAS bf16 sustained)") if src == "measured" else src
        roof = {"kernel": label, "bound": bound,
                "achieved": k["tflops"] if bound == "tensor" else k["gbs"], "peak": tf_sust if bound == "tensor" else hbm,
                "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                "frac": k["tensor_frac"] if bound == "tensor" else k["hbm_frac"], "traffic": traffic,
                "algorithmic_bytes_per_launch": nbytes, "algorithmic_flops_per_launch": flops,
                "hbm": {"achieved": k["gbs"], "peak": hbm, "unit": "GB/s", "frac": k["hbm_frac"]},
                "tensor": {"achieved": k["tflops"], "peak": tf_sust, "unit": "TFLOP/s", "frac": k["tensor_frac"]},
                "peak_source": peak_src, "avg_launch_ms": k["avg_launch_ms"], "launches_timed": k["launches"],
                "share_of_step": k["share_of_step"]}
    step_tflops = TRAIN_GFLOP_PER_LIST[args.model] * 1e-3 * head["train_lists_per_s"] / world
    del eng
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE configs (device-resident train + inference, same step definition)
    configs = {}
    if not args.no_configs:
        ksteps = max(3, min(args.steps, 5))
        for name, g in CONFIG_GROUPS.items():
            if name == args.model:
                continue
            r = measure_model(c, name, g, ksteps, 3)
            r.pop("_eng", None)
            r["tensor_frac_of_step"] = r["train_tflops"] / tf_sust
            configs[name] = r
        configs.update(measure_heads(c, args.steps))
        if rank == 0:
            configs["e2e_module_api"] = {n: measure_module_api(c, n, 10) for n in ("choopy", "bicut", "attncut", "mmoecut")}
            configs["graph_step_b64"] = {n: measure_graph_step(c, n, 30) for n in ("choopy", "bicut", "attncut", "mmoecut")}
            configs["torch_cuda_eager"] = {n: measure_torch_cuda_eager(c, n, 10) for n in ("choopy", "bicut", "attncut", "mmoecut")}
        _barrier(c)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import torch_port
        from rlt_b200.data import synthetic_lists
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cx, cy = synthetic_lists(GROUP * 2, SEQ_LEN, N_FEATURES[args.model], seed=20240229, device="cpu")
        lps, times = torch_port.time_lists_per_s(args.model, cx, cy, GROUP, "train", steps=3, warmup=1)
        cpu = {"value": lps, "unit": "lists/s", "cores": cores, "kind": "port",
               "sample": f"3 reference-style train steps (fwd + Python-loop criterion + bwd + torch Adam step + host metrics) on one batch of "
                         f"{GROUP} lists x {SEQ_LEN}, median, 1 warm-up; step times {[round(v, 3) for v in times]} s"}

    rnd = lambda v: round(v, 4) if isinstance(v, float) else v  # noqa: E731
    line = {"metric": METRIC, "value": head["train_lists_per_s"], "unit": "lists/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32 / fp16 operands (11-bit significand), fp32 accumulate, fp32 master tensors", "data": "synthetic",
            "config": {"workload": f"{args.model} train step (fwd + {CRITERION[args.model]} + bwd{' + NCCL grad all-reduce' if world > 1 else ''} + fused Adam step "
                                   f"+ argmax cut + per-list F1/DCG), {DATASET_LISTS} synthetic robust04-shaped lists x {SEQ_LEN} resident in HBM",
                       "lists_per_step_per_gpu": B, "attention_group": GROUP, "seq_len": SEQ_LEN,
                       "l2": "inputs and activations of one step (>10 GB) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"dp{world}"},
            "optimizer": {"kind": "FusedAdam (rlt_adam_step_masked: one launch over all parameter tensors, L2 weight decay, "
                                  "per-tensor step counts on the device)", "included_in_value": True, "us_per_step": head["opt_us"]},
            "step_tflops_per_gpu": step_tflops, "step_tensor_frac": step_tflops / tf_sust,
            "inference": {"value": head["inference_lists_per_s"], "unit": "lists/s", "what": "forward + fused argmax-cut + per-list F1/DCG",
                          "tflops_per_gpu": FWD_GFLOP_PER_LIST[args.model] * 1e-3 * head["inference_lists_per_s"] / world},
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "loss": head["loss"], "clocks": clocks, "roofline": roof,
            "kernels": {str(k): {kk: rnd(vv) for kk, vv in v.items()}
                        for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["share_of_step"])},
            "configs": {k: ({kk: rnd(vv) for kk, vv in v.items()} if isinstance(v, dict) else v) for k, v in configs.items()},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
