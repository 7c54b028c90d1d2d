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
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from rlt_b200 import _lib, ops, parallel
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    from rlt_b200.optim import FusedAdam
    import models

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rlt_b200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = ops.lib()
    lib.rlt_launch_count.restype = ctypes.c_ulonglong

    G = args.groups
    B = G * GROUP                                   # lists per GPU per step
    shard = DATASET_LISTS // world                  # lists resident on this GPU
    if shard < B:
        shard = B
    x_all, y_all = synthetic_lists(shard, SEQ_LEN, N_FEATURES[args.model], seed=20240229 + rank, device=dev)
    n_chunks = shard // B

    torch.manual_seed(1234)
    model = build_model(models, args.model).to(dev)
    eng = Engine(model, n_groups=G, group_size=GROUP, seq_len=SEQ_LEN, training=True)

    # run.py:104,129: Adam with L2 decay, stepped once per batch -- here one fused launch reading the (all-reduced) bucket
    opt = FusedAdam.for_engine(eng, lr=3e-5, weight_decay=1e-3)

    def step(i):
        c = i % n_chunks
        eng.train_step(x_all[c * B:(c + 1) * B], y_all[c * B:(c + 1) * B])
        parallel.allreduce_mean_(eng.grad_bucket, G, G * world)
        opt.step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident train throughput
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    _lib.set_option("time_tag", args.time_tag)
    lib.rlt_timing_reset()
    launches0 = lib.rlt_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    launches = int(lib.rlt_launch_count() - launches0)
    ms = e0.elapsed_time(e1)
    _lib.set_option("time_tag", 0)
    tot_ms, cnt = ctypes.c_double(0), ctypes.c_int(0)
    _lib.check(lib.rlt_timing_read(ctypes.byref(tot_ms), ctypes.byref(cnt)), "rlt_timing_read")
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    train_lps = world * B * args.steps / (ms * 1e-3)
    loss_val = float(eng.loss.item())

    # ---------------- the optimizer step alone (throughput is reported with the optimizer included; this is
    # the share it takes, so that the excluded figure can be read off as well)
    barrier()
    e0.record()
    for _ in range(20):
        opt.step()
    e1.record()
    barrier()
    opt_us = e0.elapsed_time(e1) / 20 * 1e3

    # ---------------- inference (forward + fused cut/F1/DCG)
    for i in range(2):
        eng.infer(x_all[:B], y_all[:B])
    barrier()
    e0.record()
    for i in range(args.steps):
        c = i % n_chunks
        k, f1, dcg = eng.infer(x_all[c * B:(c + 1) * B], y_all[c * B:(c + 1) * B])
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    infer_lps = world * B * args.steps / (float(t.item()) * 1e-3)

    # ---------------- end to end: pinned host inputs -> H2D -> train step -> loss D2H, every step
    hx = x_all[:B].cpu().pin_memory()
    hy = y_all[:B].cpu().pin_memory()
    dx, dy = torch.empty_like(x_all[:B]), torch.empty_like(y_all[:B])
    def e2e_step():
        dx.copy_(hx, non_blocking=True)
        dy.copy_(hy, non_blocking=True)
        eng.train_step(dx, dy)
        parallel.allreduce_mean_(eng.grad_bucket, G, G * world)
        opt.step()
        return eng.loss.item()        # device -> host read of the step's result
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_lps = world * B * args.steps / (float(t.item()) * 1e-3)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel, timed in situ (CUDA events on the launch stream)
    hbm, tf_burst, tf_sust, src = measured_peaks()
    T = B * SEQ_LEN
    d, dff, nh = eng.d, 2048, getattr(eng, "n_head", 8)
    # algorithmic HBM bytes per launch of every tagged call site (T tokens per launch)
    site_bytes = {
        1: ("QKV projection (gemm_tn, TF32)", T * (d + 3 * d) * 4 + 3 * d * d * 4),
        2: ("attention out-projection + residual (gemm_tn, TF32)", T * 3 * d * 4 + d * d * 4),
        3: ("FFN1: h = relu(y W1^T + b1), fp16 operands -> fp16 hidden (gemm_tn, tcgen05 kind::f16)", T * (d + dff) * 2 + dff * d * 2),
        4: ("FFN2: u2 = y + h W2^T + b2, fp16 hidden (gemm_tn, kind::f16)", T * dff * 2 + T * 2 * d * 4 + dff * d * 2),
        # d_model 128: the fused kernel (dH, db1 AND dW2 in one pass over h: ffn_bwd_fused.cuh), same bytes as dH alone
        5: ("FFN backward, one pass over h: dH = s (dU2 W2) [h > 0] (fp16), db1 += colsum, dW2 += dU2^T h "
            "(ffn_bwd_kernel, tcgen05 kind::f16, K-major + MN-major views of the same tiles)" if d == 128 else
            "dH = s (dU2 W2) [h > 0], fp16 in / fp16 out + bias-gradient column sums (gemm_tn, kind::f16)",
            T * d * 2 + T * 2 * dff * 2 + dff * d * 2),
        6: ("dY = dU2 + dH W1 / s (gemm_tn, kind::f16)", T * dff * 2 + T * 2 * d * 4 + dff * d * 2),
        7: ("dW2 += dU2^T h (gemm_dw, kind::f16, MN-major operands)", T * (d + dff) * 2),
        8: ("dW1 += dH^T y (gemm_dw, kind::f16, MN-major operands)", T * (d + dff) * 2),
        9: ("cross-list attention forward (mma.sync TF32, cp.async pipeline)", T * (3 * d + d + nh) * 4),
        10: ("cross-list attention backward (mma.sync TF32, cp.async pipeline)", T * (3 * d + d + nh + 3 * d) * 4),
        # mean over the 4 launches of a step, per token (both directions): layer-0 forward has no P tensor (fused
        # projection): saved record 4 KB (fp16 gate pairs 2 KB, c 1 KB, h_prev 1 KB) + y 1 KB; layer-1 forward adds P 4 KB;
        # each backward reads gates + c 3 KB (c_prev comes from L2) + dy 1 KB and writes dA 4 KB
        12: ("BiLSTM recurrence (tcgen05 kind::f16, unit-major; mean of 2 forward + 2 backward launches)", T * 7680),
    }
    roof, kernels = None, {}
    if args.time_tag == -1:
        for tag, (label, nbytes) in site_bytes.items():
            tm, tc = ctypes.c_double(0), ctypes.c_int(0)
            _lib.check(lib.rlt_timing_read_tag(tag, ctypes.byref(tm), ctypes.byref(tc)), "rlt_timing_read_tag")
            if tc.value:
                kernels[tag] = {"site": label, "launches": tc.value, "avg_launch_ms": tm.value / tc.value,
                                "share_of_step": tm.value / ms, "gbs": nbytes / (tm.value / tc.value * 1e-3) / 1e9}
        top = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else None
    else:
        top = args.time_tag if cnt.value > 0 and args.time_tag in site_bytes else None
        if top is not None:
            kernels[top] = {"site": site_bytes[top][0], "launches": cnt.value, "avg_launch_ms": tot_ms.value / cnt.value,
                            "share_of_step": tot_ms.value / ms,
                            "gbs": site_bytes[top][1] / (tot_ms.value / cnt.value * 1e-3) / 1e9}
    lib.rlt_timing_reset()
    # all tagged GEMM sites together (the HBM-bound part of the step): sum of algorithmic bytes / sum of device time
    gemm_sites = [t for t in kernels if t in (1, 2, 3, 4, 5, 6, 7, 8)]
    gemm_family = None
    if gemm_sites:
        tot_s = sum(kernels[t]["avg_launch_ms"] * kernels[t]["launches"] for t in gemm_sites) * 1e-3
        tot_b = sum(site_bytes[t][1] * kernels[t]["launches"] for t in gemm_sites)
        gemm_family = {"what": "all tagged tcgen05 GEMM call sites of the encoder layers", "share_of_step": tot_s * 1e3 / ms,
                       "achieved": tot_b / tot_s / 1e9, "unit": "GB/s", "peak": hbm, "frac": tot_b / tot_s / 1e9 / hbm}
    if top is not None:
        k = kernels[top]
        traffic = None
        tp = ROOT / "profiles" / "ncu_traffic.json"      # dram bytes per token of one launch, from `ncu --set full`
        if tp.exists():
            per_tok = json.loads(tp.read_text()).get(str(top), {}).get("dram_bytes_per_token")
            traffic = per_tok * T if per_tok else None
        roof = {"kernel": k["site"], "bound": "hbm", "achieved": k["gbs"], "peak": hbm, "unit": "GB/s",
                "frac": k["gbs"] / hbm, "traffic": traffic, "algorithmic_bytes_per_launch": site_bytes[top][1],
                "peak_source": src + " (STREAM-style copy, MEASURED_PEAKS.json)" if src == "measured" else src,
                "avg_launch_ms": k["avg_launch_ms"], "launches_timed": k["launches"], "share_of_step": k["share_of_step"]}

    # ---------------- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import torch_port
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cx, cy = synthetic_lists(GROUP * 2, SEQ_LEN, N_FEATURES[args.model], seed=20240229, device="cpu")
        lps, times = torch_port.time_lists_per_s(args.model, cx, cy, GROUP, "train", steps=3, warmup=1)
        cpu = {"value": lps, "unit": "lists/s", "cores": cores, "kind": "port",
               "sample": f"3 reference-style train steps (fwd + Python-loop criterion + bwd + torch Adam step + host metrics) on one batch of "
                         f"{GROUP} lists x {SEQ_LEN}, median, 1 warm-up; step times {[round(v, 3) for v in times]} s"}

    line = {"metric": METRIC, "value": train_lps, "unit": "lists/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32 / fp16 operands (11-bit significand), fp32 accumulate, fp32 master tensors", "data": "synthetic",
            "config": {"workload": f"{args.model} train step (fwd + {CRITERION[args.model]} + bwd{' + NCCL grad all-reduce' if world > 1 else ''} + fused Adam step), "
                                   f"{DATASET_LISTS} synthetic robust04-shaped lists x {SEQ_LEN} resident in HBM",
                       "lists_per_step_per_gpu": B, "attention_group": GROUP, "seq_len": SEQ_LEN,
                       "l2": "inputs and activations of one step (>10 GB) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"dp{world}"},
            "optimizer": {"kind": "FusedAdam (rlt_adam_step: one launch over all parameter tensors, L2 weight decay)",
                          "included_in_value": True, "us_per_step": opt_us,
                          "value_without_optimizer": world * B * args.steps / ((ms - args.steps * opt_us * 1e-3) * 1e-3)},
            "inference": {"value": infer_lps, "unit": "lists/s", "what": "forward + fused argmax-cut + per-list F1/DCG"},
            "e2e": {"value": e2e_lps, "unit": "lists/s", "h2d_bytes_per_step": int(hx.numel() * 4 + hy.numel() * 4),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "loss": loss_val, "clocks": sampler.summary(), "roofline": roof,
            "roofline_gemm_family": gemm_family,
            "kernels": {str(k): {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                        for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["share_of_step"])},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
