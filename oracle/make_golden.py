"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/refshim.py) in this container.

Run once here (the reference tree does not exist on the GPU box):  python -m oracle.make_golden
Fixtures are small: inputs, labels, outputs, loss and gradient DIGESTS.  The weights themselves are
not stored: they are re-created on any machine by seeding torch's CPU generator (torch.manual_seed)
and constructing the same torch submodules in the same order as the reference constructors; a
checksum of every state_dict entry is stored so a test can prove the weights were reproduced.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))

from oracle import refshim  # noqa: E402
from oracle.rlt_oracle import flat_outputs, loss_input  # noqa: E402
from rlt_b200.data import synthetic_lists  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
WEIGHT_SEED = 1234
DATA_SEED = 20240229

# model name -> (reference class name, ctor kwargs, feature count, criterion builder name)
MODELS = {
    "bicut": ("BiCut", dict(input_size=3, dropout=0.0), 3),
    "choopy": ("Choopy", dict(seq_len=300, dropout=0.0), 1),
    "attncut": ("AttnCut", dict(input_size=3, dropout=0.0), 3),
    "mtchoopy": ("MtChoopy", dict(seq_len=300, num_tasks=3, dropout=0.0), 1),
    "mtattncut": ("MtAttnCut", dict(input_size=3, num_tasks=3, dropout=0.0), 3),
    "mmoecut": ("MMOECut", dict(seq_len=300, num_tasks=3, input_size=3, dropout=0.0, num_experts=3), 3),
    # SURVEY section 8(f) row N4 (run.py:91-102)
    "moecut": ("MOECut", dict(seq_len=300, num_tasks=3, input_size=3, dropout=0.0), 3),
    "plecut": ("PLECut", dict(seq_len=300, input_size=3, dropout=0.0, num_experts=3), 3),
    # verify_probe.py:61: the base model of the probing experiment; forward returns (experts_in, experts_o, towers)
    "probebase": ("ProbeBase", dict(seq_len=300, num_tasks=3, input_size=3, dropout=0.0, num_experts=2), 3),
    # run.py --num_tasks 2.1 (class + cut) / 2.2 (rerank + cut); B = 5 only
    "mtchoopy_t21": ("MtChoopy", dict(seq_len=300, num_tasks=2.1, dropout=0.0), 1),
    "mtchoopy_t22": ("MtChoopy", dict(seq_len=300, num_tasks=2.2, dropout=0.0), 1),
    "mtattncut_t21": ("MtAttnCut", dict(input_size=3, num_tasks=2.1, dropout=0.0), 3),
    "mtattncut_t22": ("MtAttnCut", dict(input_size=3, num_tasks=2.2, dropout=0.0), 3),
    "mmoecut_t21": ("MMOECut", dict(seq_len=300, num_tasks=2.1, input_size=3, dropout=0.0, num_experts=3), 3),
    "mmoecut_t22": ("MMOECut", dict(seq_len=300, num_tasks=2.2, input_size=3, dropout=0.0, num_experts=3), 3),
}


def digest(t: torch.Tensor, rng: np.random.Generator, n: int = 512) -> dict:
    """Compact description of a tensor: norms + a fixed random subsample (full copy if it is small)."""
    a = t.detach().double().numpy().ravel()
    idx = np.arange(a.size) if a.size <= n else np.sort(rng.choice(a.size, size=n, replace=False))
    return {"l2": np.float64(np.sqrt((a * a).sum())), "sum": np.float64(a.sum()), "absmax": np.float64(np.abs(a).max()),
            "idx": idx.astype(np.int64), "val": a[idx].astype(np.float64), "size": np.int64(a.size)}


def store_output(rec: dict, key: str, t: torch.Tensor, rng: np.random.Generator, full_below: int = 50_000):
    """Model outputs are stored whole; the [B, L, 256] representations ProbeBase also returns are stored as a shape plus
    a digest of 4096 sampled entries (tests/helpers.py::output_error reads both forms)."""
    if t.numel() <= full_below:
        rec[key] = t.detach().numpy()
        return
    rec[key + "/shape"] = np.array(t.shape, dtype=np.int64)
    for k, v in digest(t, rng, n=4096).items():
        rec[f"{key}/{k}"] = v


def build_criterion(ref_losses, name: str, metric: str):
    # the dispatch of reference run.py:59-102 with its CLI defaults (div_type='js', augmented reward)
    if name == "bicut":
        return ref_losses.BiCutLoss(metric=metric)
    if name == "choopy":
        return ref_losses.ChoopyLoss(metric=metric)
    if name == "attncut":
        return ref_losses.DivLoss(metric=metric, div_type="js", augmented=True)
    num_tasks = MODELS[name][1].get("num_tasks", 3)
    if name.startswith(("mtchoopy", "mtattncut")):
        return ref_losses.MtCutLoss(metric=metric, rerank_weight=0.5, classi_weight=0.5, num_tasks=num_tasks)
    return ref_losses.MtCutLoss(metric=metric, num_tasks=num_tasks)


BIG_BATCH = ("bicut", "choopy", "attncut", "mtchoopy", "mtattncut", "mmoecut")
TRAJ_STEPS, TRAJ_B, TRAJ_LR, TRAJ_WD = 5, 16, 1e-3, 1e-3


def trajectory_goldens(ref_models, ref_losses, only=None):
    """run.py:104,120-129 for TRAJ_STEPS batches: the unmodified reference model + criterion + torch.optim.Adam(lr,
    weight_decay) (L2 decay), one batch per step.  Stored: the batches, the loss of every step, digests of the final
    parameters and of the total parameter displacement.  Pins multi-step behaviour (optimizer coupling, operands that
    must follow the parameters from step to step)."""
    for name in BIG_BATCH:
        if only and name not in only:
            continue
        cls_name, kwargs, feats = MODELS[name]
        torch.manual_seed(WEIGHT_SEED)
        model = getattr(ref_models, cls_name)(**kwargs)
        model.train()
        init = {n: p.detach().clone() for n, p in model.named_parameters()}
        torch.manual_seed(0)
        crit = build_criterion(ref_losses, name, "f1")
        opt = torch.optim.Adam(model.parameters(), lr=TRAJ_LR, weight_decay=TRAJ_WD)
        rec = {"lr": np.float64(TRAJ_LR), "weight_decay": np.float64(TRAJ_WD), "steps": np.int64(TRAJ_STEPS)}
        losses = []
        for step in range(TRAJ_STEPS):
            x, y = synthetic_lists(TRAJ_B, 300, feats, seed=DATA_SEED + 1000 + step, device="cpu")
            rec[f"x{step}"], rec[f"y{step}"] = x.numpy(), y.numpy()
            opt.zero_grad()
            loss = crit(loss_input(model(x)), y)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        rec["losses"] = np.array(losses, dtype=np.float64)
        rng = np.random.default_rng(13)
        names = []
        for pname, p in model.named_parameters():
            names.append(pname)
            d = digest(p.detach() - init[pname], rng)
            for k, v in d.items():
                rec[f"delta/{pname}/{k}"] = v
            rec[f"final/{pname}/val"] = p.detach().double().numpy().ravel()[d["idx"]]
        rec["param_names"] = np.array(names)
        np.savez_compressed(GOLDEN / f"traj_{name}.npz", **rec)
        print(f"traj_{name}: losses={losses}")


def positions_goldens(ref_models, ref_losses):
    """attend = "positions" (the papers' intent; attend_axis = 1 of rlt_encoder_desc): the UNMODIFIED
    reference module with its encoder applied to the transposed tensor -- `enc(x.transpose(0, 1)).transpose(0, 1)` -- which
    makes torch attend within each list.  Only the encoder attribute's forward is wrapped; parameters, heads and
    criterion are the reference's.  Choopy / MtChoopy at B = 5 (head dim 16)."""
    for name, attr in (("choopy", "attention_layer"), ("mtchoopy", "encoding_layer")):
        cls_name, kwargs, feats = MODELS[name]
        torch.manual_seed(WEIGHT_SEED)
        model = getattr(ref_models, cls_name)(**kwargs)
        model.train()
        enc = getattr(model, attr)
        orig = enc.forward
        enc.forward = lambda x, _f=orig: _f(x.transpose(0, 1)).transpose(0, 1)
        B = 5
        x, y = synthetic_lists(B, 300, feats, seed=DATA_SEED + B, device="cpu")
        out = model(x)
        torch.manual_seed(0)
        crit = build_criterion(ref_losses, name, "f1")
        loss = crit(loss_input(out), y)
        loss.backward()
        rng = np.random.default_rng(7)
        rec = {"x": x.numpy(), "y": y.numpy(), "loss": np.float64(loss.item())}
        outs = flat_outputs(out)
        for i, o in enumerate(outs):
            store_output(rec, f"out{i}", o, rng)
        rec["n_out"] = np.int64(len(outs))
        names = []
        for pname, p in model.named_parameters():
            names.append(pname)
            for k, v in digest(p.grad if p.grad is not None else torch.zeros_like(p), rng).items():
                rec[f"grad/{pname}/{k}"] = v
            rec[f"wsum/{pname}"] = np.float64(p.detach().double().sum().item())
            rec[f"wabs/{pname}"] = np.float64(p.detach().double().abs().sum().item())
        rec["param_names"] = np.array(names)
        np.savez_compressed(GOLDEN / f"model_{name}_positions_B{B}.npz", **rec)
        print(f"model_{name}_positions_B{B}: loss={loss.item():.8f}")


def model_goldens(ref_models, ref_losses, only=None, only_sizes=None):
    for name, (cls_name, kwargs, feats) in MODELS.items():
        if only and name not in only:
            continue
        # B = 63 / 64: the reference's real batch (run.py:307 default 63, hyper_parameter_bm25.conf:2 -> 64) for the six
        # families of run.py's dispatch; B = 5 / 16 for everything
        big = (63, 64) if name in BIG_BATCH else ()
        sizes = ((5,) if name.endswith(("_t21", "_t22")) else (5, 16)) + big
        if only_sizes:
            sizes = tuple(b for b in sizes if b in only_sizes)
        for B in sizes:
            torch.manual_seed(WEIGHT_SEED)
            model = getattr(ref_models, cls_name)(**kwargs)
            model.train()  # dropout = 0: train mode is deterministic and matches what run.py trains with
            x, y = synthetic_lists(B, 300, feats, seed=DATA_SEED + B, device="cpu")
            out = model(x)
            torch.manual_seed(0)  # MtCutLoss draws an (unused) random Parameter
            crit = build_criterion(ref_losses, name, "f1")
            loss = crit(loss_input(out), y)     # ProbeBase: the tower outputs, `output[-1]` (verify_probe.py:107)
            loss.backward()
            rng = np.random.default_rng(7)
            rec = {"x": x.numpy(), "y": y.numpy(), "loss": np.float64(loss.item())}
            outs = flat_outputs(out)
            for i, o in enumerate(outs):
                store_output(rec, f"out{i}", o, rng)
            rec["n_out"] = np.int64(len(outs))
            names = []
            for pname, p in model.named_parameters():
                names.append(pname)
                d = digest(p.grad if p.grad is not None else torch.zeros_like(p), rng)
                for k, v in d.items():
                    rec[f"grad/{pname}/{k}"] = v
                rec[f"wsum/{pname}"] = np.float64(p.detach().double().sum().item())
                rec[f"wabs/{pname}"] = np.float64(p.detach().double().abs().sum().item())
            rec["param_names"] = np.array(names)
            # float64 truth from the same reference module (tighter target for tolerance accounting)
            m64 = getattr(ref_models, cls_name)(**kwargs).double()
            m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
            m64.train()
            out64 = m64(x.double())
            outs64 = flat_outputs(out64)
            for i, o in enumerate(outs64):
                store_output(rec, f"out{i}_f64", o, rng)
            np.savez_compressed(GOLDEN / f"model_{name}_B{B}.npz", **rec)
            print(f"model_{name}_B{B}: loss={loss.item():.8f}")


PROBE_SEED = 4321


def probe_inputs(B: int, L: int = 300, d: int = 256):
    """Seeded stand-ins for ProbeBase's (experts_in, experts_o): regenerated by the tests instead of being stored."""
    g = torch.Generator().manual_seed(PROBE_SEED + B)
    return torch.randn(B, L, d, generator=g), [torch.randn(B, L, d, generator=g), torch.randn(B, L, d, generator=g)]


def probe_goldens(ref_models, ref_losses):
    """models/Probe.py:102-122 and the stand-alone towers (verify_probe.py:66-71, TaskC / TaskR) with the criteria
    verify_probe.py:82-83 pairs them with: BCELoss for the class probes, RerankLoss for the rerank probes."""
    for B in (4, 9):
        torch.manual_seed(WEIGHT_SEED)
        model = ref_models.Probe()
        e_in, e_o = probe_inputs(B)
        _, y = synthetic_lists(B, 300, 1, seed=DATA_SEED + 50 + B, device="cpu")
        e_in.requires_grad_(True)
        for t in e_o:
            t.requires_grad_(True)
        outs = model(e_in, e_o)
        bce, rr = torch.nn.BCELoss(), ref_losses.RerankLoss()
        parts = [bce(outs[0].squeeze(), y), rr(outs[1].squeeze(), y), bce(outs[2].squeeze(), y),
                 bce(outs[3].squeeze(), y), rr(outs[4].squeeze(), y), rr(outs[5].squeeze(), y)]
        loss = sum(parts)
        loss.backward()
        rng = np.random.default_rng(11)
        rec = {"y": y.numpy(), "loss": np.float64(loss.item()), "losses": np.array([p.item() for p in parts])}
        for i, o in enumerate(outs):
            rec[f"out{i}"] = o.detach().numpy()
        rec["n_out"] = np.int64(len(outs))
        names = []
        for pname, p in model.named_parameters():
            names.append(pname)
            for k, v in digest(p.grad if p.grad is not None else torch.zeros_like(p), rng).items():
                rec[f"grad/{pname}/{k}"] = v
            rec[f"wsum/{pname}"] = np.float64(p.detach().double().sum().item())
            rec[f"wabs/{pname}"] = np.float64(p.detach().double().abs().sum().item())
        rec["param_names"] = np.array(names)
        for tag, t in (("in", e_in), ("o0", e_o[0]), ("o1", e_o[1])):
            for k, v in digest(t.grad, rng).items():
                rec[f"dx/{tag}/{k}"] = v
        np.savez_compressed(GOLDEN / f"probe_B{B}.npz", **rec)
        print(f"probe_B{B}: loss={loss.item():.8f}")


def loss_goldens(ref_losses):
    """Every loss class on random probabilities / labels: value and gradient w.r.t. the model output."""
    rec = {}
    for (B, L) in ((4, 300), (3, 40)):
        x, y = synthetic_lists(B, L, 1, seed=DATA_SEED + 100 + L, device="cpu")
        if L == 300:
            y[1] = 0.0  # a list without any relevant document (N_D = 0 branch of Metric_for_Loss.f1)
        g = torch.Generator().manual_seed(99 + L)
        z = torch.randn(B, L, 1, generator=g) * 2
        rec[f"y_{L}"] = y.numpy()
        rec[f"z_{L}"] = z.numpy()
        for metric in ("f1", "dcg"):
            cases = {
                "choopy": lambda: ref_losses.ChoopyLoss(metric=metric),
                "raml": lambda: ref_losses.AttnCutLoss(metric=metric),
                "kl": lambda: ref_losses.DivLoss(metric=metric, div_type="kl", augmented=True),
                "js": lambda: ref_losses.DivLoss(metric=metric, div_type="js", augmented=True),
                "js_noaug": lambda: ref_losses.DivLoss(metric=metric, div_type="js", augmented=False),
            }
            for cname, make in cases.items():
                zz = z.clone().requires_grad_(True)
                p = torch.softmax(zz, dim=1)
                p.retain_grad()
                loss = make()(p, y)
                loss.backward()
                key = f"{cname}_{metric}_{L}"
                rec[key + "/loss"] = np.float64(loss.item())
                rec[key + "/dp"] = p.grad.numpy()
                rec[key + "/dz"] = zz.grad.numpy()
        # reward matrices straight from Metric_for_Loss
        from oracle import refshim as _r
        _, _, ref_metrics = _r.load()
        for metric in ("f1", "dcg"):
            r = torch.zeros(B, L)
            fn = getattr(ref_metrics.Metric_for_Loss, metric)
            for i in range(B):
                for j in range(L):
                    r[i, j] = fn(y[i], j + 1)
            rec[f"reward_{metric}_{L}"] = r.numpy()
        # auxiliary heads
        base = torch.randn(B, L, 1, generator=g)
        for tag, shift in (("active", -0.5), ("inactive", 0.5)):   # hinge on / hinge off (zero leaf, no grad)
            s = (base + shift * y.unsqueeze(2)).requires_grad_(True)
            rl = ref_losses.RerankLoss()(s, y)
            rl.backward()
            rec[f"rerank_{tag}_{L}/s"] = s.detach().numpy()
            rec[f"rerank_{tag}_{L}/loss"] = np.float64(rl.item())
            rec[f"rerank_{tag}_{L}/ds"] = (s.grad if s.grad is not None else torch.zeros_like(s)).numpy()
        # BiCut loss on random 2-class outputs
        u = torch.randn(B, L, 2, generator=g, requires_grad=True)
        o = torch.softmax(u, dim=2)
        o.retain_grad()
        for metric in ("f1", "nci"):
            if o.grad is not None:
                o.grad = None
                u.grad = None
            bl = ref_losses.BiCutLoss(metric=metric)(o, y)
            bl.backward(retain_graph=True)
            rec[f"bicut_{metric}_{L}/loss"] = np.float64(bl.item())
            rec[f"bicut_{metric}_{L}/do"] = o.grad.numpy().copy()
        rec[f"bicut_{L}/u"] = u.detach().numpy()
    np.savez_compressed(GOLDEN / "losses.npz", **rec)
    print("losses.npz written")


def metric_goldens(ref_metrics):
    rec = {}
    # the reference's only known-answer vector (utils/metrics.py:104-109)
    x = np.array([[1, 0, 1], [0, 0, 1], [1, 0, 0]])
    k = np.array([1, 2, 1])
    rec["known/f1"] = np.float64(ref_metrics.Metric.f1(x, k))
    rec["known/dcg"] = np.float64(ref_metrics.Metric.dcg(x, k))
    for (B, L) in ((64, 300), (7, 40)):
        _, y = synthetic_lists(B, L, 1, seed=DATA_SEED + 7 + L, device="cpu")
        y = y.numpy()
        y[3] = 0.0
        rng = np.random.default_rng(5 + L)
        p = rng.random((B, L), dtype=np.float32)
        p[5, 10] = p[5, L - 3] = 2.0  # a tie: first maximum wins
        ks = np.argmax(p, axis=1) + 1
        ks[0], ks[1] = 1, L
        rec[f"y_{L}"], rec[f"p_{L}"], rec[f"k_{L}"] = y, p, ks
        rec[f"f1_{L}"] = np.array([ref_metrics.Metric.f1(y[i:i + 1], ks[i:i + 1]) for i in range(B)], dtype=np.float64)
        rec[f"dcg_{L}"] = np.array([ref_metrics.Metric.dcg(y[i:i + 1], ks[i:i + 1]) for i in range(B)], dtype=np.float64)
        rec[f"f1_mean_{L}"] = np.float64(ref_metrics.Metric.f1(y, ks))
        rec[f"dcg_mean_{L}"] = np.float64(ref_metrics.Metric.dcg(y, ks))
        # every cut position of one list (exercises all pairwise-summation lengths)
        all_k = np.arange(1, L + 1)
        rec[f"dcg_allk_{L}"] = np.array([ref_metrics.Metric.dcg(y[2:3], all_k[j:j + 1]) for j in range(L)], dtype=np.float64)
        rec[f"f1_allk_{L}"] = np.array([ref_metrics.Metric.f1(y[2:3], all_k[j:j + 1]) for j in range(L)], dtype=np.float64)
    np.savez_compressed(GOLDEN / "metrics.npz", **rec)
    print("metrics.npz written", rec["known/f1"], rec["known/dcg"])


def rank_metric_goldens(ref_metrics):
    """Metric.taskr_metric / taskc_metric (utils/metrics.py:40-76) on seeded predictions.  The taskr inputs are
    tie-free (checked): the reference's argsort is unstable, so tied inputs have no reference answer; AUC is
    well-defined under ties and gets a tied variant."""
    rec = {}
    for (B, L) in ((16, 300), (5, 40)):
        _, y = synthetic_lists(B, L, 1, seed=DATA_SEED + 300 + L, device="cpu")
        y = y.numpy()
        y[3] = 0.0   # a one-class list: skipped by taskc_metric, all-negative for taskr_metric
        rng = np.random.default_rng(17 + L)
        p = rng.random((B, L), dtype=np.float32)
        assert all(len(np.unique(row)) == L for row in p)
        p_tied = np.round(p, 1).astype(np.float32)
        rec[f"y_{L}"], rec[f"p_{L}"], rec[f"p_tied_{L}"] = y, p, p_tied
        rec[f"taskr_{L}"] = np.array([ref_metrics.Metric.taskr_metric(y[i:i + 1], p[i:i + 1]) for i in range(B)])
        rec[f"taskr_mean_{L}"] = np.float64(ref_metrics.Metric.taskr_metric(y, p))
        for tag, pp in (("", p), ("_tied", p_tied)):
            two = [i for i in range(B) if 0 < y[i].sum() < L]
            rec[f"auc_lists{tag}_{L}"] = np.array(two, dtype=np.int64)
            rec[f"auc{tag}_{L}"] = np.array([ref_metrics.Metric.taskc_metric(y[i:i + 1], pp[i:i + 1]) for i in two])
            rec[f"taskc_mean{tag}_{L}"] = np.float64(ref_metrics.Metric.taskc_metric(y, pp))
    np.savez_compressed(GOLDEN / "rank_metrics.npz", **rec)
    print("rank_metrics.npz written", rec["taskr_mean_300"], rec["taskc_mean_300"], rec["taskc_mean_tied_300"])


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    ref_models, ref_losses, ref_metrics = refshim.load()
    torch.set_num_threads(8)
    only = set(sys.argv[1:])          # e.g. `python -m oracle.make_golden moecut plecut`: only those model fixtures
    if "positions" in only:           # `python -m oracle.make_golden positions`: attend-within-a-list fixtures
        positions_goldens(ref_models, ref_losses)
        return
    if "traj" in only:                # `python -m oracle.make_golden traj [names]`: the multi-step fixtures
        trajectory_goldens(ref_models, ref_losses, only - {"traj"})
        return
    if "big" in only:                 # `python -m oracle.make_golden big [names]`: only the B = 63 / 64 fixtures
        model_goldens(ref_models, ref_losses, (only - {"big"}) or set(BIG_BATCH), only_sizes=(63, 64))
        return
    if only == {"rank"}:
        rank_metric_goldens(ref_metrics)
        return
    if not only:
        rank_metric_goldens(ref_metrics)
        metric_goldens(ref_metrics)
        loss_goldens(ref_losses)
    if not only or "probe" in only:
        probe_goldens(ref_models, ref_losses)
    if only != {"probe"}:
        model_goldens(ref_models, ref_losses, only)


if __name__ == "__main__":
    main()
