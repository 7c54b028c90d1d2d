"""GPU diagnostic: per-tensor gradient errors of the multi-task models against the reference goldens
(tests/golden/model_*_B5.npz).  Prints, for every variant, the global figures the parity tests assert and the five
worst tensors, so that a tolerance miss can be attributed to a kernel.  Not a test: run under gpurun."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "ranked-list-truncation_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)

from helpers import MODEL_KW, build_model, grad_errors, load_golden, output_error  # noqa: E402


def main():
    from utils import losses
    names = sys.argv[1:] or ["mtattncut", "mtattncut_t21", "mtattncut_t22", "mtchoopy_t22", "mmoecut", "mmoecut_t21",
                             "mmoecut_t22"]
    for name in names:
        g = load_golden(f"model_{name}_B5.npz")
        model = build_model(name).cuda().train()
        x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
        out = model(x)
        for i, o in enumerate(out):
            err, ref_max = output_error(o, g, f"out{i}")
            print(f"{name} out{i}: err {err:.3e} ref_max {ref_max:.3e} rel {err / ref_max:.3e}")
        nt = MODEL_KW[name][1]["num_tasks"]
        crit = (losses.MtCutLoss(metric="f1", num_tasks=nt) if name.startswith("mmoecut") else
                losses.MtCutLoss(metric="f1", rerank_weight=0.5, classi_weight=0.5, num_tasks=nt)).cuda()
        loss = crit(out, y)
        print(f"{name} loss {loss.item():.7f} ref {float(g['loss']):.7f}")
        loss.backward()
        named = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
        rel_l2, rel_max, rel_norm = grad_errors(named, g)
        print(f"{name} grads: rel_l2 {rel_l2:.3e} rel_max {rel_max:.3e} rel_norm {rel_norm:.3e}")
        gmax = max(float(g[f"grad/{n}/absmax"]) for n in g["param_names"])
        rows = []
        for n in g["param_names"]:
            n = str(n)
            got = named[n].detach().double().cpu().numpy().ravel()
            d = got[g[f"grad/{n}/idx"]] - g[f"grad/{n}/val"]
            rows.append((float(np.abs(d).max()) / gmax, float(g[f"grad/{n}/absmax"]) / gmax, n))
        for e, m, n in sorted(rows, reverse=True)[:6]:
            print(f"    {n:60s} max|d|/gmax {e:.3e}  absmax/gmax {m:.3e}")


if __name__ == "__main__":
    main()
