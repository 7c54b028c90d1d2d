"""GPU diagnostic: per-tensor cosine between our 5-step Adam displacement and the reference's (tests/golden/traj_*.npz)."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "ranked-list-truncation_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
from helpers import load_golden  # noqa: E402
import test_zzzz_trajectory_gpu as T  # noqa: E402

for name in sys.argv[1:] or ["mtchoopy", "mtattncut"]:
    g = load_golden(f"traj_{name}.npz")
    model, init, losses = T._run_module(name, g)
    print(name, "losses", losses, "ref", g["losses"].tolist())
    rows = []
    for n, p in model.named_parameters():
        idx = g[f"delta/{n}/idx"]
        ours = (p.detach() - init[n]).double().cpu().numpy().ravel()[idx]
        ref = g[f"delta/{n}/val"]
        cos = float((ours * ref).sum() / (np.linalg.norm(ours) * np.linalg.norm(ref) + 1e-300))
        rows.append((cos, n, float(np.linalg.norm(ref)), float(np.abs(ours - ref).max())))
    for cos, n, nr, md in sorted(rows)[:12]:
        print(f"   cos {cos:+.4f} |ref| {nr:.3e} max|d| {md:.3e}  {n}")
