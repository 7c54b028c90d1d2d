import sys, copy, torch
sys.path.insert(0, "tests"); sys.path.insert(0, "ranked-list-truncation_b200"); sys.path.insert(0, ".")
from helpers import build_model
from rlt_b200.data import synthetic_lists
from rlt_b200.engine import Engine
from rlt_b200.optim import FusedAdam
from rlt_b200 import ops
name = "choopy"; S, L = 63, 300
model_a = build_model(name).cuda().train(); model_b = copy.deepcopy(model_a); model_c = copy.deepcopy(model_a)
batches = [synthetic_lists(S, L, 1, seed=40 + i, device="cuda") for i in range(5)]
def eager(model):
    eng = Engine(model, n_groups=1, group_size=S, seq_len=L); opt = FusedAdam.for_engine(eng, lr=1e-3, weight_decay=1e-3)
    out = []
    for x, y in batches:
        out.append(eng.train_step(x, y).item()); opt.step()
    return out
print("eager a", eager(model_a))
print("eager c", eager(model_c))
eng_b = Engine(model_b, n_groups=1, group_size=S, seq_len=L); opt_b = FusedAdam.for_engine(eng_b, lr=1e-3, weight_decay=1e-3)
xs, ys = batches[0][0].clone(), batches[0][1].clone()
replay = eng_b.capture_train_step(xs, ys, optimizer=opt_b, metrics=True)
for (n, pa), (_, pb) in zip(copy.deepcopy(build_model(name)).cuda().named_parameters(), model_b.named_parameters()):
    if not torch.equal(pa, pb): print("param differs after capture", n, (pa - pb).abs().max().item())
out = []
for x, y in batches:
    xs.copy_(x); ys.copy_(y); replay(); out.append(eng_b.loss.item())
print("graph  ", out)
