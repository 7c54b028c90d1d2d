import sys, torch
sys.path.insert(0, "ranked-list-truncation_b200"); sys.path.insert(0, "tests")
from rlt_b200 import _lib
from rlt_b200.autograd import BiLstm
B, L, F = 64 * 160, 12, 3          # 160 tiles x 2 directions = 320 CTAs > 148 SMs
torch.manual_seed(1)
lstm = torch.nn.LSTM(input_size=F, hidden_size=128, num_layers=2, batch_first=True, bidirectional=True).cuda()
x = torch.randn(B, L, F, device="cuda"); dy = torch.randn(B, L, 256, device="cuda") * 1e-4
res = {}
for be in (1, 0):
    _lib.set_option("lstm_backend", be)
    ps = [p.detach().clone().requires_grad_(True) for p in lstm._flat_weights]
    y = BiLstm.apply(x, 128, 2, *ps)
    (y * dy).sum().backward()
    torch.cuda.synchronize()
    res[be] = (y.detach(), [p.grad for p in ps])
_lib.set_option("lstm_backend", 0)
e = (res[0][0] - res[1][0]).abs().max().item()
g = max(((a - b).abs().max() / (b.abs().max() + 1e-12)).item() for a, b in zip(res[0][1], res[1][1]))
print("multi-wave grid ok: max |dy|", e, "max rel grad err", g)
