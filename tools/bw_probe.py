"""HBM ceilings for the access mixes the kernels have: write-only (fill), read-only (sum), copy. GPU box only."""
import torch
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e-3
n = 1 << 30
x = torch.empty(n, device="cuda"); y = torch.empty(n, device="cuda")
print(f"fill  (write only) {4*n/t(lambda: x.fill_(1.0))/1e9:8.0f} GB/s")
print(f"sum   (read only)  {4*n/t(lambda: x.sum())/1e9:8.0f} GB/s")
print(f"copy  (read+write) {8*n/t(lambda: y.copy_(x))/1e9:8.0f} GB/s")
print(f"relu_ (read+write in place) {8*n/t(lambda: x.relu_())/1e9:8.0f} GB/s")
