"""GPU: clock64 timeline of the tcgen05 attention forward kernel (first CTA, items 16-39): where the MMA issuers, the convert\nwarps, the softmax warps and the epilogue warps spend an item.  Uses the library's timeline hook (rlt_ffn_fused_set_timeline)."""
import sys, ctypes as C
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
from rlt_b200 import ops
G, S, L, d, nh = 32, 64, 300, 128, 8
qkv = torch.randn(G * S * L, 3 * d, device="cuda")
for _ in range(2):
    ops.attention_lists_fwd(qkv, G, S, L, d, nh)
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib = ops.lib()
lib.rlt_ffn_fused_set_timeline(C.c_void_p(buf.data_ptr()))
ops.attention_lists_fwd(qkv, G, S, L, d, nh)
torch.cuda.synchronize()
lib.rlt_ffn_fused_set_timeline(C.c_void_p(0))
t = buf.view(64, 16).cpu().numpy(); t0 = t[0, 0]
names = ["mma:top", "S waits", "S issued", "pv:top", "p_full", "PV issued", "ep:stored", "raw_full", "op_empty", "cv done", "sm:top", "s_full", "P written", "o_full", "sm done", "ep:O read"]
print("item " + " ".join(f"{n:>9s}" for n in names))
for i in range(16, 40):
    print(f"{i:4d} " + " ".join(f"{int(t[i, k] - t0) if t[i, k] else 0:9d}" for k in range(16)))
