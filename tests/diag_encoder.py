"""Diagnostic (run by hand on a GPU box: `python tests/diag_encoder.py`; not collected by pytest): per-tensor gradient
error of the encoder layer, tcgen05 vs SIMT backend vs float64 oracle.  Lives under tests/ because it imports oracle/."""
import sys, ctypes, warnings
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
from rlt_b200 import _lib, ops
from rlt_b200.autograd import EncoderStack
from oracle import rlt_oracle as O

def tf32(x):
    out = torch.empty_like(x)
    _lib.check(_lib.load().rlt_round_tf32(_lib.ptr(x), _lib.ptr(out), ctypes.c_size_t(x.numel()), _lib.stream_ptr()), "r")
    return out

# 1) raw dW GEMM on UNROUNDED inputs
for (T, M, N) in [(200, 2048, 128), (4800, 2048, 128), (200, 128, 2048)]:
    A = torch.randn(T, M, device="cuda"); B = torch.randn(T, N, device="cuda")
    for be in (0, 1):
        _lib.set_option("gemm_backend", be)
        C = torch.zeros(M, N, device="cuda")
        _lib.check(_lib.load().rlt_grad_weight(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), T, M, N, ctypes.c_float(1.0), _lib.stream_ptr()), "dw")
        ref = A.double().t() @ B.double()
        ref_r = tf32(A).double().t() @ tf32(B).double()
        print(f"dw T={T} M={M} N={N} backend={be}: err vs exact {(C.double()-ref).abs().max().item():.3e} "
              f"vs rounded-operand product {(C.double()-ref_r).abs().max().item():.3e}  max|ref| {ref.abs().max().item():.3e}")
    _lib.set_option("gemm_backend", 0)
# 1b) nn GEMM on unrounded inputs
for (M, N, K) in [(200, 2048, 128), (200, 128, 2048), (200, 128, 384)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    _lib.check(_lib.load().rlt_linear_nn(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, _lib.stream_ptr()), "nn")
    ref = A.double() @ B.double(); ref_r = tf32(A).double() @ tf32(B).double()
    print(f"nn M={M} N={N} K={K}: err vs exact {(C.double()-ref).abs().max().item():.3e} vs rounded {(C.double()-ref_r).abs().max().item():.3e} max {ref.abs().max().item():.3e}")

# 2) layer
d, n_head, S, L, G = 128, 8, 16, 300, 1
torch.manual_seed(d + S)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    layer = torch.nn.TransformerEncoderLayer(d_model=d, nhead=n_head, dropout=0.0)
sd = {k: v.detach().clone() for k, v in layer.state_dict().items()}
torch.manual_seed(7)
x = torch.randn(G * S, L, d); dy = torch.randn(G * S, L, d) * 0.01
sd64 = {("layers.0." + k): v.double().requires_grad_(True) for k, v in sd.items()}
x64 = x.double().requires_grad_(True)
out64 = O.encoder_stack(x64, sd64, "", n_head)
(out64 * dy.double()).sum().backward()
res = {}
for be in (1, 0):
    _lib.set_option("gemm_backend", be)
    params = [sd[n].cuda().requires_grad_(True) for n in ops.ENCODER_PARAM_ORDER]
    xc = x.cuda().requires_grad_(True)
    out = EncoderStack.apply(xc, n_head, G, 1e-5, *params)
    (out * dy.cuda()).sum().backward()
    res[be] = (out.detach(), [p.grad for p in params], xc.grad)
    print(f"backend {be}: out err {(out.cpu().double()-out64).abs().max().item():.3e}")
    for n, p in zip(ops.ENCODER_PARAM_ORDER, params):
        ref = sd64['layers.0.' + n].grad
        e = (p.grad.cpu().double() - ref)
        print(f"   {n:28s} max|d| {e.abs().max().item():.3e}  max|ref| {ref.abs().max().item():.3e}  relL2 {(e.norm()/ref.norm()).item():.3e}")
    e = xc.grad.cpu().double() - x64.grad
    print(f"   {'dx':28s} max|d| {e.abs().max().item():.3e}  max|ref| {x64.grad.abs().max().item():.3e}  relL2 {(e.norm()/x64.grad.norm()).item():.3e}")
_lib.set_option("gemm_backend", 0)
