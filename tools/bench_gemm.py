"""Micro-benchmark of the tcgen05 TF32 GEMM building blocks (GPU box only; prints TFLOP/s)."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
from rlt_b200 import _lib  # noqa: E402

lib = _lib.load()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    for (M, N, K) in [(262144, 2048, 128), (262144, 128, 2048), (262144, 384, 128), (262144, 128, 128),
                      (262144, 256, 256), (262144, 1024, 256)]:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        bias = torch.randn(N, device="cuda")
        C = torch.empty(M, N, device="cuda")
        s = _lib.stream_ptr()
        t = timeit(lambda: _lib.check(lib.rlt_linear(_lib.ptr(A), _lib.ptr(B), _lib.ptr(bias), _lib.ptr(C), M, N, K,
                                                     ctypes.c_float(1.0), 1, s), "linear"))
        tb = timeit(lambda: torch.relu(torch.nn.functional.linear(A, B, bias)))
        print(f"linear  M={M} N={N} K={K}: {t*1e3:8.3f} ms  {2*M*N*K/t/1e12:7.1f} TFLOP/s  "
              f"out {M*N*4/t/1e9:7.0f} GB/s | torch(cuBLAS fp32) {tb*1e3:8.3f} ms", flush=True)
    for (T, M, N) in [(262144, 2048, 128), (262144, 128, 2048), (262144, 384, 128), (262144, 256, 256)]:
        A = torch.randn(T, M, device="cuda")
        B = torch.randn(T, N, device="cuda")
        C = torch.zeros(M, N, device="cuda")
        s = _lib.stream_ptr()
        t = timeit(lambda: _lib.check(lib.rlt_grad_weight(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), T, M, N,
                                                          ctypes.c_float(1.0), s), "dw"))
        print(f"grad_w  T={T} M={M} N={N}: {t*1e3:8.3f} ms  {2*M*N*T/t/1e12:7.1f} TFLOP/s  "
              f"in {(T*(M+N))*4/t/1e9:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
