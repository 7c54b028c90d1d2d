"""tcgen05/TMA GEMM building blocks vs float64 matmul of the tf32-rounded operands (GPU)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32(x):
    from rlt_b200 import _lib
    out = torch.empty_like(x)
    _lib.check(_lib.load().rlt_round_tf32(_lib.ptr(x), _lib.ptr(out), ctypes.c_size_t(x.numel()), _lib.stream_ptr()),
               "rlt_round_tf32")
    return out


def test_round_tf32_matches_bit_definition():
    x = torch.randn(100003, device="cuda") * 3
    r = _tf32(x)
    bits = x.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    expect = ((bits + 0x1000) & 0xFFFFE000).to(torch.int64)
    expect = torch.where(expect >= 2**31, expect - 2**32, expect).to(torch.int32).view(torch.float32)
    assert torch.equal(r, expect)


def test_probe_tma_tfloat32_rounding(capsys):
    """What a TFLOAT32 tensor map does to fp32 data while TMA fills shared memory: measured = round to
    nearest (never truncation); it agrees with cvt.rna.tf32 except on exact ties (~2^-13 of random values),
    so operands can stay exact fp32 in HBM and be rounded by the copy engine for free."""
    from rlt_b200 import _lib
    n_rna = n_total = 0
    worst = 0.0
    for seed in range(16):
        g = torch.Generator(device="cuda").manual_seed(seed)
        x = (torch.randn(128, 32, device="cuda", generator=g) * 2).contiguous()
        out = torch.empty_like(x)
        _lib.check(_lib.load().rlt_probe_tma_tf32(_lib.ptr(x), _lib.ptr(out), 128, _lib.stream_ptr()), "probe")
        torch.cuda.synchronize()
        assert ((out.view(torch.int32) & 0x1FFF) == 0).all()                  # result is a tf32 value
        n_rna += int((out == _tf32(x)).sum())
        n_total += x.numel()
        worst = max(worst, ((out - x).abs() / x.abs().clamp_min(1e-30)).max().item())
    with capsys.disabled():
        print(f"\n[probe] TMA TFLOAT32: {n_rna}/{n_total} equal cvt.rna.tf32; worst relative change {worst:.3e} "
              f"(half ulp of tf32 = {2**-11:.3e})")
    assert worst <= 2 ** -11 * 1.0001          # round-to-nearest, not truncation
    assert n_rna >= n_total - 64               # differs from rna at most on ties


@pytest.mark.parametrize("backend", [1, 0])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (300, 384, 128), (1500, 2048, 128), (777, 128, 2048),
                                   (19200, 256, 256), (64, 32, 64), (4096, 512, 256)])
def test_linear(backend, M, N, K):
    from rlt_b200 import _lib
    lib = _lib.load()
    _lib.set_option("gemm_backend", backend)
    try:
        g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
        A = _tf32(torch.randn(M, K, device="cuda", generator=g))
        B = _tf32(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
        bias = torch.randn(N, device="cuda", generator=g)
        C = torch.full((M, N), float("nan"), device="cuda")
        _lib.check(lib.rlt_linear(_lib.ptr(A), _lib.ptr(B), _lib.ptr(bias), _lib.ptr(C), M, N, K,
                                  ctypes.c_float(1.0), 1, _lib.stream_ptr()), "rlt_linear")
        torch.cuda.synchronize()
        ref = torch.relu(A.double() @ B.double().t() + bias.double())
        err = (C.double() - ref).abs().max().item()
        assert err < 2e-5 * max(1.0, ref.abs().max().item()), (backend, M, N, K, err)
    finally:
        _lib.set_option("gemm_backend", 0)


@pytest.mark.parametrize("backend", [1, 0])
@pytest.mark.parametrize("T,M,N", [(256, 128, 128), (1500, 384, 128), (19200, 2048, 128), (5000, 128, 2048),
                                   (333, 256, 256), (40, 128, 32)])
def test_grad_weight(backend, T, M, N):
    from rlt_b200 import _lib
    lib = _lib.load()
    _lib.set_option("gemm_backend", backend)
    try:
        g = torch.Generator(device="cuda").manual_seed(T + M + N)
        A = _tf32(torch.randn(T, M, device="cuda", generator=g))
        B = _tf32(torch.randn(T, N, device="cuda", generator=g))
        C0 = torch.randn(M, N, device="cuda", generator=g)
        C = C0.clone()
        _lib.check(lib.rlt_grad_weight(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), T, M, N, ctypes.c_float(0.5),
                                       _lib.stream_ptr()), "rlt_grad_weight")
        torch.cuda.synchronize()
        ref = C0.double() + 0.5 * (A.double().t() @ B.double())
        err = (C.double() - ref).abs().max().item()
        assert err < 3e-5 * max(1.0, ref.abs().max().item()) * max(1.0, (T / 256) ** 0.5), (backend, T, M, N, err)
    finally:
        _lib.set_option("gemm_backend", 0)


# ---------------------------------------------------------------------------------------------
# fp16-operand building blocks of the FFN hidden path
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 2048), (777, 128, 2048), (1500, 256, 512), (300, 128, 64), (19200, 128, 2048)])
def test_linear_f16_operands(M, N, K):
    from rlt_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    B = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    C = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.rlt_linear_f16(_lib.ptr(A), _lib.ptr(B), _lib.ptr(bias), _lib.ptr(C), M, N, K, ctypes.c_float(0.5), 0,
                                  _lib.stream_ptr()), "rlt_linear_f16")
    torch.cuda.synchronize()
    ref = 0.5 * (A.double() @ B.double().t()) + bias.double()
    err = (C.double() - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), (M, N, K, err)


@pytest.mark.parametrize("T,M,N", [(256, 128, 128), (1500, 128, 2048), (19200, 2048, 128), (5000, 256, 256), (333, 64, 128)])
def test_grad_weight_f16_operands(T, M, N):
    from rlt_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(T + M + N)
    A = torch.randn(T, M, device="cuda", generator=g).half()
    B = torch.randn(T, N, device="cuda", generator=g).half()
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C = C0.clone()
    _lib.check(lib.rlt_grad_weight_f16(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), T, M, N, ctypes.c_float(0.5),
                                       _lib.stream_ptr()), "rlt_grad_weight_f16")
    torch.cuda.synchronize()
    ref = C0.double() + 0.5 * (A.double().t() @ B.double())
    err = (C.double() - ref).abs().max().item()
    assert err < 3e-5 * max(1.0, ref.abs().max().item()) * max(1.0, (T / 256) ** 0.5), (T, M, N, err)


@pytest.mark.parametrize("M,N,K", [(300, 2048, 128), (1500, 256, 256)])
def test_linear_half_precision_output(M, N, K):
    from rlt_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K + 1)
    A = _tf32(torch.randn(M, K, device="cuda", generator=g))
    B = _tf32(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    bias = torch.randn(N, device="cuda", generator=g)
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
    _lib.check(lib.rlt_linear_out_f16(_lib.ptr(A), _lib.ptr(B), _lib.ptr(bias), _lib.ptr(C), M, N, K, 1, _lib.stream_ptr()),
               "rlt_linear_out_f16")
    torch.cuda.synchronize()
    ref = torch.relu(A.double() @ B.double().t() + bias.double())
    err = ((C.double() - ref).abs() / ref.abs().clamp_min(1.0)).max().item()
    assert err < 2 ** -11 * 1.01, (M, N, K, err)          # one fp16 rounding of an fp32-accurate value


def test_convert_f16_with_device_scale():
    from rlt_b200 import _lib
    lib = _lib.load()
    x = torch.randn(4096 * 12, device="cuda") * 1e-5
    scale = torch.tensor([2.0 ** 20, 2.0 ** -20], device="cuda")
    out = torch.empty(x.numel(), device="cuda", dtype=torch.float16)
    _lib.check(lib.rlt_convert_f16(_lib.ptr(x), _lib.ptr(out), ctypes.c_size_t(x.numel()), _lib.ptr(scale), _lib.stream_ptr()),
               "rlt_convert_f16")
    torch.cuda.synchronize()
    assert torch.equal(out, (x * 2.0 ** 20).half())
