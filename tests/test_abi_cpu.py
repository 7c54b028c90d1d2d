"""CPU: the C-ABI library loads and exports every symbol declared in include/rlt_b200.h; the host
mirror refuses CPU tensors (no fallback)."""
import ctypes

import pytest
import torch

from rlt_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/rlt_b200.h but not exported: {missing}"
    assert b"sm_100a" in lib.rlt_version()


def test_every_export_is_bound_with_the_header_prototype():
    """ctypes argtypes / restype of every function come from include/rlt_b200.h (rlt_b200/_lib.py: prototypes()): a call
    whose arguments do not convert to the declared C types raises ArgumentError on the host."""
    lib = _lib.load()
    protos = _lib.prototypes()
    assert set(protos) == set(_lib.declared_symbols())
    for name, (restype, argtypes) in protos.items():
        fn = getattr(lib, name)
        assert fn.restype is restype and list(fn.argtypes) == argtypes, name
    assert protos["rlt_version"] == (ctypes.c_char_p, [])
    assert protos["rlt_encoder_layer_saved_bytes"] == (ctypes.c_size_t, [ctypes.c_void_p])
    assert protos["rlt_adam_step_masked"][1][12:18] == [ctypes.c_double] * 6
    assert protos["rlt_pair_softmax_fwd"][1].count(ctypes.c_float) == 1
    with pytest.raises(ctypes.ArgumentError):
        lib.rlt_set_option(b"gemm_backend", 1.5)                 # a float where the header says int
    with pytest.raises(ctypes.ArgumentError):
        lib.rlt_round_tf32(None, None, "4", None)                # a str where the header says size_t
    with pytest.raises(TypeError):
        lib.rlt_set_option(b"gemm_backend")                      # too few arguments


def test_status_codes_and_last_error_without_gpu():
    lib = _lib.load()
    assert lib.rlt_set_option(b"no_such_option", 1) == -1
    assert b"unknown option" in lib.rlt_last_error()
    assert lib.rlt_round_tf32(None, None, ctypes.c_size_t(4), None) == -1
    assert _lib.get_option("gemm_backend") == 0
    assert _lib.get_option("tma_round") == 1


def test_modules_refuse_cpu_tensors():
    import models
    from utils import losses
    torch.manual_seed(0)
    m = models.Choopy(seq_len=40, dropout=0.0)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.randn(2, 40, 1))
    with pytest.raises(RuntimeError, match="no CPU"):
        losses.ChoopyLoss()(torch.rand(2, 40, 1), torch.zeros(2, 40))


def test_state_dict_keys_match_reference_layout():
    import models
    torch.manual_seed(0)
    keys = set(models.MMOECut(seq_len=40, dropout=0.0).state_dict())
    for k in ("pre_encoding.weight_ih_l0", "pre_encoding.weight_hh_l1_reverse",
              "experts.2.attention_layer.layers.0.self_attn.in_proj_weight", "w_gates.0", "w_gates.2",
              "towers.0.classification_layer.0.weight", "towers.1.rerank_layer.0.bias", "towers.2.cut_layer.0.weight"):
        assert k in keys, k
    keys = set(models.Choopy(seq_len=40).state_dict())
    assert {"position_encoding", "attention_layer.layers.2.norm2.bias", "decison_layer.0.weight"} <= keys
    keys = set(models.BiCut(input_size=3).state_dict())
    assert {"bilstm.weight_ih_l0", "bilstm.bias_hh_l1_reverse", "fc.weight", "softmax.1.bias"} <= keys
    keys = set(models.MtAttnCut().state_dict())
    assert {"pre_encoding.weight_hh_l0", "encoding_layer.layers.0.linear1.weight", "classi.0.weight", "rerank.bias",
            "decison_layer.0.bias"} <= keys


def test_probe_modules_match_reference_layout():
    """models/Probe.py, Classification.py, Rerank.py: state_dict keys, constructor defaults and the CPU refusal."""
    import models
    torch.manual_seed(0)
    base = models.ProbeBase(seq_len=40, dropout=0.0)
    keys = set(base.state_dict())
    assert {"pre_encoding.weight_ih_l0", "experts.1.attention_layer.layers.0.linear2.weight", "w_gates.2",
            "towers.0.classification_layer.0.weight", "towers.1.rerank_layer.0.bias", "towers.2.cut_layer.0.weight"} <= keys
    assert "experts.2.attention_layer.layers.0.linear2.weight" not in keys          # num_experts defaults to 2
    assert base.w_gates[0].shape == (40 * 256, 2)
    keys = set(models.Probe().state_dict())
    assert keys == {f"probe_{n}.{layer}.0.{p}" for n, layer in (("c1", "classification_layer"), ("r1", "rerank_layer"),
                    ("ce1", "classification_layer"), ("ce2", "classification_layer"), ("re1", "rerank_layer"),
                    ("re2", "rerank_layer")) for p in ("weight", "bias")}
    assert set(models.TaskC().state_dict()) == {"classification_layer.0.weight", "classification_layer.0.bias"}
    assert models.TaskR().rerank_layer[0].in_features == 128
    with pytest.raises(RuntimeError, match="no CPU"):
        models.TowerClass(d_model=256)(torch.randn(2, 40, 256))


def test_python_call_sites_match_header_arity():
    """Static twin of the run-time argtypes check (which only fires when a call site executes): every
    `<lib>.rlt_*(...)` call in the host package must pass exactly as many arguments as include/rlt_b200.h declares."""
    import ast
    import re
    from pathlib import Path
    text = re.sub(r"/\*.*?\*/", "", _lib.HEADER_PATH.read_text(), flags=re.S)
    arity = {}
    for name, params in re.findall(r"\b(rlt_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        params = params.strip()
        arity[name] = 0 if params in ("", "void") else params.count(",") + 1
    assert len(arity) >= 20 and arity["rlt_rank_metrics"] == 9 and arity["rlt_gather_lists"] == 12
    pkg = Path(_lib.__file__).resolve().parent.parent
    checked = 0
    for path in sorted(pkg.rglob("*.py")):
        for node in ast.walk(ast.parse(path.read_text())):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in arity:
                assert not node.keywords and not any(isinstance(a, ast.Starred) for a in node.args), (path.name, node.func.attr)
                assert len(node.args) == arity[node.func.attr], (path.name, node.lineno, node.func.attr, len(node.args))
                checked += 1
    assert checked >= 30, checked


def test_rank_metrics_refuse_a_host_without_cuda():
    import numpy as np
    from utils.metrics import Metric
    if torch.cuda.is_available():
        pytest.skip("needs a host without CUDA")
    for fn in (Metric.taskr_metric, Metric.taskc_metric):
        with pytest.raises(RuntimeError, match="no CPU"):
            fn(np.zeros((1, 4), np.float32), np.zeros((1, 4), np.float32))
