"""GPU parity of the per-list kernels (K3 cut losses, K4 eval, aux heads, BiCut loss) against the oracle
and the reference goldens, called through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import rlt_oracle as O

pytestmark = pytest.mark.gpu

KINDS = {"choopy": ("choopy", 1.0), "raml": ("raml", 0.95), "kl": ("kl", 0.85), "js": ("js", 0.85), "js_noaug": ("js", 1.0)}


@pytest.mark.parametrize("L", [300, 40])
@pytest.mark.parametrize("metric", ["f1", "dcg"])
@pytest.mark.parametrize("kind", list(KINDS))
def test_cut_loss_kernel_vs_reference_golden(kind, metric, L):
    """fp32 kernel vs the reference's autograd: loss rel <= 1e-5, gradients max|d| <= 1e-4 * max|ref| (the DCG
    reward is a float32 prefix scan instead of L separate float32 sums)."""
    from rlt_b200 import ops
    g = load_golden("losses.npz")
    y = torch.from_numpy(g[f"y_{L}"]).cuda()
    z = torch.from_numpy(g[f"z_{L}"]).cuda().reshape(y.shape).contiguous()
    B = y.shape[0]
    lk, tau = KINDS[kind]
    key = f"{kind}_{metric}_{L}"
    ref_loss = float(g[key + "/loss"])
    for input_kind, ref_grad in ((0, g[key + "/dz"]), (1, g[key + "/dp"])):
        inp = z if input_kind == 0 else torch.softmax(z, dim=1).contiguous()
        grad = torch.empty_like(z)
        per = torch.empty(B, device="cuda")
        loss = torch.empty((), device="cuda")
        probs = torch.empty_like(z)
        ops.cut_loss(inp, y, loss_kind=lk, metric=metric, tau=tau, input_kind=input_kind, probs_out=probs, grad=grad,
                     loss_per_list=per, loss_out=loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
        assert abs(loss.item() - ref_loss) <= 2e-5 * max(1.0, abs(ref_loss)), (input_kind, loss.item(), ref_loss)
        ref = ref_grad.reshape(B, L)
        err = np.abs(grad.cpu().numpy() - ref).max()
        assert err <= 2e-4 * max(np.abs(ref).max(), 1e-6), (input_kind, err, np.abs(ref).max())
        assert torch.allclose(probs, torch.softmax(z, dim=1), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("L", [300, 40])
@pytest.mark.parametrize("metric", ["f1", "dcg"])
@pytest.mark.parametrize("kind", list(KINDS))
def test_cut_loss_from_label_bit_masks_vs_reference_golden(kind, metric, L):
    """rlt_cut_loss_bits (labels as the words of rlt_pack_labels) against the same reference goldens, and bit-identical
    to the float-label call: both run the packed kernel of cut_loss_pair.cuh, only the label fetch differs."""
    from rlt_b200 import ops
    g = load_golden("losses.npz")
    y = torch.from_numpy(g[f"y_{L}"]).cuda()
    z = torch.from_numpy(g[f"z_{L}"]).cuda().reshape(y.shape).contiguous()
    B = y.shape[0]
    lk, tau = KINDS[kind]
    key = f"{kind}_{metric}_{L}"
    bits = ops.pack_labels(y)
    out = {}
    for fmt in ("bits", "float"):
        grad, per, loss, probs = torch.empty_like(z), torch.empty(B, device="cuda"), torch.empty((), device="cuda"), torch.empty_like(z)
        ops.cut_loss(z, y if fmt == "float" else None, label_bits=bits if fmt == "bits" else None, loss_kind=lk, metric=metric,
                     tau=tau, probs_out=probs, grad=grad, loss_per_list=per, loss_out=loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
        out[fmt] = (grad, per, loss, probs)
    grad, per, loss, probs = out["bits"]
    ref_loss = float(g[key + "/loss"])
    assert abs(loss.item() - ref_loss) <= 2e-5 * max(1.0, abs(ref_loss)), (loss.item(), ref_loss)
    ref = g[key + "/dz"].reshape(B, L)
    err = np.abs(grad.cpu().numpy() - ref).max()
    assert err <= 2e-4 * max(np.abs(ref).max(), 1e-6), (err, np.abs(ref).max())
    if metric == "f1":     # DCG: the sign is applied by a select instead of a multiply -- same values, checked to 1 ulp below
        for a, b in zip(out["bits"], out["float"]):
            assert torch.equal(a, b)
    else:
        for a, b in zip(out["bits"], out["float"]):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-9)


def test_cut_loss_odd_length_and_probability_input_keep_the_scalar_kernel():
    """The packed kernel needs even L and logits: L = 301 and input_kind = 1 go through cut_loss_kernel and agree with the
    oracle's float64 loss; rlt_cut_loss_bits refuses an odd length instead of falling back."""
    from rlt_b200 import ops
    torch.manual_seed(4)
    B, L = 7, 301
    z = torch.randn(B, L, device="cuda")
    y = (torch.rand(B, L, device="cuda") < 0.15).float()
    p64 = torch.softmax(z.double().cpu(), dim=1)
    y64 = y.cpu().double()
    for kind, (lk, tau) in KINDS.items():
        for metric in ("f1", "dcg"):
            per, loss = torch.empty(B, device="cuda"), torch.empty((), device="cuda")
            ops.cut_loss(z, y, loss_kind=lk, metric=metric, tau=tau, loss_per_list=per, loss_out=loss, loss_scale=1.0 / B)
            ref = (O.choopy_loss(p64, y64, metric) if lk == "choopy" else
                   O.attncut_loss(p64, y64, metric, tau) if lk == "raml" else O.div_loss(p64, y64, metric, tau, lk))
            assert abs(loss.item() - float(ref)) <= 2e-5 * max(1.0, abs(float(ref))), (kind, metric, loss.item(), float(ref))
    # ... and even lengths of every other size class of the packed kernel (NP = 8, 16, 1, the 64 / 66 boundary, and lengths that leave whole slots of NP = 5 empty),
    # floats and bit masks
    for L2 in (500, 1000, 64, 66, 100, 200, 258):
        _packed_lengths(B, L2)
    with pytest.raises(RuntimeError):
        ops.cut_loss(z, None, label_bits=ops.pack_labels(y), loss_kind="js", metric="f1", tau=0.85,
                     loss_per_list=torch.empty(B, device="cuda"))


def _packed_lengths(B, L2):
    """Loss AND gradient of the packed kernel against float64 autograd through the oracle's criterion."""
    from rlt_b200 import ops
    torch.manual_seed(L2)
    z2 = torch.randn(B, L2, device="cuda")
    y2 = (torch.rand(B, L2, device="cuda") < 0.1).float()
    y64 = y2.cpu().double()
    for kind, (lk, tau) in KINDS.items():
        for metric in ("f1", "dcg"):
            z64 = z2.double().cpu().requires_grad_(True)
            p2 = torch.softmax(z64, dim=1)
            ref = (O.choopy_loss(p2, y64, metric) if lk == "choopy" else
                   O.attncut_loss(p2, y64, metric, tau) if lk == "raml" else O.div_loss(p2, y64, metric, tau, lk))
            ref.backward()
            gref = z64.grad
            for bits in (None, ops.pack_labels(y2)):
                per, loss = torch.empty(B, device="cuda"), torch.empty((), device="cuda")
                grad, probs = torch.empty_like(z2), torch.empty_like(z2)
                ops.cut_loss(z2, y2 if bits is None else None, label_bits=bits, loss_kind=lk, metric=metric, tau=tau,
                             probs_out=probs, grad=grad, loss_per_list=per, loss_out=loss, loss_scale=1.0 / B, grad_scale=1.0 / B)
                assert abs(loss.item() - float(ref)) <= 2e-5 * max(1.0, abs(float(ref))), (kind, metric, L2, loss.item(), float(ref))
                err = (grad.double().cpu() - gref).abs().max().item()
                assert err <= 2e-4 * max(gref.abs().max().item(), 1e-6), (kind, metric, L2, err, gref.abs().max().item())
                assert torch.allclose(probs.double().cpu(), p2.detach(), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("L", [300, 40])
def test_reward_matrix_kernel(L):
    from rlt_b200 import ops
    g = load_golden("losses.npz")
    y = torch.from_numpy(g[f"y_{L}"]).cuda()
    r = torch.empty_like(y)
    ops.reward_matrix(y, r, "f1")
    assert np.array_equal(r.cpu().numpy(), g[f"reward_f1_{L}"])          # same op order => bit-exact
    ops.reward_matrix(y, r, "dcg")
    ref = g[f"reward_dcg_{L}"]
    assert np.abs(r.cpu().numpy() - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("L", [300, 40])
def test_eval_cut_bit_exact_vs_reference_golden(L):
    from rlt_b200 import ops
    g = load_golden("metrics.npz")
    y = torch.from_numpy(g[f"y_{L}"]).cuda()
    p = torch.from_numpy(g[f"p_{L}"]).cuda()
    ks = g[f"k_{L}"]
    k, cnt, nrel, f1, dcg = ops.eval_cut(p, y, mode=0)
    k = k.cpu().numpy()
    assert np.array_equal(k[2:], ks[2:])                                  # rows 0,1 had k overridden in the golden
    assert k[5] == 11                                                     # tie -> first maximum
    ref_f1 = np.array(O.f1_per_list(g[f"y_{L}"], k.astype(np.int64)), dtype=np.float64)
    ref_dcg = np.array(O.dcg_per_list(g[f"y_{L}"], k.astype(np.int64)), dtype=np.float64)
    assert np.array_equal(f1.cpu().numpy(), ref_f1)
    assert np.array_equal(dcg.cpu().numpy(), ref_dcg)
    assert np.array_equal(f1.cpu().numpy()[2:], g[f"f1_{L}"][2:]) and np.array_equal(dcg.cpu().numpy()[2:], g[f"dcg_{L}"][2:])
    assert np.array_equal(nrel.cpu().numpy(), g[f"y_{L}"].sum(1).astype(np.int32))


@pytest.mark.parametrize("L", [300, 40])
def test_eval_cut_from_label_bit_masks_is_bit_identical(L):
    """rlt_eval_cut_bits (labels as rlt_pack_labels words) == rlt_eval_cut == the reference golden, every output."""
    from rlt_b200 import ops
    g = load_golden("metrics.npz")
    y = torch.from_numpy(g[f"y_{L}"]).cuda()
    p = torch.from_numpy(g[f"p_{L}"]).cuda()
    a = ops.eval_cut(p, y, mode=0)
    b = ops.eval_cut(p, None, mode=0, label_bits=ops.pack_labels(y))
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    assert np.array_equal(b[3].cpu().numpy()[2:], g[f"f1_{L}"][2:]) and np.array_equal(b[4].cpu().numpy()[2:], g[f"dcg_{L}"][2:])


def test_eval_cut_every_cut_position_and_ragged_batches():
    """Every k = 1..L (each leaf shape of numpy's pairwise tree: 1, 2, 3, 4 and 8 leaves, tails of 0..7 terms) at L = 300,
    500 and 1000, odd L through the scalar kernel, and batch sizes that leave warps / CTAs partly empty."""
    from rlt_b200 import ops
    for L, B in ((300, 300), (500, 500), (1000, 1000), (301, 301), (300, 131), (40, 5), (66, 66), (100, 100), (258, 131),
                 (600, 200)):     # ... and lengths that leave whole slots of a size class empty
        torch.manual_seed(L + B)
        y = (torch.rand(B, L, device="cuda") < 0.2).float()
        p = torch.rand(B, L, device="cuda") * 0.5
        rows = torch.arange(B, device="cuda")
        p[rows, rows % L] = 1.0                              # list b cuts at k = b % L + 1
        k, cnt, nrel, f1, dcg = ops.eval_cut(p, y, mode=0)
        kk = k.cpu().numpy().astype(np.int64)
        assert np.array_equal(kk, np.arange(B) % L + 1)
        yn = y.cpu().numpy()
        assert np.array_equal(dcg.cpu().numpy(), np.array(O.dcg_per_list(yn, kk)))
        assert np.array_equal(f1.cpu().numpy(), np.array(O.f1_per_list(yn, kk), dtype=np.float64))
        assert np.array_equal(cnt.cpu().numpy(), np.array([int(yn[i, :kk[i]].sum()) for i in range(B)], dtype=np.int32))
        if L % 2 == 0:
            for u, v in zip((k, cnt, nrel, f1, dcg), ops.eval_cut(p, None, mode=0, label_bits=ops.pack_labels(y))):
                assert torch.equal(u, v)


@pytest.mark.parametrize("L", [300, 40])
def test_metric_api_bit_exact(L):
    """Drop-in Metric.f1 / Metric.dcg (host numpy in, float out) equal the reference's numbers exactly."""
    from utils.metrics import Metric
    g = load_golden("metrics.npz")
    y, ks = g[f"y_{L}"], g[f"k_{L}"]
    assert Metric.f1(y, ks) == float(g[f"f1_mean_{L}"])
    assert Metric.dcg(y, ks) == float(g[f"dcg_mean_{L}"])
    all_k = np.arange(1, L + 1)
    got = np.array([Metric.dcg(np.repeat(y[2:3], L, 0), all_k)])
    assert got[0] == np.mean(g[f"dcg_allk_{L}"])
    f1s = Metric.f1(np.repeat(y[2:3], L, 0), all_k)
    assert f1s == O.metric_f1(np.repeat(y[2:3], L, 0), all_k)
    # mixed Python-int / numpy-int cut positions (BiCut, run.py:132-136) and the reference's known answer
    mixed = [L, np.int64(3), L, np.int64(1)]
    assert Metric.f1(y[:4], mixed) == O.metric_f1(y[:4], mixed)
    x = np.array([[1, 0, 1], [0, 0, 1], [1, 0, 0]])
    assert Metric.f1(x, np.array([1, 2, 1])) == 0.5555555555555555
    assert Metric.dcg(x, np.array([1, 2, 1])) == 0.1230234154761809


def test_eval_cut_long_lists_and_bicut_rule():
    from rlt_b200 import ops
    from rlt_b200.data import synthetic_lists
    for L in (500, 1000):
        _, y = synthetic_lists(300, L, 1, seed=L, device="cuda")
        p = torch.rand(300, L, device="cuda")
        k, cnt, nrel, f1, dcg = ops.eval_cut(p, y, mode=0)
        kk = k.cpu().numpy().astype(np.int64)
        assert np.array_equal(kk, O.cut_positions(p.cpu().numpy()))
        assert np.array_equal(dcg.cpu().numpy(), np.array(O.dcg_per_list(y.cpu().numpy(), kk)))
        assert np.array_equal(f1.cpu().numpy(), np.array(O.f1_per_list(y.cpu().numpy(), kk), dtype=np.float64))
    # BiCut rule: first position whose argmax class is 0, Python-int L when there is none
    L = 300
    _, y = synthetic_lists(64, L, 1, seed=3, device="cuda")
    out = torch.softmax(torch.randn(64, L, 2, device="cuda") + torch.tensor([0.0, 2.5], device="cuda"), dim=2).contiguous()
    out[0, :, 1] = 0.9; out[0, :, 0] = 0.1        # never truncates -> k = L (Python int)
    out[1, 4, :] = 0.5                             # tie -> class 0 -> k = 5
    ks_ref = O.bicut_cut_positions(out.cpu().numpy())
    k, cnt, nrel, f1, dcg = ops.eval_cut(out, y, mode=1)
    assert [int(v) for v in k.cpu().numpy()] == [int(v) for v in ks_ref]
    ref_f1 = O.f1_per_list(y.cpu().numpy(), ks_ref)
    assert np.array_equal(f1.cpu().numpy(), np.array([float(v) for v in ref_f1]))
    assert np.array_equal(dcg.cpu().numpy(), np.array(O.dcg_per_list(y.cpu().numpy(), ks_ref)))


@pytest.mark.parametrize("L", [300, 40])
def test_aux_heads_and_bicut_loss_kernels(L):
    from rlt_b200 import ops
    g = load_golden("losses.npz")
    y = torch.from_numpy(g[f"y_{L}"]).cuda()
    B = y.shape[0]
    for tag in ("active", "inactive"):
        s = torch.from_numpy(g[f"rerank_{tag}_{L}/s"]).cuda().reshape(B, L).contiguous()
        dzr = torch.empty_like(s)
        lg = torch.empty(1, device="cuda"); st = torch.zeros(1, dtype=torch.int32, device="cuda")
        loss = torch.empty((), device="cuda")
        ops.aux_heads_loss(None, s, y, n_groups=1, group_size=B, seq_len=L, class_weight=0.0, rerank_weight=1.0, dzr=dzr,
                           loss_group=lg, status=st, loss_out=loss)
        assert abs(loss.item() - float(g[f"rerank_{tag}_{L}/loss"])) <= 1e-6
        assert np.abs(dzr.cpu().numpy() - g[f"rerank_{tag}_{L}/ds"].reshape(B, L)).max() <= 1e-8
        assert st.item() == (2 if tag == "active" else 0)        # bit 1: the hinge is active
    # class head: BCE on sigmoid(z) vs torch, both input conventions
    zc = torch.randn(B, L, device="cuda") * 3
    ref = torch.nn.functional.binary_cross_entropy(torch.sigmoid(zc.double()), y.double()).item()
    dzc = torch.empty_like(zc); lg = torch.empty(1, device="cuda"); loss = torch.empty((), device="cuda")
    ops.aux_heads_loss(zc, None, y, n_groups=1, group_size=B, seq_len=L, class_weight=0.5, rerank_weight=0.5, dzc=dzc,
                       loss_group=lg, loss_out=loss)
    assert abs(loss.item() - 0.5 * ref) <= 1e-5 * ref
    ref_g = 0.5 * (torch.sigmoid(zc.double()) - y.double()) / (B * L)
    assert (dzc.double() - ref_g).abs().max().item() <= 1e-6 * ref_g.abs().max().item()
    # degenerate group -> status flag (the reference raises)
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.aux_heads_loss(None, zc, torch.zeros_like(y), n_groups=1, group_size=B, seq_len=L, dzr=dzc, loss_group=lg, status=st)
    assert st.item() == 1
    # BiCut loss
    u = torch.from_numpy(g[f"bicut_{L}/u"]).cuda().contiguous()
    for metric in ("f1", "nci"):
        o = torch.softmax(u, dim=2).contiguous()
        grad = torch.empty_like(o); per = torch.empty(B, device="cuda"); loss = torch.empty((), device="cuda")
        ops.bicut_loss(o, y, input_kind=1, metric_nci=(metric == "nci"), grad=grad, loss_per_list=per, loss_out=loss,
                       grad_scale=1.0 / B, loss_scale=1.0 / B)
        ref = float(g[f"bicut_{metric}_{L}/loss"])
        assert abs(loss.item() - ref) <= 1e-5 * max(1.0, abs(ref)), metric
        rd = g[f"bicut_{metric}_{L}/do"]
        assert np.abs(grad.cpu().numpy() - rd).max() <= 1e-5 * np.abs(rd).max()
        # logits mode: gradient through the 2-way softmax vs autograd of the oracle
        uu = u.detach().cpu().clone().requires_grad_(True)
        O.bicut_loss(torch.softmax(uu, dim=2), y.cpu(), metric=metric).backward()
        ops.bicut_loss(u, y, input_kind=0, metric_nci=(metric == "nci"), grad=grad, loss_per_list=per, loss_out=loss,
                       grad_scale=1.0 / B, loss_scale=1.0 / B)
        assert np.abs(grad.cpu().numpy() - uu.grad.numpy()).max() <= 2e-5 * np.abs(uu.grad.numpy()).max()
