"""Pins the CPU oracle (oracle/*.py) to golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import adnssd_oracle as AO
from oracle import block_oracle as BO
from oracle import metrics_oracle as MO
from oracle import wtconv_oracle as WO


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    params = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    return z, params, grads


@pytest.mark.parametrize("name", sorted(cases.MIXER_CASES))
def test_mixer_oracle_matches_reference_fp64(golden_dir, name):
    D, P, N, B, g = cases.MIXER_CASES[name]
    z, params, grads = load(golden_dir, name)
    p = {k: v.double() for k, v in params.items()}
    u, dout = cases.mixer_inputs(name, torch.float32)
    if "u" in z.files:  # inputs stored explicitly for the small cases: the PCG64 rebuild must agree bit for bit
        assert np.array_equal(z["u"], u.numpy()) and np.array_equal(z["dout"], dout.numpy())
    u, dout = u.double(), dout.double()
    s = cases.SUBSAMPLE_STRIDE if g >= 64 else 1
    out = AO.mixer_forward(p, u, g, g, P, N)
    du, og = AO.mixer_backward(p, u, g, g, P, N, dout)
    assert rel(out[:, ::s], z["out"]) < 1e-12
    assert rel(du[:, ::s], z["du"]) < 1e-12
    assert set(grads) == set(AO.PARAM_NAMES) - set(AO.UNUSED_PARAMS)
    for k, ref in grads.items():
        assert rel(og[k].reshape(ref.shape), ref) < 1e-11, k


@pytest.mark.parametrize("name", sorted(cases.MIXER_BF16_CASES))
def test_mixer_oracle_matches_reference_model_scale(golden_dir, name):
    D, P, N, B, g, _ = cases.MIXER_BF16_CASES[name]
    z, params, grads = load(golden_dir, name)          # stored as float32
    p = {k: v.double() for k, v in params.items()}
    u, dout = cases.mixer_bf16_inputs(name, torch.float32)
    out = AO.mixer_forward(p, u.double(), g, g, P, N)
    du, og = AO.mixer_backward(p, u.double(), g, g, P, N, dout.double())
    assert rel(out, z["out"]) < 1e-6 and rel(du, z["du"]) < 1e-6
    for k, ref in grads.items():
        assert rel(og[k].reshape(ref.shape), ref) < 1e-6, k


@pytest.mark.parametrize("name", sorted(cases.MIXER_CORR_CASES))
def test_mixer_oracle_matches_reference_headline_shapes(golden_dir, name):
    """128 x 128 goldens of the unmodified reference with a correlated upstream gradient (make_golden.mixer_corr_case)."""
    D, P, N, B, g, _ = cases.MIXER_CORR_CASES[name]
    z, params, grads = load(golden_dir, name)
    p = {k: v.double() for k, v in params.items()}
    u = cases.mixer_corr_u(name, torch.float32).double()
    out = AO.mixer_forward(p, u, g, g, P, N)
    dout = cases.mixer_corr_dout(name, out).double()
    du, og = AO.mixer_backward(p, u, g, g, P, N, dout)
    s = cases.SUBSAMPLE_STRIDE
    assert rel(out[:, ::s], z["out"]) < 1e-12 and rel(du[:, ::s], z["du"]) < 1e-11
    assert set(grads) == set(AO.PARAM_NAMES) - set(AO.UNUSED_PARAMS)
    for k, ref in grads.items():
        assert rel(og[k].reshape(ref.shape), ref) < 1e-10, k


def test_mixer_oracle_explicit_backward_equals_autograd():
    D, P, N, g = 16, 4, 8, 5
    p = AO.init_params(D, P, N, seed=3, perturb=0.3, dtype=torch.float64)
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    u = cases.rng_normal(5, (2, g * g, D)).requires_grad_(True)
    dout = cases.rng_normal(6, (2, g * g, D))
    AO.mixer_forward(q, u, g, g, P, N).backward(dout)
    du, og = AO.mixer_backward(p, u.detach(), g, g, P, N, dout)
    assert rel(du, u.grad) < 1e-12
    for k in AO.PARAM_NAMES:
        if k in AO.UNUSED_PARAMS:
            assert q[k].grad is None
        else:
            assert rel(og[k].reshape(q[k].shape), q[k].grad) < 1e-11, k


@pytest.mark.parametrize("name", sorted(cases.WTCONV_CASES))
def test_wtconv_oracle_matches_reference_fp64(golden_dir, name):
    C, k, L, B, H, W, bias = cases.WTCONV_CASES[name]
    z, params, grads = load(golden_dir, name)
    p = {n: v.double() for n, v in params.items()}
    x, dy = cases.wtconv_inputs(name, torch.float32)
    assert np.array_equal(z["x"], x.numpy()) and np.array_equal(z["dy"], dy.numpy())
    out, dx, og = WO.wtconv_forward_backward(p, x.double(), L, dy.double())
    assert rel(out, z["out"]) < 1e-12
    assert rel(dx, z["dx"]) < 1e-12
    assert set(grads) == set(og)
    for n, ref in grads.items():
        assert rel(og[n], ref) < 1e-11, n


def test_haar_filters_match_reference_parameters(golden_dir):
    z, params, _ = load(golden_dir, "wtconv_c5_k5_l3_64")
    wt, iwt = WO.haar_filters(5)
    assert torch.allclose(params["wt_filter"], wt, atol=1e-7) and torch.allclose(params["iwt_filter"], iwt, atol=1e-7)
    x = cases.rng_normal(1, (1, 3, 8, 10))
    assert rel(WO.haar_idwt(WO.haar_dwt(x)), x) < 1e-14  # orthonormal round trip


def test_metric_counts_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, cases.METRIC_CASE[0] + ".npz"))
    obs, sim = cases.metric_inputs()
    table = MO.counts(obs, sim)
    assert np.array_equal(table, z["table"])
    assert int(MO.float2int(obs).astype(np.int64).sum()) == int(z["obs_int_checksum"])
    assert (table.sum(1) == obs.size).all()
    sc = MO.scores(table)
    assert np.allclose(sc["CSI"], table[:, 0] / (table[:, 0] + table[:, 1] + table[:, 2]))
    # CSI and HSS are symmetric under the FP<->FN swap the reference's call order introduces
    sw = MO.scores(table[:, [0, 2, 1, 3]])
    assert np.allclose(sw["CSI"], sc["CSI"]) and np.allclose(sw["HSS"], sc["HSS"])


@pytest.mark.parametrize("name", sorted(cases.BLOCK_CASES))
def test_block_oracle_matches_reference_fp64(golden_dir, name):
    """oracle/block_oracle.py against the unmodified reference Block (goldens made by tests/golden/make_golden.py)."""
    dim, out_dim, B, g, skip = cases.BLOCK_CASES[name]
    z, params, grads = load(golden_dir, name)
    p = {k: v.double().requires_grad_(True) for k, v in params.items()}
    x, res, feat = (None if t is None else t.requires_grad_(True) for t in cases.block_inputs(name, torch.float64))
    out = BO.block_forward(p, x, g, g, 4, 16, residual=res, features=feat)
    assert rel(out.detach(), z["out"]) < 1e-12
    out.backward(cases.block_dout(name, torch.from_numpy(z["out"])).double())
    assert rel(x.grad, z["dx"]) < 1e-11
    if skip:
        assert rel(res.grad, z["dresidual"]) < 1e-11 and rel(feat.grad, z["dfeatures"]) < 1e-11
    live = {k for k, v in p.items() if v.grad is not None and float(v.grad.abs().max()) > 0}
    assert live == set(grads), live ^ set(grads)
    for k, ref in grads.items():
        assert rel(p[k].grad, ref) < 1e-10, k


def test_evaluator_oracle_matches_reference_evaluator():
    """oracle.metrics_oracle.evaluator_done against the reference's own SimplifiedEvaluator.evaluate + done
    (datasets/Shanghai_metrics.py:49-103,218-290; LPIPS stubbed out: it needs a pretrained AlexNet)."""
    import importlib.util
    import sys
    import types
    import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference sources not available")
    sys.modules.setdefault("lpips", types.ModuleType("lpips"))
    spec = importlib.util.spec_from_file_location("ref_shanghai_metrics", os.path.join(ref_loader.reference_root(), "datasets", "Shanghai_metrics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class Ev(mod.SimplifiedEvaluator):
        def __init__(self, seq_len, value_scale, thresholds):
            self.lpips_fn = None      # the constructor proper (:15-43) loads lpips.LPIPS(net='alex')
            self.metrics, self.thresholds, self.seq_len, self.value_scale = {}, thresholds, seq_len, value_scale
            self.TP, self.TN, self.FP, self.FN = [], [], [], []
            mod.SimplifiedEvaluator.reset(self)

        def _cal_batch_lpips(self, preds, trues):
            return [[0.0] * preds.shape[1]] * preds.shape[0]

    ev = Ev(6, 90, [20, 30, 35, 40])
    rng = np.random.default_rng(5)
    batches = [(rng.random((2, 6, 24, 24), dtype=np.float32) * 1.2 - 0.1, rng.random((2, 6, 24, 24), dtype=np.float32) * 1.2 - 0.1) for _ in range(2)]
    for tb, pb in batches:
        ev.evaluate(tb, pb)
    ref = ev.done()
    got = MO.evaluator_done(batches)
    for thr in (20, 30, 35, 40):
        for k in ("TP", "TN", "FP", "FN"):
            assert got["threshold_metrics"][thr][k] == ref["threshold_metrics"][thr][k], (thr, k)
        for k in ("CSI", "POD", "HSS"):
            assert abs(got["threshold_metrics"][thr][k] - ref["threshold_metrics"][thr][k]) < 1e-12
    assert abs(got["FAR"] - ref["FAR"]) < 1e-12 and abs(got["RMSE"] - ref["RMSE"]) < 1e-5 * ref["RMSE"]


@pytest.mark.parametrize("cfg", [(32, 8, 4, 2, 25), (48, 6, 8, 1, 16)], ids=lambda c: "dim%d_h%d_dh%d_B%d_L%d" % c)
def test_attention_oracle_matches_reference(cfg):
    """oracle/attention_oracle.py against the unmodified reference StandardAttention (models/ADNssd.py:26-47), fp64."""
    import ref_loader
    from oracle import attention_oracle as AT
    if not ref_loader.reference_available():
        pytest.skip("reference sources not available")
    dim, heads, dh, B, L = cfg
    ref = ref_loader.load_reference()
    m = ref.ADNssd.StandardAttention(dim, heads=heads, dim_head=dh, dropout=0.).double()
    p = AT.init_params(dim, heads, dh, seed=3)
    m.load_state_dict(p, strict=True)
    x = cases.rng_normal(71, (B, L, dim)).requires_grad_(True)
    dy = cases.rng_normal(72, (B, L, dim))
    y = m(x, 5, 5)
    y.backward(dy)
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    x2 = x.detach().clone().requires_grad_(True)
    y2 = AT.attention_forward(pp, x2, heads, dh)
    y2.backward(dy)
    assert rel(y2.detach(), y.detach()) < 1e-12 and rel(x2.grad, x.grad) < 1e-12
    for k, v in m.named_parameters():
        assert rel(pp[k].grad, v.grad) < 1e-12, k


def convstage_oracle_run(name, params, x, second, dtype=torch.float64):
    """Forward of oracle/convstage_oracle.py for a CONVSTAGE_CASES entry; returns (out, leaves) with autograd leaves."""
    from oracle import convstage_oracle as CO
    kind, kw, B, g, skip = cases.CONVSTAGE_CASES[name]
    p = {k: (v.to(dtype).clone().requires_grad_(not k.endswith("wt_filter"))) for k, v in params.items()}
    x = x.to(dtype).clone().requires_grad_(kind != "PatchEmbed")
    if second is not None:
        second = second.to(dtype).clone().requires_grad_(kind == "WTLayer")
    if kind == "WTLayer":
        out = CO.wtlayer_forward(p, x, kw["wt_levels"], residual=second, features=second.detach() * 0.5 if skip else None)
    elif kind == "PatchEmbed":
        out, res = CO.patchembed_forward(p, x, kw["wt_levels"])
        assert torch.equal(res, x.view(B, g, g, -1)[..., -1])
    else:
        out = CO.outproj_forward(p, x, second, g, g)
    return out, p, x, second


@pytest.mark.parametrize("name", sorted(cases.CONVSTAGE_CASES))
def test_convstage_oracle_matches_reference_fp64(golden_dir, name):
    """oracle/convstage_oracle.py against goldens of the unmodified reference WTLayer / PatchEmbed / OutProj
    (models/model_untils.py:226-426,799-892; tests/golden/make_golden.py::convstage_case)."""
    z, params, grads = load(golden_dir, name)
    x, second = cases.convstage_inputs(name, torch.float64)
    out, p, x, second = convstage_oracle_run(name, params, x, second)
    assert rel(out.detach(), z["out"]) < 1e-12
    out.backward(cases.convstage_dout(name, torch.from_numpy(z["out"])).double())
    if "dx" in z.files:
        assert rel(x.grad, z["dx"]) < 1e-11
    if "dsecond" in z.files:
        assert rel(second.grad, z["dsecond"]) < 1e-11
    assert grads, "golden holds no parameter gradients"
    for k, g in grads.items():
        assert p[k].grad is not None, k
        if g.abs().max() < 1e-14:      # a bias in front of an InstanceNorm: its gradient is identically zero, both sides hold rounding noise
            assert p[k].grad.abs().max() < 1e-14, k
            continue
        assert rel(p[k].grad, g) < 1e-10, k
    unused = [k for k, v in p.items() if v.requires_grad and v.grad is None]
    assert sorted(unused) == sorted(k for k in p if k not in grads and not k.endswith("wt_filter")), unused
