"""Threshold counts on the GPU are bit-identical to the reference's numpy path (golden) and to the oracle."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


def test_counts_match_reference_golden(golden_dir):
    import adnm_unet_b200 as A
    z = np.load(os.path.join(golden_dir, cases.METRIC_CASE[0] + ".npz"))
    obs, sim = cases.metric_inputs()
    table = A.threshold_counts(torch.from_numpy(obs).cuda(), torch.from_numpy(sim).cuda())
    assert np.array_equal(table.cpu().numpy(), z["table"])
    csi, hss = A.csi_hss(table)
    sc = MO.scores(z["table"])
    assert np.allclose(csi.cpu().numpy(), sc["CSI"]) and np.allclose(hss.cpu().numpy(), sc["HSS"])


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 20 * 256 * 256 + 3])
def test_counts_ragged_sizes_and_boundaries(n):
    import adnm_unet_b200 as A
    rng = np.random.default_rng(n)
    obs = (rng.random(n, dtype=np.float32) * 1.4 - 0.2)
    sim = (rng.integers(0, 91, n).astype(np.float32) / np.float32(90))  # exact k/90 boundaries
    o = torch.from_numpy(obs).cuda() if n else torch.zeros(0, device="cuda")
    s = torch.from_numpy(sim).cuda() if n else torch.zeros(0, device="cuda")
    table = A.threshold_counts(o, s).cpu().numpy()
    assert np.array_equal(table, MO.counts(obs, sim))
    assert (table.sum(1) == n).all()


@pytest.mark.parametrize("shape", [(2, 6, 24, 24), (3, 20, 1, 33, 47), (64, 20, 16, 16)], ids=lambda s: "x".join(map(str, s)))
def test_device_evaluator_matches_oracle(shape):
    """evaluate() x 2 batches + done() on the device against the numpy restatement of
    SimplifiedEvaluator.evaluate / done (pinned to the reference class in tests/test_oracle_vs_golden.py):
    integer tables identical, CSI / POD / HSS / FAR to 1e-12, RMSE to 1e-6."""
    import numpy as np
    from adnm_unet_b200.evaluator import SimplifiedEvaluator
    from oracle import metrics_oracle as MO
    rng = np.random.default_rng(17)
    batches = [(rng.random(shape, dtype=np.float32) * 1.2 - 0.1, rng.random(shape, dtype=np.float32) * 1.2 - 0.1) for _ in range(2)]
    for tb, pb in batches:      # exact k/90 boundaries
        tb.reshape(-1)[::97] = np.float32(30) / np.float32(90)
        pb.reshape(-1)[::89] = np.float32(35) / np.float32(90)
    ev = SimplifiedEvaluator(seq_len=shape[1], value_scale=90, thresholds=[20, 30, 35, 40])
    for tb, pb in batches:
        ev.evaluate(torch.from_numpy(tb).cuda(), torch.from_numpy(pb).cuda())
    got = ev.done()
    sq = [(tb.reshape(tb.shape[0], tb.shape[1], *tb.shape[-2:]), pb.reshape(pb.shape[0], pb.shape[1], *pb.shape[-2:])) for tb, pb in batches]
    ref = MO.evaluator_done(sq)
    for thr in (20, 30, 35, 40):
        for k in ("TP", "TN", "FP", "FN"):
            assert got["threshold_metrics"][thr][k] == ref["threshold_metrics"][thr][k], (thr, k)
        for k in ("CSI", "POD", "HSS"):
            assert abs(got["threshold_metrics"][thr][k] - ref["threshold_metrics"][thr][k]) < 1e-12
    assert abs(got["FAR"] - ref["FAR"]) < 1e-12
    assert abs(got["RMSE"] - ref["RMSE"]) < 1e-6 * ref["RMSE"]
    ev.reset()
    assert int(ev.counts().sum()) == 0 and ev.total == 0
