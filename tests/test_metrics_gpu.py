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
