"""CPU oracle for the CSI/HSS threshold counts (TEST INFRASTRUCTURE - never on the product path).

numpy restatement of datasets/Shanghai_metrics.py of the reference:
  float2int (:45-47)     clip(0,1) * value_scale -> astype(uint16)  (truncation)
  _cal_frame (:105-114)  per frame TP/FN/FP/TN with `>= threshold`
  done (:240-274)        sums over (batch, frame); CSI, POD, HSS, FAR
`evaluate(true_batch, pred_batch)` is declared in that order (:49) and called as
`evaluate(preds, gts)` by validate.py:117 / train.py:241, so FP and FN are swapped in the
reference's printed tables; `counts(obs, sim)` keeps the declared meaning (obs = first argument).
"""
import numpy as np

THRESHOLDS = (20, 30, 35, 40)
VALUE_SCALE = 90


def float2int(arr, value_scale=VALUE_SCALE):
    return (np.clip(arr, 0.0, 1.0) * value_scale).astype(np.uint16)


def counts(obs, sim, thresholds=THRESHOLDS, value_scale=VALUE_SCALE):
    """obs, sim: float arrays (N,T,H,W) -> int64 (len(thresholds), 4) columns TP, FN, FP, TN."""
    o = float2int(np.asarray(obs, dtype=np.float32), value_scale)
    s = float2int(np.asarray(sim, dtype=np.float32), value_scale)
    out = np.zeros((len(thresholds), 4), dtype=np.int64)
    for i, t in enumerate(thresholds):
        ob, sb = o >= t, s >= t
        out[i] = [(ob & sb).sum(), (ob & ~sb).sum(), (~ob & sb).sum(), (~ob & ~sb).sum()]
    return out


def scores(table):
    """CSI / POD / HSS / FAR per threshold from a counts() table (float64, NaN where 0/0 like the reference)."""
    t = np.asarray(table, dtype=np.float64)
    TP, FN, FP, TN = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        return {"CSI": TP / (TP + FP + FN), "POD": TP / (TP + FN),
                "HSS": (2 * (TP * TN - FP * FN)) / (FP ** 2 + FN ** 2 + 2 * TP * TN + (FP + FN) * (TP + TN)),
                "FAR": FP / (TP + FP)}
