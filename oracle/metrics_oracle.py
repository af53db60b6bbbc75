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


def evaluator_done(batches, thresholds=THRESHOLDS, value_scale=VALUE_SCALE):
    """`SimplifiedEvaluator.evaluate` over `batches` = [(true_batch, pred_batch), ...] (declared argument order, :49) and
    then `done` (:218-290), restated without the per-frame Python loops: threshold_metrics (TP/TN/FP/FN/CSI/POD/HSS), FAR and
    RMSE = mean_t sqrt(mean_b mse[b][t]) with mse on the clipped frames times value_scale, fp32 like the reference (:116-121,276)."""
    table = np.zeros((len(thresholds), 4), dtype=np.int64)
    mse = []
    for tb, pb in batches:
        tb = np.clip(np.asarray(tb, dtype=np.float32), 0.0, 1.0)
        pb = np.clip(np.asarray(pb, dtype=np.float32), 0.0, 1.0)
        table += counts(tb, pb, thresholds, value_scale)
        d = pb * np.float32(value_scale) - tb * np.float32(value_scale)
        mse.append(np.mean(d.reshape(d.shape[0], d.shape[1], -1) ** 2, axis=2))
    mse = np.concatenate(mse, axis=0)
    sc = scores(table)
    tm = {thr: {"TP": float(table[i, 0]), "TN": float(table[i, 3]), "FP": float(table[i, 2]), "FN": float(table[i, 1]),
                "CSI": sc["CSI"][i], "POD": sc["POD"][i], "HSS": sc["HSS"][i]} for i, thr in enumerate(thresholds)}
    return {"threshold_metrics": tm, "FAR": float(np.mean(sc["FAR"])), "RMSE": float(np.mean(np.sqrt(np.mean(mse, axis=0))))}
