"""On-device restatement of the reference's threshold counts (datasets/Shanghai_metrics.py:45-47 float2int,
:105-114 _cal_frame, :240-274 done): identical integer TP/FN/FP/TN tables from one pass over the data."""
import ctypes as C

import torch

from adnm_unet_b200 import _lib

THRESHOLDS = (20, 30, 35, 40)


def threshold_counts(obs, sim, thresholds=THRESHOLDS, value_scale=90.0):
    """obs, sim: same-shape CUDA tensors -> int64 tensor (len(thresholds), 4) with columns TP, FN, FP, TN."""
    _lib.require_cuda(obs, "obs")
    _lib.require_cuda(sim, "sim")
    if obs.shape != sim.shape:
        raise RuntimeError("threshold_counts: shape mismatch")
    lib = _lib.load()
    obs, sim = obs.float().contiguous(), sim.float().contiguous()
    table = torch.empty(len(thresholds), 4, dtype=torch.int64, device=obs.device)
    thr = (C.c_int32 * len(thresholds))(*thresholds)
    with torch.cuda.device(obs.device):
        _lib.check(lib.adn_threshold_counts(_lib.ptr(obs), _lib.ptr(sim), obs.numel(), thr, len(thresholds),
                                            float(value_scale), _lib.ptr(table), _lib.stream_ptr()), "adn_threshold_counts")
    return table


def csi_hss(table):
    """CSI and HSS per threshold from a counts table (Shanghai_metrics.py:259-265), float64 on the table's device."""
    t = table.double()
    TP, FN, FP, TN = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    csi = TP / (TP + FP + FN)
    hss = (2 * (TP * TN - FP * FN)) / (FP ** 2 + FN ** 2 + 2 * TP * TN + (FP + FN) * (TP + TN))
    return csi, hss
