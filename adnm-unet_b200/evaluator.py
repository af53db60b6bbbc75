"""On-device evaluation path: the reference's `SimplifiedEvaluator` (datasets/Shanghai_metrics.py:14-290)
without the per-batch `.cpu().numpy()` of validate.py:103-106 and without its Python loops over batch x frame x threshold.

Same constructor, `evaluate(true_batch, pred_batch)`, `done()` and `reset()`; `evaluate` takes CUDA tensors
(B, T, H, W) or (B, T, 1, H, W) and only enqueues one kernel (adn_eval_batch) that accumulates, on the device,
  * the integer TP / FN / FP / TN table of every threshold - bit-identical to float2int + _cal_frame (:45-47,105-114),
  * the per-lead-time mean squared error of the clipped, value_scale-d frames (:116-121) behind `RMSE` (:276).
`done()` reads the 4 x 4 table and seq_len doubles back (the only device -> host copy of an evaluation) and returns the
reference's dictionary: threshold_metrics[thr] = {TP, TN, FP, FN, CSI, POD, HSS}, FAR, RMSE.  SSIM (cv2 Gaussian filter on
the host) and LPIPS (a pretrained AlexNet) are not on the path BASELINE.json names and are reported as NaN."""
import ctypes as C

import numpy as np
import torch

from adnm_unet_b200 import _lib


class SimplifiedEvaluator:
    def __init__(self, seq_len, value_scale, thresholds=(20, 30, 35, 40), device=None):
        self.seq_len = int(seq_len)
        self.value_scale = value_scale
        self.thresholds = list(thresholds)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        _lib.load()
        self._thr = (C.c_int32 * len(self.thresholds))(*self.thresholds)
        self.table = torch.zeros(len(self.thresholds), 4, dtype=torch.int64, device=self.device)
        self.mse_t = torch.zeros(self.seq_len, dtype=torch.float64, device=self.device)
        self.total = 0
        self.TP, self.TN, self.FP, self.FN = [], [], [], []

    def float2int(self, arr):
        """datasets/Shanghai_metrics.py:45-47 on a tensor (kept for API parity; `evaluate` does this inside the kernel)."""
        return (arr.clamp(0.0, 1.0) * self.value_scale).to(torch.int32)

    def evaluate(self, true_batch, pred_batch):
        if not isinstance(pred_batch, torch.Tensor):
            pred_batch, true_batch = torch.as_tensor(pred_batch), torch.as_tensor(true_batch)
        pred_batch, true_batch = pred_batch.to(self.device), true_batch.to(self.device)
        if pred_batch.dim() == 5:
            pred_batch, true_batch = pred_batch.squeeze(2), true_batch.squeeze(2)
        if pred_batch.shape != true_batch.shape or pred_batch.dim() != 4:
            raise RuntimeError(f"evaluate: (B, T, H, W) tensors of one shape expected, got {tuple(true_batch.shape)} / {tuple(pred_batch.shape)}")
        B, T, H, W = pred_batch.shape
        if T != self.seq_len:
            raise RuntimeError(f"evaluate: seq_len {T} != {self.seq_len}")
        t = true_batch.detach().float().contiguous()
        p = pred_batch.detach().float().contiguous()
        lib = _lib.load()
        per_call = max(1, 65535 // T)
        with _lib.on_device(self.device):
            for b0 in range(0, B, per_call):
                nb = min(per_call, B - b0)
                _lib.check(lib.adn_eval_batch(_lib.ptr(t[b0:]), _lib.ptr(p[b0:]), nb, T, H * W, self._thr, len(self.thresholds),
                                              float(self.value_scale), _lib.ptr(self.table), _lib.ptr(self.mse_t),
                                              _lib.stream_ptr(self.device)), "adn_eval_batch")
        self.total += B

    def counts(self):
        """The accumulated integer table, rows = thresholds, columns TP, FN, FP, TN (device tensor)."""
        return self.table

    def done(self):
        table = self.table.cpu().numpy().astype(np.float64)
        threshold_metrics, all_far = {}, []
        sums = np.zeros(4)
        with np.errstate(divide="ignore", invalid="ignore"):
            for i, thr in enumerate(self.thresholds):
                TP, FN, FP, TN = table[i]
                sums += (TP, TN, FP, FN)
                all_far.append(FP / (TP + FP))
                threshold_metrics[thr] = {
                    "TP": TP, "TN": TN, "FP": FP, "FN": FN, "CSI": TP / (TP + FP + FN), "POD": TP / (TP + FN),
                    "HSS": (2 * (TP * TN - FP * FN)) / (FP ** 2 + FN ** 2 + 2 * TP * TN + (FP + FN) * (TP + TN))}
            n = len(self.thresholds)
            self.TP.append(sums[0] / n); self.TN.append(sums[1] / n); self.FP.append(sums[2] / n); self.FN.append(sums[3] / n)
            rmse = float(np.mean(np.sqrt(self.mse_t.cpu().numpy() / max(self.total, 1))))
        return {"threshold_metrics": threshold_metrics, "FAR": float(np.mean(all_far)), "RMSE": rmse,
                "SSIM": float("nan"), "LPIPS": float("nan")}

    def reset(self):
        self.table.zero_()
        self.mse_t.zero_()
        self.total = 0
