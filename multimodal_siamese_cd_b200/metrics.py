"""Drop-in for the metric the reference's evaluation loop uses (`utils/metrics.py:5-66`, called from
`utils/evaluation.py:12-28`): same constructor, `add_sample(y_true, y_pred)`, `precision`, `recall`,
`compute_basic_metrics()`, `compute_f1()` and the same (swapped) meaning of its FP / FN counters — but the eight
elementwise / reduction launches of `add_sample` are ONE kernel (`b200cd_confusion_counts`) accumulating exact integer
counts on the device. `add_logits(y_true, logits)` additionally folds the `torch.sigmoid` of `evaluation.py:22` in.
CUDA tensors only.
"""
from __future__ import annotations

import torch

from . import _lib


class MultiThresholdMetric(object):
    def __init__(self, threshold: torch.Tensor):
        if threshold.dim() != 1 or not 1 <= threshold.numel() <= 8:
            raise ValueError("MultiThresholdMetric (b200cd): 1..8 thresholds")
        if not threshold.is_cuda:
            raise RuntimeError("MultiThresholdMetric (b200cd) runs on CUDA tensors only; there is no CPU path")
        self._thr = threshold.detach().float().contiguous()
        self._counts = torch.zeros(self._thr.numel(), 4, dtype=torch.int64, device=threshold.device)

    def _add(self, y_true: torch.Tensor, y_pred: torch.Tensor, from_logits: bool) -> None:
        if not (y_true.is_cuda and y_pred.is_cuda):
            raise RuntimeError("MultiThresholdMetric (b200cd) runs on CUDA tensors only; there is no CPU path")
        if y_true.numel() != y_pred.numel():
            raise ValueError("y_true and y_pred differ in size")
        yt = y_true.detach().float().contiguous()
        yp = y_pred.detach().float().contiguous()
        dev = yp.device.index
        _lib.init(dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().b200cd_confusion_counts(yp.data_ptr(), yt.data_ptr(), yp.numel(), int(from_logits),
                                                           self._thr.data_ptr(), self._thr.numel(),
                                                           self._counts.data_ptr(), torch.cuda.current_stream().cuda_stream))
        for name in ("_precision", "_recall"):
            self.__dict__.pop(name, None)

    def add_sample(self, y_true: torch.Tensor, y_pred: torch.Tensor) -> None:
        """y_pred: probabilities (the reference passes torch.sigmoid(logits))."""
        self._add(y_true, y_pred, False)

    def add_logits(self, y_true: torch.Tensor, logits: torch.Tensor) -> None:
        self._add(y_true, logits, True)

    # counters under the reference's names (float tensors of shape [thresholds], utils/metrics.py:27-30)
    @property
    def TP(self):
        return self._counts[:, 0].float()

    @property
    def TN(self):
        return self._counts[:, 1].float()

    @property
    def FP(self):
        return self._counts[:, 2].float()

    @property
    def FN(self):
        return self._counts[:, 3].float()

    @property
    def precision(self):
        if not hasattr(self, "_precision"):
            self._precision = self.TP / (self.TP + self.FP).clamp(10e-05)
        return self._precision

    @property
    def recall(self):
        if not hasattr(self, "_recall"):
            self._recall = self.TP / (self.TP + self.FN).clamp(10e-05)
        return self._recall

    def compute_basic_metrics(self):
        return self.FP / (self.FP + self.TN), self.FN / (self.FN + self.TP)

    def compute_f1(self):
        denom = (self.precision + self.recall).clamp(10e-05)
        return 2 * self.precision * self.recall / denom
