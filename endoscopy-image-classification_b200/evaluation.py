"""Evaluation head: the per-batch tail of ``evaluate_one``
(``code/fixmatch.py:148-168``, same in ``comatch.py`` / ``semiformer.py``) and ``utils.calculate_metrics``
(``code/utils.py:38-55``) without per-batch host round trips.

The reference synchronises twice per validation batch (``losses.item()``, ``outputs.cpu()``), keeps every
probability row on the host and lets scikit-learn recount the predictions a dozen times.  Here one launch per batch
(``b200ssl_eval_head``) computes the batch's mean cross-entropy, the arg-max of the soft-max (first index on ties, like
``np.argmax``) and adds the rows to a ``[C, C]`` confusion matrix in device memory; ``EvalAccumulator.finalize`` makes ONE
device-to-host copy (matrix + per-batch losses) and derives every number of ``calculate_metrics`` from the matrix.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as N
from .utils import AverageMeter

__all__ = ["EvalAccumulator", "metrics_from_confusion"]


def metrics_from_confusion(conf: np.ndarray) -> Dict[str, object]:
    """``utils.calculate_metrics`` (``code/utils.py:38-55``) as a function of the confusion matrix ``conf[target, pred]``:
    micro / macro precision, recall, F1 with scikit-learn's conventions (classes that occur neither in the targets nor in
    the predictions are left out of the macro averages; 0 where a denominator is 0; F1 = 2 TP / (true + predicted)) and the
    per-class sensitivity / specificity table (one-vs-rest recall of the positive / negative class)."""
    import pandas as pd
    conf = np.asarray(conf, dtype=np.int64)
    C = conf.shape[0]
    total = float(conf.sum())
    tp = np.diag(conf).astype(np.float64)
    true = conf.sum(axis=1).astype(np.float64)          # rows: targets
    pred = conf.sum(axis=0).astype(np.float64)          # columns: predictions

    def div(a, b):
        return np.divide(a, b, out=np.zeros_like(a, dtype=np.float64), where=b != 0)

    present = (true + pred) > 0
    precision, recall, f1 = div(tp, pred), div(tp, true), div(2.0 * tp, true + pred)
    micro = float(tp.sum() / total) if total else 0.0
    out: Dict[str, object] = {"micro/precision": micro, "micro/recall": micro, "micro/f1": micro}
    for name, v in (("precision", precision), ("recall", recall), ("f1", f1)):
        out[f"macro/{name}"] = float(v[present].mean()) if present.any() else 0.0
    # utils.py:42-46: sensitivity = recall of class l, specificity = recall of "not l"
    tn = total - true - pred + tp
    spec = div(tn, total - true)
    out["sen/spec"] = pd.DataFrame({"class": np.arange(C), "sensitivity": div(tp, true), "specificity": spec})
    return out


class EvalAccumulator:
    """Device-side state of one evaluation pass.

    >>> acc = EvalAccumulator(num_classes, device)
    >>> for images, targets in valid_dl:
    ...     acc.update(model(images), targets)          # one launch, no synchronisation
    >>> meter, metric = acc.finalize(config.DATA.BATCH_SIZE)
    """

    def __init__(self, num_classes: int, device, max_batches: int = 4096, keep_predictions: bool = False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("EvalAccumulator needs a CUDA device: the evaluation head is a CUDA kernel, there is no CPU path")
        N.lib()
        self.C, self.cap = int(num_classes), int(max_batches)
        # one buffer, one D2H copy: [C*C] int64 confusion counts, then `cap` fp32 batch losses (two per int64 slot)
        self._state = torch.zeros(self.C * self.C + (self.cap + 1) // 2, dtype=torch.int64, device=self.device)
        self._losses = self._state[self.C * self.C:].view(torch.float32)
        self.n = 0
        self._preds = [] if keep_predictions else None
        self._targets = [] if keep_predictions else None

    def update(self, logits: torch.Tensor, targets: torch.Tensor) -> None:
        """``ce_loss(outputs, targets, reduction='mean')`` + ``softmax`` + ``argmax`` of one batch (fixmatch.py:154-162)."""
        N.require_cuda(logits, targets, what="EvalAccumulator.update")
        if self.n >= self.cap:
            raise RuntimeError(f"more than max_batches={self.cap} validation batches")
        x = logits.detach()
        if x.dim() != 2 or x.shape[1] != self.C:
            raise ValueError(f"expected [rows, {self.C}] logits, got {tuple(x.shape)}")
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        y = targets.detach().reshape(-1).to(torch.int64).contiguous()
        if y.shape[0] != x.shape[0]:
            raise ValueError("one target per row")
        pred = torch.empty(x.shape[0], dtype=torch.int64, device=self.device) if self._preds is not None else None
        ws, wsb = N.workspace(self.device, x.shape[0], self.C)
        N.check(N.lib().b200ssl_eval_head(x.data_ptr(), y.data_ptr(), x.shape[0], self.C, N.dtype_enum(x), self._state.data_ptr(),
                                          self._losses[self.n:].data_ptr(), N.ptr(pred), ws, wsb, N.stream_ptr(self.device)),
                "eval_head")
        if pred is not None:
            self._preds.append(pred)
            self._targets.append(y)
        self.n += 1

    def finalize(self, batch_size: int) -> Tuple[AverageMeter, Dict[str, object]]:
        """One device-to-host copy; returns ``(summary_loss, metric)`` like ``evaluate_one`` (fixmatch.py:178).  The meter
        weighs every batch loss by ``DATA.BATCH_SIZE`` exactly as the reference does (:158)."""
        host = self._state.cpu()                         # the only synchronisation of the pass
        self.confusion = host[: self.C * self.C].view(self.C, self.C).numpy().copy()
        losses = host[self.C * self.C:].view(torch.float32)[: self.n].double().tolist()
        meter = AverageMeter()
        for v in losses:
            meter.update(v, batch_size)
        return meter, metrics_from_confusion(self.confusion)

    def predictions(self) -> Tuple[np.ndarray, np.ndarray]:
        """``(pred, target)`` of every row seen (``keep_predictions=True``), for ``classification_report``."""
        if self._preds is None:
            raise RuntimeError("built without keep_predictions")
        return torch.cat(self._preds).cpu().numpy(), torch.cat(self._targets).cpu().numpy()
