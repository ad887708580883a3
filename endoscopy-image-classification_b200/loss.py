"""Drop-in for the hot-path criteria of the reference ``code/loss.py``.

Same names, argument meaning and return shapes as the reference:

* ``consistency_loss`` -- ``code/loss.py:126-168`` (FixMatch unlabeled head),
* ``ce_loss``          -- ``code/loss.py:90-124`` (hard / poly / soft branches),
* ``PolyLoss``         -- ``code/loss.py:308-364`` (hard integer targets).

Underneath, one fused sm_100a kernel per call computes the forward scalars AND
the gradient w.r.t. the logits (``b200ssl_fixmatch_head_fwd_bwd`` /
``b200ssl_labeled_ce_fwd_bwd`` in ``include/b200ssl.h``); autograd only chains
the upstream scalar.  CUDA tensors only -- there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _native as N

__all__ = ["consistency_loss", "consistency_loss_dual", "ce_loss", "PolyLoss", "fixmatch_head", "bad_label_count"]


def _as_rows(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dim() != 2:
        raise ValueError(f"{what}: expected [rows, classes], got {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


class _FixMatchHead(torch.autograd.Function):
    """loss.py:150-164 forward + F.cross_entropy backward in ONE launch."""

    @staticmethod
    def forward(ctx, logits_w, logits_s, logits_s2, p_cutoff, T, use_hard_labels, want_details):
        dev = N.require_cuda(logits_w, logits_s, logits_s2, what="consistency_loss")
        w = _as_rows(logits_w.detach(), "logits_w")
        s = _as_rows(logits_s.detach(), "logits_s")
        if w.shape != s.shape or w.dtype != s.dtype:
            raise ValueError(f"logits_w {tuple(w.shape)}/{w.dtype} vs logits_s {tuple(s.shape)}/{s.dtype}")
        s2 = None
        if logits_s2 is not None:
            s2 = _as_rows(logits_s2.detach(), "logits_s2")
            if s2.shape != s.shape or s2.dtype != s.dtype:
                raise ValueError("second strong head must match the first")
        rows, classes = w.shape
        grad_s = torch.empty_like(s)
        grad_s2 = torch.empty_like(s2) if s2 is not None else None
        out = torch.empty(4, dtype=torch.float32, device=dev)
        idx = torch.empty(rows, dtype=torch.int64, device=dev) if want_details else None
        mask = torch.empty(rows, dtype=torch.float32, device=dev) if want_details else None
        ws, ws_bytes = N.workspace(dev, rows, classes)
        N.check(N.lib().b200ssl_fixmatch_head_fwd_bwd(
            w.data_ptr(), s.data_ptr(), N.ptr(s2), grad_s.data_ptr(), N.ptr(grad_s2), rows, classes,
            N.dtype_enum(w), float(p_cutoff), 1.0 / float(T), 1 if use_hard_labels else 0,
            out.data_ptr(), N.ptr(idx), N.ptr(mask), ws, ws_bytes, N.stream_ptr(dev)), "fixmatch_head_fwd_bwd")
        ctx.save_for_backward(grad_s, grad_s2)
        ctx.has_s2 = s2 is not None
        ctx.done = False
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(*(t for t in (idx, mask) if t is not None))
        return out[0], out[1], out[2], idx, mask

    @staticmethod
    def backward(ctx, g_loss, g_mask_mean, g_loss2, g_idx, g_mask):
        if ctx.done:
            raise RuntimeError("consistency_loss: backward through the fused head twice (its stashed gradient is scaled in "
                               "place by the first backward; recompute the loss instead of retain_graph=True)")
        ctx.done = True
        grad_s, grad_s2 = ctx.saved_tensors
        dev = grad_s.device
        lib = N.lib()

        def chain(stash, g):
            if stash is None:
                return None
            if g is None:
                return torch.zeros_like(stash)
            g = g.detach().to(torch.float32).reshape(1).contiguous()
            # in place: the stash is consumed exactly once by the first backward
            N.check(lib.b200ssl_scale_inplace(stash.data_ptr(), stash.numel(), N.dtype_enum(stash),
                                              g.data_ptr(), 1.0, N.stream_ptr(dev)), "scale_inplace")
            return stash

        gs = chain(grad_s, g_loss) if ctx.needs_input_grad[1] else None
        gs2 = chain(grad_s2, g_loss2) if (ctx.has_s2 and ctx.needs_input_grad[2]) else None
        return None, gs, gs2, None, None, None, None


def fixmatch_head(logits_w, logits_s, logits_s2=None, p_cutoff=0.0, T=1.0, use_hard_labels=True):
    """Full output of the fused head: (loss, mask_mean, loss2, idx int64[rows], mask f32[rows])."""
    return _FixMatchHead.apply(logits_w, logits_s, logits_s2, p_cutoff, T, use_hard_labels, True)


def consistency_loss(logits_w, logits_s, name="ce", T=1.0, p_cutoff=0.0, use_hard_labels=True,
                     device=None, loss_fc=None, fc=None):
    """Drop-in for ``code/loss.py:126-168``.

    Returns ``(masked_loss.mean(), mask.mean())`` for ``name='ce'`` (``:164``) and a
    bare tensor for ``name='L2'`` (``:143-145``, quirk Q8).  ``T`` is ignored in
    hard-label mode exactly like the reference (quirk Q4); ``use_hard_labels=False``
    implements the evident intent of the reference's (crashing) soft branch.
    ``device`` is accepted for signature parity; tensors must already be on CUDA.
    The angular-margin branch (``loss_fc and fc``, ``:131-139``) is dead code in
    the reference and not on the hot path.
    """
    assert name in ["ce", "L2"]
    if loss_fc and fc:
        raise NotImplementedError("angular-margin consistency branch (loss.py:131-139) is outside the SSL hot path")
    if name == "L2":
        assert logits_w.size() == logits_s.size()
        return torch.nn.functional.mse_loss(logits_s, logits_w.detach(), reduction="mean")
    loss, mask_mean, _, _, _ = _FixMatchHead.apply(logits_w, logits_s, None, p_cutoff,
                                                   1.0 if use_hard_labels else T, use_hard_labels, False)
    return loss, mask_mean


def consistency_loss_dual(logits_w, logits_s_a, logits_s_b, T=1.0, p_cutoff=0.0, use_hard_labels=True
                          ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The two ``consistency_loss`` calls of ``code/semiformer.py:129-130`` (same weak
    logits, conv-head and trans-head strong logits) in one launch.
    Returns ``(lu_conv, lu_trans, mask_mean)``."""
    la, mask_mean, lb, _, _ = _FixMatchHead.apply(logits_w, logits_s_a, logits_s_b, p_cutoff,
                                                  1.0 if use_hard_labels else T, use_hard_labels, False)
    return la, lb, mask_mean


class _LabeledCE(torch.autograd.Function):
    """loss.py:103-119 / 308-364 forward + backward in one launch."""

    @staticmethod
    def forward(ctx, logits, targets, class_weights, poly, epsilon):
        dev = N.require_cuda(logits, targets, class_weights, what="ce_loss")
        x = _as_rows(logits.detach(), "logits")
        y = targets.detach()
        if y.dim() == 2 and y.shape[1] == 1:
            y = y.squeeze(1)
        if y.dim() != 1 or y.shape[0] != x.shape[0]:
            raise ValueError(f"targets {tuple(targets.shape)} do not match logits {tuple(x.shape)}")
        y = y.to(torch.int64).contiguous()
        cw = None
        if class_weights is not None:
            cw = class_weights.detach().to(torch.float32).contiguous()
            if cw.numel() != x.shape[1]:
                raise ValueError("class_weights must have one entry per class")
        rows, classes = x.shape
        grad = torch.empty_like(x)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        ws, ws_bytes = N.workspace(dev, rows, classes)
        N.check(N.lib().b200ssl_labeled_ce_fwd_bwd(
            x.data_ptr(), y.data_ptr(), N.ptr(cw), grad.data_ptr(), rows, classes, N.dtype_enum(x),
            1 if poly else 0, float(epsilon), out.data_ptr(), ws, ws_bytes, N.stream_ptr(dev)), "labeled_ce_fwd_bwd")
        ctx.save_for_backward(grad)
        ctx.done = False
        return out[0]

    @staticmethod
    def backward(ctx, g):
        if ctx.done:
            raise RuntimeError("ce_loss: backward through the fused criterion twice (its stashed gradient is scaled in "
                               "place by the first backward; recompute the loss instead of retain_graph=True)")
        ctx.done = True
        (grad,) = ctx.saved_tensors
        g = g.detach().to(torch.float32).reshape(1).contiguous()
        N.check(N.lib().b200ssl_scale_inplace(grad.data_ptr(), grad.numel(), N.dtype_enum(grad), g.data_ptr(), 1.0,
                                              N.stream_ptr(grad.device)), "scale_inplace")
        return grad, None, None, None, None


class _RowCE(torch.autograd.Function):
    """Un-reduced criterion (loss.py:118-124; PolyLoss ``reduction='none'``): per-row losses, with the per-row
    gradient stashed by the same launch and chained out of place in backward (safe under ``retain_graph``)."""

    @staticmethod
    def forward(ctx, logits, targets, class_weights, poly, epsilon, soft):
        dev = N.require_cuda(logits, targets, class_weights, what="ce_loss")
        x = _as_rows(logits.detach(), "logits")
        rows, classes = x.shape
        y = t = cw = None
        if soft:
            if targets.shape != x.shape:
                raise ValueError(f"soft targets {tuple(targets.shape)} must match logits {tuple(x.shape)}")
            t = targets.detach().to(torch.float32).contiguous()
        else:
            y = targets.detach()
            if y.dim() == 2 and y.shape[1] == 1:
                y = y.squeeze(1)
            if y.dim() != 1 or y.shape[0] != rows:
                raise ValueError(f"targets {tuple(targets.shape)} do not match logits {tuple(x.shape)}")
            y = y.to(torch.int64).contiguous()
            if class_weights is not None:
                cw = class_weights.detach().to(torch.float32).contiguous()
                if cw.numel() != classes:
                    raise ValueError("class_weights must have one entry per class")
        grad = torch.empty_like(x)
        loss_rows = torch.empty(rows, dtype=torch.float32, device=dev)
        ws, ws_bytes = N.workspace(dev, rows, classes)
        N.check(N.lib().b200ssl_ce_rows_fwd_bwd(x.data_ptr(), N.ptr(y), N.ptr(t), N.ptr(cw), grad.data_ptr(),
                                                loss_rows.data_ptr(), rows, classes, N.dtype_enum(x), 1 if poly else 0,
                                                float(epsilon), ws, ws_bytes, N.stream_ptr(dev)), "ce_rows_fwd_bwd")
        ctx.save_for_backward(grad)
        return loss_rows if x.dtype == torch.float32 else loss_rows.to(x.dtype)

    @staticmethod
    def backward(ctx, g_rows):
        (grad,) = ctx.saved_tensors
        g = g_rows.detach().to(torch.float32).contiguous()
        out = torch.empty_like(grad)
        N.check(N.lib().b200ssl_scale_rows(grad.data_ptr(), out.data_ptr(), grad.shape[0], grad.shape[1], N.dtype_enum(grad),
                                           g.data_ptr(), N.stream_ptr(grad.device)), "scale_rows")
        return out, None, None, None, None, None


def bad_label_count(device=None, reset: bool = True) -> int:
    """Labels outside ``[0, classes)`` (other than the ignore index -100) seen by the labeled kernels on ``device``
    since the last reset.  Such rows are dropped instead of read out of bounds; this call synchronises, so use it as a
    debugging check (``assert bad_label_count() == 0`` once per epoch), not per step."""
    import ctypes
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ws, _ = N.workspace(dev, 1, 2)
    n = ctypes.c_uint32()
    N.check(N.lib().b200ssl_bad_label_count(ws, ctypes.byref(n), 1 if reset else 0), "bad_label_count")
    return int(n.value)


def ce_loss(logits, targets, class_weights=None, use_hard_labels=True, reduction="none", type_loss="none",
            cls_num_list=None):
    """Drop-in for ``code/loss.py:90-124``: hard labels with ``type_loss`` ``'none'`` (``:118-119``,
    ``F.cross_entropy`` with ``weight`` and ``reduction`` in ``'none' | 'mean' | 'sum'``) or ``'poly'`` (``:103-114``,
    PolyLoss epsilon = 2 with the same reductions), and soft targets (``:120-124``, always un-reduced like the
    reference).  ``reduction='mean'`` is one fused launch; the un-reduced forms return per-row losses from one launch.
    The focal / LDAM variants (``:98-117``) are not used by the SSL trainers and raise."""
    if not use_hard_labels:
        assert logits.shape == targets.shape
        return _RowCE.apply(logits, targets, None, False, 0.0, True)
    if type_loss not in ("none", "poly"):
        raise NotImplementedError(f"ce_loss(type_loss={type_loss!r}) is outside the SSL hot path (supported: 'none', 'poly')")
    if reduction == "mean":
        return _LabeledCE.apply(logits, targets, class_weights, type_loss == "poly", 2.0)
    if reduction not in ("none", "sum"):
        raise ValueError(f"reduction {reduction!r}")
    rows = _RowCE.apply(logits, targets, class_weights, type_loss == "poly", 2.0, False)
    return rows.sum() if reduction == "sum" else rows


class PolyLoss(nn.Module):
    """``code/loss.py:308-364`` for hard integer targets, ``reduction='mean'``."""

    def __init__(self, softmax: bool = True, ce_weight: Optional[torch.Tensor] = None, reduction: str = "mean",
                 epsilon: float = 1.0) -> None:
        super().__init__()
        if not softmax:
            raise NotImplementedError("fused PolyLoss supports softmax=True (logits in)")
        if reduction not in ("mean", "sum", "none"):
            raise ValueError(f'Unsupported reduction: {reduction}, available options are ["mean", "sum", "none"].')
        self.ce_weight = ce_weight
        self.epsilon = epsilon
        self.reduction = reduction

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.reduction == "mean":
            return _LabeledCE.apply(input, target, self.ce_weight, True, self.epsilon)
        rows = _RowCE.apply(input, target, self.ce_weight, True, self.epsilon, False)
        return rows.sum() if self.reduction == "sum" else rows
