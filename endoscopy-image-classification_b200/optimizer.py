"""Host-side optimizer factory with the reference's contract (``code/optimizer.py``):
``build_optimizer(model, opt_func, lr)`` -> SGD-nesterov (wd 0.05) / AdamW (wd 0.05) / Adam,
with 1-D tensors, biases and ``model.no_weight_decay()`` names exempt from weight decay.
Stock ``torch.optim`` -- the optimizer is outside the accelerated hot path."""
from __future__ import annotations

from torch import optim

__all__ = ["build_optimizer", "set_weight_decay"]


def set_weight_decay(model, skip_list=(), skip_keywords=()):
    """Two parameter groups: decayed, and (1-D | bias | skipped) with weight_decay 0 (optimizer.py:13-27)."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        exempt = p.dim() == 1 or name.endswith(".bias") or name in skip_list or any(k in name for k in skip_keywords)
        (no_decay if exempt else decay).append(p)
    return [{"params": decay}, {"params": no_decay, "weight_decay": 0.0}]


def build_optimizer(model, opt_func="Adam", lr=1e-3):
    skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else {}
    skip_kw = model.no_weight_decay_keywords() if hasattr(model, "no_weight_decay_keywords") else {}
    groups = set_weight_decay(model, skip, skip_kw)
    kind = opt_func.lower()
    if kind == "sgd":
        return optim.SGD(groups, momentum=0.9, nesterov=True, lr=lr, weight_decay=0.05)
    if kind == "adamw":
        return optim.AdamW(groups, eps=1e-8, betas=(0.9, 0.999), lr=lr, weight_decay=0.05)
    if kind == "adam":
        return optim.Adam(groups, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    return None          # the reference returns None for unknown names (optimizer.py:41-52)
