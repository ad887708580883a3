"""Host-side optimizer factory with the reference's contract (``code/optimizer.py``):
``build_optimizer(model, opt_func, lr)`` -> SGD-nesterov (wd 0.05) / AdamW (wd 0.05) / Adam,
with 1-D tensors, biases and ``model.no_weight_decay()`` names exempt from weight decay.
Stock ``torch.optim`` objects: they stay the owners of all optimizer state, also when the step itself
runs through ``fused_step.FusedOptimizerEMA``."""
from __future__ import annotations

from functools import partial

from torch import optim

__all__ = ["build_optimizer", "set_weight_decay"]

# name (lower case) -> constructor with the reference's hyper-parameters (optimizer.py:43-51)
_FACTORIES = {
    "sgd": partial(optim.SGD, momentum=0.9, nesterov=True, weight_decay=0.05),
    "adamw": partial(optim.AdamW, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05),
    "adam": partial(optim.Adam, betas=(0.9, 0.999), eps=1e-8, weight_decay=0),
}


def _exempt_from_decay(name: str, param, skip_names, skip_substrings) -> bool:
    """optimizer.py:19-20: vectors (norm scales, biases), explicit names and name fragments carry no weight decay."""
    return param.dim() == 1 or name.endswith(".bias") or name in skip_names or any(s in name for s in skip_substrings)


def set_weight_decay(model, skip_list=(), skip_keywords=()):
    """Two parameter groups over the trainable parameters: decayed, and exempt with ``weight_decay=0``
    (optimizer.py:13-27).  Frozen parameters are left out."""
    trainable = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    exempt = [p for n, p in trainable if _exempt_from_decay(n, p, skip_list, skip_keywords)]
    exempt_ids = {id(p) for p in exempt}
    decayed = [p for _, p in trainable if id(p) not in exempt_ids]
    return [{"params": decayed}, {"params": exempt, "weight_decay": 0.0}]


def build_optimizer(model, opt_func="Adam", lr=1e-3):
    """``None`` for an unknown name, like the reference (optimizer.py:41-52)."""
    factory = _FACTORIES.get(opt_func.lower())
    if factory is None:
        return None
    groups = set_weight_decay(model, getattr(model, "no_weight_decay", dict)(), getattr(model, "no_weight_decay_keywords", dict)())
    return factory(groups, lr=lr)
