"""Synthetic Hyper-Kvasir-shaped inputs for benchmarks, smoke tests and examples
(SURVEY section 8d): there is no network for datasets or checkpoints, so the head is fed
random-init-like logits / unit-norm embeddings and the EMA runs over random-init
weights of the reference architectures.

``modelwemb_like`` reproduces the *state_dict contract* of the reference's
``ModelwEmb`` (``code/models/custom_model.py:147-213``): a stock backbone whose
modules are registered under three names (``model``, ``fc``, ``backbone``) plus the
``head_emb`` projector, so ``state_dict()`` lists every backbone tensor twice
(quirk Q2).  The backbone itself is stock torchvision -- it is the baseline, not the
product.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

NUM_CLASSES = 23   # Hyper-Kvasir (reference README.md:8-23)


class Normalize(nn.Module):
    """L2 row normalisation (custom_model.py:136-145)."""

    def __init__(self, power: int = 2):
        super().__init__()
        self.power = power

    def forward(self, x):
        return x / x.pow(self.power).sum(1, keepdim=True).pow(1.0 / self.power)


def complex_head(in_fts: int, out_fts: int) -> nn.Sequential:
    """custom_model.py:107-116 (is_complex=True)."""
    return nn.Sequential(nn.Linear(in_fts, in_fts // 4), nn.ReLU(), nn.Dropout(0.2), nn.BatchNorm1d(in_fts // 4),
                         nn.Linear(in_fts // 4, out_fts))


class ModelwEmbLike(nn.Module):
    def __init__(self, arch: str = "resnet50", num_classes: int = NUM_CLASSES, low_dim: int = 64):
        super().__init__()
        import torchvision
        self.model = getattr(torchvision.models, arch)(weights=None)
        in_fts = self.model.fc.in_features
        self.model.fc = complex_head(in_fts, num_classes)
        self.fc = self.model.fc                                              # alias (custom_model.py:195)
        self.backbone = nn.Sequential(*(list(self.model.children())[:-1]))   # alias (custom_model.py:199)
        self.head_emb = nn.Sequential(nn.Linear(in_fts, low_dim * 3), nn.LeakyReLU(inplace=True, negative_slope=0.1),
                                      nn.Linear(low_dim * 3, low_dim), Normalize(2))

    def forward(self, x):
        fts = torch.flatten(self.backbone(x), 1)
        return self.fc(fts), fts, self.head_emb(fts)


def modelwemb_like(arch: str = "resnet50", num_classes: int = NUM_CLASSES, low_dim: int = 64) -> nn.Module:
    return ModelwEmbLike(arch, num_classes, low_dim)


def perturb_(model: nn.Module, gen: torch.Generator, scale: float = 1e-3) -> None:
    """m <- m + scale*N(0,1) on every floating state tensor (stands in for an optimizer step)."""
    with torch.no_grad():
        seen = set()
        for v in model.state_dict().values():
            if v.data_ptr() in seen or not v.is_floating_point():
                continue
            seen.add(v.data_ptr())
            v.add_(scale * torch.randn(v.shape, generator=gen, device=gen.device, dtype=torch.float32).to(v.dtype))


def rownorm(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=1, keepdim=True)


def comatch_step_inputs(gen: torch.Generator, batch: int, mu: int, low_dim: int = 64, num_classes: int = NUM_CLASSES,
                        protos: torch.Tensor = None, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """One CoMatch step worth of head inputs (CPU tensors), clustered around class
    prototypes so that the confidence mask fires."""
    bu = batch * mu
    if protos is None:
        protos = rownorm(torch.randn(num_classes, low_dim, generator=gen))
    y_u = torch.randint(0, num_classes, (bu,), generator=gen)
    y_x = torch.randint(0, num_classes, (batch,), generator=gen)

    def feats(y):
        return rownorm(protos[y] + 0.075 * torch.randn(len(y), low_dim, generator=gen)).to(dtype)

    def logits(y, scale):
        return (scale * torch.nn.functional.one_hot(y, num_classes).float()
                + 2.0 * torch.randn(len(y), num_classes, generator=gen)).to(dtype)

    return {"logits_x": logits(y_x, 5.0), "logits_u_w": logits(y_u, 5.0), "logits_u_s0": logits(y_u, 4.0),
            "feats_u_w": feats(y_u), "feats_u_s0": feats(y_u), "feats_u_s1": feats(y_u), "feats_x": feats(y_x),
            "targets_x": y_x}


def fixmatch_step_inputs(gen: torch.Generator, batch: int, mu: int, num_classes: int = NUM_CLASSES, scale: float = 6.0,
                         dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """logits = 6*N(0,1): mask rate ~0.30 at threshold 0.95."""
    bu = batch * mu
    return {"logits_x": (scale * torch.randn(batch, num_classes, generator=gen)).to(dtype),
            "logits_u_w": (scale * torch.randn(bu, num_classes, generator=gen)).to(dtype),
            "logits_u_s": (scale * torch.randn(bu, num_classes, generator=gen)).to(dtype),
            "targets_x": torch.randint(0, num_classes, (batch,), generator=gen)}
