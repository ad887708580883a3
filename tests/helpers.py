"""Scripted model / loaders shared by the trainer tests (same harness the pin script drives the
REAL reference trainers with, oracle/pin_against_reference.py)."""
import torch
import torch.nn as nn


class Loader:
    def __init__(self, batches):
        self.batches = list(batches)

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


class NoSched:
    def step_update(self, it):
        pass

    def state_dict(self):
        return {}

    def load_state_dict(self, sd):
        pass


class ScriptedModel(nn.Module):
    """Replays one (logits, feats) pair per forward; the tensors are leaves so that
    ``losses.backward()`` leaves d(loss)/d(outputs) in ``.grad``."""

    def __init__(self, steps, with_feats, two_heads=False):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.steps, self.with_feats, self.two_heads = steps, with_feats, two_heads
        self.calls, self.seen = 0, []

    def forward(self, imgs):
        st = self.steps[self.calls]
        self.calls += 1
        dev = self.dummy.device
        logits = st["logits"].to(dev).clone().requires_grad_(True)
        if self.two_heads:
            logits2 = st["logits2"].to(dev).clone().requires_grad_(True)
            self.seen.append((logits, logits2))
            return logits + 0.0 * self.dummy, logits2 + 0.0 * self.dummy
        if self.with_feats:
            feats = st["feats"].to(dev).clone().requires_grad_(True)
            self.seen.append((logits, feats))
            return logits + 0.0 * self.dummy, None, feats + 0.0 * self.dummy
        self.seen.append((logits,))
        return logits + 0.0 * self.dummy
