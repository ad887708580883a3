"""Host-side helpers the SSL trainers read (reference ``code/utils.py:16-36, 128-134``):
``AttrDict``, ``AverageMeter`` and the two-level YAML ``get_config``.  Metrics,
plots and image helpers of the reference's utils.py are outside the hot path."""
from __future__ import annotations

import yaml

__all__ = ["AttrDict", "AverageMeter", "get_config", "COMATCH_OPTIONAL_KNOBS"]


class AttrDict(dict):
    """dict with attribute access (utils.py:16-19)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


class AverageMeter(object):
    """Running average fed by ``losses.item()`` (utils.py:21-36)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def get_config(config_file):
    """YAML -> two-level AttrDict (``DATA`` / ``MODEL`` / ``TRAIN``), utils.py:128-134.
    No validation, like the reference; an unquoted ``None`` stays the string 'None'."""
    with open(config_file) as f:
        raw = yaml.safe_load(f)
    config = AttrDict(raw)
    for k in list(config.keys()):
        if isinstance(config[k], dict):
            config[k] = AttrDict(config[k])
    return config


# Optional TRAIN.* keys (absent from the reference YAMLs, whose CoMatch knobs are
# hard-coded attributes at comatch.py:29-39) -> (attribute, default)
COMATCH_OPTIONAL_KNOBS = {
    "ALPHA": ("alpha", 0.9),
    "TEMPERATURE": ("temperature", 0.2),
    "CONTRAST_TH": ("contrast_th", 0.8),
    "GAMMA": ("gamma", 2),
    "QUEUE_BATCH": ("queue_batch", 5),
    "QUEUE_SIZE": ("queue_size", None),
    "ENQUEUE_MODE": ("enqueue_mode", "reference"),
}
