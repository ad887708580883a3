"""Per-iteration LR schedules with the interface the trainers call (``code/lr_scheduler.py``):
``build_scheduler(config, optimizer, n_iter_per_epoch)`` -> object with ``step_update(num_updates)``,
``state_dict()`` / ``load_state_dict()``.  The reference wraps timm's Cosine/Step schedulers and a
local linear one; timm is not a dependency here, so the three shapes are implemented directly
(warm-up from ``TRAIN.WARMUP_LR`` over ``WARMUP_EPOCHS``, then cosine to 5e-6 / linear to 1% /
step decay by ``LR_DECAY`` every ``DECAY_EPOCHS``).  Host code, outside the hot path."""
from __future__ import annotations

import math

__all__ = ["build_scheduler", "IterScheduler"]


class IterScheduler:
    def __init__(self, optimizer, kind, t_initial, warmup_t, warmup_lr_init, lr_min=5e-6, lr_min_rate=0.01, decay_t=1,
                 decay_rate=0.9):
        self.optimizer, self.kind = optimizer, kind
        self.t_initial, self.warmup_t, self.warmup_lr_init = int(t_initial), int(warmup_t), float(warmup_lr_init)
        self.lr_min, self.lr_min_rate, self.decay_t, self.decay_rate = lr_min, lr_min_rate, max(int(decay_t), 1), decay_rate
        for g in optimizer.param_groups:           # like timm: the schedule is anchored on initial_lr, which survives a
            g.setdefault("initial_lr", g["lr"])    # resume (optimizer.load_state_dict restores the decayed g["lr"])
        self.base_values = [g["initial_lr"] for g in optimizer.param_groups]
        self.last_update = -1
        if self.warmup_t:
            self._set([self.warmup_lr_init] * len(self.base_values))

    def _set(self, values):
        for g, v in zip(self.optimizer.param_groups, values):
            g["lr"] = v

    def _lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * (v - self.warmup_lr_init) / self.warmup_t for v in self.base_values]
        if self.kind == "cosine":
            if t >= self.t_initial:          # cycle_limit = 1
                return [self.lr_min for _ in self.base_values]
            return [self.lr_min + 0.5 * (v - self.lr_min) * (1 + math.cos(math.pi * t / self.t_initial)) for v in self.base_values]
        if self.kind == "linear":
            tt, total = t - self.warmup_t, max(self.t_initial - self.warmup_t, 1)
            return [v - (v - v * self.lr_min_rate) * (tt / total) for v in self.base_values]
        if self.kind == "step":
            return [v * (self.decay_rate ** (t // self.decay_t)) for v in self.base_values]
        raise ValueError(self.kind)

    def step_update(self, num_updates, metric=None):
        self.last_update = int(num_updates)
        self._set(self._lr(self.last_update))

    def step(self, epoch, metric=None):      # schedules are per-iteration (t_in_epochs=False)
        pass

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


def build_scheduler(config, optimizer, n_iter_per_epoch):
    """lr_scheduler.py:13-58: same knobs (TRAIN.EPOCHS / WARMUP_EPOCHS / DECAY_EPOCHS / WARMUP_LR / SCH_NAME / LR_DECAY)."""
    T = config.TRAIN
    num_steps = int(T.EPOCHS * n_iter_per_epoch)
    warmup = int(T.WARMUP_EPOCHS * n_iter_per_epoch)
    decay = int(T.DECAY_EPOCHS * n_iter_per_epoch)
    name = T.SCH_NAME
    if name not in ("cosine", "linear", "step"):
        return None
    return IterScheduler(optimizer, name, num_steps, warmup, T.WARMUP_LR, decay_t=decay, decay_rate=getattr(T, "LR_DECAY", 0.9))
