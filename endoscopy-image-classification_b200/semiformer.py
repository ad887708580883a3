"""Drop-in for the reference ``code/semiformer.py``: FixMatch on a two-head Conformer
(``model(x) -> (out_conv, out_trans)``).  Epochs below ``TRAIN.EVAL_STEP_SUP`` run the labeled-only
loop (``semiformer.py:72-100``); afterwards the weak pseudo-labels come from the conv head only
and BOTH strong heads are trained against them (``:122-131``) -- here with one launch
(``consistency_loss_dual``) instead of two ``consistency_loss`` calls."""
from __future__ import annotations

import torch

from .loss import ce_loss, consistency_loss_dual
from .trainer import SemiSupervisedTrainer

__all__ = ["SemiFormer"]


class SemiFormer(SemiSupervisedTrainer):
    TRAINING_MODE = "SemiFormer"
    EMA_BEFORE_FREEZE = True           # semiformer.py:39-49

    def _supervised_phase(self, epoch):
        return epoch < self.config.TRAIN.EVAL_STEP_SUP

    def _steps_in_epoch(self, epoch):
        return len(self.train_labeled_dl) if self._supervised_phase(epoch) else self.config.TRAIN.EVAL_STEP

    def _schedule_index(self, epoch, batch_idx, steps):
        if self._supervised_phase(epoch):
            return epoch * steps + batch_idx                                   # semiformer.py:93 (num_steps = len(labeled loader))
        return super()._schedule_index(epoch, batch_idx, steps)                # semiformer.py:139

    def _eval_forward(self, model, images):
        out_conv, out_trans = model(images)
        return out_conv + out_trans

    def _train_step(self, epoch, batch_idx):
        cw = self.class_weights
        if self._supervised_phase(epoch):
            images, targets = self._labeled.next()
            images, targets = images.to(self.device, non_blocking=True), targets.to(self.device, non_blocking=True)
            with self._autocast():
                out_conv, out_trans = self.net(images)
            return (ce_loss(out_conv, targets, class_weights=cw, reduction="mean")
                    + ce_loss(out_trans, targets, class_weights=cw, reduction="mean"))
        inputs_x, targets_x = self._labeled.next()
        (inputs_u_w, inputs_u_s), _ = self._unlabeled.next()
        bs_lb = inputs_x.shape[0]
        targets_x = targets_x.to(self.device, non_blocking=True)
        inputs = self.to_device_views(inputs_x, inputs_u_w, inputs_u_s)
        with self._autocast():
            out_conv, out_trans = self.net(inputs)
        outputs_u_w, outputs_u_s_conv = out_conv[bs_lb:].chunk(2)
        outputs_u_s_trans = out_trans[bs_lb:].chunk(2)[1]
        lx = (ce_loss(out_conv[:bs_lb], targets_x, class_weights=cw, reduction="mean")
              + ce_loss(out_trans[:bs_lb], targets_x, class_weights=cw, reduction="mean"))
        lu_conv, lu_trans, self.last_mask_mean = consistency_loss_dual(
            outputs_u_w, outputs_u_s_conv, outputs_u_s_trans, T=self.config.TRAIN.T, p_cutoff=self.config.TRAIN.THRES)
        return lx + self.config.TRAIN.LAMBDA_U * (lu_conv + lu_trans)          # semiformer.py:131-133
