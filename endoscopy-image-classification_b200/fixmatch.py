"""Drop-in for the reference ``code/fixmatch.py``: class ``FixMatch`` with the same constructor,
``get_dataloader`` / ``get_config`` / ``train_one`` / ``evaluate_one`` / ``save_checkpoint`` /
``load_checkpoint`` / ``fit``.  The step (``fixmatch.py:101-127``) concatenates labeled, weak and
strong views for ONE backbone forward (stock PyTorch), then calls the fused sm_100a criteria:
``ce_loss(type_loss='poly')`` for the labeled rows and ``consistency_loss`` for the unlabeled
pair, followed by one multi-tensor ``ModelEMA.update``."""
from __future__ import annotations

import torch

from .loss import ce_loss, consistency_loss
from .trainer import SemiSupervisedTrainer

__all__ = ["FixMatch"]


class FixMatch(SemiSupervisedTrainer):
    TRAINING_MODE = "FixMatch"
    EMA_BEFORE_FREEZE = False          # fixmatch.py:40-55: freeze / unfreeze first, then the EMA copy

    def _train_step(self, epoch, batch_idx):
        inputs_x, targets_x = self._labeled.next()
        (inputs_u_w, inputs_u_s), _ = self._unlabeled.next()
        bs_lb = inputs_x.shape[0]
        targets_x = targets_x.to(self.device, non_blocking=True)
        inputs = self.to_device_views(inputs_x, inputs_u_w, inputs_u_s)
        with self._autocast():
            outputs = self.net(inputs)
        outputs_x = outputs[:bs_lb]
        outputs_u_w, outputs_u_s = outputs[bs_lb:].chunk(2)                    # contiguous row blocks
        lx = ce_loss(outputs_x, targets_x, class_weights=self.class_weights, reduction="mean", type_loss="poly")
        lu, self.last_mask_mean = consistency_loss(outputs_u_w, outputs_u_s, T=self.config.TRAIN.T,
                                                   p_cutoff=self.config.TRAIN.THRES, device=self.device)
        return lx + self.config.TRAIN.LAMBDA_U * lu                            # fixmatch.py:118
