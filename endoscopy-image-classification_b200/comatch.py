"""Drop-in for the reference ``code/comatch.py``: class ``CoMatch`` with the same public
attributes (``queue_batch, alpha, temperature, contrast_th, gamma`` -- ``comatch.py:29-39``;
``low_dim, queue_size, queue_feats, queue_probs, queue_ptr, prob_list`` -- ``:90-96``) and methods.
The inline head of ``train_one`` (``:162-220``) is the device-resident ``CoMatchHead``.

Optional YAML knobs (absent from the reference, whose values are hard-coded attributes):
``TRAIN.ALPHA, TEMPERATURE, CONTRAST_TH, GAMMA, QUEUE_BATCH, QUEUE_SIZE, ENQUEUE_MODE``
(``utils.COMATCH_OPTIONAL_KNOBS``).  ``ENQUEUE_MODE: 'reference'`` (default) keeps the guard of
``comatch.py:192`` (quirk Q1: with queue_batch = 5 the bank is never written); ``'always'`` is the
upstream CoMatch ring buffer.  With ``torch.distributed`` initialised and
``TRAIN.SHARD_BANK: True`` the bank is sharded over the ranks (``bank.py``)."""
from __future__ import annotations

import torch

from .comatch_head import CoMatchHead
from .loss import ce_loss
from .trainer import SemiSupervisedTrainer, _cfg
from .utils import COMATCH_OPTIONAL_KNOBS

__all__ = ["CoMatch"]


class CoMatch(SemiSupervisedTrainer):
    TRAINING_MODE = "CoMatch"
    EMA_BEFORE_FREEZE = True           # comatch.py:53-73

    def __init__(self, model, opt_func="Adam", lr=1e-3, device="cpu"):
        super().__init__(model, opt_func, lr, device)
        self.queue_batch = 5           # number of batches stored in the memory bank
        self.alpha = 0.9
        self.temperature = 0.2         # softmax temperature
        self.contrast_th = 0.8         # pseudo-label graph threshold
        self.gamma = 2                 # focal exponent of the unlabeled branch
        self.enqueue_mode = "reference"
        self.head = None

    def _freeze_backbone(self):
        super()._freeze_backbone()
        self.model.head_emb.requires_grad_(True)                # comatch.py:73

    def get_config(self, config, optimizer=None, lr_scheduler=None):
        super().get_config(config, optimizer, lr_scheduler)
        queue_size = None
        for key, (attr, _default) in COMATCH_OPTIONAL_KNOBS.items():
            if key in config.TRAIN:
                if attr == "queue_size":
                    queue_size = int(config.TRAIN[key])
                else:
                    setattr(self, attr, config.TRAIN[key])
        self.low_dim = config.MODEL.LOW_DIM
        pg, world = None, 1
        # one ring for the whole job, spread over the ranks, whenever the trainer runs data parallel (TRAIN.SHARD_BANK: False
        # keeps an independent per-rank bank of the per-rank size instead)
        if _cfg(config.TRAIN, "SHARD_BANK", True) and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            pg = torch.distributed.group.WORLD
            world = torch.distributed.get_world_size(pg)
        # comatch.py:91 with the GLOBAL batch: one ring for the whole job, every rank contributes its block per step
        # (TRAIN.QUEUE_SIZE, when given, is the global size)
        self.queue_size = queue_size or self.queue_batch * (config.DATA.MU + 1) * config.DATA.BATCH_SIZE * world
        self.head = CoMatchHead(config.MODEL.NUM_CLASSES, self.low_dim, self.queue_size, config.TRAIN.THRES,
                                alpha=self.alpha, temperature=self.temperature, contrast_th=self.contrast_th,
                                gamma=self.gamma, enqueue_mode=self.enqueue_mode, device=self.device, process_group=pg,
                                exchange=_cfg(config.TRAIN, "BANK_EXCHANGE", "auto"))

    # ---- reference-visible bank state (comatch.py:92-96) lives in the head -------------------
    queue_feats = property(lambda self: self.head.queue_feats)
    queue_probs = property(lambda self: self.head.queue_probs)
    prob_list = property(lambda self: self.head.prob_list)

    @property
    def queue_ptr(self):
        return self.head.queue_ptr

    @queue_ptr.setter
    def queue_ptr(self, v):
        self.head.queue_ptr = v

    def _steps_in_epoch(self, epoch):
        return len(self.train_unlabeled_dl)                     # the loop runs over the unlabeled loader (comatch.py:131)

    def _train_step(self, epoch, batch_idx):
        (inputs_u_w, inputs_u_s_0, inputs_u_s_1), _ = self._unlabeled.next()
        inputs_x, targets_x = self._labeled.next()
        bt, btu = inputs_x.size(0), inputs_u_w.size(0)
        imgs = self.to_device_views(inputs_x, inputs_u_w, inputs_u_s_0, inputs_u_s_1)
        targets_x = targets_x.to(self.device, non_blocking=True)
        with self._autocast():
            logits, _, features = self.net(imgs)
        logits_x = logits[:bt]
        logits_u_w, logits_u_s0, _ = torch.split(logits[bt:], btu)              # logits_u_s1 is unused (comatch.py:151)
        feats_x = features[:bt]
        feats_u_w, feats_u_s0, feats_u_s1 = torch.split(features[bt:], btu)
        loss_x = ce_loss(logits=logits_x, targets=targets_x, class_weights=self.class_weights, reduction="mean",
                         type_loss="poly")
        T = self.config.TRAIN
        total_u, self.last_loss_u, self.last_loss_contrast, self.last_mask_mean = self.head.total_loss(
            logits_u_w, logits_u_s0, feats_u_w, feats_u_s0, feats_u_s1, feats_x, targets_x,
            lambda_u=float(T.LAMBDA_U), lambda_c=float(T.LAMBDA_C),
            smooth=(epoch > 0 or batch_idx > self.queue_batch))                 # the gate of comatch.py:179
        return loss_x + total_u                                                 # comatch.py:222

    # the reference does not checkpoint the bank / DA history (comatch.py:285-306); we add them
    def _extra_state(self):
        """Collective in a multi-rank job: the shards of the bank are gathered so that rank 0 stores the whole ring (the
        checkpoint does not depend on the number of ranks that wrote it); the DA history is rank 0's."""
        return {"comatch_head": {k: (v.detach().cpu() if torch.is_tensor(v) else v)
                                 for k, v in self.head.state_dict(full=True).items()}}

    def _load_extra_state(self, checkpoint):
        if "comatch_head" in checkpoint and self.head is not None:
            self.head.load_state_dict(checkpoint["comatch_head"])
