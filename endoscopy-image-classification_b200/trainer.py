"""Shared host-side shell of the three semi-supervised trainers.

The reference repeats the same ~250 lines in ``code/fixmatch.py``, ``code/comatch.py`` and
``code/semiformer.py`` (constructor, ``get_dataloader``, ``get_config``, ``train_one``,
``evaluate_one``, ``save_checkpoint`` / ``load_checkpoint``, ``fit``).  Here the common
orchestration lives once; each trainer only supplies its step.  Public names, attributes,
config keys and checkpoint keys are the reference's, so driver code written against it
(``learn.py``-style) keeps working.  Everything in this file is stock PyTorch host code --
the accelerated path is what the steps call: ``loss.ce_loss`` / ``loss.consistency_loss``
/ ``CoMatchHead`` / ``ModelEMA``.
"""
from __future__ import annotations

import os
from contextlib import nullcontext
from datetime import date, datetime

import numpy as np
import torch

from .ema import ModelEMA
from .loss import ce_loss
from .lr_scheduler import build_scheduler
from .optimizer import build_optimizer
from .utils import AverageMeter

try:  # progress bars are cosmetic
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kw):
        return it


def _cfg(section, key, default=None):
    return section[key] if key in section else default


class BatchSource:
    """Endless ``next()`` over a DataLoader / iterable (the reference re-creates the iterator
    inside a bare ``except`` when it is exhausted, ``fixmatch.py:91-100``; quirk Q5: it calls the
    Python-2 style ``.next()`` -- here plain ``next(it)`` so modern loaders work)."""

    def __init__(self, loader):
        self.loader, self.it = loader, None

    def next(self):
        if self.it is None:
            self.it = iter(self.loader)
        try:
            return next(self.it)
        except StopIteration:
            self.it = iter(self.loader)
            return next(self.it)

    __next__ = next


class SemiSupervisedTrainer:
    TRAINING_MODE = "semi-supervised"
    EMA_BEFORE_FREEZE = False          # CoMatch / SemiFormer build the EMA copy before freezing (quirk Q9)

    def __init__(self, model, opt_func="Adam", lr=1e-3, device="cpu"):
        self.model = model
        self.opt_func = opt_func
        self.device = device
        self.model.to(self.device)
        self.epoch_start = 1
        self.best_valid_perf = None
        self._fused = None                 # fused_step.FusedOptimizerEMA when TRAIN.FUSED_OPT_EMA is set
        self._ddp = None                   # DistributedDataParallel wrapper of self.model in a multi-rank job

    # ---- reference API ---------------------------------------------------------------
    def get_dataloader(self, train_dl, valid_dl, test_dl=None):
        self.train_labeled_dl, self.train_unlabeled_dl = train_dl
        self.valid_dl = valid_dl
        self.test_dl = test_dl

    def _freeze_backbone(self):
        for p in self.model.parameters():
            p.requires_grad = False
        head = self.model.classifier if self.config.MODEL.NAME == "densenet161" else self.model.fc
        head.requires_grad_(True)

    def _apply_freeze(self):
        if _cfg(self.config.TRAIN, "IS_FREEZE", False):
            print("Freeze backbone")
            self._freeze_backbone()
        elif not self.EMA_BEFORE_FREEZE:
            print("Unfreeze backbone")
            for p in self.model.parameters():
                p.requires_grad = True

    def _build_ema(self):
        if self.config.TRAIN.USE_EMA:
            # TRAIN.EMA_OVERLAP: the update runs on a side stream next to the following step's forward / head and is joined
            # ahead of the next optimizer step (EMA(t) only has to be finished before the weights change again)
            self.ema_model = ModelEMA(model=self.model, decay=self.config.TRAIN.EMA_DECAY, device=self.device,
                                      overlap=bool(_cfg(self.config.TRAIN, "EMA_OVERLAP", False)))

    def _class_weights(self):
        if not _cfg(self.config.TRAIN, "CLS_WEIGHT", False):
            return None
        from sklearn.utils import class_weight
        df = self.train_labeled_dl.dataset.df
        y = list(df[self.config.DATA.TARGET_NAME])
        w = class_weight.compute_class_weight(class_weight="balanced", classes=np.unique(y).tolist(), y=y)
        return torch.tensor(w, dtype=torch.float).to(self.device)

    # ---- data parallel (the reference is single process) -------------------------------------------
    @staticmethod
    def _dist():
        import torch.distributed as dist
        return dist if (dist.is_available() and dist.is_initialized()) else None

    @property
    def rank(self) -> int:
        d = self._dist()
        return d.get_rank() if d is not None else 0

    @property
    def world_size(self) -> int:
        d = self._dist()
        return d.get_world_size() if d is not None else 1

    @property
    def net(self):
        """What the step calls: the DDP wrapper in a multi-rank job (its backward all-reduces the gradients), else the
        module itself.  ``self.model`` always stays the bare module -- EMA, checkpoints and freezing see the reference's
        parameter names."""
        return self._ddp if self._ddp is not None else self.model

    def _wrap_ddp(self):
        """``TRAIN.DDP`` (default: on whenever torch.distributed is initialised with more than one rank): every rank
        holds a replica, gradients are averaged by DistributedDataParallel during ``backward()``; the optimizer step
        and ``ModelEMA.update`` then run identically on every rank (no further communication: SURVEY 8e)."""
        self._ddp = None
        want = _cfg(self.config.TRAIN, "DDP", None)          # None: whenever there is more than one rank; True: also a 1-rank group
        if self._dist() is not None and (want or (want is None and self.world_size > 1)):
            from torch.nn.parallel import DistributedDataParallel
            if not any(p.requires_grad for p in self.model.parameters()):
                return
            dev = torch.device(self.device)
            self._ddp = DistributedDataParallel(self.model, device_ids=[dev.index] if dev.type == "cuda" else None,
                                                broadcast_buffers=_cfg(self.config.TRAIN, "DDP_BROADCAST_BUFFERS", True),
                                                find_unused_parameters=_cfg(self.config.TRAIN, "DDP_FIND_UNUSED", False))

    def distributed_loader(self, dataset, batch_size, shuffle=True, drop_last=True, **kw):
        """A DataLoader whose sampler deals the dataset over the ranks (the reference's loaders are single-process
        RandomSampler ones, ``dataset.py``); ``train_one`` calls ``sampler.set_epoch``."""
        from torch.utils.data import DataLoader
        from torch.utils.data.distributed import DistributedSampler
        sampler = DistributedSampler(dataset, num_replicas=self.world_size, rank=self.rank, shuffle=shuffle, drop_last=drop_last)
        return DataLoader(dataset, batch_size=batch_size, sampler=sampler, drop_last=drop_last, **kw)

    def get_config(self, config, optimizer=None, lr_scheduler=None):
        """``fixmatch.py:36-70`` / ``comatch.py:48-96`` / ``semiformer.py:37-62``.  ``optimizer`` /
        ``lr_scheduler`` may be injected; by default they are built like the reference does."""
        self.config = config
        if self.rank == 0:
            print(f"Training mode: {self.TRAINING_MODE}")
        if self.EMA_BEFORE_FREEZE:
            self._build_ema()
            self._apply_freeze()
        else:
            self._apply_freeze()
            self._build_ema()
        self.optimizer = optimizer or build_optimizer(self.model, opt_func=self.opt_func, lr=self.config.TRAIN.BASE_LR)
        self.lr_scheduler = lr_scheduler or build_scheduler(config=self.config, optimizer=self.optimizer,
                                                            n_iter_per_epoch=config.TRAIN.EVAL_STEP)
        self.class_weights = self._class_weights()
        self._wrap_ddp()
        amp = _cfg(self.config.TRAIN, "AMP", False)
        # TRAIN.FUSED_OPT_EMA: optimizer.step() + ema.update() as one multi-tensor launch (fused_step.py)
        self._fused = None
        if _cfg(self.config.TRAIN, "FUSED_OPT_EMA", False):
            from .fused_step import FusedOptimizerEMA
            self._fused = FusedOptimizerEMA(self.optimizer, self.ema_model if self.config.TRAIN.USE_EMA else None, self.model)
        self._autocast = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if amp else nullcontext
        self._labeled = BatchSource(self.train_labeled_dl) if hasattr(self, "train_labeled_dl") else None
        self._unlabeled = BatchSource(self.train_unlabeled_dl) if hasattr(self, "train_unlabeled_dl") else None

    def to_device_views(self, *batches):
        """Concatenate image batches on the device.  fp32 NCHW tensors (the reference's loaders, ``dataset.py:51-53``) are
        simply moved; uint8 ``[N, H, W, 3]`` batches -- loaders that stop before ``ToTensor`` -- are moved at one byte per
        sample and finished there by ``views.normalize_views`` (``ToTensor`` + ``Normalize`` in one launch, bit-exact)."""
        if all(b.dtype == torch.uint8 and b.dim() == 4 and b.shape[-1] == 3 for b in batches):
            from .views import normalize_views
            x = torch.cat([b.to(self.device, non_blocking=True) for b in batches], dim=0)
            return normalize_views(x)
        return torch.cat(batches, dim=0).to(self.device, non_blocking=True)

    def _after_backward(self, epoch, step_index, losses, summary_loss):
        """optimizer step, per-iteration LR schedule, EMA, bookkeeping (``fixmatch.py:120-131``)."""
        if self._fused is not None:
            self._fused.step()                         # optimizer + EMA: one launch (the EMA sees the new weights)
            if self.lr_scheduler is not None:
                self.lr_scheduler.step_update(step_index)
            self._fused.zero_grad()
        else:
            if self.config.TRAIN.USE_EMA:
                self.ema_model.join()                  # an overlapped EMA(t-1) still reads the weights this step rewrites
            self.optimizer.step()
            if self.lr_scheduler is not None:
                self.lr_scheduler.step_update(step_index)
            if self.config.TRAIN.USE_EMA:
                self.ema_model.update(self.model)      # one multi-tensor launch
            self.model.zero_grad()
        summary_loss.update(losses.item(), self.config.DATA.BATCH_SIZE)

    def train_one(self, epoch):
        self.model.train()
        summary_loss = AverageMeter()
        steps = self._steps_in_epoch(epoch)
        for dl in (getattr(self, "train_labeled_dl", None), getattr(self, "train_unlabeled_dl", None)):
            sampler = getattr(dl, "sampler", None)
            if hasattr(sampler, "set_epoch"):
                sampler.set_epoch(epoch)                # DistributedSampler: a different shuffle every epoch
        bar = tqdm(range(steps), total=steps, disable=self.rank != 0)
        for batch_idx in bar:
            losses = self._train_step(epoch, batch_idx)
            if self._fused is not None:
                self._fused.zero_grad()                 # in place: gradient addresses are in the device table
            else:
                self.optimizer.zero_grad()
            losses.backward()
            self._after_backward(epoch, self._schedule_index(epoch, batch_idx, steps), losses, summary_loss)
            if hasattr(bar, "set_postfix"):
                bar.set_postfix(loss=summary_loss.avg)
        return summary_loss

    def _steps_in_epoch(self, epoch):
        return self.config.TRAIN.EVAL_STEP

    def _schedule_index(self, epoch, batch_idx, steps):
        """Argument of ``lr_scheduler.step_update``: the reference always counts ``TRAIN.EVAL_STEP`` updates per epoch
        (``fixmatch.py:124``, ``comatch.py:228``, ``semiformer.py:139``) -- also in CoMatch, whose loop runs over the
        unlabeled loader -- because the scheduler was built with ``n_iter_per_epoch = EVAL_STEP``."""
        return epoch * self.config.TRAIN.EVAL_STEP + batch_idx

    def _train_step(self, epoch, batch_idx):  # pragma: no cover - abstract
        raise NotImplementedError

    def _eval_forward(self, model, images):
        return model(images)

    def evaluate_one(self, show_metric=False, show_report=False, show_cf_matrix=False):
        """Validation on the EMA weights when ``USE_EMA`` (``fixmatch.py:135-178``).  Returns ``(AverageMeter, metrics
        dict)`` with the keys of ``utils.calculate_metrics`` (micro / macro precision, recall, F1 and the ``sen/spec``
        table).  Per batch ONE launch of the evaluation head (``evaluation.EvalAccumulator``: mean CE, softmax arg-max,
        confusion matrix on the device); the pass synchronises once, for one device-to-host copy (the reference does
        two round trips per batch and recounts the predictions on the host).  Plots are out of scope; ``show_cf_matrix``
        prints the matrix."""
        from .evaluation import EvalAccumulator
        if self.config.TRAIN.USE_EMA:
            self.ema_model.join()
        eval_model = self.ema_model.ema if self.config.TRAIN.USE_EMA else self.model
        eval_model.eval()
        acc = EvalAccumulator(self.config.MODEL.NUM_CLASSES, self.device, max_batches=max(len(self.valid_dl), 1),
                              keep_predictions=bool(show_report))
        with torch.no_grad():
            for images, targets in tqdm(self.valid_dl, total=len(self.valid_dl), disable=self.rank != 0):
                images = images.to(self.device, non_blocking=True)
                targets = targets.to(self.device, non_blocking=True)
                acc.update(self._eval_forward(eval_model, images), targets)
        summary_loss, metric = acc.finalize(self.config.DATA.BATCH_SIZE)
        if show_metric:
            print("Metric:\n", metric)
        if show_report:
            from sklearn.metrics import classification_report
            preds, targs = acc.predictions()
            print("Classification Report:\n", classification_report(targs, preds))
        if show_cf_matrix:
            print("Confusion matrix (rows: actual, columns: predicted):\n", acc.confusion)
        return summary_loss, metric

    # ---- checkpoints (same dict keys as fixmatch.py:181-236) -------------------------------
    def _extra_state(self):
        return {}

    def _load_extra_state(self, checkpoint):
        pass

    def save_checkpoint(self, foldname):
        """Rank 0 writes the file (same dict keys as ``fixmatch.py:181-202``); the other ranks only take part in
        gathering state that is spread over the ranks (a sharded CoMatch bank) and wait at the barrier."""
        extra = self._extra_state()                     # collective when the bank is sharded
        d = self._dist()
        if self.rank != 0:
            if d is not None:
                d.barrier()
            return None
        checkpoint = {}
        if self.config.TRAIN.USE_EMA:
            self.ema_model.join()
            checkpoint["ema_state_dict"] = self.ema_model.ema.state_dict()
        stamp = date.today().strftime("%m_%d_%Y") + "_" + datetime.now().strftime("%H_%M_%S")
        checkpoint["epoch"] = self.epoch
        checkpoint["best_valid_perf"] = self.best_valid_perf
        checkpoint["model_state_dict"] = self.model.state_dict()
        checkpoint["optimizer"] = self.optimizer.state_dict()
        checkpoint["scheduler"] = self.lr_scheduler.state_dict() if self.lr_scheduler is not None else None
        checkpoint.update(extra)
        os.makedirs(foldname, exist_ok=True)
        path = os.path.join(foldname, f"{stamp}_epoch_{self.epoch}.pth")
        torch.save(checkpoint, path)
        print("Saved checkpoint")
        if d is not None:
            d.barrier()
        return path

    def load_checkpoint(self, checkpoint_dir, is_train=False):
        checkpoint = torch.load(checkpoint_dir, map_location="cpu", weights_only=False)
        self.model.load_state_dict(checkpoint["model_state_dict"])
        if is_train:
            # fixmatch.py:208-219: resuming re-applies TRAIN.IS_FREEZE (a frozen backbone stays frozen, so no stale
            # gradients pile up on parameters the optimizer never zeroes)
            for p in self.model.parameters():
                p.requires_grad = True
            if _cfg(self.config.TRAIN, "IS_FREEZE", False):
                self._freeze_backbone()
        else:
            for p in self.model.parameters():
                p.requires_grad = False
        if self.config.TRAIN.USE_EMA:
            self.ema_model.join()
            self.ema_model.ema.load_state_dict(checkpoint["ema_state_dict"])     # in place: pointers stay valid
            for p in self.ema_model.ema.parameters():
                p.requires_grad = bool(is_train)
        self.epoch_start = checkpoint["epoch"]
        self.best_valid_perf = checkpoint["best_valid_perf"]
        self.optimizer.load_state_dict(checkpoint["optimizer"])
        if self.lr_scheduler is not None and checkpoint.get("scheduler") is not None:
            self.lr_scheduler.load_state_dict(checkpoint["scheduler"])
        self._load_extra_state(checkpoint)

    def fit(self):
        for epoch in range(self.epoch_start, self.config.TRAIN.EPOCHS + 1):
            self.epoch = epoch
            best = f"{float(self.best_valid_perf):.3f}" if self.best_valid_perf else "inf"
            if self.rank == 0:
                print(f'Training epoch: {self.epoch} | Current LR: {self.optimizer.param_groups[0]["lr"]:.6f} | The best loss: {best}')
            train_loss = self.train_one(self.epoch)
            if self.rank == 0:
                print(f"\tTrain Loss: {train_loss.avg:.3f}")
            if epoch % self.config.TRAIN.FREQ_EVAL == 0:
                valid_loss, _ = self.evaluate_one()
                if not self.best_valid_perf or self.best_valid_perf > valid_loss.avg:
                    self.best_valid_perf = valid_loss.avg
                self.save_checkpoint(self.config.TRAIN.SAVE_CP)
                print(f"\tValid Loss: {valid_loss.avg:.3f}")
