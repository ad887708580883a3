"""CPU oracle for the SSL-head + EMA hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU (torch fp32 eager / numpy) *restatement* of the reference
algorithm for the one hot path this repo accelerates.  It is the checker, never
the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``endoscopy-image-classification_b200/`` imports it, and the product raises
when the CUDA extension is missing instead of falling back to this code.

Parity status: the reference repository ships no tests and no golden vectors
for this path ("parity unpinned" by the reference's own tests).  The oracle is
therefore pinned against *outputs of the reference itself run in the build
container*: ``oracle/pin_against_reference.py`` imports ``loss.py`` / ``ema.py``
from ``/root/reference/code`` and drives the real ``CoMatch.train_one`` /
``FixMatch.train_one`` through a stub-import harness, asserts this file
reproduces them, and writes ``tests/golden/*.npz``.  The CPU test-suite then
checks this file against those committed fixtures (the reference tree does not
exist on the GPU box).

Every function cites the reference lines (relative to ``/root/reference/``) it
follows.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "ce_loss", "poly_loss", "consistency_loss", "fixmatch_head_details",
    "ema_update_", "ema_set_", "ema_update_numpy",
    "CoMatchState", "comatch_da", "comatch_smooth", "bank_enqueue",
    "comatch_contrast", "comatch_focal_softce", "comatch_head",
    "comatch_head_sharded",
]


# --------------------------------------------------------------------------
# loss.py criteria
# --------------------------------------------------------------------------
def poly_loss(logits: torch.Tensor, targets: torch.Tensor,
              ce_weight: Optional[torch.Tensor] = None, reduction: str = "mean",
              epsilon: float = 2.0) -> torch.Tensor:
    """Poly-1 CE for hard integer targets.

    code/loss.py:308-364 (PolyLoss.forward, hard-target branch :338-342) as
    configured by ce_loss(type_loss='poly') at code/loss.py:103-114 (epsilon=2).
    Class weights scale the CE term only (:319); the mean is a plain mean
    over rows (:355-356), not the weight-normalised mean of F.cross_entropy.
    """
    ce = F.cross_entropy(logits, targets.long(), weight=ce_weight, reduction="none")
    pt = torch.softmax(logits, dim=1).gather(1, targets.long().view(-1, 1)).squeeze(1)
    out = ce + epsilon * (1.0 - pt)
    if reduction == "mean":
        return out.mean()
    if reduction == "sum":
        return out.sum()
    if reduction == "none":
        return out
    raise ValueError(f"Unsupported reduction: {reduction}")


def ce_loss(logits, targets, class_weights=None, use_hard_labels=True,
            reduction="none", type_loss="none", cls_num_list=None):
    """code/loss.py:90-124.  Only the branches on the hot path are restated:
    hard/plain (:118-119), hard/poly (:103-114) and soft targets (:120-124)."""
    if use_hard_labels:
        if type_loss == "poly":
            return poly_loss(logits, targets, class_weights, reduction, 2.0)
        if type_loss in ("focal", "ldam"):
            raise NotImplementedError("not on the SSL hot path (SURVEY §2 row 13)")
        return F.cross_entropy(logits, targets, weight=class_weights, reduction=reduction)
    assert logits.shape == targets.shape
    return torch.sum(-targets * F.log_softmax(logits, dim=-1), dim=1)


def fixmatch_head_details(logits_w: torch.Tensor, logits_s: torch.Tensor,
                          p_cutoff: float, T: float = 1.0,
                          use_hard_labels: bool = True) -> Dict[str, torch.Tensor]:
    """Everything the FixMatch unlabeled head computes, plus d(loss)/d(logits_s).

    code/loss.py:150-164: softmax of the detached weak logits (:128,:151), max /
    argmax (:153), mask = max_probs.ge(p_cutoff).float() (:154), masked
    hard-label CE against the strong logits (:157-160 -> :119), means (:164).
    Quirk Q4: T is not applied in hard-label mode.  The soft-label branch of the
    reference raises TypeError (loss.py:162-163); ``use_hard_labels=False`` here
    implements its evident intent and is an extension without reference oracle.
    """
    w = logits_w.detach().float()
    s = logits_s.detach().float().clone().requires_grad_(True)
    probs = torch.softmax(w, dim=-1)
    pmax, idx = torch.max(probs, dim=-1)
    mask = pmax.ge(p_cutoff).float()
    if use_hard_labels:
        per_row = F.cross_entropy(s, idx, reduction="none") * mask
    else:
        soft = torch.softmax(w / T, dim=-1)
        per_row = torch.sum(-soft * F.log_softmax(s, dim=-1), dim=1) * mask
    loss = per_row.mean()
    loss.backward()
    return {"loss": loss.detach(), "mask_mean": mask.mean(), "idx": idx,
            "mask": mask, "pmax": pmax, "grad_s": s.grad.detach()}


def consistency_loss(logits_w, logits_s, name="ce", T=1.0, p_cutoff=0.0,
                     use_hard_labels=True, device=None, loss_fc=None, fc=None):
    """code/loss.py:126-168 (live branches).  Returns what the reference returns:
    a (loss, mask.mean()) tuple for 'ce', a bare tensor for 'L2' (quirk Q8)."""
    assert name in ["ce", "L2"]
    logits_w = logits_w.detach()
    if name == "L2":
        assert logits_w.size() == logits_s.size()
        return F.mse_loss(logits_s, logits_w, reduction="mean")
    probs = torch.softmax(logits_w, dim=-1)
    pmax, idx = torch.max(probs, dim=-1)
    mask = pmax.ge(p_cutoff).float()
    if use_hard_labels:
        per_row = F.cross_entropy(logits_s, idx, reduction="none") * mask
    else:
        soft = torch.softmax(logits_w / T, dim=-1)
        per_row = torch.sum(-soft * F.log_softmax(logits_s, dim=-1), dim=1) * mask
    return per_row.mean(), mask.mean()


# --------------------------------------------------------------------------
# ema.py
# --------------------------------------------------------------------------
@torch.no_grad()
def ema_update_(ema_tensors: Sequence[torch.Tensor], model_tensors: Sequence[torch.Tensor],
                decay: float) -> None:
    """code/ema.py:51-59: for every state-dict entry (parameters AND buffers, in
    state_dict order, aliased entries visited as often as they appear),
    e <- decay*e + (1-decay)*m written back with copy_ (so integer buffers are
    computed in fp32 and truncated, quirk Q3)."""
    for e, m in zip(ema_tensors, model_tensors):
        e.copy_(decay * e + (1.0 - decay) * m)


@torch.no_grad()
def ema_set_(ema_tensors, model_tensors) -> None:
    """code/ema.py:61-62."""
    for e, m in zip(ema_tensors, model_tensors):
        e.copy_(m)


def ema_update_numpy(e: np.ndarray, m: np.ndarray, decay: float, repeat: int = 1) -> np.ndarray:
    """The rounding sequence of ema.py:59 spelled out for fp32 tensors: both
    python scalars are rounded to fp32 first ((1.-decay) is formed in double),
    then mul, mul, add are each rounded to fp32 -- no FMA contraction.
    ``repeat`` applies the update that many times (state_dict aliasing, Q2)."""
    d32 = np.float32(decay)
    o32 = np.float32(1.0 - decay)
    e = e.astype(np.float32, copy=True)
    m = m.astype(np.float32, copy=False)
    for _ in range(repeat):
        e = (d32 * e).astype(np.float32) + (o32 * m).astype(np.float32)
        e = e.astype(np.float32)
    return e


# --------------------------------------------------------------------------
# comatch.py:162-220 unlabeled head over explicit state
# --------------------------------------------------------------------------
@dataclass
class CoMatchState:
    """code/comatch.py:90-96: the memory bank, its write pointer and the
    distribution-alignment history."""
    queue_feats: torch.Tensor            # [K, D] fp32
    queue_probs: torch.Tensor            # [K, C] fp32
    queue_ptr: int = 0
    prob_list: List[torch.Tensor] = field(default_factory=list)

    @classmethod
    def zeros(cls, K: int, D: int, C: int) -> "CoMatchState":
        return cls(torch.zeros(K, D), torch.zeros(K, C), 0, [])

    @property
    def queue_size(self) -> int:
        return self.queue_feats.shape[0]

    def clone(self) -> "CoMatchState":
        return CoMatchState(self.queue_feats.clone(), self.queue_probs.clone(),
                            int(self.queue_ptr), [p.clone() for p in self.prob_list])


@torch.no_grad()
def comatch_da(logits_u_w: torch.Tensor, prob_list: List[torch.Tensor],
               window: int = 32) -> torch.Tensor:
    """code/comatch.py:167-176: softmax, push the batch column-mean on the
    history (keep the newest ``window``), divide by the history mean (stacked
    oldest -> newest), renormalise rows.  Mutates ``prob_list``."""
    probs = torch.softmax(logits_u_w.float(), dim=1)
    prob_list.append(probs.mean(0))
    if len(prob_list) > window:
        prob_list.pop(0)
    prob_avg = torch.stack(prob_list, dim=0).mean(0)
    probs = probs / prob_avg
    probs = probs / probs.sum(dim=1, keepdim=True)
    return probs


@torch.no_grad()
def comatch_smooth(probs: torch.Tensor, feats_u_w: torch.Tensor, queue_feats: torch.Tensor,
                   queue_probs: torch.Tensor, alpha: float, temperature: float) -> torch.Tensor:
    """code/comatch.py:180-182: A = exp(F_w Q_f^T / tau); A /= rowsum;
    probs = alpha*probs + (1-alpha) * A Q_p.  No running max (H2)."""
    A = torch.exp(torch.mm(feats_u_w.float(), queue_feats.float().t()) / temperature)
    A = A / A.sum(1, keepdim=True)
    return alpha * probs + (1 - alpha) * torch.mm(A, queue_probs.float())


@torch.no_grad()
def bank_enqueue(state: CoMatchState, feats_rows: torch.Tensor, probs_rows: torch.Tensor,
                 mode: str = "reference") -> None:
    """code/comatch.py:191-196.  ``mode='reference'`` keeps the guard
    ``n == queue_size`` (quirk Q1: with queue_batch=5 nothing is ever written);
    ``mode='always'`` is the same three lines without the guard, generalised
    to a wrapping ring write (upstream CoMatch semantics)."""
    n, K = feats_rows.shape[0], state.queue_size
    if mode == "reference":
        if n != K:
            return
    elif mode != "always":
        raise ValueError(mode)
    if n > K:
        raise ValueError("enqueue block larger than the bank")
    rows = (state.queue_ptr + torch.arange(n)) % K
    state.queue_feats[rows] = feats_rows.float()
    state.queue_probs[rows] = probs_rows.float()
    state.queue_ptr = (state.queue_ptr + n) % K


def comatch_contrast(feats_u_s0: torch.Tensor, feats_u_s1: torch.Tensor, probs: torch.Tensor,
                     temperature: float, contrast_th: float) -> torch.Tensor:
    """code/comatch.py:199-213: graph-contrastive loss (differentiable w.r.t.
    the two strong-view embeddings only; ``probs`` is a no-grad product)."""
    sim = torch.exp(torch.mm(feats_u_s0, feats_u_s1.t()) / temperature)
    sim_probs = sim / sim.sum(1, keepdim=True)
    Q = torch.mm(probs, probs.t())
    Q.fill_diagonal_(1)
    pos_mask = (Q >= contrast_th).float()
    Q = Q * pos_mask
    Q = Q / Q.sum(1, keepdim=True)
    return (-(torch.log(sim_probs + 1e-7) * Q).sum(1)).mean()


def comatch_focal_softce(logits_u_s0: torch.Tensor, probs: torch.Tensor, mask: torch.Tensor,
                         gamma: float) -> torch.Tensor:
    """code/comatch.py:216-220: focal-modulated soft CE (quirk Q7)."""
    logp = -torch.sum(F.log_softmax(logits_u_s0, dim=1) * probs, dim=1) * mask
    p = torch.exp(-logp)
    return ((1 - p) ** gamma * logp).mean()


def comatch_head(state: CoMatchState, logits_u_w, logits_u_s0, feats_u_w, feats_u_s0,
                 feats_u_s1, feats_x, targets_x, *, thr: float, num_classes: int,
                 alpha: float = 0.9, temperature: float = 0.2, contrast_th: float = 0.8,
                 gamma: float = 2, da_window: int = 32, enqueue_mode: str = "reference",
                 smoothing: bool = True, do_enqueue: bool = True) -> Dict[str, torch.Tensor]:
    """code/comatch.py:162-220 as one function over explicit state.

    Order as in the reference: DA (:167-176) -> smoothing against the bank
    *before* this step's enqueue (:179-182; the gate is always true, quirk Q6)
    -> max/mask (:184-185) -> enqueue rows = [unlabeled-weak ; labeled]
    (:187-196) -> contrastive (:199-213) -> focal soft-CE (:216-220).
    Returns losses, intermediates and the gradients w.r.t. logits_u_s0,
    feats_u_s0, feats_u_s1 (autograd, as the reference's ``losses.backward()``
    would produce for LAMBDA_U = LAMBDA_C = 1).
    """
    logits_u_w = logits_u_w.detach().float()
    feats_x = feats_x.detach().float()
    feats_u_w = feats_u_w.detach().float()
    s0 = logits_u_s0.detach().float().clone().requires_grad_(True)
    f0 = feats_u_s0.detach().float().clone().requires_grad_(True)
    f1 = feats_u_s1.detach().float().clone().requires_grad_(True)

    with torch.no_grad():
        probs = comatch_da(logits_u_w, state.prob_list, da_window)
        probs_orig = probs.clone()
        if smoothing:
            probs = comatch_smooth(probs, feats_u_w, state.queue_feats, state.queue_probs,
                                   alpha, temperature)
        scores, lbs = torch.max(probs, dim=1)
        mask = scores.ge(thr).float()
        bt = feats_x.shape[0]
        feats_w = torch.cat([feats_u_w, feats_x], dim=0)
        onehot = torch.zeros(bt, num_classes).scatter(1, targets_x.view(-1, 1).long(), 1)
        probs_w = torch.cat([probs_orig, onehot], dim=0)
        if do_enqueue:
            bank_enqueue(state, feats_w, probs_w, enqueue_mode)

    loss_c = comatch_contrast(f0, f1, probs, temperature, contrast_th)
    loss_u = comatch_focal_softce(s0, probs, mask, gamma)
    (loss_u + loss_c).backward()
    return {"loss_u": loss_u.detach(), "loss_contrast": loss_c.detach(),
            "probs": probs, "probs_orig": probs_orig, "scores": scores, "lbs": lbs,
            "mask": mask, "feats_w": feats_w, "probs_w": probs_w,
            "grad_logits_s0": s0.grad.detach(), "grad_feats_s0": f0.grad.detach(),
            "grad_feats_s1": f1.grad.detach()}


def comatch_head_sharded(state: CoMatchState, da_histories: List[List[torch.Tensor]],
                         rank_inputs: List[dict], **kw) -> List[Dict[str, torch.Tensor]]:
    """Single-process oracle for the R-rank data-parallel step with a global
    bank: every rank smooths its own queries against the *whole*
    pre-step bank; DA history, contrastive graph and focal CE are rank-local;
    afterwards the R enqueue blocks are written in rank-major order at the
    global pointer (what one process would do for the concatenated batch,
    comatch.py:187-196 applied R times)."""
    outs = []
    pre = state.clone()
    for r, inp in enumerate(rank_inputs):
        view = CoMatchState(pre.queue_feats, pre.queue_probs, pre.queue_ptr, da_histories[r])
        outs.append(comatch_head(view, **inp, do_enqueue=False, **kw))
    for o in outs:
        bank_enqueue(state, o["feats_w"], o["probs_w"], "always")
    return outs


# --------------------------------------------------------------------------
# evaluation
# --------------------------------------------------------------------------
def calculate_metrics(pred: np.ndarray, target: np.ndarray, num_classes: int) -> Dict[str, object]:
    """code/utils.py:38-55 verbatim in behaviour: per-class sensitivity / specificity from one-vs-rest
    ``precision_recall_fscore_support`` and micro / macro precision, recall, F1 through scikit-learn."""
    import pandas as pd
    from sklearn.metrics import f1_score, precision_recall_fscore_support, precision_score, recall_score
    res = []
    for l in range(num_classes):
        _, recall, _, _ = precision_recall_fscore_support(np.array(target) == l, np.array(pred) == l, pos_label=True, average=None)
        res.append([l, recall[1], recall[0]])
    df = pd.DataFrame(res, columns=["class", "sensitivity", "specificity"])
    out = {"sen/spec": df}
    for avg in ("micro", "macro"):
        out[f"{avg}/precision"] = precision_score(y_true=target, y_pred=pred, average=avg)
        out[f"{avg}/recall"] = recall_score(y_true=target, y_pred=pred, average=avg)
        out[f"{avg}/f1"] = f1_score(y_true=target, y_pred=pred, average=avg)
    return out


def evaluate_one(logits_batches: Sequence[torch.Tensor], target_batches: Sequence[torch.Tensor], num_classes: int,
                 batch_size: int) -> Dict[str, object]:
    """code/fixmatch.py:135-178 for scripted model outputs: per batch ``ce_loss(outputs, targets, reduction='mean')`` into
    an AverageMeter weighted by ``DATA.BATCH_SIZE`` (:158, the configured size, also for a ragged last batch), softmax ->
    argmax over the whole set (:165-168), then ``calculate_metrics``.  Also returns the confusion matrix (rows = target,
    columns = prediction) every metric above is a function of."""
    s = cnt = 0.0
    val = 0.0
    preds, targs = [], []
    for x, y in zip(logits_batches, target_batches):
        val = float(F.cross_entropy(x.float(), y.long(), reduction="mean"))
        s += val * batch_size
        cnt += batch_size
        preds.append(np.argmax(torch.softmax(x.float(), dim=1).numpy(), axis=1))
        targs.append(y.numpy())
    pred, targ = np.concatenate(preds), np.concatenate(targs)
    conf = np.zeros((num_classes, num_classes), dtype=np.int64)
    np.add.at(conf, (targ, pred), 1)
    return {"loss_avg": s / cnt, "loss_val": val, "loss_sum": s, "loss_count": cnt, "pred": pred, "target": targ,
            "confusion": conf, "metric": calculate_metrics(pred, targ, num_classes)}

