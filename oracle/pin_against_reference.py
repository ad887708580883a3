"""Pin oracle/ssl_oracle.py against the REAL reference and write tests/golden/.

Runs only in the build container (needs /root/reference; the GPU box does not
have it).  It
  1. imports ``loss.py`` / ``ema.py`` unmodified from /root/reference/code and
     calls ``consistency_loss``, ``ce_loss`` (plain / poly / soft) and
     ``ModelEMA`` directly;
  2. runs the real ``CoMatch.train_one`` and ``FixMatch.train_one`` verbatim
     through a stub-import harness (seaborn / matplotlib / timm / cv2 are not
     installed here) with a scripted model that replays fixed (logits, feats),
     capturing losses, gradients, bank rows, pointer and DA history;
  3. asserts the oracle reproduces every one of those outputs (bit-exact for
     argmax / mask / bank / ptr / EMA, <=1e-6 rel for losses & grads), and
  4. stores inputs + reference outputs as small .npz fixtures.

Usage:  python oracle/pin_against_reference.py [--out tests/golden]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

sys.dont_write_bytecode = True
REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/code")
sys.path.insert(0, str(REPO))

from oracle import ssl_oracle as O  # noqa: E402

C = 23


# ----------------------------------------------------------------------------
# stub-import harness
# ----------------------------------------------------------------------------
def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return None

    for name in ["seaborn", "matplotlib", "cv2"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod(name)
    if "matplotlib.pyplot" not in sys.modules:
        try:
            __import__("matplotlib.pyplot")
        except Exception:
            sys.modules["matplotlib"].pyplot = mod("matplotlib.pyplot")
    try:
        __import__("timm")
    except Exception:
        mod("timm")
        mod("timm.loss", SoftTargetCrossEntropy=_Anything)
        mod("timm.scheduler")
        mod("timm.scheduler.cosine_lr", CosineLRScheduler=_Anything)
        mod("timm.scheduler.step_lr", StepLRScheduler=_Anything)
        mod("timm.scheduler.scheduler", Scheduler=object)


def import_reference():
    _install_stubs()
    sys.path.insert(0, str(REF))
    import loss as ref_loss  # noqa
    import ema as ref_ema  # noqa
    import comatch as ref_comatch  # noqa
    import fixmatch as ref_fixmatch  # noqa
    import utils as ref_utils  # noqa
    return ref_loss, ref_ema, ref_comatch, ref_fixmatch, ref_utils


class _Loader:
    """Iterable with len, whose iterators expose .next (quirk Q5)."""

    def __init__(self, batches):
        self.batches = list(batches)

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        outer = self

        class It:
            def __init__(self):
                self.i = 0

            def __iter__(self):
                return self

            def __next__(self):
                if self.i >= len(outer.batches):
                    raise StopIteration
                b = outer.batches[self.i]
                self.i += 1
                return b

            next = __next__
        return It()


class _Sched:
    def step_update(self, it):
        pass


class ScriptedModel(nn.Module):
    """Replays one (logits, feats) pair per forward; the tensors are leaves so
    the reference's own ``losses.backward()`` leaves d(loss)/d(outputs) in .grad."""

    def __init__(self, steps, with_feats):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.steps = steps
        self.with_feats = with_feats
        self.calls = 0
        self.seen = []

    def forward(self, imgs):
        st = self.steps[self.calls]
        self.calls += 1
        logits = st["logits"].clone().requires_grad_(True)
        # keep the graph attached to a parameter so optimizer.step() is legal
        out_logits = logits + 0.0 * self.dummy
        if self.with_feats:
            feats = st["feats"].clone().requires_grad_(True)
            self.seen.append((logits, feats))
            return out_logits, None, feats + 0.0 * self.dummy
        self.seen.append((logits,))
        return out_logits


def rownorm(x):
    return x / x.norm(dim=1, keepdim=True)


def clustered_step(g, B, MU, D, protos):
    """SURVEY §8d clustered CoMatch inputs (so that the mask is exercised)."""
    Bu = B * MU
    y_u = torch.randint(0, C, (Bu,), generator=g)
    y_x = torch.randint(0, C, (B,), generator=g)

    def feats(y):
        return rownorm(protos[y] + 0.075 * torch.randn(len(y), D, generator=g))

    def logit(y, scale):
        return scale * torch.nn.functional.one_hot(y, C).float() + 2.0 * torch.randn(len(y), C, generator=g)

    logits = torch.cat([logit(y_x, 5.0), logit(y_u, 5.0), logit(y_u, 4.0), logit(y_u, 4.0)])
    fts = torch.cat([feats(y_x), feats(y_u), feats(y_u), feats(y_u)])
    return {"logits": logits, "feats": fts, "targets_x": y_x}


def sha(arrs):
    h = hashlib.sha256()
    for k in sorted(arrs):
        h.update(k.encode())
        h.update(np.ascontiguousarray(arrs[k]).tobytes())
    return h.hexdigest()


def close(a, b, rtol=1e-6, atol=1e-7, what=""):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    if not torch.allclose(a, b, rtol=rtol, atol=atol):
        raise AssertionError(f"oracle != reference for {what}: max abs {float((a - b).abs().max())}")


def exact(a, b, what=""):
    if not torch.equal(torch.as_tensor(a), torch.as_tensor(b)):
        raise AssertionError(f"oracle != reference (bit-exact) for {what}")


# ----------------------------------------------------------------------------
def pin_fixmatch_criteria(ref_loss, out):
    for seed, Bu, thr, scale in [(0, 112, 0.95, 6.0), (1, 448, 0.95, 6.0), (2, 448, 0.7, 1.0), (3, 37, 0.0, 6.0)]:
        g = torch.Generator().manual_seed(seed)
        w = scale * torch.randn(Bu, C, generator=g)
        s = scale * torch.randn(Bu, C, generator=g)
        s_ref = s.clone().requires_grad_(True)
        lu, mm = ref_loss.consistency_loss(w, s_ref, T=1.0, p_cutoff=thr, device="cpu")
        lu.backward()
        pl = torch.softmax(w, dim=-1)
        pmax, idx = pl.max(-1)
        d = O.fixmatch_head_details(w, s, thr)
        close(d["loss"], lu.detach(), what="consistency loss")
        close(d["mask_mean"], mm, what="mask mean")
        close(d["grad_s"], s_ref.grad, what="grad_s")
        exact(d["idx"], idx, "argmax")
        exact(d["mask"], pmax.ge(thr).float(), "mask")
        lo, mo = O.consistency_loss(w, s, p_cutoff=thr)
        close(lo, lu.detach(), what="oracle.consistency_loss")
        # quirk Q4: T ignored in hard-label mode
        lu_T, _ = ref_loss.consistency_loss(w, s, T=0.5, p_cutoff=thr, device="cpu")
        exact(lu_T, lu.detach(), "T ignored")
        arrs = dict(logits_w=w.numpy(), logits_s=s.numpy(), thr=np.float64(thr),
                    loss=lu.detach().numpy(), mask_mean=mm.numpy(), idx=idx.numpy(),
                    mask=d["mask"].numpy(), pmax=pmax.numpy(), grad_s=s_ref.grad.numpy())
        np.savez_compressed(out / f"fixmatch_head_seed{seed}.npz", **arrs)
    # L2 branch returns a bare tensor (Q8)
    l2 = ref_loss.consistency_loss(w, s, name="L2")
    close(O.consistency_loss(w, s, name="L2"), l2, what="L2")

    # ce_loss branches
    g = torch.Generator().manual_seed(11)
    x = 3.0 * torch.randn(64, C, generator=g)
    y = torch.randint(0, C, (64,), generator=g)
    cw = 0.5 + torch.rand(C, generator=g)
    soft = torch.softmax(torch.randn(64, C, generator=g), -1)
    cases = {}
    for nm, kw in {"plain_none": dict(reduction="none"),
                   "plain_mean_w": dict(reduction="mean", class_weights=cw),
                   "poly_mean_w": dict(reduction="mean", class_weights=cw, type_loss="poly"),
                   "poly_mean": dict(reduction="mean", type_loss="poly"),
                   "poly_none": dict(reduction="none", type_loss="poly")}.items():
        xr = x.clone().requires_grad_(True)
        r = ref_loss.ce_loss(xr, y, **kw)
        r.sum().backward()
        xo = x.clone().requires_grad_(True)
        o = O.ce_loss(xo, y, **kw)
        o.sum().backward()
        close(o.detach(), r.detach(), what=f"ce_loss {nm}")
        close(xo.grad, xr.grad, what=f"ce_loss grad {nm}")
        cases[nm] = r.detach().numpy()
        cases[nm + "_grad"] = xr.grad.numpy()
    r = ref_loss.ce_loss(x, soft, use_hard_labels=False)
    close(O.ce_loss(x, soft, use_hard_labels=False), r, what="soft ce")
    cases["soft"] = r.numpy()
    np.savez_compressed(out / "ce_loss.npz", logits=x.numpy(), targets=y.numpy(),
                        class_weights=cw.numpy(), soft_targets=soft.numpy(), **cases)


class _TinyNet(nn.Module):
    def __init__(self, alias=False):
        super().__init__()
        self.conv = nn.Conv2d(3, 8, 3, bias=False)
        self.bn = nn.BatchNorm2d(8)
        self.fc = nn.Linear(8, C)
        self.model = nn.Sequential(self.conv, self.bn)  # alias -> state_dict repeats (Q2)
        if alias:
            self.backbone = nn.Sequential(self.conv, self.bn)

    def forward(self, x):
        return self.fc(self.bn(self.conv(x)).mean((2, 3)))


def pin_ema(ref_ema, out):
    for alias in (False, True):
        torch.manual_seed(5 + alias)
        model = _TinyNet(alias)
        ema = ref_ema.ModelEMA(model, decay=0.999, device="cpu")
        keys = list(ema.ema.state_dict().keys())
        e0 = {k: v.clone() for k, v in ema.ema.state_dict().items()}
        ours = {k: v.clone() for k, v in e0.items()}
        # aliased names must share storage in our copy too
        ptr_to_key = {}
        for k, v in ema.ema.state_dict().items():
            if v.data_ptr() in ptr_to_key:
                ours[k] = ours[ptr_to_key[v.data_ptr()]]
            else:
                ptr_to_key[v.data_ptr()] = k
        snaps = {}
        for step in range(3):
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1e-2 * torch.randn_like(p))
                model.bn.running_mean.add_(0.1 * torch.randn(8))
                model.bn.num_batches_tracked.add_(7)
            snaps[f"m{step}"] = {k: v.clone() for k, v in model.state_dict().items()}
            ema.update(model)
            O.ema_update_([ours[k] for k in keys], [model.state_dict()[k] for k in keys], 0.999)
            for k in keys:
                exact(ours[k], ema.ema.state_dict()[k], f"EMA {k} step {step}")
        # numpy rounding model (repeat = multiplicity of the storage)
        mult = {}
        for k, v in ema.ema.state_dict().items():
            mult[v.data_ptr()] = mult.get(v.data_ptr(), 0) + 1
        k = "conv.weight"
        rep = mult[ema.ema.state_dict()[k].data_ptr()]
        e = e0[k].numpy()
        for step in range(3):
            e = O.ema_update_numpy(e, snaps[f"m{step}"][k].numpy(), 0.999, rep)
        exact(torch.from_numpy(e), ema.ema.state_dict()[k], "numpy EMA rounding model")
        arrs = {}
        for k in keys:
            arrs[f"e0/{k}"] = e0[k].numpy()
            arrs[f"e3/{k}"] = ema.ema.state_dict()[k].numpy()
            for step in range(3):
                arrs[f"m{step}/{k}"] = snaps[f"m{step}"][k].numpy()
        arrs["keys"] = np.array(keys)
        arrs["decay"] = np.float64(0.999)
        np.savez_compressed(out / f"ema_tinynet_alias{int(alias)}.npz", **arrs)
        # set()
        ema.set(model)
        for k in keys:
            exact(ema.ema.state_dict()[k], model.state_dict()[k], "EMA set")


def _make_comatch(ref_comatch, ref_utils, ref_ema, model, B, MU, D, thr, queue_batch):
    tr = ref_comatch.CoMatch(model, device="cpu")
    tr.queue_batch = queue_batch
    cfg = ref_utils.AttrDict(
        DATA=ref_utils.AttrDict(BATCH_SIZE=B, MU=MU),
        MODEL=ref_utils.AttrDict(NUM_CLASSES=C, LOW_DIM=D),
        TRAIN=ref_utils.AttrDict(THRES=thr, USE_EMA=False, LAMBDA_U=1.0, LAMBDA_C=1.0, EVAL_STEP=8))
    tr.config = cfg
    tr.class_weights = None
    tr.optimizer = torch.optim.SGD(model.parameters(), lr=0.0)
    tr.lr_scheduler = _Sched()
    # comatch.py:90-96
    tr.low_dim = D
    tr.queue_size = tr.queue_batch * (MU + 1) * B
    tr.queue_feats = torch.zeros(tr.queue_size, D)
    tr.queue_probs = torch.zeros(tr.queue_size, C)
    tr.queue_ptr = 0
    tr.prob_list = []
    return tr


def pin_comatch(ref_comatch, ref_utils, ref_ema, out):
    B, MU, D, thr = 8, 7, 64, 0.9
    Bu = B * MU
    nsteps = 4
    for queue_batch in (1, 5):
        g = torch.Generator().manual_seed(100 + queue_batch)
        protos = rownorm(torch.randn(C, D, generator=g))
        steps = [clustered_step(g, B, MU, D, protos) for _ in range(nsteps)]
        model = ScriptedModel(steps, with_feats=True)
        tr = _make_comatch(ref_comatch, ref_utils, ref_ema, model, B, MU, D, thr, queue_batch)
        img = torch.zeros(1)
        lab = _Loader([(torch.zeros(B, 1), st["targets_x"]) for st in steps])
        unl = _Loader([((torch.zeros(Bu, 1), torch.zeros(Bu, 1), torch.zeros(Bu, 1)), None) for _ in steps])
        tr.get_dataloader((lab, unl), None)
        # quirk Q5: next(self.train_labeled_dl) fails -> fresh iterator every step
        # => every step sees labeled batch 0.  Reproduce by feeding targets of step 0.
        state = O.CoMatchState.zeros(tr.queue_size, D, C)
        arrs = {}
        # drive the reference one step at a time so that we can snapshot state
        for i, st in enumerate(steps):
            one_unl = _Loader([unl.batches[i]])
            tr.train_unlabeled_dl = one_unl
            model.calls = i
            import io, contextlib
            with contextlib.redirect_stderr(io.StringIO()):
                meter = tr.train_one(epoch=1)
            logits_leaf, feats_leaf = model.seen[-1]
            tx = steps[0]["targets_x"]          # Q5
            lg, ft = st["logits"], st["feats"]
            lx, (luw, lus0, lus1) = lg[:B], torch.split(lg[B:], Bu)
            fx, (fuw, fus0, fus1) = ft[:B], torch.split(ft[B:], Bu)
            o = O.comatch_head(state, luw, lus0, fuw, fus0, fus1, fx, tx, thr=thr, num_classes=C,
                               enqueue_mode="reference")
            loss_x = O.ce_loss(lx, tx, None, reduction="mean", type_loss="poly")
            total = loss_x + o["loss_u"] + o["loss_contrast"]
            close(total, meter.avg, rtol=2e-6, what=f"CoMatch total loss step {i}")
            g_lg, g_ft = logits_leaf.grad, feats_leaf.grad
            close(o["grad_logits_s0"], g_lg[B + Bu:B + 2 * Bu], rtol=1e-5, atol=1e-8, what="grad logits_u_s0")
            close(o["grad_feats_s0"], g_ft[B + Bu:B + 2 * Bu], rtol=1e-5, atol=1e-8, what="grad feats_u_s0")
            close(o["grad_feats_s1"], g_ft[B + 2 * Bu:], rtol=1e-5, atol=1e-8, what="grad feats_u_s1")
            assert float(g_lg[B:B + Bu].abs().max()) == 0.0 and float(g_lg[B + 2 * Bu:].abs().max()) == 0.0
            assert float(g_ft[:B + Bu].abs().max()) == 0.0
            exact(state.queue_feats, tr.queue_feats, f"queue_feats step {i}")
            exact(state.queue_probs, tr.queue_probs, f"queue_probs step {i}")
            assert state.queue_ptr == tr.queue_ptr
            for a, b in zip(state.prob_list, tr.prob_list):
                exact(a, b, "DA history")
            arrs[f"s{i}/logits"] = lg.numpy()
            arrs[f"s{i}/feats"] = ft.numpy()
            arrs[f"s{i}/targets_x"] = tx.numpy()
            arrs[f"s{i}/total_loss"] = np.float64(meter.avg)
            arrs[f"s{i}/grad_logits"] = g_lg.numpy()
            arrs[f"s{i}/grad_feats"] = g_ft.numpy()
            arrs[f"s{i}/queue_feats"] = tr.queue_feats.numpy().copy()
            arrs[f"s{i}/queue_probs"] = tr.queue_probs.numpy().copy()
            arrs[f"s{i}/queue_ptr"] = np.int64(tr.queue_ptr)
            arrs[f"s{i}/mask"] = o["mask"].numpy()
            arrs[f"s{i}/lbs"] = o["lbs"].numpy()
            arrs[f"s{i}/probs"] = o["probs"].numpy()
            arrs[f"s{i}/loss_u"] = o["loss_u"].numpy()
            arrs[f"s{i}/loss_contrast"] = o["loss_contrast"].numpy()
            arrs[f"s{i}/loss_x"] = loss_x.numpy()
        if queue_batch == 5:      # quirk Q1
            assert float(tr.queue_feats.abs().max()) == 0.0 and tr.queue_ptr == 0
        else:
            assert float(tr.queue_feats.abs().max()) > 0.0
            assert sum(float(arrs[f"s{i}/mask"].sum()) for i in range(nsteps)) > 0, "mask never fires"
        arrs["meta"] = np.array(json.dumps(dict(B=B, MU=MU, D=D, C=C, thr=thr, queue_batch=queue_batch,
                                                nsteps=nsteps, alpha=0.9, temperature=0.2,
                                                contrast_th=0.8, gamma=2)))
        np.savez_compressed(out / f"comatch_train_one_qb{queue_batch}.npz", **arrs)


def pin_fixmatch_train_one(ref_fixmatch, ref_utils, out):
    B, MU, thr = 16, 7, 0.95
    Bu = B * MU
    g = torch.Generator().manual_seed(42)
    steps = [{"logits": 6.0 * torch.randn(B + 2 * Bu, C, generator=g),
              "targets_x": torch.randint(0, C, (B,), generator=g)} for _ in range(2)]
    model = ScriptedModel(steps, with_feats=False)
    tr = ref_fixmatch.FixMatch(model, device="cpu")
    tr.config = ref_utils.AttrDict(
        DATA=ref_utils.AttrDict(BATCH_SIZE=B, MU=MU),
        MODEL=ref_utils.AttrDict(NUM_CLASSES=C),
        TRAIN=ref_utils.AttrDict(THRES=thr, T=1.0, USE_EMA=False, LAMBDA_U=1.0, EVAL_STEP=1))
    tr.class_weights = None
    tr.optimizer = torch.optim.SGD(model.parameters(), lr=0.0)
    tr.lr_scheduler = _Sched()
    arrs = {}
    for i, st in enumerate(steps):
        lab = _Loader([(torch.zeros(B, 1), st["targets_x"])])
        unl = _Loader([((torch.zeros(Bu, 1), torch.zeros(Bu, 1)), None)])
        tr.get_dataloader((lab, unl), None)
        model.calls = i
        import io, contextlib
        with contextlib.redirect_stderr(io.StringIO()):
            meter = tr.train_one(epoch=0)
        (leaf,) = model.seen[-1]
        lg = st["logits"]
        lx = O.ce_loss(lg[:B], st["targets_x"], None, reduction="mean", type_loss="poly")
        w, s = lg[B:].chunk(2)
        d = O.fixmatch_head_details(w, s, thr)
        close(lx + d["loss"], meter.avg, rtol=2e-6, what="FixMatch step loss")
        close(d["grad_s"], leaf.grad[B + Bu:], rtol=1e-5, atol=1e-9, what="FixMatch grad")
        arrs[f"s{i}/logits"] = lg.numpy()
        arrs[f"s{i}/targets_x"] = st["targets_x"].numpy()
        arrs[f"s{i}/total_loss"] = np.float64(meter.avg)
        arrs[f"s{i}/grad_logits"] = leaf.grad.numpy()
    arrs["meta"] = np.array(json.dumps(dict(B=B, MU=MU, C=C, thr=thr)))
    np.savez_compressed(out / "fixmatch_train_one.npz", **arrs)


def pin_evaluate_one(ref_fixmatch, ref_utils, out):
    """The REAL ``FixMatch.evaluate_one`` (fixmatch.py:135-178) + ``utils.calculate_metrics`` (utils.py:38-55) on scripted
    validation logits: summary loss, micro / macro precision / recall / F1 and the per-class sensitivity / specificity
    table.  Pins ``oracle.evaluate_one`` and gives the evaluation head its golden vector."""
    import contextlib
    import io
    N, BS = 500, 32                                   # a ragged last batch (20 rows)
    g = torch.Generator().manual_seed(2024)
    y = torch.randint(0, C, (N,), generator=g)
    y[:C] = torch.arange(C)                           # every class present (the reference indexes recall[1] per class)
    logits = 3.0 * torch.nn.functional.one_hot(y, C).float() + 2.0 * torch.randn(N, C, generator=g)
    batches = [(torch.zeros(len(y[i:i + BS]), 1), y[i:i + BS]) for i in range(0, N, BS)]
    model = ScriptedModel([{"logits": logits[i:i + BS]} for i in range(0, N, BS)], with_feats=False)
    tr = ref_fixmatch.FixMatch(model, device="cpu")
    tr.config = ref_utils.AttrDict(DATA=ref_utils.AttrDict(BATCH_SIZE=BS), MODEL=ref_utils.AttrDict(NUM_CLASSES=C),
                                   TRAIN=ref_utils.AttrDict(USE_EMA=False))
    tr.get_dataloader((None, None), _Loader(batches))
    with contextlib.redirect_stderr(io.StringIO()):
        meter, metric = tr.evaluate_one()
    mine = O.evaluate_one([logits[i:i + BS] for i in range(0, N, BS)], [b[1] for b in batches], C, BS)
    close(mine["loss_avg"], meter.avg, rtol=1e-6, what="evaluate_one loss")
    keys = ["micro/precision", "micro/recall", "micro/f1", "macro/precision", "macro/recall", "macro/f1"]
    for k in keys:
        close(mine["metric"][k], metric[k], rtol=1e-12, what=k)
    sen = metric["sen/spec"]
    close(torch.tensor(mine["metric"]["sen/spec"]["sensitivity"].values), torch.tensor(sen["sensitivity"].values), rtol=1e-12, what="sensitivity")
    close(torch.tensor(mine["metric"]["sen/spec"]["specificity"].values), torch.tensor(sen["specificity"].values), rtol=1e-12, what="specificity")
    arrs = {"logits": logits.numpy(), "targets": y.numpy(), "loss_avg": np.float64(meter.avg), "loss_val": np.float64(meter.val),
            "loss_sum": np.float64(meter.sum), "loss_count": np.float64(meter.count),
            "metrics": np.array([metric[k] for k in keys], dtype=np.float64),
            "sensitivity": sen["sensitivity"].values.astype(np.float64), "specificity": sen["specificity"].values.astype(np.float64),
            "confusion": mine["confusion"].astype(np.int64),
            "meta": np.array(json.dumps(dict(N=N, batch_size=BS, C=C, metric_keys=keys)))}
    np.savez_compressed(out / "evaluate_one.npz", **arrs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(REPO / "tests" / "golden"))
    args = ap.parse_args()
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)   # fixed reduction order for the fixtures
    ref_loss, ref_ema, ref_comatch, ref_fixmatch, ref_utils = import_reference()
    pin_fixmatch_criteria(ref_loss, out)
    pin_ema(ref_ema, out)
    pin_comatch(ref_comatch, ref_utils, ref_ema, out)
    pin_fixmatch_train_one(ref_fixmatch, ref_utils, out)
    pin_evaluate_one(ref_fixmatch, ref_utils, out)
    manifest = {}
    for f in sorted(out.glob("*.npz")):
        with np.load(f, allow_pickle=False) as z:
            manifest[f.name] = sha({k: z[k] for k in z.files})
    (out / "MANIFEST.json").write_text(json.dumps(
        {"generator": "oracle/pin_against_reference.py", "torch": torch.__version__,
         "reference": "taindp98/Endoscopy-Image-Classification @ /root/reference (code/loss.py, ema.py, comatch.py, fixmatch.py, utils.py)",
         "sha256": manifest}, indent=1) + "\n")
    print("oracle pinned against the reference; fixtures:", ", ".join(manifest))


if __name__ == "__main__":
    main()
