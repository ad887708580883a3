"""Optimizer step + EMA update as one multi-tensor launch.

``FusedOptimizerEMA(optimizer, ema)`` replaces the pair

    self.optimizer.step()                  # code/fixmatch.py:123
    self.ema_model.update(self.model)      # code/fixmatch.py:127

for the ``torch.optim.SGD`` / ``Adam`` / ``AdamW`` instances that ``build_optimizer`` creates
(``code/optimizer.py:43-51``, incl. the no-weight-decay group of ``:13-27``).  The trainable fp32
parameters go through ``b200ssl_opt_ema_multi_tensor`` (read p, g, state, e once; write p, state,
e once); everything else in ``state_dict()`` -- buffers, frozen parameters -- keeps going through
``b200ssl_ema_multi_tensor``.  The wrapped optimizer stays the owner of all state
(``momentum_buffer`` / ``exp_avg`` / ``exp_avg_sq`` / ``step``): ``optimizer.state_dict()``,
``load_state_dict`` and LR schedulers keep working, and the plain ``optimizer.step()`` can be
resumed at any time.

Gradients must stay at fixed addresses: use ``FusedOptimizerEMA.zero_grad()`` (in-place zero) or
``optimizer.zero_grad(set_to_none=False)``.  A gradient that moved is detected and the device
tables are rebuilt.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _native as N
from .ema import ModelEMA, _EmaPlan

__all__ = ["FusedOptimizerEMA", "CapturedFusedStep"]

OPT_SGD, OPT_ADAM, OPT_ADAMW = 0, 1, 2
BLOCK = np.dtype([("param", "<u8"), ("grad", "<u8"), ("s1", "<u8"), ("s2", "<u8"), ("ema", "<u8"), ("count", "<i4"),
                  ("group", "<i4"), ("ema_repeat", "<i4"), ("rsv", "<i4"), ("rsv2", "<i8")])
GROUP = np.dtype([("lr", "<f4"), ("beta1", "<f4"), ("beta2", "<f4"), ("eps", "<f4"), ("weight_decay", "<f4"),
                  ("step_size", "<f4"), ("bias2_sqrt", "<f4"), ("momentum", "<f4"), ("kind", "<i4"), ("nesterov", "<i4"),
                  ("first_step", "<i4"), ("rsv", "<i4"), ("one_minus_beta1", "<f4"), ("one_minus_beta2", "<f4"),
                  ("decay_factor", "<f4"), ("pad", "<f4")])
assert BLOCK.itemsize == 64 and GROUP.itemsize == 64
CHUNK = N.EMA_BLOCK_ELEMS


def _kind_of(opt) -> int:
    if type(opt) is torch.optim.SGD:
        return OPT_SGD
    if type(opt) is torch.optim.AdamW:
        return OPT_ADAMW
    if type(opt) is torch.optim.Adam:
        return OPT_ADAM
    raise NotImplementedError(f"FusedOptimizerEMA supports torch.optim.SGD / Adam / AdamW, not {type(opt).__name__}")


def group_row(kind: int, g: dict, step: int) -> tuple:
    """One ``b200ssl_opt_group`` row for parameter group ``g`` at (1-based) step ``step``: every derived scalar is
    formed in double exactly like torch.optim does and rounded to fp32 once.  Pure host logic (CPU-tested)."""
    lr, wd = float(g["lr"]), float(g.get("weight_decay", 0.0))
    if kind == OPT_SGD:
        return (lr, 0.0, 0.0, 0.0, wd, 0.0, 1.0, float(g.get("momentum", 0.0)), kind, int(bool(g.get("nesterov", False))),
                int(step == 1), 0, 0.0, 0.0, 1.0, 0.0)
    b1, b2 = (float(b) for b in g["betas"])
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    if g.get("decoupled_weight_decay", False):       # torch.optim.Adam(decoupled_weight_decay=True) is AdamW
        kind = OPT_ADAMW
    return (lr, b1, b2, float(g["eps"]), wd, lr / bc1, math.sqrt(bc2), 0.0, kind, 0, int(step == 1), 0, 1.0 - b1, 1.0 - b2,
            1.0 - lr * wd, 0.0)


class FusedOptimizerEMA:
    def __init__(self, optimizer: torch.optim.Optimizer, ema: Optional[ModelEMA] = None, model: Optional[torch.nn.Module] = None):
        self.optimizer, self.ema, self.model = optimizer, ema, model
        if ema is not None and model is None:
            raise ValueError("pass the live model together with its ModelEMA")
        self.kind = _kind_of(optimizer)
        for g in optimizer.param_groups:
            bad = [k for k in ("amsgrad", "maximize", "capturable", "differentiable") if g.get(k)]
            if bad or (self.kind == OPT_SGD and float(g.get("dampening", 0.0)) != 0.0):
                raise NotImplementedError(f"param group options not supported by the fused step: {bad or 'dampening'}")
        if len(optimizer.param_groups) > 8:
            raise NotImplementedError("the fused step takes at most 8 parameter groups")
        self._tables = None          # block table, n_blocks, device, grad pointers, EMA plan of the remaining state
        self._steps = 0
        # torch.optim keeps one CPU `step` tensor per parameter; bumping 161 of them costs more host time than the
        # whole launch, so fused steps are counted here and written back just before anybody can look
        # (state_dict(), a plain optimizer.step(), load_state_dict()).
        self._pending = 0
        optimizer.register_state_dict_pre_hook(lambda opt: self._flush_steps())
        optimizer.register_step_pre_hook(lambda opt, args, kwargs: self._flush_steps())
        optimizer.register_load_state_dict_post_hook(lambda opt: self._invalidate())

    def _flush_steps(self) -> None:
        t = self._tables
        if self._pending and t is not None and t["steps"]:
            torch._foreach_add_(t["steps"], float(self._pending))
            t["base"] = [b + self._pending for b in t["base"]]
        self._pending = 0

    def _invalidate(self) -> None:
        self._pending = 0
        self._tables = None

    # ---- state, owned by the wrapped optimizer (same lazy init as torch.optim) -----------------
    def _state_for(self, p: torch.Tensor, group: dict):
        st = self.optimizer.state[p]
        if self.kind == OPT_SGD:
            if float(group.get("momentum", 0.0)) == 0.0:
                return None, None
            if st.get("momentum_buffer") is None:
                st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["_b200_first"] = True           # torch clones the first gradient into the buffer
            return st["momentum_buffer"], None
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st["exp_avg"], st["exp_avg_sq"]

    def _build(self):
        self._flush_steps()
        params: List[tuple] = []
        for gi, g in enumerate(self.optimizer.param_groups):
            for p in g["params"]:
                if p.grad is None:
                    continue
                N.require_cuda(p, p.grad, what="FusedOptimizerEMA")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise NotImplementedError("the fused step handles dense fp32 parameters")
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise ValueError("the fused step needs contiguous parameters and gradients")
                params.append((gi, g, p))
        if not params:
            raise RuntimeError("no parameter has a gradient: call backward() before the first fused step")
        device = params[0][2].device
        # EMA pairing by storage: parameter storage -> (ema tensor, multiplicity in state_dict())
        ema_of: Dict[int, tuple] = {}
        rest_plan = None
        if self.ema is not None:
            e_vals, m_vals = list(self.ema.ema.state_dict().values()), list(self.model.state_dict().values())
            if len(e_vals) != len(m_vals):
                raise ValueError("EMA / model state_dict mismatch")
            for e, m in zip(e_vals, m_vals):
                if m.data_ptr() in ema_of:
                    ema_of[m.data_ptr()][1] += 1
                else:
                    ema_of[m.data_ptr()] = [e, 1]
        parts, fused_ptrs = [], set()
        for gi, g, p in params:
            s1, s2 = self._state_for(p, g)
            e, rep = ema_of.get(p.data_ptr(), (None, 0))
            if e is not None and (e.dtype != torch.float32 or e.shape != p.shape or not e.is_contiguous()):
                e, rep = None, 0                   # leave unusual entries to the plain EMA launch
            if e is not None:
                fused_ptrs.add(p.data_ptr())
            n = p.numel()
            off = np.arange(0, n, CHUNK, dtype=np.int64)
            part = np.zeros(len(off), dtype=BLOCK)
            byte = (off * 4).astype(np.uint64)
            part["param"] = np.uint64(p.data_ptr()) + byte
            part["grad"] = np.uint64(p.grad.data_ptr()) + byte
            part["s1"] = (np.uint64(s1.data_ptr()) + byte) if s1 is not None else 0
            part["s2"] = (np.uint64(s2.data_ptr()) + byte) if s2 is not None else 0
            part["ema"] = (np.uint64(e.data_ptr()) + byte) if e is not None else 0
            part["count"] = np.minimum(CHUNK, n - off)
            part["group"], part["ema_repeat"] = gi, rep
            parts.append(part)
        tbl = np.concatenate(parts)
        blocks = torch.from_numpy(tbl.view(np.uint8).copy()).to(device)
        if self.ema is not None:
            rest_plan = _EmaPlan(self.ema.ema, self.model, exclude=fused_ptrs)
        # host-side bookkeeping, prepared once: the gradient OBJECTS (kept alive: a gradient that was dropped and
        # re-allocated is a different object), the per-group live parameters and the Adam step counters
        plist = [p for _, _, p in params]
        live = [[p for gi2, _, p in params if gi2 == gi] for gi in range(len(self.optimizer.param_groups))]
        steps = [] if self.kind == OPT_SGD else [self.optimizer.state[p]["step"] for p in plist]
        base = [int(self.optimizer.state[ps[0]]["step"]) if (ps and self.kind != OPT_SGD) else 0 for ps in live]
        if self.kind != OPT_SGD:
            # the bias corrections are per-group scalars: parameters of one group that have been stepped a different number
            # of times (a head that received its first gradient later) cannot share them
            for gi, ps in enumerate(live):
                if any(int(self.optimizer.state[p]["step"]) != base[gi] for p in ps):
                    raise NotImplementedError("Adam: parameters of one group must have been stepped equally often; put "
                                              "late-starting parameters into their own param group")
        self._tables = dict(base=base, blocks=blocks, n_blocks=len(tbl), device=device, rest=rest_plan, params=plist, live=live, steps=steps,
                            grads=[(p, p.grad) for p in plist], ptrs=[(p, p.data_ptr()) for p in (plist[0], plist[-1])])

    def _valid(self) -> bool:
        t = self._tables
        if t is None:
            return False
        if self.kind != OPT_SGD and any(self.optimizer.state[p].get("step") is not s for p, s in zip(t["params"][:1], t["steps"][:1])):
            return False                             # optimizer state was replaced (load_state_dict)
        if not (all(p.grad is g for p, g in t["grads"]) and all(p.data_ptr() == a for p, a in t["ptrs"])):
            return False
        # a trainable parameter that received its first gradient since the tables were built must join them
        return sum(1 for g in self.optimizer.param_groups for p in g["params"] if p.grad is not None) == len(t["params"])

    # ---- public API ---------------------------------------------------------------------------
    def _group_rows(self, t):
        """Per-group scalars of the coming step (the step counters live in the optimizer's state like torch keeps them);
        returns ``(rows, any_first)``."""
        rows = np.zeros(len(self.optimizer.param_groups), dtype=GROUP)
        any_first = False
        for gi, g in enumerate(self.optimizer.param_groups):
            step = 1
            live = t["live"][gi]
            if self.kind == OPT_SGD:
                flags = [self.optimizer.state[p].get("_b200_first", False) for p in live]
                first = any(flags)
                if first and not all(flags):
                    raise NotImplementedError("SGD: parameters of one group must receive their first gradient in the same step")
                any_first |= first
                step = 1 if first else 2
            elif live:
                step = t["base"][gi] + self._pending + 1
            rows[gi] = group_row(self.kind, g, step)
        return rows, any_first

    def _after_step(self, t, any_first) -> None:
        """Bookkeeping torch.optim would have done."""
        if self.kind == OPT_SGD:
            if any_first:
                for p in t["params"]:
                    self.optimizer.state[p].pop("_b200_first", None)
        else:
            self._pending += 1                       # written back to the per-parameter step tensors by _flush_steps()
        self._steps += 1

    @property
    def bytes_per_step(self) -> int:
        """Algorithmic HBM bytes of one fused launch: p, g, state, e read once; p, state, e written once."""
        if not self._valid():
            self._build()
        t, total = self._tables, 0
        for p in t["params"]:
            st = self.optimizer.state[p]
            n_state = 0 if self.kind == OPT_SGD and st.get("momentum_buffer") is None else (1 if self.kind == OPT_SGD else 2)
            total += p.numel() * 4 * (3 + 2 * n_state + 2)   # p r/w, g r, states r/w, ema r/w (parameters without an EMA copy are rare)
        return total

    @torch.no_grad()
    def step(self) -> None:
        if not self._valid():
            self._build()
        t = self._tables
        rows, any_first = self._group_rows(t)
        d = self.ema.decay if self.ema is not None else 0.0
        N.check(N.lib().b200ssl_opt_ema_multi_tensor(t["blocks"].data_ptr(), t["n_blocks"], rows.ctypes.data,
                                                     len(rows), float(np.float32(d)), float(np.float32(1.0 - d)),
                                                     N.stream_ptr(t["device"])), "opt_ema_multi_tensor")
        if t["rest"] is not None and t["rest"].n_blocks > 0:
            t["rest"].launch(self.ema.decay, 0)
        self._after_step(t, any_first)

    @torch.no_grad()
    def capture(self) -> "CapturedFusedStep":
        """The fused step as a CUDA graph (the step of a captured training loop): the group rows live in device memory and
        are refreshed by ``CapturedFusedStep.replay()`` with one 64-byte-per-group copy ahead of the graph launch, so the
        learning-rate schedule and Adam's bias corrections keep changing while the launches are replayed."""
        if not self._valid():
            self._build()
        return CapturedFusedStep(self)

    def zero_grad(self) -> None:
        """In-place zero of the gradients (their addresses are baked into the device table)."""
        grads = [p.grad for g in self.optimizer.param_groups for p in g["params"] if p.grad is not None]
        if grads:
            torch._foreach_zero_(grads)

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd) -> None:
        self.optimizer.load_state_dict(sd)      # the post hook drops the tables: state tensors were re-created


class CapturedFusedStep:
    """``optimizer.step(); ema.update(model)`` as one replayed graph node sequence.  ``replay()`` == ``FusedOptimizerEMA.step()``
    bit for bit (same kernel, same scalars); the tables are those of the owner at capture time -- a gradient that moves
    afterwards invalidates the graph (``replay`` raises)."""
    RING = 64

    def __init__(self, owner: FusedOptimizerEMA):
        self.owner = owner
        t = owner._tables
        self.tables = t
        dev = t["device"]
        G = len(owner.optimizer.param_groups)
        self._host = torch.zeros(self.RING, G * GROUP.itemsize, dtype=torch.uint8).pin_memory()
        self._host_np = self._host.numpy()
        self._dev = torch.zeros(G * GROUP.itemsize, dtype=torch.uint8, device=dev)
        self._slot = 0
        self._guard = [None] * self.RING           # event recorded after the copy out of a pinned slot
        d = owner.ema.decay if owner.ema is not None else 0.0
        rows, _ = owner._group_rows(t)
        self._dev.copy_(torch.from_numpy(rows.view(np.uint8).copy()))
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # the capture only records the launches: parameters, optimizer state and EMA are NOT stepped by it
        with torch.cuda.graph(self.graph):
            N.check(N.lib().b200ssl_opt_ema_multi_tensor_dev(t["blocks"].data_ptr(), t["n_blocks"], self._dev.data_ptr(), G,
                                                             float(np.float32(d)), float(np.float32(1.0 - d)),
                                                             N.stream_ptr(dev)), "opt_ema_multi_tensor_dev")
            if t["rest"] is not None and t["rest"].n_blocks > 0:
                t["rest"].launch(owner.ema.decay, 0)

    @torch.no_grad()
    def replay(self) -> None:
        o = self.owner
        if o._tables is not self.tables or not o._valid():
            raise RuntimeError("the captured fused step is stale (gradients or optimizer state were re-allocated): capture() again")
        rows, any_first = o._group_rows(self.tables)
        i = self._slot
        if self._guard[i] is not None:
            self._guard[i].synchronize()             # the copy that last read this pinned slot has run (64 replays ago)
        self._host_np[i, :] = rows.view(np.uint8)
        self._dev.copy_(self._host[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._guard[i] = ev
        self._slot = (i + 1) % self.RING
        self.graph.replay()
        o._after_step(self.tables, any_first)
