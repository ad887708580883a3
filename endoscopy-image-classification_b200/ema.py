"""Drop-in for the reference ``code/ema.py`` (``ModelEMA``, lines 40-62).

``update`` / ``set`` run ONE multi-tensor sm_100a kernel
(``b200ssl_ema_multi_tensor``) over a device-resident block table instead of
the reference's per-tensor ``mul, mul, add, copy_`` loop (``ema.py:51-56``),
with bit-identical results:

* every ``state_dict()`` entry is covered -- parameters AND buffers, integer
  buffers computed in fp32 and truncated (``ema.py:56`` ``copy_``; quirk Q3);
* a storage that appears under several names (``ModelwEmb`` registers the
  backbone three times, ``models/custom_model.py:194-200``) is updated as many
  times as it appears, in registers (quirk Q2);
* mul, mul, add are rounded separately (no FMA), with ``decay`` and
  ``1.-decay`` rounded to fp32 after the double subtraction.
"""
from __future__ import annotations

import ctypes as C
import os
from copy import deepcopy
from typing import List, Optional

import numpy as np
import torch

from . import _native as N

__all__ = ["ModelEMA"]

_FLOAT_ENUMS = (N.F32, N.BF16, N.F16)


BLOCK_DTYPE = np.dtype([("ema", "<u8"), ("model", "<u8"), ("count", "<i4"), ("dtype", "<i4"),
                        ("repeat", "<i4"), ("rsv", "<i4")])
assert BLOCK_DTYPE.itemsize == C.sizeof(N.EmaBlock)


def build_block_table(entries, block_elems: int = N.EMA_BLOCK_ELEMS) -> np.ndarray:
    """Chunk ``(ema_ptr, model_ptr, numel, elem_size, dtype_enum, repeat)`` entries into
    ``b200ssl_ema_block`` rows of at most ``block_elems`` elements (pure host logic)."""
    parts = []
    for e_ptr, m_ptr, n, es, dt, rep in entries:
        if n <= 0:
            continue
        off = np.arange(0, n, block_elems, dtype=np.int64)
        part = np.zeros(len(off), dtype=BLOCK_DTYPE)
        part["ema"] = np.uint64(e_ptr) + (off * es).astype(np.uint64)
        part["model"] = np.uint64(m_ptr) + (off * es).astype(np.uint64)
        part["count"] = np.minimum(block_elems, n - off)
        part["dtype"], part["repeat"] = dt, rep
        parts.append(part)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=BLOCK_DTYPE)


class _EmaPlan:
    """Device block table for one (ema module, model) pair."""

    def __init__(self, ema_mod: torch.nn.Module, model: torch.nn.Module, exclude=()):
        """``exclude``: model storages (``data_ptr``) that another launch updates (``fused_step.FusedOptimizerEMA``)."""
        e_vals = list(ema_mod.state_dict().values())
        m_vals = list(model.state_dict().values())
        if len(e_vals) != len(m_vals):
            raise ValueError(f"EMA has {len(e_vals)} state entries, model has {len(m_vals)}")
        uniq = {}                      # ema data_ptr -> [e, m, repeat]
        for e, m in zip(e_vals, m_vals):
            dev = N.require_cuda(e, m, what="ModelEMA")
            if e.shape != m.shape or e.dtype != m.dtype:
                raise ValueError(f"EMA/model entry mismatch: {tuple(e.shape)}/{e.dtype} vs {tuple(m.shape)}/{m.dtype}")
            if not (e.is_contiguous() and m.is_contiguous()):
                raise ValueError("ModelEMA needs contiguous state tensors")
            if e.numel() == 0 or m.data_ptr() in exclude:
                continue
            key = e.data_ptr()
            if key in uniq:
                if uniq[key][1].data_ptr() != m.data_ptr():
                    raise ValueError("EMA storage aliasing differs from the model's")
                uniq[key][2] += 1
            else:
                uniq[key] = [e, m, 1]
        self.device = dev
        self.n_entries = len(e_vals)
        self.n_unique = len(uniq)
        self.unique_elems = 0
        self.bytes_per_update = 0
        entries = []
        float_dtypes = set()
        self.has_ints = False
        for e, m, rep in uniq.values():
            dt = N.dtype_enum(e)
            if dt in _FLOAT_ENUMS:
                float_dtypes.add(dt)
            else:
                self.has_ints = True
            n, es = e.numel(), e.element_size()
            self.unique_elems += n
            self.bytes_per_update += 3 * n * es
            entries.append((e.data_ptr(), m.data_ptr(), n, es, dt, rep))
        tbl = build_block_table(entries)
        self.n_blocks = len(tbl)
        self.table = torch.from_numpy(tbl.view(np.uint8).copy()).to(self.device)
        self.float_dtypes: List[int] = sorted(float_dtypes) or [N.F32]
        # cheap identity check for later calls: the live first / last parameters and
        # buffers of both modules must still sit where the table says they do
        self._model_id = id(model)
        self._sentinels = []
        for mod in (model, ema_mod):
            ps, bs = list(mod.parameters()), list(mod.buffers())
            self._sentinels += [(t, t.data_ptr()) for t in (ps[:1] + ps[-1:] + bs[:1] + bs[-1:])]

    def matches(self, model: torch.nn.Module) -> bool:
        if id(model) != self._model_id:
            return False
        return all(t.data_ptr() == p for t, p in self._sentinels)

    def launch(self, decay: float, mode: int, max_ctas: int = 0, masked=None) -> None:
        """``max_ctas`` > 0: cap the grid (``b200ssl_ema_multi_tensor_ctas``); ``masked = (sm_mask, sched)``: leave the marked
        SMs alone (``b200ssl_ema_multi_tensor_masked``)."""
        d32 = float(np.float32(decay))
        o32 = float(np.float32(1.0 - decay))
        lib, st = N.lib(), N.stream_ptr(self.device)
        for i, fd in enumerate(self.float_dtypes):
            ints = 1 if (i == 0 and self.has_ints) else 0
            if masked is not None:
                N.check(lib.b200ssl_ema_multi_tensor_masked(self.table.data_ptr(), self.n_blocks, fd, ints, d32, o32, mode,
                                                            masked[0].data_ptr(), masked[1].data_ptr(), st), "ema_multi_tensor_masked")
            elif max_ctas > 0:
                N.check(lib.b200ssl_ema_multi_tensor_ctas(self.table.data_ptr(), self.n_blocks, fd, ints, d32, o32, mode, int(max_ctas), st),
                        "ema_multi_tensor_ctas")
            else:
                N.check(lib.b200ssl_ema_multi_tensor(self.table.data_ptr(), self.n_blocks, fd, ints, d32, o32, mode, st),
                        "ema_multi_tensor")


class ModelEMA(object):
    """Same constructor, attributes and methods as ``code/ema.py:40-62``.

    ``revalidate_every``: the block table caches raw pointers of ``model``'s
    state tensors; every call checks two sentinel pointers (O(1)) and every
    ``revalidate_every``-th call rebuilds the table from ``state_dict()`` so a
    re-allocated parameter is picked up.  Call ``refresh()`` after replacing
    parameters by hand (``load_state_dict`` copies in place and needs nothing).
    """

    OVERLAP_MODE = "capped"            # 'masked': the update leaves a probed set of SMs alone (b200ssl_ema_multi_tensor_masked)
    OVERLAP_FREE_CLUSTERS = 6          # ... of this many clusters of 8 SMs
    OVERLAP_DELAY_NS = 500             # head start (ns) given to the kernel queued next to an overlapped update (0: 2-3 slow blocks of 25; >= 2500: the whole PDL-chained head is resident first and the update starves, 90 us)
    OVERLAP_CTAS = 4 * (148 - 48)      # grid of an overlapped update: 4 CTAs per SM on all but 48 SMs (see __init__; 336..464 measured, bench cfg 2)

    def __init__(self, model, decay=0.9999, device=None, revalidate_every: int = 1024, overlap: bool = False,
                 overlap_ctas: Optional[int] = None):
        """``overlap=True``: ``update`` is launched on a side stream forked from the current one, so the 300 MB weight
        stream runs next to whatever the caller queues afterwards (the next step's forward / SSL head); ``join()`` makes the
        current stream wait for it and has to be called before the model's weights are written again (the trainer does so
        ahead of ``optimizer.step()``), before the EMA weights are read, and before the end of a CUDA-graph capture.
        The grid of an overlapped update is capped at ``overlap_ctas`` (default ``OVERLAP_CTAS``): its CTAs are persistent and
        four of them fill an SM, so the SMs that the kernel launched just ahead of it holds (the head's first kernel, on a
        higher-priority stream) are never handed to the update and stay free for the rest of the head, whose tensor-core
        kernels need whole SMs."""
        super(ModelEMA, self).__init__()
        self.ema = deepcopy(model)
        self.ema.eval()
        self.decay = decay
        self.device = device
        if self.device is not None:
            self.ema.to(device=device)
        self._plan: Optional[_EmaPlan] = None
        self._calls = 0
        self._revalidate_every = int(revalidate_every)
        self.overlap = bool(overlap)
        self.overlap_ctas = int(self.OVERLAP_CTAS if overlap_ctas is None else overlap_ctas)
        self.overlap_delay_ns = int(os.environ.get("B200SSL_EMA_DELAY_NS", self.OVERLAP_DELAY_NS))
        self.overlap_mode = os.environ.get("B200SSL_EMA_OVERLAP_MODE", self.OVERLAP_MODE)      # 'masked' | 'capped'
        self._masked = None
        self.free_clusters = int(os.environ.get("B200SSL_EMA_FREE_CLUSTERS", self.OVERLAP_FREE_CLUSTERS))
        if self.overlap:
            # the head's launch planners keep to the SMs the capped update leaves free (process-wide setting)
            N.lib().b200ssl_set_head_sm_budget(max(148 - self.overlap_ctas // 4, 8))
        self._side: Optional[torch.cuda.Stream] = None
        self._pending = False

    def refresh(self) -> None:
        self._plan = None

    def _get_plan(self, model) -> _EmaPlan:
        self._calls += 1
        if self._plan is not None and torch.cuda.is_current_stream_capturing():
            return self._plan                  # CUDA-graph capture: no table rebuild (it would synchronise)
        if (self._plan is None or not self._plan.matches(model)
                or (self._revalidate_every > 0 and self._calls % self._revalidate_every == 0)):
            with torch.no_grad():
                self._plan = _EmaPlan(self.ema, model)
        return self._plan

    def _update(self, model, update_fn):
        """Generic per-tensor path of ``ema.py:51-56`` for an arbitrary ``update_fn``
        (device eager ops; ``update`` / ``set`` below do not use it)."""
        with torch.no_grad():
            for ema_v, model_v in zip(self.ema.state_dict().values(), model.state_dict().values()):
                N.require_cuda(ema_v, model_v, what="ModelEMA._update")
                ema_v.copy_(update_fn(ema_v, model_v))

    def update(self, model):
        """``ema.py:58-59``: e <- decay*e + (1-decay)*m for every state entry."""
        plan = self._get_plan(model)
        if not self.overlap:
            plan.launch(self.decay, 0)
            return
        dev = plan.device
        if self._side is None:
            self._side = torch.cuda.Stream(dev)                 # default (low) priority: the head's stream may be created above it
        cur = torch.cuda.current_stream(dev)
        self._side.wait_stream(cur)                               # after everything queued so far (the optimizer step)
        if self.overlap_mode == "masked" and self._masked is None and not torch.cuda.is_current_stream_capturing():
            self._probe_sm_set(dev)
        with torch.cuda.stream(self._side):
            if self._masked is not None:                          # a fixed set of SMs left alone: placement order does not matter
                plan.launch(self.decay, 0, masked=self._masked)
            else:
                if self.overlap_delay_ns > 0:                     # the kernel queued next on the caller's stream places its CTAs first
                    N.check(N.lib().b200ssl_stream_delay(self.overlap_delay_ns, N.stream_ptr(dev)), "stream_delay")
                plan.launch(self.decay, 0, max_ctas=self.overlap_ctas)
        self._pending = True

    def _probe_sm_set(self, dev) -> None:
        """Find ``OVERLAP_FREE_CLUSTERS`` x 8 SMs in which that many clusters of 8 one-CTA-per-SM blocks fit side by side (the
        head's tensor-core kernels run as such clusters) and keep them free of the update from now on."""
        mask = torch.zeros(8, dtype=torch.int32, device=dev)
        scratch = torch.zeros(1, dtype=torch.int32, device=dev)
        N.check(N.lib().b200ssl_probe_sm_set(self.free_clusters, 8, mask.data_ptr(), scratch.data_ptr(), N.stream_ptr(dev)),
                "probe_sm_set")
        bits = sum(bin(int(w) & 0xFFFFFFFF).count("1") for w in mask.tolist())       # synchronises (once)
        if bits == 8 * self.free_clusters:
            self._masked = (mask, torch.zeros(2, dtype=torch.int32, device=dev))
            N.lib().b200ssl_set_head_sm_budget(bits)
        else:                                                     # e.g. a busy device: keep the capped-grid form
            self.overlap_mode = "capped"

    def join(self):
        """Make the current stream wait for an overlapped ``update`` (no-op otherwise)."""
        if self._pending and self._side is not None:
            torch.cuda.current_stream(self._side.device).wait_stream(self._side)
            self._pending = False

    def set(self, model):
        """``ema.py:61-62``: e <- m."""
        self.join()
        self._get_plan(model).launch(self.decay, 1)

    # introspection used by bench.py / tests
    @property
    def plan(self) -> Optional[_EmaPlan]:
        return self._plan
