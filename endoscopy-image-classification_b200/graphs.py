"""CUDA-graph replay of the launch-bound SSL step.

The head is 5-8 tiny launches plus one 300 MB EMA stream; issued eagerly from
Python the step is bound by host launch overhead (~0.4 ms), not by the GPU
(~0.07 ms).  ``GraphedStep`` captures ``step_fn`` (head forward, ``backward()``,
``ModelEMA.update``) once and replays it with a single ``cudaGraphLaunch``.  Every
kernel of ``libb200ssl.so`` is capture safe by construction: no allocation, no
synchronisation, state that changes between steps (bank write pointer, DA history,
ticket counters) lives in device memory.

Two graphs are captured over the same static tensors:

* ``replay()``       -- compute only; inputs already in the static device tensors;
* ``replay_host()``  -- one H2D copy of all inputs from a pinned staging buffer, the
  compute, and a D2H copy of the scalar result into pinned memory (the trainer's
  ``losses.item()``), i.e. the end-to-end step.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Dict, Optional

import torch

__all__ = ["GraphedStep"]


class GraphedStep:
    def __init__(self, step_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], example: Dict[str, torch.Tensor],
                 device, warmup: int = 3, on_replay: Optional[Callable[[], None]] = None,
                 after_capture: Optional[Callable[[], None]] = None, capture_host_io: bool = True,
                 high_priority: bool = False, pre_fn: Optional[Callable[[], None]] = None):
        """``step_fn(static_inputs) -> scalar loss tensor`` must do all its work on the current
        stream.  ``example`` gives shapes/dtypes (and the initial contents) of the inputs.
        ``on_replay`` runs after every replay (host mirrors of device state, e.g.
        ``CoMatchHead.note_graph_replay``); ``after_capture`` runs once after the captures, whose
        host code ran without device work (e.g. ``CoMatchHead.sync_ptr_from_device``).  ``high_priority``: capture on a
        high-priority stream, so the kernel nodes of ``step_fn``'s own stream outrank work it forks onto default-priority
        side streams (an overlapped ``ModelEMA.update``: the head's few CTAs are placed as soon as an SM has room).
        ``pre_fn``: work that does not depend on the inputs (forking that update); it is queued AHEAD of the H2D copy of the
        end-to-end graph, so the copy runs under it.  Only for work whose placement does not have to follow the head's first
        kernel (``ModelEMA.overlap_mode == 'masked'``)."""
        self.device = torch.device(device)
        self.on_replay = None
        # All inputs live in ONE device buffer (256-byte aligned slices) mirrored by ONE pinned host buffer, so the
        # end-to-end graph needs a single H2D copy node instead of one per tensor (each costs ~2 us of stream time).
        offsets, total = {}, 0
        for k, v in example.items():
            offsets[k] = total
            total += (v.numel() * v.element_size() + 255) // 256 * 256
        self._packed_dev = torch.zeros(max(total, 256), dtype=torch.uint8, device=self.device)
        self._packed_host = torch.zeros(max(total, 256), dtype=torch.uint8).pin_memory()

        def carve(buf, k, v):
            nbytes = v.numel() * v.element_size()
            return buf[offsets[k]:offsets[k] + nbytes].view(v.dtype).view(v.shape)

        self.static = {k: carve(self._packed_dev, k, v) for k, v in example.items()}
        self.staging = {k: carve(self._packed_host, k, v) for k, v in example.items()}
        for k, v in example.items():
            self.static[k].copy_(v.detach())
        self._stage_plan = [(k, b.data_ptr(), b.numel() * b.element_size(), b.dtype, b.shape) for k, b in self.staging.items()]
        self.result_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.h2d_bytes = int(self._packed_dev.numel())       # what the H2D node moves per step (tensors + alignment padding)
        self._calls = 0
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                if pre_fn is not None:
                    pre_fn()
                step_fn(self.static)
                self._note()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        cap = {"stream": torch.cuda.Stream(self.device, priority=-1)} if high_priority else {}
        with torch.cuda.graph(self.graph, **cap):
            if pre_fn is not None:
                pre_fn()
            self.result = step_fn(self.static).detach().reshape(1).float()
        # gradients the captured backward writes (static buffers of THIS graph)
        self.grads = {k: v.grad for k, v in self.static.items() if v.grad is not None}
        self._note()
        self.graph_host = None
        if capture_host_io:
            self.graph_host = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_host, pool=self.graph.pool(), **cap):
                if pre_fn is not None:
                    pre_fn()
                with torch.no_grad():
                    self._packed_dev.copy_(self._packed_host, non_blocking=True)      # one H2D node for all inputs
                res = step_fn(self.static).detach().reshape(1).float()
                self.result_host.copy_(res, non_blocking=True)
            self.grads_host = {k: v.grad for k, v in self.static.items() if v.grad is not None}
            self._note()
        torch.cuda.synchronize(self.device)
        if after_capture is not None:
            after_capture()
        self.on_replay = on_replay

    def _note(self):
        self._calls += 1
        if self.on_replay is not None:
            self.on_replay()

    def replay(self) -> torch.Tensor:
        """Inputs are whatever ``self.static`` holds; returns the (static) device scalar."""
        self.graph.replay()
        self._note()
        return self.result

    def replay_host(self, host_batch: Optional[Dict[str, torch.Tensor]] = None) -> float:
        """Stage ``host_batch`` (CPU tensors) into pinned memory, replay H2D + step + D2H,
        wait for the result and return it as a Python float."""
        if self.graph_host is None:
            raise RuntimeError("captured without host I/O")
        if host_batch is not None:
            # ordinary host tensors -> the pinned staging buffer: a plain memmove per tensor (a torch ``copy_`` of a few KB
            # is ~2 us of dispatch; seven of them were an eighth of the end-to-end step)
            for k, dst, nbytes, dtype, shape in self._stage_plan:
                src = host_batch[k]
                if src.dtype == dtype and src.shape == shape and src.device.type == "cpu" and src.is_contiguous():
                    ctypes.memmove(dst, src.data_ptr(), nbytes)
                else:
                    self.staging[k].copy_(src)
        self.graph_host.replay()
        self._note()
        torch.cuda.current_stream(self.device).synchronize()
        return float(self.result_host[0])
