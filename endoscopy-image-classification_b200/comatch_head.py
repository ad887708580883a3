"""The CoMatch unlabeled head (reference ``code/comatch.py:162-220``, inline in
``CoMatch.train_one``) as a stateful device-side object.

State mirrors ``code/comatch.py:90-96``: the memory bank ``queue_feats [K, D]`` /
``queue_probs [K, C]``, the write pointer ``queue_ptr`` and the distribution-
alignment history (``prob_list``; kept on the device as a ``[window, C]`` ring
so a step has no host synchronisation).

A bf16 step at the reference's size is four launches of ``libb200ssl.so`` (three in forward, one in backward):

====  ===================================  ==========================================
K3    ``b200ssl_bank_smooth_partial``      rowsum / numer of exp(F_w Q_f^T / tau) Q_p                         (:180-181)
rows  ``b200ssl_comatch_rows_fused``       DA history + prob_avg (:167-173), DA divide, alpha-mix, max/mask, focal
                                           soft-CE + grad (:174-176,182-185,216-220), ring enqueue (:187-196)
K6    ``b200ssl_contrast_fwd`` / ``_bwd``  graph-contrastive loss (:199-213), LAMBDA-weighted total; its gradient
====  ===================================  ==========================================

Larger batches / more than 32 classes use the un-fused row kernels (``b200ssl_comatch_da``,
``b200ssl_comatch_finalize``, ``b200ssl_bank_enqueue``).

With ``process_group`` given the ring spans the ranks of one node; ``exchange`` selects whether it
lives in NVLink peer memory (write-through copies or shards read in place, ``peer.py``) or is sharded behind
own / NCCL row exchanges (``bank.py``).
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from . import _native as N
from .bank import ShardGeometry, all_gather_rows, reduce_scatter_rows

def C_byref(struct):
    """ctypes pointer to an optional struct (None -> NULL)."""
    import ctypes
    return None if struct is None else ctypes.byref(struct)


__all__ = ["CoMatchHead", "lockstep_total_loss"]


class _HeadFn(torch.autograd.Function):
    """Outputs: loss_u, loss_contrast, mask_mean, total (= lambda_u*loss_u + lambda_c*loss_contrast,
    comatch.py:222 without loss_x), then the non-differentiable mask, lbs, scores, probs."""

    @staticmethod
    def forward(ctx, head, lambda_u, lambda_c, logits_u_w, logits_u_s0, feats_u_w, feats_u_s0, feats_u_s1, feats_x,
                targets_x, smooth=None):
        f0, f1 = feats_u_s0.detach().contiguous(), feats_u_s1.detach().contiguous()
        out = head._step(logits_u_w.detach(), logits_u_s0.detach(), feats_u_w.detach(), f0, f1,
                         feats_x.detach(), targets_x, float(lambda_u), float(lambda_c), smooth)
        ctx.head, ctx.lambda_u, ctx.lambda_c = head, float(lambda_u), float(lambda_c)
        ctx.save_for_backward(out["grad_s0"], f0, f1, out["probs"], out["stats"], out["probs_hl"])
        ctx.done = False
        ctx.set_materialize_grads(False)
        aux = (out["mask"], out["lbs"], out["scores"], out["probs"])
        ctx.mark_non_differentiable(*aux)
        sc = out["scalars"]
        return (sc[0], sc[2], sc[1], sc[3]) + aux

    @staticmethod
    def backward(ctx, g_u, g_c, g_mm, g_total, *unused):
        if ctx.done:
            raise RuntimeError("CoMatchHead: backward through the fused head twice (the stashed gradient is consumed)")
        ctx.done = True
        grad_s0, f0, f1, probs, stats, probs_hl = ctx.saved_tensors
        head = ctx.head

        def upstream(g_own, lam):
            """(device scalar, host factor) with d(loss)/d(own) = g_own + lam * g_total."""
            if g_total is None:
                return g_own, 1.0
            if g_own is None:
                return g_total, lam
            return g_own + lam * g_total, 1.0

        gs0 = gf0 = gf1 = None
        gu, fac_u = upstream(g_u, ctx.lambda_u)
        gc, fac_c = upstream(g_c, ctx.lambda_c)
        want_s0 = ctx.needs_input_grad[4]
        want_f = ctx.needs_input_grad[6] or ctx.needs_input_grad[7]
        if want_f and gc is not None:
            # one launch: contrastive gradients + (piggy-backed) scaling of the stashed focal-CE gradient
            piggy = (grad_s0, gu, fac_u) if (want_s0 and gu is not None) else None
            gf0, gf1 = head._k_contrast_bwd(f0, f1, probs, stats, gc, fac_c, probs_hl, scale=piggy)
            if piggy is not None:
                gs0, want_s0 = grad_s0, False
        elif want_f:
            gf0, gf1 = torch.zeros_like(f0), torch.zeros_like(f1)
        if want_s0:
            gs0 = torch.zeros_like(grad_s0) if gu is None else head._k_scale(grad_s0, gu, fac_u)
        return None, None, None, None, gs0, None, gf0, gf1, None, None, None


class CoMatchHead:
    """Device-resident CoMatch head state + fused step.

    Parameters mirror the attributes of the reference trainer
    (``comatch.py:29-39, 90-96``): ``alpha=0.9``, ``temperature=0.2``,
    ``contrast_th=0.8``, ``gamma=2``; ``queue_size`` = K (global rows).

    ``enqueue_mode``: ``'reference'`` keeps the guard ``n == queue_size`` of
    ``comatch.py:192`` (quirk Q1: with ``queue_batch=5`` the bank is never
    written); ``'always'`` is the ring write without the guard (upstream CoMatch).

    Up to ``FUSED_ROWS_MAX`` unlabeled rows (classes <= 32) the row phase (DA + finalize + single-GPU
    enqueue) is ONE thread-block-cluster launch; set ``fuse_rows = False`` to force the three separate
    kernels (larger batches always use them).  ``exchange``: see ``__init__`` and DESIGN section 5.
    """
    FUSED_ROWS_MAX = 2048
    REPLICATE_MAX_BYTES = 256 << 20      # 'auto' keeps a full copy of the ring per rank up to this size

    def __init__(self, num_classes: int, low_dim: int, queue_size: int, thr: float, *, alpha: float = 0.9,
                 temperature: float = 0.2, contrast_th: float = 0.8, gamma: float = 2.0, da_window: int = 32,
                 enqueue_mode: str = "reference", smoothing: bool = True, device="cuda",
                 dtype: torch.dtype = torch.float32, process_group=None, exchange: str = "auto", local_ranks=None):
        if enqueue_mode not in ("reference", "always"):
            raise ValueError(enqueue_mode)
        if exchange not in ("auto", "replicated", "direct", "peer", "collective"):
            raise ValueError(exchange)
        self.num_classes, self.low_dim, self.queue_size = int(num_classes), int(low_dim), int(queue_size)
        self.thr, self.alpha, self.temperature = float(thr), float(alpha), float(temperature)
        self.contrast_th, self.gamma, self.da_window = float(contrast_th), float(gamma), int(da_window)
        self.enqueue_mode, self.smoothing = enqueue_mode, bool(smoothing)
        self.device = torch.device(device)
        self._check_backend()
        self.pg = process_group
        world, rank = 1, 0
        # local_ranks = (peer.LocalArenaSet, rank): this head is one of several emulated ranks of ONE process on ONE GPU
        # (single-GPU validation of the peer-memory bank; drive the heads with lockstep_total_loss)
        self._local = local_ranks
        if local_ranks is not None:
            if process_group is not None:
                raise ValueError("process_group and local_ranks exclude each other")
            world, rank = int(local_ranks[0].world), int(local_ranks[1])
        elif process_group is not None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        self.geom = ShardGeometry(self.queue_size, world, rank)
        # How the bank of a multi-rank job (world > 1) is kept:
        #   'replicated' every rank keeps the whole ring in NVLink peer memory; the only exchange of a step is the
        #                enqueue block, written through into every copy by the row kernel (two epoch flags per step).
        #                K3 never leaves local memory -- the fastest layout while the ring fits (REPLICATE_MAX_BYTES)
        #   'direct'     the shards live in NVLink peer memory; K3 reads every shard in place and the enqueue stores
        #                into the owning shard -- same launches as one GPU, two epoch flags per step (bf16 banks with
        #                the tensor-core layout, <= 8 ranks of one node)
        #   'peer'       all-gather / reduce-scatter / all-gather by own kernels over peer memory (peer.py)
        #   'collective' the same three exchanges through torch.distributed (NCCL; gloo in the CPU tests)
        #   'auto'       replicated / direct when the bank qualifies (bf16 tensor-core layout), else peer; collective on CPU
        self._exchange_req = exchange
        self.exchange = "collective"
        self._arena = None
        self._shards = None
        self._alloc_bank(dtype)
        # write pointer: device-resident {ptr, ticket} (graph-replay safe) + host mirror
        self.ptr_state = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._queue_ptr = 0
        self.da_ring = torch.zeros(self.da_window, self.num_classes, dtype=torch.float32, device=self.device)
        self.da_state = torch.zeros(2, dtype=torch.int32, device=self.device)   # count, head
        self.prob_avg = torch.empty(self.num_classes, dtype=torch.float32, device=self.device)
        self._pristine = True
        self.fuse_rows = True
        self._presmoothed = None
        self.last = {}

    def _check_backend(self) -> None:
        if self.device.type != "cuda":
            raise RuntimeError("CoMatchHead needs a CUDA device: the head is CUDA-only, there is no CPU path")
        N.lib()

    def _alloc_bank(self, dtype) -> None:
        self.dtype = dtype
        Ks, C, D, R = self.geom.shard_rows, self.num_classes, self.low_dim, self.geom.world_size
        req, on_gpu = self._exchange_req, self.device.type == "cuda"
        self.queue_feats = self.queue_probs = self.queue_probs_t = self._shards = None
        if self._arena is not None:
            self._arena.close()
            self._arena = None
        direct_ok = (R > 1 and R <= 8 and on_gpu and dtype == torch.bfloat16 and C <= 31 and D == 64 and Ks % 8 == 0
                     and self.smoothing and self.enqueue_mode == "always")    # the epoch flags need one enqueue per step
        if req in ("direct", "replicated") and not direct_ok:
            raise ValueError(f"exchange={req!r} needs 2..8 CUDA ranks, a bf16 bank, low_dim 64, <= 31 classes, shard rows % 8 == 0, "
                             "smoothing and enqueue_mode='always'")
        small = self.queue_size * (D + C + 32) * 2 <= self.REPLICATE_MAX_BYTES
        self.exchange = ((("replicated" if small else "direct") if direct_ok else "peer" if on_gpu else "collective")
                         if req == "auto" else req)
        if self._local is not None and self.exchange not in ("direct", "replicated"):
            raise ValueError("emulated ranks (local_ranks) run the peer-memory bank only: exchange 'direct' or 'replicated'")
        if R > 1 and self.exchange in ("direct", "replicated"):
            rep = self.exchange == "replicated"
            Ks = self.queue_size if rep else Ks                      # rows held by this rank
            import ctypes
            import torch.distributed as dist
            from .peer import PeerArena
            named = {"qf": Ks * D * 2, "qp": Ks * C * 2, "qpt": 32 * Ks * 2}
            self._arena = a = (PeerArena(self.pg, self.device, {}, named=named) if self._local is None else
                               self._local[0].arena(self.geom.rank, {}, named))
            self.queue_feats = a.tensor("qf", (Ks, D), dtype)            # zero-filled by the allocation
            self.queue_probs = a.tensor("qp", (Ks, C), dtype)
            self.queue_probs_t = a.tensor("qpt", (32, Ks), dtype)
            self.queue_probs_t[C].fill_(1.0)
            self._shards = N.BankShards(R, self.geom.rank, Ks, ctypes.addressof(a.bases_host), a.bases.data_ptr(),
                                        a.named_offset["qf"], a.named_offset["qp"], a.named_offset["qpt"], 1 if rep else 0, 0)
            torch.cuda.synchronize(self.device)
            if self._local is None:
                dist.barrier(group=self.pg)      # the ones row of every shard is in place before any peer reads it
            return
        self.queue_feats = torch.zeros(self.geom.shard_rows, self.low_dim, dtype=dtype, device=self.device)
        self.queue_probs = torch.zeros(self.geom.shard_rows, self.num_classes, dtype=dtype, device=self.device)
        # transposed, class-padded copy [32, K_local] for the tensor-core smoothing kernel (bf16 banks)
        self.queue_probs_t = None
        if dtype == torch.bfloat16 and self.num_classes <= 31 and self.low_dim == 64 and self.geom.shard_rows % 8 == 0:
            self.queue_probs_t = torch.zeros(32, self.geom.shard_rows, dtype=dtype, device=self.device)
            self.queue_probs_t[self.num_classes].fill_(1.0)      # "ones" row: the second MMA also yields the row sums

    # ---- reference-visible state ------------------------------------------------
    @property
    def queue_ptr(self) -> int:
        """``self.queue_ptr`` of the reference (comatch.py:94).  Host mirror of the device pointer;
        it advances deterministically, so reading it never synchronises."""
        return self._queue_ptr

    @queue_ptr.setter
    def queue_ptr(self, value: int) -> None:
        self._queue_ptr = int(value) % self.queue_size
        self.ptr_state.copy_(torch.tensor([self._queue_ptr, 0], dtype=torch.int64))

    def sync_ptr_from_device(self) -> int:
        """Re-read the device pointer into the host mirror (synchronises; use after graph capture)."""
        self._queue_ptr = int(self.ptr_state[0].item())
        return self._queue_ptr

    def note_graph_replay(self, n_per_rank: int) -> None:
        """Advance the host mirror after a CUDA-graph replay of a step that enqueued."""
        self._queue_ptr = self.geom.next_ptr(self._queue_ptr, n_per_rank)

    @property
    def prob_list(self) -> List[torch.Tensor]:
        """The DA history as the reference keeps it (``comatch.py:96``): oldest first."""
        count, head = (int(v) for v in self.da_state.tolist())
        return [self.da_ring[(head - count + a) % self.da_window].clone() for a in range(count)]

    @prob_list.setter
    def prob_list(self, entries) -> None:
        entries = list(entries)[-self.da_window:]
        self.da_ring.zero_()
        for i, e in enumerate(entries):
            self.da_ring[i].copy_(e.to(self.device, torch.float32))
        self.da_state.copy_(torch.tensor([len(entries), len(entries) % self.da_window], dtype=torch.int32))

    def _holds_whole_ring(self) -> bool:
        return self.geom.world_size == 1 or self.exchange == "replicated"

    def state_dict(self, full: bool = False) -> dict:
        """``full=True``: the whole ring (collective in a multi-rank job with a sharded bank: an all-gather of the
        shards, rank-major = ring order); otherwise the rows this rank holds."""
        qf, qp = self.queue_feats, self.queue_probs
        if full and not self._holds_whole_ring():
            if self._local is not None:
                raise RuntimeError("emulated ranks: gather the shards of the heads yourself")
            qf, qp = all_gather_rows(qf, self.pg), all_gather_rows(qp, self.pg)
        return {"queue_feats": qf, "queue_probs": qp, "queue_ptr": self.queue_ptr, "da_ring": self.da_ring, "da_state": self.da_state}

    def load_state_dict(self, sd: dict) -> None:
        """Accepts the rows this rank holds or the whole ring (a sharded bank then keeps its own slice)."""
        if sd["queue_feats"].dtype != self.dtype:          # otherwise copy in place: captured graphs
            self._alloc_bank(sd["queue_feats"].dtype)      # keep pointing at the same storage
        qf, qp = sd["queue_feats"], sd["queue_probs"]
        if qf.shape[0] == self.queue_size and not self._holds_whole_ring():
            lo, hi = self.geom.shard_begin, self.geom.shard_begin + self.geom.shard_rows
            qf, qp = qf[lo:hi], qp[lo:hi]
        self.queue_feats.copy_(qf)
        self.queue_probs.copy_(qp)
        if self.queue_probs_t is not None:
            self.queue_probs_t[: self.num_classes].copy_(self.queue_probs.t())
        self.queue_ptr = int(sd["queue_ptr"])
        self.da_ring.copy_(sd["da_ring"])
        self.da_state.copy_(sd["da_state"])
        self._pristine = False

    # ---- the step -----------------------------------------------------------------
    def __call__(self, logits_u_w, logits_u_s0, feats_u_w, feats_u_s0, feats_u_s1, feats_x, targets_x, smooth=None):
        """Returns ``(loss_u, loss_contrast, mask_mean, mask, lbs_u_guess, scores, probs)``;
        the two losses carry grad w.r.t. ``logits_u_s0`` / ``feats_u_s0`` / ``feats_u_s1``.
        ``smooth``: the per-step gate of comatch.py:179 (``epoch > 0 or batch_idx > queue_batch``);
        ``None`` = the head's ``smoothing`` attribute."""
        o = _HeadFn.apply(self, 1.0, 1.0, logits_u_w, logits_u_s0, feats_u_w, feats_u_s0, feats_u_s1, feats_x, targets_x,
                          smooth)
        return o[:3] + o[4:]

    def total_loss(self, logits_u_w, logits_u_s0, feats_u_w, feats_u_s0, feats_u_s1, feats_x, targets_x,
                   lambda_u: float = 1.0, lambda_c: float = 1.0, smooth=None):
        """``LAMBDA_U*loss_u + LAMBDA_C*loss_contrast`` (the unlabeled part of comatch.py:222) formed on
        the device by the last forward kernel, with both weights folded into the backward launches --
        no eager elementwise kernels around the head.  Returns ``(total, loss_u, loss_contrast, mask_mean)``."""
        o = _HeadFn.apply(self, lambda_u, lambda_c, logits_u_w, logits_u_s0, feats_u_w, feats_u_s0, feats_u_s1,
                          feats_x, targets_x, smooth)
        return o[3], o[0].detach(), o[1].detach(), o[2].detach()

    def presmooth(self, feats_u_w) -> None:
        """Phase 1 of a lock-step step over emulated ranks: K3 of this rank now, consumed by its next step."""
        if self._shards is None:
            raise RuntimeError("presmooth is the lock-step protocol of a peer-memory bank")
        self._presmoothed = self._k_smooth(feats_u_w.detach().contiguous())

    def _step(self, lw, ls0, fw, fs0, fs1, fx, tx, lambda_u: float = 1.0, lambda_c: float = 1.0, smooth=None):
        lw, ls0, fw, fx = (t.contiguous() for t in (lw, ls0, fw, fx))
        tx = tx.to(torch.int64).contiguous()
        rows, C = lw.shape
        D, n_x = fw.shape[1], fx.shape[0]
        if C != self.num_classes or D != self.low_dim:
            raise ValueError(f"head built for C={self.num_classes}, D={self.low_dim}; got C={C}, D={D}")
        if not (lw.dtype == ls0.dtype == fw.dtype == fs0.dtype == fs1.dtype == fx.dtype):
            raise TypeError("all logits / embeddings of one step must share a dtype")
        if lw.dtype != self.dtype:
            if not self._pristine:
                raise TypeError(f"bank holds {self.dtype} rows but the step is {lw.dtype}")
            self._alloc_bank(lw.dtype)          # still all-zero: re-create in the step's dtype
        R, geom = self.geom.world_size, self.geom

        n = rows + n_x
        do_enqueue = geom.should_enqueue(n, self.enqueue_mode)
        fused = self.fuse_rows and C <= 32 and rows <= self.FUSED_ROWS_MAX
        rowsum = numer = join_side = None
        lds = (0, 0)                                                         # (rowsum_ld, numer_ld): 0 = dense
        if R == 1 or self._shards is not None:
            if R > 1 and not fused:
                raise RuntimeError(f"exchange={self.exchange!r} needs the fused row kernel (<= {self.FUSED_ROWS_MAX} unlabeled rows per "
                                   "rank, fuse_rows=True); build the head with exchange='peer' for larger batches")
            if not fused:
                self._k_da(lw)                                               # K2
            multi = self._shards is not None
            smooth = self.smoothing if smooth is None else bool(smooth)
            if self._presmoothed is not None:                                # lock-step emulation: K3 ran in phase 1
                rowsum, numer = self._presmoothed
                self._presmoothed = None
            elif smooth or multi:                                            # K3, bank as of *before* this step's enqueue
                rowsum, numer = self._k_smooth(fw)                           # (peer-memory bank: K3 also carries the epoch flags)
            if not smooth:
                rowsum = numer = None                                        # comatch.py:179 gate closed: un-smoothed probs
            if fused:                                                        # ONE cluster launch: DA + finalize + enqueue
                out = self._k_rows_fused(lw, ls0, rowsum, numer, lds, fw, fx, tx, do_enqueue and not multi, False)
                if multi:
                    # peer-memory bank: this rank's rows go out (over NVLink) in a wide launch on a side stream --
                    # nothing of this step depends on them; the stream is joined after the contrastive forward
                    side, cur = self._side_stream(), torch.cuda.current_stream(self.device)
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        self._k_enqueue_peer(fw, fx, out["probs_orig"], tx)
                    join_side = side
            else:
                out = self._k_finalize(lw, ls0, rowsum, numer, lds)          # K2b + K4 + K7
                if do_enqueue:
                    self._k_enqueue(fw, fx, out["probs_orig"], tx, 0, n)     # K5
        else:
            # sharded bank (bank.py): 3 collectives per step.  (1) all-gather of the enqueue blocks -- they contain
            # every rank's queries; (2) reduce-scatter of the packed [R*n, W] partial sums computed against the
            # local shard (all block rows, labeled ones included: their 1/MU of extra work avoids a strided copy);
            # (3) all-gather of the probability blocks [probs_orig ; onehot], then ONE sharded ring write: the
            # gathered buffers are already in global (rank-major) row order.
            W = (C + 1 + 3) & ~3
            arena = self._peer_arena(n, D * fw.element_size(), W * 4, C * 4)
            gathered_f = self._x_all_gather(arena, 0, [fw, fx])              # [R*n, D]
            if self.smoothing if smooth is None else bool(smooth):
                packed = self._k_smooth(gathered_f, packed_ld=W)             # [R*n, W]: numer | rowsum
                mine = self._x_reduce_scatter(arena, 1, packed)              # [n, W]; rows [0, rows) are this rank's queries
                numer, rowsum, lds = mine, mine[:, C:], (W, W)
            if fused:
                out = self._k_rows_fused(lw, ls0, rowsum, numer, lds, fw, fx, tx, False, True)
                probs_block = out["probs_block"]
            else:
                self._k_da(lw)
                out = self._k_finalize(lw, ls0, rowsum, numer, lds)
                onehot = torch.zeros(n_x, C, dtype=torch.float32, device=lw.device).scatter_(1, tx.view(-1, 1), 1.0)
                probs_block = torch.cat([out["probs_orig"], onehot], dim=0)
            if do_enqueue:
                pb_all = self._x_all_gather(arena, 2, [probs_block])         # [R*n, C]
                self._k_enqueue(gathered_f, gathered_f[:0], pb_all, tx[:0], 0, R * n)
        if do_enqueue:
            self._queue_ptr = geom.next_ptr(self._queue_ptr, n)
            self._pristine = False
        stats, loss_c = self._k_contrast_fwd(fs0, fs1, out["probs"], out["scalars"], lambda_u, lambda_c,
                                             probs_hl=out["probs_hl"])  # K6
        if join_side is not None:
            torch.cuda.current_stream(self.device).wait_stream(join_side)
        self.last = {"probs_orig": out["probs_orig"], "rowsum": rowsum, "numer": numer, "probs": out["probs"],
                     "mask": out["mask"], "lbs": out["lbs"], "scores": out["scores"]}
        out["stats"] = stats
        return out

    # ---- row exchanges of the sharded bank ------------------------------------------
    def _peer_arena(self, n: int, feat_row_bytes: int, packed_row_bytes: int, prob_row_bytes: int):
        """The NVLink peer-memory arena for blocks of ``n`` rows, (re)built collectively when the block grows;
        ``None`` when the exchanges go through torch.distributed."""
        if self.exchange != "peer":
            return None
        need = {0: n * feat_row_bytes, 1: n * packed_row_bytes, 2: n * prob_row_bytes}
        if self._arena is None or not all(self._arena.fits(x, b) for x, b in need.items()):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the peer arena must exist before CUDA graph capture: run one eager step first")
            from .peer import PeerArena
            if self._arena is not None:
                self._arena.close()
            self._arena = PeerArena(self.pg, self.device, need)
        return self._arena

    def _x_all_gather(self, arena, x: int, parts):
        if arena is not None and all((t.numel() * t.element_size()) % 16 == 0 for t in parts):
            return arena.all_gather(x, parts)
        return all_gather_rows(parts[0] if len(parts) == 1 else torch.cat(list(parts), dim=0), self.pg)

    def _x_reduce_scatter(self, arena, x: int, packed):
        if arena is not None and (packed.numel() // self.geom.world_size * 4) % 16 == 0:
            return arena.reduce_scatter(x, packed)
        return reduce_scatter_rows(packed, self.pg)

    def peer_timeouts(self) -> int:
        """Peer waits that gave up since the arena was built (0 in a healthy job; synchronises)."""
        return self._arena.timeouts() if self._arena is not None else 0

    def close(self) -> None:
        """Collective: release the peer arena (before ``destroy_process_group``)."""
        if self._arena is not None:
            torch.cuda.synchronize(self.device)
            self.queue_feats = self.queue_probs = self.queue_probs_t = self._shards = None    # they alias the arena
            self._arena.close()
            self._arena = None

    # ---- kernel wrappers (one C-ABI call each) --------------------------------------
    def _ws(self, rows):
        return N.workspace(self.device, rows, self.num_classes,
                           self.queue_size if self._shards is not None else self.geom.shard_rows)

    def _k_da(self, lw) -> None:
        ws, wsb = self._ws(lw.shape[0])
        N.check(N.lib().b200ssl_comatch_da(lw.data_ptr(), lw.shape[0], lw.shape[1], N.dtype_enum(lw),
                                           self.da_ring.data_ptr(), self.da_state.data_ptr(), self.da_window,
                                           self.prob_avg.data_ptr(), None, ws, wsb, N.stream_ptr(self.device)),
                "comatch_da")

    def _k_smooth(self, queries, packed_ld: int = 0):
        """Returns ``(rowsum [rows], numer [rows, C])`` or, with ``packed_ld = W``, one ``[rows, W]`` tensor holding
        numer in columns [0, C) and rowsum in column C (what the sharded bank reduce-scatters)."""
        rows, D = queries.shape
        C = self.num_classes
        queries = queries.contiguous()
        if packed_ld:
            packed = torch.empty(rows, packed_ld, dtype=torch.float32, device=self.device)
            rowsum, numer, lds = packed[:, C:], packed, (packed_ld, packed_ld)
        else:
            rowsum = torch.empty(rows, dtype=torch.float32, device=self.device)
            numer = torch.empty(rows, C, dtype=torch.float32, device=self.device)
            lds = (0, 0)
        ws, wsb = self._ws(rows)
        sh = self._shards
        N.check(N.lib().b200ssl_bank_smooth_partial(queries.data_ptr(), self.queue_feats.data_ptr(),
                                                    self.queue_probs.data_ptr(), N.ptr(self.queue_probs_t), rows,
                                                    self.queue_size if sh is not None else self.geom.shard_rows, D, C,
                                                    N.dtype_enum(queries), self.temperature,
                                                    rowsum.data_ptr(), numer.data_ptr(), lds[0], lds[1],
                                                    C_byref(sh), ws, wsb,
                                                    N.stream_ptr(self.device)), "bank_smooth_partial")
        return packed if packed_ld else (rowsum, numer)

    def _k_finalize(self, lw, ls0, rowsum, numer, lds=(0, 0)) -> dict:
        rows, C = lw.shape
        f32 = dict(dtype=torch.float32, device=self.device)
        out = {"probs": torch.empty(rows, C, **f32), "probs_orig": torch.empty(rows, C, **f32),
               "scores": torch.empty(rows, **f32), "mask": torch.empty(rows, **f32),
               "lbs": torch.empty(rows, dtype=torch.int64, device=self.device),
               "grad_s0": torch.empty_like(ls0),
               "scalars": torch.empty(4, **f32),      # loss_u, mask_mean, loss_contrast, total
               # bf16 hi/lo split of probs: operand of the tensor-core graph kernel (contrast_tc.cu)
               "probs_hl": (torch.empty(rows, 64, dtype=torch.bfloat16, device=self.device)
                            if (lw.dtype == torch.bfloat16 and C <= 32 and self.low_dim == 64) else None)}
        ws, wsb = self._ws(rows)
        N.check(N.lib().b200ssl_comatch_finalize(
            lw.data_ptr(), ls0.data_ptr(), self.prob_avg.data_ptr(), N.ptr(rowsum), N.ptr(numer), lds[0], lds[1], rows, C,
            N.dtype_enum(lw), float(np.float32(self.alpha)), float(np.float32(1.0 - self.alpha)), self.thr, self.gamma,
            out["probs"].data_ptr(), out["probs_orig"].data_ptr(), N.ptr(out["probs_hl"]), out["scores"].data_ptr(),
            out["lbs"].data_ptr(),
            out["mask"].data_ptr(), out["grad_s0"].data_ptr(), out["scalars"].data_ptr(), ws, wsb,
            N.stream_ptr(self.device)), "comatch_finalize")
        return out

    def _k_rows_fused(self, lw, ls0, rowsum, numer, lds, fw, fx, tx, do_enqueue: bool, onehot_tail: bool) -> dict:
        """DA statistics + finalisation (+ enqueue) in one cluster launch (``b200ssl_comatch_rows_fused``).
        ``onehot_tail``: also emit ``probs_block = [probs_orig ; onehot(targets_x)]`` (sharded enqueue)."""
        rows, C = lw.shape
        f32 = dict(dtype=torch.float32, device=self.device)
        block = torch.empty(rows + (fx.shape[0] if onehot_tail else 0), C, **f32)
        out = {"probs": torch.empty(rows, C, **f32), "probs_orig": block[:rows], "probs_block": block,
               "scores": torch.empty(rows, **f32), "mask": torch.empty(rows, **f32),
               "lbs": torch.empty(rows, dtype=torch.int64, device=self.device),
               "grad_s0": torch.empty_like(ls0), "scalars": torch.empty(4, **f32),
               "probs_hl": (torch.empty(rows, 64, dtype=torch.bfloat16, device=self.device)
                            if (lw.dtype == torch.bfloat16 and C <= 32 and self.low_dim == 64) else None)}
        enq = do_enqueue
        N.check(N.lib().b200ssl_comatch_rows_fused(
            lw.data_ptr(), ls0.data_ptr(), N.ptr(rowsum), N.ptr(numer), lds[0], lds[1], rows, C, N.dtype_enum(lw),
            float(np.float32(self.alpha)), float(np.float32(1.0 - self.alpha)), self.thr, self.gamma,
            self.da_ring.data_ptr(), self.da_state.data_ptr(), self.da_window, self.prob_avg.data_ptr(),
            out["probs"].data_ptr(), out["probs_orig"].data_ptr(), N.ptr(out["probs_hl"]), out["scores"].data_ptr(),
            out["lbs"].data_ptr(), out["mask"].data_ptr(), out["grad_s0"].data_ptr(), out["scalars"].data_ptr(),
            self.queue_feats.data_ptr() if enq else None, self.queue_probs.data_ptr() if enq else None,
            N.ptr(self.queue_probs_t) if enq else None, fw.data_ptr(), fx.data_ptr(), tx.data_ptr(), fx.shape[0],
            self.low_dim, self.ptr_state.data_ptr(), self.queue_size, 1 if onehot_tail else 0,
            N.stream_ptr(self.device)), "comatch_rows_fused")
        return out

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def _k_enqueue_peer(self, fw, fx, probs_orig, tx) -> None:
        """This rank's block -> the peer-memory bank (owning shard, or every copy of a replicated ring)."""
        N.check(N.lib().b200ssl_bank_enqueue_peer(fw.data_ptr(), fx.data_ptr(), probs_orig.data_ptr(), tx.data_ptr(), fw.shape[0],
                                                  fx.shape[0], self.low_dim, self.num_classes, N.dtype_enum(fw),
                                                  self.ptr_state.data_ptr(), C_byref(self._shards),
                                                  N.stream_ptr(self.device)), "bank_enqueue_peer")

    def _k_enqueue(self, fw, fx, probs_orig, tx, block_offset: int, advance: int) -> None:
        g = self.geom
        fw, fx, probs_orig, tx = (t.contiguous() for t in (fw, fx, probs_orig, tx))
        N.check(N.lib().b200ssl_bank_enqueue(self.queue_feats.data_ptr(), self.queue_probs.data_ptr(),
                                             N.ptr(self.queue_probs_t), fw.data_ptr(),
                                             fx.data_ptr(), probs_orig.data_ptr(), tx.data_ptr(), fw.shape[0],
                                             fx.shape[0], self.low_dim, self.num_classes, N.dtype_enum(fw),
                                             0, self.ptr_state.data_ptr(), advance, block_offset, g.queue_size,
                                             g.shard_begin, g.shard_rows,
                                             N.stream_ptr(self.device)), "bank_enqueue")

    def _k_contrast_fwd(self, fs0, fs1, probs, scalars, lambda_u: float = 1.0, lambda_c: float = 1.0, probs_hl=None):
        """scalars: [loss_u (in), mask_mean, loss_contrast (out), total (out)]."""
        rows, D = fs0.shape
        stats = torch.empty(3, rows, dtype=torch.float32, device=self.device)
        ws, wsb = self._ws(rows)
        N.check(N.lib().b200ssl_contrast_fwd(fs0.data_ptr(), fs1.data_ptr(), probs.data_ptr(), N.ptr(probs_hl), rows, D,
                                             self.num_classes, N.dtype_enum(fs0), self.temperature, self.contrast_th,
                                             stats.data_ptr(), scalars[2:].data_ptr(), scalars.data_ptr(), lambda_u,
                                             lambda_c, scalars[3:].data_ptr(), ws, wsb,
                                             N.stream_ptr(self.device)), "contrast_fwd")
        return stats, scalars[2]

    def _k_contrast_bwd(self, f0, f1, probs, stats, g_c, factor: float = 1.0, probs_hl=None, scale=None):
        g = g_c.detach().to(torch.float32).reshape(1).contiguous()
        sg, sn, su, sf = None, 0, None, 1.0
        if scale is not None:
            sg, su, sf = scale[0], scale[1].detach().to(torch.float32).reshape(1).contiguous(), float(scale[2])
            sn = sg.numel()
        gf0, gf1 = torch.empty_like(f0), torch.empty_like(f1)
        rows, D = f0.shape
        ws, wsb = self._ws(rows)
        N.check(N.lib().b200ssl_contrast_bwd(f0.data_ptr(), f1.data_ptr(), probs.data_ptr(), N.ptr(probs_hl),
                                             stats.data_ptr(), rows,
                                             D, self.num_classes, N.dtype_enum(f0), self.temperature, self.contrast_th,
                                             g.data_ptr(), factor, gf0.data_ptr(), gf1.data_ptr(), N.ptr(sg), sn, N.ptr(su), sf,
                                             ws, wsb,
                                             N.stream_ptr(self.device)), "contrast_bwd")
        return gf0, gf1

    def _k_scale(self, grad, g, factor: float = 1.0):
        g = g.detach().to(torch.float32).reshape(1).contiguous()
        N.check(N.lib().b200ssl_scale_inplace(grad.data_ptr(), grad.numel(), N.dtype_enum(grad), g.data_ptr(), factor,
                                              N.stream_ptr(self.device)), "scale_inplace")
        return grad


def lockstep_total_loss(heads, batches, lambda_u: float = 1.0, lambda_c: float = 1.0):
    """One step of several emulated ranks (``CoMatchHead(local_ranks=...)``) on ONE GPU and ONE stream, phase by phase:
    every rank's K3 first (it needs all enqueues of the previous step, which are complete), then every rank's row
    kernel, enqueue and contrastive forward (the enqueue needs every rank's K3 of this step).  No kernel is ever
    launched before one whose flag it waits for, so nothing spins on the device.  Returns the per-rank ``total_loss`` tuples."""
    for h, b in zip(heads, batches):
        h.presmooth(b["feats_u_w"])
    return [h.total_loss(**b, lambda_u=lambda_u, lambda_c=lambda_c) for h, b in zip(heads, batches)]
