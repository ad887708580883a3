"""Host-side geometry and collectives of the rank-sharded CoMatch memory bank.

The reference keeps one bank per process (``code/comatch.py:90-96``) and has no
distributed code.  Here the global ring of ``K`` rows is split into contiguous
shards, rank ``r`` owning rows ``[r*K/R, (r+1)*K/R)``.  Per step

1. all-gather of each rank's enqueue block ``[n, D]`` (it contains the queries);
2. every rank runs K3 for all ``R*B_u`` queries against its shard;
3. reduce-scatter (sum) of the ``[R*B_u, 1+C]`` partial row-sums / numerators;
4. every rank writes the slice of global rows ``[ptr, ptr + R*n)`` that falls
   into its shard; ``ptr`` advances identically everywhere.

Everything in this file is device agnostic (NCCL on GPUs, gloo in the CPU
tests); the kernels it feeds live in ``csrc/bank.cu``.  This is the ``exchange='collective'``
path of ``CoMatchHead``; ``peer.py`` carries the same exchanges -- or removes them by keeping
the ring in NVLink peer memory -- without a communication library on the data path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import torch

__all__ = ["ShardGeometry", "all_gather_rows", "reduce_scatter_rows", "local_segments"]


@dataclass(frozen=True)
class ShardGeometry:
    queue_size: int       # K, global rows
    world_size: int
    rank: int

    def __post_init__(self):
        if self.queue_size <= 0 or self.world_size <= 0 or not (0 <= self.rank < self.world_size):
            raise ValueError(f"bad shard geometry {self}")
        if self.queue_size % self.world_size:
            raise ValueError("queue_size must be divisible by the number of ranks (contiguous equal shards)")

    @property
    def shard_rows(self) -> int:
        return self.queue_size // self.world_size

    @property
    def shard_begin(self) -> int:
        return self.rank * self.shard_rows

    def next_ptr(self, ptr: int, n_per_rank: int) -> int:
        """``queue_ptr = (queue_ptr + n) % queue_size`` (comatch.py:196) for the R concatenated blocks."""
        return (ptr + n_per_rank * self.world_size) % self.queue_size

    def should_enqueue(self, n_per_rank: int, mode: str) -> bool:
        """``'reference'``: the guard of comatch.py:192 on the concatenated batch; ``'always'``: ring write."""
        total = n_per_rank * self.world_size
        if total > self.queue_size:
            raise ValueError("enqueue block larger than the bank")
        return mode == "always" or total == self.queue_size


def local_segments(ptr: int, total_rows: int, geom: ShardGeometry) -> List[Tuple[int, int, int]]:
    """Which of the ``total_rows`` rows written at global position ``ptr`` (with
    wrap) land in this rank's shard: list of ``(src_row, local_dst_row, length)``.
    Pure description of what ``b200ssl_bank_enqueue`` does row by row; used to
    validate the kernel and by the CPU tests."""
    K, lo, hi = geom.queue_size, geom.shard_begin, geom.shard_begin + geom.shard_rows
    segs = []
    src = 0
    while src < total_rows:
        g = (ptr + src) % K
        run = min(total_rows - src, K - g)           # until the ring wraps
        a, b = max(g, lo), min(g + run, hi)
        if a < b:
            segs.append((src + (a - g), a - lo, b - a))
        src += run
    return segs


def all_gather_rows(block: torch.Tensor, pg=None) -> torch.Tensor:
    """Rank-major concatenation ``[R*n, ...]`` of every rank's ``[n, ...]`` block."""
    import torch.distributed as dist
    R = dist.get_world_size(pg)
    block = block.contiguous()
    out = torch.empty((R * block.shape[0],) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, block, group=pg)
    return out


def reduce_scatter_rows(part: torch.Tensor, pg=None) -> torch.Tensor:
    """Sum ``[R*rows, ...]`` over ranks and keep this rank's ``[rows, ...]`` slice."""
    import torch.distributed as dist
    R = dist.get_world_size(pg)
    part = part.contiguous()
    rows = part.shape[0] // R
    out = torch.empty((rows,) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
    dist.reduce_scatter_tensor(out, part, op=dist.ReduceOp.SUM, group=pg)
    return out
