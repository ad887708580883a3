"""NVLink peer-memory exchanges of the rank-sharded memory bank.

``PeerArena`` owns one device allocation per rank that every rank of the node maps through CUDA IPC
(``b200ssl_peer_alloc`` / ``b200ssl_peer_open``) and runs the bank's row exchanges as ONE kernel launch each
(``csrc/peer.cu``): push over NVLink, publish an epoch flag, wait, copy out / fold in rank order.  The process
group is used only at construction (to swap the 64-byte IPC handles) and never on the data path.

The reference has no distributed code (``code/comatch.py:90-96`` keeps one bank per process); this module and
``bank.py`` are the multi-GPU extension of that bank.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _native as N

__all__ = ["PeerArena", "LocalArenaSet"]


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class PeerArena:
    """``regions``: ``{exchange_id: bytes_per_rank}`` -- the largest block one rank contributes to that exchange."""

    def __init__(self, pg, device, regions: Dict[int, int], named: Optional[Dict[str, int]] = None, *, _local=None):
        import torch.distributed as dist
        self.pg, self.device = pg, torch.device(device)
        self._local = _local                       # (LocalArenaSet, rank): all arenas live in this process (single-GPU validation)
        if _local is not None:
            self.rank, self.world = int(_local[1]), _local[0].world
        else:
            self.rank, self.world = dist.get_rank(pg), dist.get_world_size(pg)
        lib = N.lib()
        self.slot: Dict[int, int] = {}
        self.offset: Dict[int, int] = {}
        off = int(lib.b200ssl_peer_control_bytes())
        for x in sorted(regions):
            self.slot[x] = _round_up(int(regions[x]), 256)
            self.offset[x] = off
            off += 2 * self.world * self.slot[x]
        # named areas (e.g. the bank shard itself): same offset in every arena, 256-byte aligned
        self.named_offset: Dict[str, int] = {}
        self.named_bytes: Dict[str, int] = dict(named or {})
        for name, nbytes in self.named_bytes.items():
            self.named_offset[name] = off
            off += _round_up(int(nbytes), 256)
        self.bytes = off
        self._mapped = []
        if _local is not None:
            bases = _local[0]._allocate(self.bytes)
            self._own = C.c_void_p(bases[self.rank])
        else:
            self._own = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            with torch.cuda.device(self.device):
                N.check(lib.b200ssl_peer_alloc(self.bytes, C.byref(self._own), handle), "peer_alloc")
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle), group=pg)
                bases = []
                for r, h in enumerate(handles):
                    if r == self.rank:
                        bases.append(self._own.value)
                        continue
                    p = C.c_void_p()
                    N.check(lib.b200ssl_peer_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(p)), f"peer_open(rank {r})")
                    self._mapped.append(p)
                    bases.append(p.value)
        self.bases_host = (C.c_uint64 * self.world)(*bases)
        self.bases = torch.tensor(bases, dtype=torch.int64).to(self.device)
        if _local is None:
            dist.barrier(group=pg)              # nobody pushes before every rank has mapped every arena
        torch.cuda.synchronize(self.device)

    def tensor(self, name: str, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
        """A torch tensor aliasing the named area of the OWN arena (zero-copy; valid until ``close``)."""
        numel = 1
        for d in shape:
            numel *= int(d)
        itemsize = torch.empty(0, dtype=dtype).element_size()
        if numel * itemsize > self.named_bytes[name]:
            raise ValueError(f"area {name!r} holds {self.named_bytes[name]} bytes, asked for {numel * itemsize}")
        typestr = {torch.bfloat16: "<i2", torch.float32: "<f4", torch.float16: "<f2"}[dtype]

        class _View:
            __cuda_array_interface__ = {"shape": tuple(int(d) for d in shape), "typestr": typestr, "version": 2,
                                        "data": (self._own.value + self.named_offset[name], False)}

        t = torch.as_tensor(_View(), device=self.device)
        return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t

    def fits(self, exchange_id: int, nbytes: int) -> bool:
        return exchange_id in self.slot and nbytes <= self.slot[exchange_id]

    def all_gather(self, exchange_id: int, parts: Sequence[torch.Tensor]) -> torch.Tensor:
        """Rank-major ``[R*n, cols]`` concatenation of every rank's block ``cat(parts)`` (one or two row blocks)."""
        parts = [t.contiguous() for t in parts if t.shape[0] > 0]
        if not 1 <= len(parts) <= 2:
            raise ValueError("all_gather takes one or two non-empty row blocks")
        a = parts[0]
        b = parts[1] if len(parts) == 2 else None
        n = a.shape[0] + (b.shape[0] if b is not None else 0)
        nb = [t.numel() * t.element_size() for t in parts]
        if any(x % 16 for x in nb) or not self.fits(exchange_id, sum(nb)):
            raise ValueError(f"exchange {exchange_id}: blocks of {nb} bytes do not fit the arena (16-byte multiples, slot "
                             f"{self.slot.get(exchange_id)})")
        out = torch.empty((self.world * n,) + tuple(a.shape[1:]), dtype=a.dtype, device=self.device)
        N.check(N.lib().b200ssl_peer_all_gather(a.data_ptr(), nb[0], N.ptr(b), nb[1] if b is not None else 0, out.data_ptr(),
                                                self.bases.data_ptr(), self.offset[exchange_id], self.slot[exchange_id],
                                                exchange_id, self.rank, self.world, N.stream_ptr(self.device)), "peer_all_gather")
        return out

    def reduce_scatter(self, exchange_id: int, part: torch.Tensor) -> torch.Tensor:
        """Rank-ordered fp32 sum of ``[R*rows, cols]`` over ranks; returns this rank's ``[rows, cols]``."""
        part = part.contiguous()
        if part.dtype != torch.float32 or part.shape[0] % self.world:
            raise ValueError("reduce_scatter takes fp32 [R*rows, cols]")
        rows = part.shape[0] // self.world
        count = rows * part[0].numel()
        if (count * 4) % 16 or not self.fits(exchange_id, count * 4):
            raise ValueError(f"exchange {exchange_id}: {count * 4} bytes per rank do not fit the arena")
        out = torch.empty((rows,) + tuple(part.shape[1:]), dtype=torch.float32, device=self.device)
        N.check(N.lib().b200ssl_peer_reduce_scatter_f32(part.data_ptr(), out.data_ptr(), count, self.bases.data_ptr(),
                                                        self.offset[exchange_id], self.slot[exchange_id], exchange_id,
                                                        self.rank, self.world, N.stream_ptr(self.device)),
                "peer_reduce_scatter_f32")
        return out

    def timeouts(self) -> int:
        """Number of peer waits that gave up (0 in a healthy job).  Synchronises the device."""
        n = C.c_uint32()
        N.check(N.lib().b200ssl_peer_timeouts(self._own, C.byref(n)), "peer_timeouts")
        return int(n.value)

    def close(self) -> None:
        """Collective: unmap the peers' arenas, then free the own one."""
        if self._own is None:
            return
        torch.cuda.synchronize(self.device)
        if self._local is not None:             # the set owns the allocations
            self._own = None
            return
        import torch.distributed as dist
        lib = N.lib()
        for p in self._mapped:
            lib.b200ssl_peer_close(p)
        self._mapped = []
        dist.barrier(group=self.pg)             # every mapping is gone before any arena is freed
        lib.b200ssl_peer_free(self._own)
        self._own = None


class LocalArenaSet:
    """``world`` arenas of ONE process on ONE device: the multi-rank data path of the peer-memory bank (shard addressing, ring
    positions, epoch flags) run rank after rank on a single GPU.  For validation only -- kernels of different ranks that wait
    for each other must never be left to the mercy of one GPU's scheduler, so the caller launches the ranks phase by phase
    (``comatch_head.lockstep_total_loss``): every flag a kernel looks at has been published by an earlier launch."""

    def __init__(self, world: int, device):
        if not 2 <= int(world) <= 8:
            raise ValueError("2..8 emulated ranks")
        self.world, self.device = int(world), torch.device(device)
        self._bytes = None
        self._ptrs = []

    def _allocate(self, nbytes: int):
        if self._bytes is None:
            lib = N.lib()
            with torch.cuda.device(self.device):
                for _ in range(self.world):
                    p, handle = C.c_void_p(), (C.c_ubyte * 64)()
                    N.check(lib.b200ssl_peer_alloc(int(nbytes), C.byref(p), handle), "peer_alloc")
                    self._ptrs.append(p)
            self._bytes = int(nbytes)
        elif int(nbytes) != self._bytes:
            raise ValueError("every emulated rank must ask for the same arena layout")
        return [p.value for p in self._ptrs]

    def arena(self, rank: int, regions: Dict[int, int], named: Optional[Dict[str, int]] = None) -> "PeerArena":
        return PeerArena(None, self.device, regions, named, _local=(self, int(rank)))

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        for p in self._ptrs:
            N.lib().b200ssl_peer_free(p)
        self._ptrs, self._bytes = [], None
