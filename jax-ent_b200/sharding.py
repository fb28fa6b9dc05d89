"""Frame sharding of the ensemble across ranks.

Every per-frame quantity (feature columns, logits, Adam moments, prior, weights) is local to the rank
that owns the frame; ranks couple only through three tiny collectives:

  * once, at initialisation:  max(z0) (MAX) and sum(|prior|+eps) (SUM)
  * every step, after `reduce`:  the (4R+4)-double partial-sum vector  JAXENT_BUF_RED   (SUM)
  * every step, after `tail`  :  the 4-double MaxEnt vector             JAXENT_BUF_KLSUM (SUM)

On GPUs with peer access (NVLink / NVSwitch) the two per-step exchanges are fused into the kernels
(`PeerExchange`: every rank's reduce / tail kernel stores its partial sums straight into the peers'
inboxes and the consumers wait on flags; include/jaxent_b200.h "frame-sharded exchange"), so a step
launches no collective at all; NCCL all-reduces remain the fallback when peer mapping is unavailable.

BV parameters, the R x T tail and the plateau decision are replicated: every rank computes them from
the same all-reduced numbers, so they stay bit-identical without a broadcast.

The same code drives NCCL on GPUs (engine.py) and gloo on CPU (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

FRAME_ALIGN = 8  # shards start on a tile boundary (8 frames) so every rank packs whole tiles


@dataclass(frozen=True)
class FrameShard:
    rank: int
    world: int
    start: int
    stop: int
    total: int

    @property
    def size(self) -> int:
        return self.stop - self.start

    def slice(self) -> slice:
        return slice(self.start, self.stop)


def frame_shard(total_frames: int, rank: int, world: int, align: int = FRAME_ALIGN) -> FrameShard:
    """Contiguous, tile-aligned block of frames owned by `rank`; remainders go to the last ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"invalid rank {rank} / world {world}")
    if total_frames < world:
        raise ValueError(f"cannot shard {total_frames} frames over {world} ranks")
    tiles = (total_frames + align - 1) // align
    lo = (tiles * rank // world) * align
    hi = (tiles * (rank + 1) // world) * align
    lo, hi = min(lo, total_frames), min(hi, total_frames)
    if rank == world - 1:
        hi = total_frames
    if hi <= lo:
        raise ValueError(f"rank {rank} would own no frames ({total_frames} frames over {world} ranks)")
    return FrameShard(rank, world, lo, hi, total_frames)


class ShardComm:
    """The collectives of the sharded step; a no-op for a single rank."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.active = bool(group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1)
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.calls = 0

    def sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.active:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
        return t

    def max_(self, t: torch.Tensor) -> torch.Tensor:
        if self.active:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self.calls += 1
        return t


class PeerExchange:
    """Maps one exchange buffer per rank into every process (CUDA IPC over NVLink peer access) for the
    fused in-kernel all-reduce.  Set-up only; raises if any rank cannot map its peers (the caller then
    falls back to NCCL collectives on all ranks together)."""

    # Exchange buffers belong to the process group, not to one engine: mapping them (cudaMalloc + IPC handle exchange +
    # cudaIpcOpenMemHandle on every peer) costs tens of milliseconds, so a released buffer set is kept and handed to the next
    # engine of the same group / device that fits.  Every rank constructs and releases its engines in the same order, so the
    # pools stay identical across ranks (acquire / release are collective in that sense).
    _pool: dict = {}

    @classmethod
    def acquire(cls, lib, group, problem, device: torch.device) -> "PeerExchange":
        import ctypes as C

        need = lib.jaxent_bv_exchange_bytes_for(C.byref(problem), dist.get_world_size(group))
        key = (id(group), torch.device(device).index)
        for px in cls._pool.get(key, []):
            if not px.in_use and px.nbytes >= need:
                px.in_use = True
                return px
        px = cls(lib, group, problem, device)
        cls._pool.setdefault(key, []).append(px)
        return px

    def release(self) -> None:
        """the owning engine is done (its ranks have synchronised): keep the mapping for the next engine"""
        self.in_use = False

    @classmethod
    def purge(cls) -> None:
        """unmap and free every pooled buffer set (collective: call on all ranks, after a barrier)"""
        for lst in cls._pool.values():
            for px in lst:
                px.close()
        cls._pool.clear()

    def __init__(self, lib, group, problem, device: torch.device):
        import ctypes as C

        from . import _lib as L

        self.in_use = True
        self.nbytes = 0
        self.lib, self.group = lib, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.own = None
        self.opened = []
        self.ptrs = [None] * self.world
        if self.world > L.MAX_RANKS:
            raise RuntimeError(f"the fused exchange supports at most {L.MAX_RANKS} ranks")
        nbytes = lib.jaxent_bv_exchange_bytes_for(C.byref(problem), self.world)
        self.nbytes = nbytes
        ok, err = 1, ""
        handle = C.create_string_buffer(L.IPC_HANDLE_BYTES)
        ptr = C.c_void_p()
        with torch.cuda.device(device):
            if lib.jaxent_peer_alloc(nbytes, C.byref(ptr), handle) != 0:
                ok, err = 0, lib.jaxent_last_error().decode()
            else:
                self.own = ptr.value
            handles = [None] * self.world
            dist.all_gather_object(handles, (ok, handle.raw), group=group)
            if all(h[0] for h in handles):
                for r, (_, raw) in enumerate(handles):
                    if r == self.rank:
                        self.ptrs[r] = self.own
                        continue
                    q = C.c_void_p()
                    if lib.jaxent_peer_open(raw, C.byref(q)) != 0:
                        ok, err = 0, lib.jaxent_last_error().decode()
                        break
                    self.opened.append(q.value)
                    self.ptrs[r] = q.value
            else:
                ok = 0
            flag = torch.tensor([ok], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.close()
                raise RuntimeError(f"peer mapping of the exchange buffers failed on some rank ({err or 'see other ranks'})")

    def attach(self, plan) -> None:
        import ctypes as C

        from . import _lib as L

        arr = (C.c_void_p * self.world)(*self.ptrs)
        L.check(self.lib.jaxent_bv_plan_set_exchange(plan, self.rank, self.world, arr))

    def close(self) -> None:
        for q in self.opened:
            self.lib.jaxent_peer_close(q)
        self.opened = []
        if self.own is not None:
            self.lib.jaxent_peer_free(self.own)
            self.own = None
