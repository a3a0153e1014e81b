"""Multi-GPU plumbing for the selective-pose glue path.  One process per GPU.

* Frames (and everything derived from them: head maps, crops, heatmaps) are data-parallel: rank r owns
  its own frames, no data-path collective.
* A large gallery is sharded by rows: rank r holds identities ``[offset_r, offset_r + N_r)``.  One step:
  ``all_gather`` of the probes (``M_local x 512`` fp32 per rank), local fused GEMM + top-1 of ALL probes
  against the local shard, then a top-1 (value, index) reduction: every (similarity, global id) pair is
  packed into one int64 key whose integer order is "higher similarity first, lower id on ties", so a
  plain ``all_reduce(MAX)`` over NCCL/NVLink is the arg-max reduce (NCCL has no arg-max operator).

The key layout is the one ``spp_match_top1`` writes on the device (include/spp.h):
``key = (orderable_i32(sim) << 32) | (0xFFFFFFFF - global_id)`` interpreted as a SIGNED 64-bit integer.
``pack_keys`` / ``unpack_keys`` are the host mirror (used by the CPU tests over gloo).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

NOT_FOUND_ID = 0x7FFFFFFF


def float_to_ordered(x: torch.Tensor) -> torch.Tensor:
    """Monotone fp32 -> int32 map (a < b <=> key(a) < key(b)); same bit trick as the kernels."""
    i = x.contiguous().view(torch.int32)
    return i ^ ((i >> 31) & 0x7FFFFFFF)


def ordered_to_float(i: torch.Tensor) -> torch.Tensor:
    i = i.to(torch.int32)
    return (i ^ ((i >> 31) & 0x7FFFFFFF)).view(torch.float32)


def pack_keys(sims: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """(fp32 similarity, global id >= 0 or -1 for "none") -> signed int64 key."""
    gid = torch.where(ids < 0, torch.full_like(ids, NOT_FOUND_ID), ids).to(torch.int64)
    s = torch.where(ids < 0, torch.full_like(sims, float("-inf")), sims).float()
    hi = float_to_ordered(s).to(torch.int64)
    return (hi << 32) | (0xFFFFFFFF - gid)


def unpack_keys(keys: torch.Tensor, threshold: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inverse of :func:`pack_keys` + the similarity gate: returns ``(ids int64, sims fp32)``, id -1 when
    nothing was found or ``sim < threshold``."""
    hi = (keys >> 32).to(torch.int32)
    gid = 0xFFFFFFFF - (keys & 0xFFFFFFFF)
    sims = ordered_to_float(hi)
    ok = gid != NOT_FOUND_ID
    if threshold is not None:
        ok = ok & ~(sims < threshold)
    return torch.where(ok, gid, torch.full_like(gid, -1)), sims


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of gallery shard ``rank``."""
    per, rem = divmod(n, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


class ShardedGalleryMatcher:
    """Top-1 cosine match of data-parallel probes against a row-sharded gallery.

    ``local_match(all_probes [G*M, 512]) -> int64 keys [G*M]`` scores every probe against THIS rank's
    shard.  On a GPU rank it is the tcgen05 kernel (``ops.match_top1(..., want_keys=True)``); the CPU
    tests inject an oracle-based function so the collective logic runs over gloo.
    Every rank must bring the same number of probes per step (pad with zeros otherwise).
    """

    def __init__(self, local_match: Callable[[torch.Tensor], torch.Tensor], threshold: Optional[float] = None,
                 group: Optional[dist.ProcessGroup] = None,
                 unpack: Callable[[torch.Tensor, Optional[float]], Tuple[torch.Tensor, torch.Tensor]] = unpack_keys):
        self.local_match, self.threshold, self.group, self.unpack = local_match, threshold, group, unpack
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gather: Optional[torch.Tensor] = None

    def match(self, probes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        m = probes.shape[0]
        if self.world == 1:
            return self.unpack(self.local_match(probes), self.threshold)
        if self._gather is None or self._gather.shape[0] != self.world * m or self._gather.device != probes.device:
            self._gather = torch.empty((self.world * m, probes.shape[1]), dtype=probes.dtype, device=probes.device)
        dist.all_gather_into_tensor(self._gather, probes.contiguous(), group=self.group)      # exchange 1: probes
        keys = self.local_match(self._gather)                                                 # [G*M] int64
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=self.group)                         # exchange 2: top-1 reduce
        mine = keys[self.rank * m:(self.rank + 1) * m]
        return self.unpack(mine, self.threshold)


def gpu_matcher(gallery_shard_bf16: torch.Tensor, id_offset: int, threshold: Optional[float] = None,
                group: Optional[dist.ProcessGroup] = None) -> ShardedGalleryMatcher:
    """The production wiring: local match = tcgen05 GEMM + fused top-2 + fp32 re-score on this GPU."""
    from . import ops

    def local(p: torch.Tensor) -> torch.Tensor:
        return ops.match_top1(p, gallery_shard_bf16, None, id_offset, want_keys=True)[2]

    def unpack(k: torch.Tensor, thr: Optional[float]):
        ids, sims = ops.match_unpack_keys(k.contiguous(), thr)
        return ids.long(), sims

    return ShardedGalleryMatcher(local, threshold, group, unpack)
