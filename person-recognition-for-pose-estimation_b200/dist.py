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

    capturable = False     # NCCL calls: issued eagerly beside the pipeline's CUDA graph

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


# ------------------------------------------------------------------------------------------------
# The same exchange WITHOUT NCCL calls in the step: the match kernels write over NVLink peer memory
# ------------------------------------------------------------------------------------------------

STAGE_PUSH, STAGE_WAIT, STAGE_SEARCH, STAGE_FINALIZE, STAGE_REDUCE, STAGE_ALL = 1, 2, 4, 8, 16, 31


def exchange_handles(handle: bytes, group: Optional[dist.ProcessGroup] = None) -> list:
    """All ranks' 64-byte IPC handles, in rank order (the only host-side exchange of the peer path; set-up time)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [handle]
    out = [None] * world
    dist.all_gather_object(out, handle, group=group)
    return out


class PeerGroup:
    """One exchange buffer per rank, all of them mapped into this process (``spp_peer_alloc`` / ``spp_peer_open``,
    CUDA IPC over NVLink).  One process per GPU; ``m_local`` probes per rank and step."""

    def __init__(self, m_local: int, group: Optional[dist.ProcessGroup] = None,
                 alloc: Optional[Callable[[int], Tuple[int, bytes]]] = None, open_: Optional[Callable[[bytes], int]] = None):
        import ctypes
        from . import _lib
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > _lib.MAX_PEERS:
            raise ValueError(f"PeerGroup: at most {_lib.MAX_PEERS} ranks")
        self.m_local, self.group = int(m_local), group
        self._alloc, self._open = alloc or _ipc_alloc, open_ or _ipc_open
        self._injected = alloc is not None
        nbytes = 1 if self._injected else int(_lib.lib().spp_peer_buffer_bytes(self.world, self.m_local, 512))
        if nbytes == 0:
            raise ValueError(f"PeerGroup: unsupported world={self.world} m_local={self.m_local}")
        self.nbytes = nbytes
        own, handle = self._alloc(nbytes)
        self.handles = exchange_handles(handle, group)
        self.ptrs = [own if r == self.rank else self._open(self.handles[r]) for r in range(self.world)]
        self._opened = [p for r, p in enumerate(self.ptrs) if r != self.rank]
        self._own = own
        self.struct = _lib.PeerGroupStruct(self.world, self.rank, self.m_local, (ctypes.c_void_p * _lib.MAX_PEERS)(*self.ptrs))
        if self.world > 1:
            dist.barrier(group)      # every buffer is mapped everywhere before the first kernel touches a peer

    @classmethod
    def virtual(cls, world: int, m_local: int):
        """``world`` ranks inside ONE process on the current device (each with its own buffer): the single-GPU tests run
        the real kernels, flags and waits this way, one stream per virtual rank."""
        import ctypes
        from . import _lib
        nbytes = int(_lib.lib().spp_peer_buffer_bytes(world, m_local, 512))
        ptrs = [_ipc_alloc(nbytes)[0] for _ in range(world)]
        out = []
        for r in range(world):
            g = cls.__new__(cls)
            g.world, g.rank, g.m_local, g.group, g.nbytes = world, r, m_local, None, nbytes
            g.ptrs, g._opened, g._own, g._injected = ptrs, [], ptrs[r], False
            g.struct = _lib.PeerGroupStruct(world, r, m_local, (ctypes.c_void_p * _lib.MAX_PEERS)(*ptrs))
            out.append(g)
        return out

    def close(self, barrier: bool = True) -> None:
        """Unmap the peers' buffers and free this rank's.  ``barrier=False`` only when no rank can still be using them and the
        ranks may not all be calling (a failed collective set-up)."""
        from . import _lib
        if self._injected or self._own is None:
            return
        if barrier and self.world > 1 and dist.is_initialized():
            dist.barrier(self.group)
        for p in self._opened:
            _lib.check(_lib.lib().spp_peer_close(p), "spp_peer_close")
        _lib.check(_lib.lib().spp_peer_free(self._own), "spp_peer_free")
        self._own, self._opened = None, []


def _ipc_alloc(nbytes: int) -> Tuple[int, bytes]:
    import ctypes
    from . import _lib
    ptr = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES)()
    _lib.check(_lib.lib().spp_peer_alloc(nbytes, ctypes.byref(ptr), handle), "spp_peer_alloc")
    return int(ptr.value), bytes(handle)


def _ipc_open(handle: bytes) -> int:
    import ctypes
    from . import _lib
    ptr = ctypes.c_void_p()
    buf = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES).from_buffer_copy(handle)
    _lib.check(_lib.lib().spp_peer_open(buf, ctypes.byref(ptr)), "spp_peer_open")
    return int(ptr.value)


class PeerShardedMatcher:
    """Top-1 of this rank's probes over a row-sharded gallery, the exchange done by the kernels themselves
    (``spp_sharded_match_top1``): probes are normalised and stored into every rank's buffer, every rank scores ALL
    probes against its shard (tcgen05 GEMM + fp32 re-score), the packed (similarity, id) keys are stored into their
    owner's buffer, the owner takes the integer maximum.  Five kernels, no NCCL call, CUDA-graph capturable.
    Every rank must call ``match`` the same number of times."""

    capturable = True

    def __init__(self, peers: PeerGroup, shard_bf16: torch.Tensor, id_offset: int, threshold: Optional[float] = None,
                 shard_f32: Optional[torch.Tensor] = None, max_row_norm: float = 1.0):
        from . import _lib, ops
        self.peers, self.shard, self.shard_f32 = peers, shard_bf16.contiguous(), shard_f32
        self.id_offset, self.threshold, self.max_row_norm = int(id_offset), threshold, float(max_row_norm)
        dev = shard_bf16.device
        nbytes = int(_lib.lib().spp_sharded_match_workspace_bytes(peers.world, peers.m_local, shard_bf16.shape[0], 512))
        if nbytes == 0:
            raise ValueError("PeerShardedMatcher: unsupported shape")
        self._ws, self._ws_bytes = ops.alloc_workspace(dev, nbytes), nbytes
        self.ids = torch.empty((peers.m_local,), dtype=torch.int32, device=dev)
        self.sims = torch.empty((peers.m_local,), dtype=torch.float32, device=dev)
        self.launches = 5

    def match(self, probes: torch.Tensor, stages: int = STAGE_ALL) -> Tuple[torch.Tensor, torch.Tensor]:
        import ctypes
        from . import _lib
        if probes.shape != (self.peers.m_local, 512) or probes.dtype != torch.float32 or not probes.is_contiguous():
            raise ValueError(f"PeerShardedMatcher.match: probes must be a contiguous fp32 [{self.peers.m_local}, 512] tensor")
        thr = float("nan") if self.threshold is None else float(self.threshold)
        p = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
        _lib.check(_lib.lib().spp_sharded_match_top1(
            ctypes.byref(self.peers.struct), p(probes), p(self.shard), p(self.shard_f32), self.max_row_norm, self.shard.shape[0], 512,
            self.id_offset, thr, stages, p(self.ids), p(self.sims), None, p(self._ws), self._ws_bytes,
            ctypes.c_void_p(torch.cuda.current_stream(probes.device).cuda_stream)), "spp_sharded_match_top1")
        return self.ids, self.sims
