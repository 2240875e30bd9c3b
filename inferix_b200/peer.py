"""Peer-memory exchange for the sequence-parallel path (one process per GPU, one NVSwitch box).

The reference's SP attention ships K/V between ranks with NCCL P2P every layer (ring pass-kv,
inferix/models/attention/distributed.py:564-712).  The first native variant replaced that by one NCCL all-gather of the
block's new K/V per layer into a replicated paged cache.  This module removes the collective from the layer altogether:
every rank maps the other ranks' cache buffers through CUDA IPC once, and `ifx_qk_norm_rope_append_peers` (the fused
QK-norm + RoPE kernel that already produces K and V) stores its rows straight into all caches over NVLink and publishes
an epoch flag; `ifx_peer_wait` orders the attention after every rank's flag.  Per layer: 3 launches, no staging buffer,
no all-gather, no re-interleave pass.

PyTorch's part is plumbing only: `torch.distributed.all_gather_object` carries the 64-byte IPC handles once at setup.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib

ENABLED = os.environ.get("IFX_SP_PEER", "1") != "0"
WAIT_TIMEOUT_MS = int(os.environ.get("IFX_SP_PEER_TIMEOUT_MS", "60000"))


class PeerGroup:
    """Flag array + registry of peer-mapped allocations of one sequence-parallel group."""

    def __init__(self, world: int, rank: int, group, device):
        if world > _lib.IFX_MAX_PEERS:
            raise ValueError(f"peer exchange supports up to {_lib.IFX_MAX_PEERS} ranks (one NVSwitch box)")
        self.world, self.rank, self.group, self.device = world, rank, group, torch.device(device)
        self.epoch = 0
        self._opened: Dict[Tuple[int, int], int] = {}      # (source rank, its allocation base) -> base mapped here
        self.flags = torch.zeros(world, dtype=torch.int64, device=self.device)
        self.flag_ptrs: List[int] = []

    # ------------------------------------------------------------------ IPC
    def _export(self, t: torch.Tensor):
        h = (C.c_ubyte * _lib.IFX_PEER_HANDLE_BYTES)()
        off, base, size = C.c_int64(), C.c_uint64(), C.c_uint64()
        _lib.check(_lib.load().ifx_peer_export(t.data_ptr(), h, C.byref(off), C.byref(base), C.byref(size)))
        return bytes(h), off.value, base.value

    def exchange(self, tensors: Sequence[torch.Tensor]) -> List[List[int]]:
        """Collective.  For every tensor (same list, same order on all ranks) returns the `world` device pointers
        under which this rank can address that tensor on each rank (own entry: the local pointer)."""
        try:
            local, err = [self._export(t) for t in tensors], None
        except Exception as e:   # noqa: BLE001 — still take part in the collective below, then fail on every rank
            local, err = None, e
        gathered: list = [None] * self.world
        dist.all_gather_object(gathered, local, group=self.group)
        if err is not None:
            raise err
        if any(g is None for g in gathered):
            raise RuntimeError("a peer rank could not export its buffers")
        lib = _lib.load()
        out = []
        for i, t in enumerate(tensors):
            ptrs = []
            for src in range(self.world):
                if src == self.rank:
                    ptrs.append(t.data_ptr())
                    continue
                handle, off, base = gathered[src][i]
                key = (src, base)
                if key not in self._opened:
                    mapped = C.c_void_p()
                    _lib.check(lib.ifx_peer_open(handle, C.byref(mapped)))
                    self._opened[key] = mapped.value
                ptrs.append(self._opened[key] + off)
            out.append(ptrs)
        return out

    def peer_dst(self, k: torch.Tensor, v: torch.Tensor, k_ptrs: List[int], v_ptrs: List[int]) -> "_lib.PeerDst":
        d = _lib.PeerDst()
        d.world, d.rank, d.epoch = self.world, self.rank, 0
        for r in range(self.world):
            d.k[r], d.v[r], d.flags[r] = k_ptrs[r], v_ptrs[r], self.flag_ptrs[r]
        assert d.k[self.rank] == k.data_ptr() and d.v[self.rank] == v.data_ptr()
        return d

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def wait(self, epoch: int) -> None:
        _lib.check(_lib.load().ifx_peer_wait(self.flags.data_ptr(), self.world, epoch, WAIT_TIMEOUT_MS,
                                             torch.cuda.current_stream().cuda_stream))

    def release(self) -> None:
        """Unmap every peer allocation (call before the owners free their caches)."""
        if self._opened:
            torch.cuda.synchronize(self.device)
            lib = _lib.load()
            for mapped in self._opened.values():
                lib.ifx_peer_close(mapped)
            self._opened.clear()


def try_setup(stores, world: int, rank: int, group, device):
    """Collective.  Map every store's K / V buffers on every rank and attach `store.peer` (an ifx_peer_dst) and
    `store.peer_group`.  Returns the PeerGroup, or None (on ALL ranks) if any rank could not set the mapping up —
    the caller then keeps the NCCL all-gather path."""
    pg, err = None, ""
    try:
        pg = PeerGroup(world, rank, group, device)
        bufs = [pg.flags]
        for s in stores:
            bufs += [s.k, s.v]
        ptrs = pg.exchange(bufs)                      # the ONE collective before the vote below
        pg.flag_ptrs = ptrs[0]
        for i, s in enumerate(stores):
            s.peer = pg.peer_dst(s.k, s.v, ptrs[1 + 2 * i], ptrs[2 + 2 * i])
            s.peer_group = pg
    except Exception as e:   # noqa: BLE001 — any failure (no IPC in the container, no P2P) means "use NCCL"
        err = f"{type(e).__name__}: {e}"
    ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        return pg
    for s in stores:
        s.peer = s.peer_group = None
    if pg is not None:
        pg.release()
    if err and rank == 0:
        import warnings
        warnings.warn(f"inferix_b200: peer-memory KV exchange unavailable ({err}); using the NCCL all-gather path")
    return None


def setup_for_pipeline(model, kv_cache_manager, kv_cache_requests, parallel_config):
    """Called by the block schedulers right after the self-attention caches are allocated (collective under SP).
    Returns the PeerGroup or None (single GPU, IFX_SP_PEER=0, or mapping not possible -> NCCL all-gather path)."""
    pc = parallel_config
    if pc is None or pc.world_size == 1 or not ENABLED or not dist.is_initialized():
        return None
    stores = [blk.kv_cache_manager.store(kv_cache_manager, req) for blk in model.blocks for req in kv_cache_requests]
    return try_setup(stores, pc.world_size, pc.rank, pc.group, kv_cache_manager.device)
