"""KVCacheManager with the reference's API (inferix/kvcache_manager/kvcache_manager.py:20-244) on native paged storage.

Same dataclasses, same method names / arguments / exceptions.  What differs is the storage: the reference keeps one
torch tensor ``(2, num_blocks, block_size, H, D)`` per (request, layer) and moves bytes on every roll; here each
layer is a ``PagedKV`` (two bf16 buffers ``[num_blocks * block_size, H*D]`` plus the native block table), the
hot path appends / evicts / attends through the C ABI without ever materialising the reference layout, and the
tensor-returning methods (get / get_range / select / get_raw) gather a copy in the reference's logical order.

``kv_offload=True`` (the reference's pinned-CPU tier, :222-244; its model default).  A B200 holds the whole window
(15.9 GB for Self-Forcing 720p x 8 blocks of 180 GB), so by default the request is NOT honoured: the cache stays in HBM
and a one-time warning says so.  ``KVCacheManager(device, offload_tier=True)`` (or IFX_KV_OFFLOAD=1) honours it: paged
layer caches then live in pinned host memory and are staged through ``offload_slots`` device slots — layer i + 1 is
copied in on a side stream while layer i computes, the pages a forward writes are copied back (ops.PagedKV.stage /
write_back).  HBM use drops from all layers to `offload_slots` layers; PCIe carries the window once per layer per forward.
"""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass
from typing import KeysView, List, Optional, Sequence, Union

import torch

from ..ops import OffloadSlots, PagedKV


def cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


def align(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def get_dtype_size(dtype: torch.dtype) -> int:
    return torch.tensor([], dtype=dtype).element_size()


@dataclass(frozen=True)
class KVCacheRequest:
    request_id: str


@dataclass(frozen=True)
class KVCacheSpec:
    num_kv_heads: int
    head_size: int
    dtype: torch.dtype
    kv_offload: bool
    use_mla: bool


@dataclass
class KVCacheRequestSpec:
    num_tokens: int
    block_size: int
    specs: dict


@dataclass(frozen=True)
class KVCacheTensorSpec:
    size: int
    num_tokens: int
    num_blocks: int
    block_size: int
    spec: KVCacheSpec


@dataclass
class KVCaches:
    tensors: dict   # layer_name -> PagedKV (the reference stores torch tensors here)
    specs: dict     # layer_name -> KVCacheTensorSpec


class KVCacheManager:
    def __init__(self, device: Union[str, torch.device, int], offload_tier: Optional[bool] = None,
                 offload_slots: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError(f"inferix_b200.KVCacheManager needs a CUDA device, got {self.device} (no CPU path)")
        self.offload_device = torch.device("cpu")
        self.request_to_kv_caches: dict = {}
        self.offload_tier = (os.environ.get("IFX_KV_OFFLOAD", "0") == "1") if offload_tier is None else bool(offload_tier)
        if offload_slots < 2:
            raise ValueError("offload_slots must be >= 2 (one layer computing, one being staged)")
        self.offload_slots = offload_slots
        self._slots: dict = {}          # (rows, width) -> OffloadSlots
        self._warned_offload = False

    # ------------------------------------------------------------------ allocation (reference :62-125)
    def allocate_slots(self, req: KVCacheRequest, spec: KVCacheRequestSpec) -> KVCaches:
        num_blocks = cdiv(spec.num_tokens, spec.block_size)
        num_tokens_aligned = align(spec.num_tokens, spec.block_size)
        kv_caches = self.request_to_kv_caches.setdefault(req.request_id, KVCaches(tensors={}, specs={}))
        for layer_name in spec.specs:
            if layer_name in kv_caches.tensors:
                raise ValueError(f"Layer {layer_name} already exists")
        for layer_name, s in spec.specs.items():
            if s.use_mla:
                raise NotImplementedError("use_mla caches are not on the Wan / MAGI hot path")
            if s.dtype != torch.bfloat16:
                raise ValueError(f"the native cache is bf16 (production dtype of the reference path), got {s.dtype}")
            size = 2 * num_tokens_aligned * s.num_kv_heads * s.head_size * get_dtype_size(s.dtype)
            kv_caches.specs[layer_name] = KVCacheTensorSpec(size=size, num_tokens=num_tokens_aligned,
                                                            num_blocks=num_blocks, block_size=spec.block_size, spec=s)
            slots = None
            if s.kv_offload and spec.block_size > 1:          # paged (frame-sized pages) layer caches only
                if self.offload_tier:
                    key = (num_blocks * spec.block_size, s.num_kv_heads * s.head_size)
                    slots = self._slots.get(key)
                    if slots is None:
                        slots = self._slots[key] = OffloadSlots(self.offload_slots, key[0], key[1], self.device)
                elif not self._warned_offload:
                    self._warned_offload = True
                    warnings.warn("inferix_b200: kv_offload=True requested but the KV window stays in HBM (it fits a "
                                  "B200); pass KVCacheManager(device, offload_tier=True) or IFX_KV_OFFLOAD=1 to keep it "
                                  "in pinned host memory with device staging slots", stacklevel=2)
            kv_caches.tensors[layer_name] = PagedKV(num_blocks, spec.block_size, s.num_kv_heads, s.head_size,
                                                    self.device, offload=slots)
        return kv_caches

    def free(self, req: KVCacheRequest):
        caches = self.request_to_kv_caches.pop(req.request_id)   # KeyError like the reference's `del`
        for store in caches.tensors.values():
            store.free()

    def free_layer(self, req: KVCacheRequest, layer_name: str):
        caches = self.request_to_kv_caches[req.request_id]
        caches.tensors.pop(layer_name).free()
        del caches.specs[layer_name]

    # ------------------------------------------------------------------ native access used by the hot path
    def store(self, req: KVCacheRequest, layer_name: str) -> PagedKV:
        return self.request_to_kv_caches[req.request_id].tensors[layer_name]

    # ------------------------------------------------------------------ tensor views (reference :127-184)
    def _materialise(self, req, layer_name, start_block: int, num_blocks: int) -> torch.Tensor:
        store = self.store(req, layer_name)
        spec = self.layer_spec(req, layer_name)
        bs, h, d = spec.block_size, spec.spec.num_kv_heads, spec.spec.head_size
        out = torch.zeros((2, num_blocks * bs, h * d), dtype=torch.bfloat16, device=self.device)
        _, _, table = store.state()
        mapped = len(table) * store.page_tokens       # native pages (frame-sized after PagedKV.repage), not `bs`
        lo, hi = start_block * bs, min((start_block + num_blocks) * bs, mapped)
        if hi > lo:   # blocks never written are "uninitialised" in the reference; they read as zeros here
            k, v = store.export(lo, hi - lo)
            out[0, : hi - lo], out[1, : hi - lo] = k, v
        return out.view(2, num_blocks, bs, h, d)

    def select(self, req: KVCacheRequest, layer_name: str, block_indices: List[int]):
        full = self.get(req, layer_name)
        return full[:, block_indices, ...]

    def layers(self, req: KVCacheRequest) -> KeysView[str] | Sequence[str]:
        if req.request_id not in self.request_to_kv_caches:
            return ()
        return self.request_to_kv_caches[req.request_id].tensors.keys()

    def get(self, req: KVCacheRequest, layer_name: str):
        return self._materialise(req, layer_name, 0, self.layer_spec(req, layer_name).num_blocks)

    def get_range(self, req: KVCacheRequest, layer_name: str, start: int, length: int) -> torch.Tensor:
        return self._materialise(req, layer_name, start, length)

    def get_raw(self, req: KVCacheRequest, layer_name: str):
        return self.get(req, layer_name)

    def get_range_raw(self, req: KVCacheRequest, layer_name: str, start: int, length: int):
        return self._materialise(req, layer_name, start, length)

    def layer_spec(self, req: KVCacheRequest, layer_name: str):
        return self.request_to_kv_caches[req.request_id].specs[layer_name]

    # ------------------------------------------------------------------ partial set (reference :192-220)
    def set(self, req: KVCacheRequest, layer_name: str, start: int, size: int, new_kv: torch.Tensor) -> None:
        spec = self.request_to_kv_caches[req.request_id].specs[layer_name]
        assert len(new_kv) == 2
        store = self.store(req, layer_name)
        width = spec.spec.num_kv_heads * spec.spec.head_size
        rows = size * spec.block_size
        k = new_kv[0].to(device=self.device, dtype=torch.bfloat16).reshape(rows, width).contiguous()
        v = new_kv[1].to(device=self.device, dtype=torch.bfloat16).reshape(rows, width).contiguous()
        store.import_(start * spec.block_size, k, v)
