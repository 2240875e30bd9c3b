"""CausVid per-layer cache adapter (inferix/kvcache_manager/model/causvid_kv_cache_manager.py).

Identical to the Self-Forcing adapter except ``get_kv_cache`` takes an explicit ``(start_index, length)`` range
(reference :110-127), because the CausVid block passes kv_start / kv_end instead of relying on end indices.
"""
from __future__ import annotations

import torch

from ..kvcache_manager import KVCacheManager, KVCacheRequest
from .self_forcing_kv_cache_manager import SelfForcingKVCacheManager


class CausVidKVCacheManager(SelfForcingKVCacheManager):
    def get_kv_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest, start_index: int = 0,
                     length: int = None) -> torch.Tensor:
        spec = kv_cache_manager.layer_spec(kv_cache_request, self.layer_name)
        if length is None:
            length = spec.num_tokens - start_index
        if start_index % spec.block_size or length % spec.block_size:
            raise ValueError("get_kv_cache: start_index / length must be whole pages")
        t = kv_cache_manager.get_range(kv_cache_request, self.layer_name, start_index // spec.block_size,
                                       length // spec.block_size)
        return t.reshape(2, length, spec.spec.num_kv_heads, spec.spec.head_size)


class KVCacheManagerFactory:
    @staticmethod
    def create_manager(layer_number: int, num_query_groups_per_partition: int, hidden_size_per_attention_head: int,
                       enable_kv_offload: bool) -> CausVidKVCacheManager:
        return CausVidKVCacheManager(layer_number, num_query_groups_per_partition, hidden_size_per_attention_head,
                                     enable_kv_offload)
