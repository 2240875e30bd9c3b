"""Per-layer cache adapter with the reference's surface
(inferix/kvcache_manager/model/self_forcing_kv_cache_manager.py:8-285).

Layer names (``layer_{i}`` / ``crossattn_layer_{i}``), method names and arguments are the reference's.  Two
additions serve the native path: ``page_tokens`` on ``allocate_kv_cache`` (frame-sized pages instead of the
reference's block_size=1, :47) and ``store()`` which hands the block its ``PagedKV``.  Under sequence parallelism
the native cache is REPLICATED (every rank holds all tokens and all heads), so ``ring_size`` /
``ulysses_size`` do not shrink the allocation unless ``replicated=False`` is asked for.
"""
from __future__ import annotations

from typing import Optional

import torch

from ..kvcache_manager import KVCacheManager, KVCacheRequest, KVCacheRequestSpec, KVCacheSpec


class SelfForcingKVCacheManager:
    def __init__(self, layer_number: int, num_query_groups_per_partition: int, hidden_size_per_attention_head: int,
                 enable_kv_offload: bool = False):
        self.layer_number = layer_number
        self.num_query_groups_per_partition = num_query_groups_per_partition
        self.hidden_size_per_attention_head = hidden_size_per_attention_head
        self.enable_kv_offload = enable_kv_offload

    # -- names
    @property
    def layer_name(self) -> str:
        return f"layer_{self.layer_number}"

    @property
    def crossattn_layer_name(self) -> str:
        return f"crossattn_layer_{self.layer_number}"

    # -- allocation (reference :33-87)
    def allocate_kv_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest,
                          sequence_length: int, dtype: torch.dtype, ulysses_size: int = 1, ring_size: int = 1,
                          page_tokens: Optional[int] = None, replicated: bool = True) -> None:
        tokens = sequence_length if replicated else sequence_length // ring_size
        heads = self.num_query_groups_per_partition if replicated else self.num_query_groups_per_partition // ulysses_size
        spec = KVCacheRequestSpec(
            num_tokens=tokens,
            block_size=page_tokens or 1,
            specs={self.layer_name: KVCacheSpec(num_kv_heads=heads, head_size=self.hidden_size_per_attention_head,
                                                dtype=dtype, kv_offload=self.enable_kv_offload, use_mla=False)},
        )
        kv_cache_manager.allocate_slots(kv_cache_request, spec)

    def allocate_crossattn_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest,
                                 crossattn_length: int, dtype: torch.dtype) -> None:
        spec = KVCacheRequestSpec(
            num_tokens=crossattn_length,
            block_size=1,
            specs={self.crossattn_layer_name: KVCacheSpec(num_kv_heads=self.num_query_groups_per_partition,
                                                          head_size=self.hidden_size_per_attention_head, dtype=dtype,
                                                          kv_offload=self.enable_kv_offload, use_mla=False)},
        )
        kv_cache_manager.allocate_slots(kv_cache_request, spec)

    # -- native handles for the hot path
    def store(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest):
        return kv_cache_manager.store(kv_cache_request, self.layer_name)

    def crossattn_store(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest):
        return kv_cache_manager.store(kv_cache_request, self.crossattn_layer_name)

    # -- resets (reference :89-110 are no-ops; here the block table really is reset)
    def reset_kv_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest,
                       device: Optional[torch.device] = None) -> None:
        if self.layer_name in kv_cache_manager.layers(kv_cache_request):
            self.store(kv_cache_manager, kv_cache_request).reset()

    def reset_crossattn_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest) -> None:
        if self.crossattn_layer_name in kv_cache_manager.layers(kv_cache_request):
            pass   # validity of the text K/V lives in crossattn_cache_meta["is_init"], as in the reference

    # -- tensor access (reference :112-172)
    def get_kv_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest) -> torch.Tensor:
        spec = kv_cache_manager.layer_spec(kv_cache_request, self.layer_name)
        t = kv_cache_manager.get(kv_cache_request, self.layer_name)          # (2, nblk, blk, H, D)
        return t.reshape(2, spec.num_tokens, spec.spec.num_kv_heads, spec.spec.head_size)

    def set_kv_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest, start_index: int,
                     k_data: torch.Tensor, v_data: torch.Tensor) -> None:
        spec = kv_cache_manager.layer_spec(kv_cache_request, self.layer_name)
        if start_index % spec.block_size or k_data.shape[0] % spec.block_size:
            raise ValueError("set_kv_cache: start_index / length must be whole pages")
        combined = torch.stack([k_data, v_data], dim=0)
        kv_cache_manager.set(kv_cache_request, self.layer_name, start_index // spec.block_size,
                             k_data.shape[0] // spec.block_size, combined)

    def get_crossattn_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest) -> torch.Tensor:
        return kv_cache_manager.get(kv_cache_request, self.crossattn_layer_name)

    def set_crossattn_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest,
                            k_data: torch.Tensor, v_data: torch.Tensor) -> None:
        combined = torch.stack([k_data, v_data], dim=0).unsqueeze(2)
        kv_cache_manager.set(kv_cache_request, self.crossattn_layer_name, 0, combined.shape[1], combined)

    # -- bookkeeping (reference :174-220)
    def clear_cache(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest) -> None:
        if self.layer_name in kv_cache_manager.layers(kv_cache_request):
            kv_cache_manager.free_layer(kv_cache_request, self.layer_name)
        if self.crossattn_layer_name in kv_cache_manager.layers(kv_cache_request):
            kv_cache_manager.free_layer(kv_cache_request, self.crossattn_layer_name)

    def get_cache_size(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest) -> Optional[int]:
        if self.layer_name in kv_cache_manager.layers(kv_cache_request):
            s = kv_cache_manager.layer_spec(kv_cache_request, self.layer_name)
            return 2 * s.num_tokens * s.spec.num_kv_heads * s.spec.head_size
        return None

    def is_cached(self, kv_cache_manager: KVCacheManager, kv_cache_request: KVCacheRequest) -> bool:
        return self.layer_name in kv_cache_manager.layers(kv_cache_request)


class SelfForcingKVCacheManagerFactory:
    @staticmethod
    def create_manager(layer_number: int, num_query_groups_per_partition: int = 12,
                       hidden_size_per_attention_head: int = 128,
                       enable_kv_offload: bool = False) -> SelfForcingKVCacheManager:
        return SelfForcingKVCacheManager(layer_number, num_query_groups_per_partition,
                                         hidden_size_per_attention_head, enable_kv_offload)

    @staticmethod
    def create_managers(num_layers: int, num_query_groups_per_partition: int = 12,
                        hidden_size_per_attention_head: int = 128, enable_kv_offload: bool = False):
        return [SelfForcingKVCacheManagerFactory.create_manager(i, num_query_groups_per_partition,
                                                                hidden_size_per_attention_head, enable_kv_offload)
                for i in range(num_layers)]
