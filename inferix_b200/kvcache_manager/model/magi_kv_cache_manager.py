"""MAGI-1 per-layer cache adapter with the reference's surface
(inferix/kvcache_manager/model/magi_kv_cache_manager.py:20-210) on the native paged cache.

Semantics (reference :76-187): the layer receives K and V interleaved per head, ``[tokens, kv_heads, 2*D]``; the cache
holds ``max_sequence_length`` tokens; a forward loads the clean history ``[0, slice_point * clip_token_nums)``,
optionally stores this forward's first ``clip_size`` tokens right behind it (``update_kv_cache``), and returns
``cat(history, new)`` split into K and V.  Here the de-interleave is one strided copy per side into the native cache
(`ifx_kv_import`), the history comes back through `ifx_kv_export`, and nothing is rearranged twice.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from ..kvcache_manager import KVCacheManager, KVCacheRequest, KVCacheRequestSpec, KVCacheSpec


class InferenceParams:
    """inferix/core/types/inference.py:88-101."""

    def __init__(self, max_batch_size, max_sequence_length, device=None):
        self.max_sequence_length = max_sequence_length
        self.max_batch_size = max_batch_size
        self.sequence_len_offset = 0
        self.kv_cache_request = KVCacheRequest(request_id="magi")
        self.kv_cache_manager = KVCacheManager(device=device if device is not None else torch.cuda.current_device())
        self.key_value_memory_dict = {}
        self.update_kv_cache = False


@dataclass
class KVMetaArgs:
    """The fields of ModelMetaArgs (core/types/inference.py:70-85) this adapter reads."""
    slice_point: int
    clip_token_nums: int
    extract_prefix_video_feature: bool = False
    fwd_extra_1st_chunk: bool = False
    distill_nearly_clean_chunk: bool = False


class MagiKVCacheManager:
    def __init__(self, layer_number: int, num_query_groups_per_partition: int, hidden_size_per_attention_head: int,
                 engine_config=None):
        self.layer_number = layer_number
        self.num_query_groups_per_partition = num_query_groups_per_partition
        self.hidden_size_per_attention_head = hidden_size_per_attention_head
        self.engine_config = engine_config

    @property
    def layer_name(self) -> str:
        return f"layer_{self.layer_number}"

    def allocate_key_value_memory(self, inference_params, sequence_length: int, batch_size: int, dtype) -> None:
        spec = KVCacheRequestSpec(
            num_tokens=sequence_length, block_size=1,
            specs={self.layer_name: KVCacheSpec(num_kv_heads=self.num_query_groups_per_partition,
                                                head_size=self.hidden_size_per_attention_head, dtype=dtype,
                                                kv_offload=bool(getattr(self.engine_config, "kv_offload", False)),
                                                use_mla=False)})
        inference_params.kv_cache_manager.allocate_slots(inference_params.kv_cache_request, spec)

    def native_store(self, inference_params, dtype=torch.bfloat16):
        """The layer's native cache (allocated on first use, reference :104-112).  The MAGI cache is only ever written
        in place (reference :110-146), so logical row i is physical row i: the native layer writes the new K / V rows
        at [start, start + n) through `store.map_rows` and attends rows [0, start + n) directly — the reference's
        get_range copy (:118-123) and torch.cat (:149) disappear."""
        mgr, req = inference_params.kv_cache_manager, inference_params.kv_cache_request
        if self.layer_name not in mgr.layers(req):
            self.allocate_key_value_memory(inference_params, inference_params.max_sequence_length,
                                           inference_params.max_batch_size, dtype)
        return mgr.store(req, self.layer_name)

    def _full_adjust_key_and_value(self, inference_params, key_and_value: torch.Tensor, meta_args):
        """reference :76-151.  Returns (key, value), each [history + new, kv_heads, D]."""
        mgr, req = inference_params.kv_cache_manager, inference_params.kv_cache_request
        hn, d = self.num_query_groups_per_partition, self.hidden_size_per_attention_head
        if self.layer_name not in mgr.layers(req):
            self.allocate_key_value_memory(inference_params, inference_params.max_sequence_length,
                                           inference_params.max_batch_size, key_and_value.dtype)
        store = mgr.store(req, self.layer_name)
        k_new = key_and_value[..., :d].reshape(-1, hn * d).contiguous()      # '(nb bls) hn (coef d)' -> coef ...
        v_new = key_and_value[..., d:].reshape(-1, hn * d).contiguous()
        start = meta_args.slice_point * meta_args.clip_token_nums * inference_params.max_batch_size
        if start > 0:
            k_hist, v_hist = store.export(0, start)
        else:
            k_hist = v_hist = k_new.new_empty((0, hn * d))
        if inference_params.update_kv_cache:
            clip = (k_new.shape[0] - meta_args.clip_token_nums * inference_params.max_batch_size
                    if meta_args.distill_nearly_clean_chunk else k_new.shape[0])
            assert start + clip <= inference_params.max_sequence_length
            if clip > 0:
                store.import_(start, k_new[:clip], v_new[:clip])
        key = torch.cat([k_hist, k_new], dim=0).view(-1, hn, d)
        value = torch.cat([v_hist, v_new], dim=0).view(-1, hn, d)
        return key, value

    def adjust_key_and_value_for_inference(self, key_and_value: torch.Tensor, inference_params,
                                           meta_args) -> Tuple[torch.Tensor, torch.Tensor]:
        """reference :153-187."""
        if inference_params is None:
            return torch.chunk(key_and_value, 2, dim=-1)
        if meta_args.extract_prefix_video_feature or meta_args.fwd_extra_1st_chunk or meta_args.slice_point > 0:
            return self._full_adjust_key_and_value(inference_params, key_and_value, meta_args)
        key, value = torch.chunk(key_and_value, 2, dim=-1)
        return key.contiguous(), value.contiguous()

    def clear_cache(self, inference_params) -> None:
        mgr, req = inference_params.kv_cache_manager, inference_params.kv_cache_request
        if self.layer_name in mgr.layers(req):
            mgr.free_layer(req, self.layer_name)

    def is_cached(self, inference_params) -> bool:
        return self.layer_name in inference_params.kv_cache_manager.layers(inference_params.kv_cache_request)
