"""Shared by the CPU and GPU MAGI-layer tests: rebuild the reference-typed meta objects a golden file describes."""
import types

import numpy as np
import torch


def meta_from_plain(m: dict):
    """The fields of ModelMetaArgs (core/types/inference.py:70-85) the layer reads, from a golden's plain dict."""
    core = types.SimpleNamespace(np_q_range=np.asarray(m["q_range"]), np_k_range=np.asarray(m["k_range"]),
                                 q_range=torch.tensor(m["q_range"], dtype=torch.int32),
                                 k_range=torch.tensor(m["k_range"], dtype=torch.int32))
    cu_q = torch.tensor(m["cu_seqlens_q"], dtype=torch.int32)
    cu_k = torch.tensor(m["cu_seqlens_kv"], dtype=torch.int32)
    cross = types.SimpleNamespace(cu_seqlens_q=cu_q, cu_seqlens_kv=cu_k,
                                  q_ranges=torch.stack([cu_q[:-1], cu_q[1:]], 1),
                                  kv_ranges=torch.stack([cu_k[:-1], cu_k[1:]], 1))
    return types.SimpleNamespace(slice_point=m["slice_point"], denoising_range_num=m["denoising_range_num"],
                                 clip_token_nums=m["clip_token_nums"],
                                 extract_prefix_video_feature=m["extract_prefix_video_feature"],
                                 fwd_extra_1st_chunk=m["fwd_extra_1st_chunk"],
                                 distill_nearly_clean_chunk=m["distill_nearly_clean_chunk"], cp_split_sizes=None,
                                 core_attn_params=core, cross_attn_params=cross)
