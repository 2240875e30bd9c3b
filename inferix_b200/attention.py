"""attention() / flash_attention() with the reference's signatures
(inferix/models/attention/flash_attention.py:42-56,153-167) on the tcgen05 kernel.

q [B, Lq, N, D], k/v [B, Lk, N, D] -> [B, Lq, N, D] in q's dtype.  Supported envelope = what the hot path uses:
bf16/fp16->bf16 compute, full (non-causal) attention, no dropout, no sliding window, Nq == Nk, D == 128.  Anything
else raises instead of silently taking another path (the reference would dispatch to FA3 / FA2 / SDPA).
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["flash_attention", "attention"]


def flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
                    window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
    half_dtypes = (torch.float16, torch.bfloat16)
    assert dtype in half_dtypes
    assert q.device.type == "cuda" and q.size(-1) <= 256
    if causal or dropout_p != 0. or tuple(window_size) != (-1, -1):
        raise NotImplementedError("inferix_b200 attention: only full, dropout-free attention is on the hot path")
    if q.size(2) != k.size(2) or q.size(-1) != 128:
        raise NotImplementedError("inferix_b200 attention: needs Nq == Nk and head_dim == 128")
    b, lq, n, d = q.shape
    out_dtype = q.dtype
    outs = []
    for i in range(b):
        ql = lq if q_lens is None else int(q_lens[i])
        kl = k.size(1) if k_lens is None else int(k_lens[i])
        qi = q[i, :ql].reshape(ql, n * d).to(torch.bfloat16)
        if q_scale is not None:
            qi = qi * q_scale
        ki = k[i, :kl].reshape(kl, n * d).to(torch.bfloat16).contiguous()
        vi = v[i, :kl].reshape(kl, n * d).to(torch.bfloat16).contiguous()
        o = ops.attention(qi.contiguous(), ki, vi, n, softmax_scale=softmax_scale)
        if ql < lq:
            o = torch.cat([o, o.new_zeros(lq - ql, n * d)])
        outs.append(o.view(lq, n, d))
    return torch.stack(outs).type(out_dtype)


def attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
              window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, fa_version=None):
    return flash_attention(q=q, k=k, v=v, q_lens=q_lens, k_lens=k_lens, dropout_p=dropout_p,
                           softmax_scale=softmax_scale, q_scale=q_scale, causal=causal, window_size=window_size,
                           deterministic=deterministic, dtype=dtype, version=fa_version)
