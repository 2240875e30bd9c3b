"""attention() / flash_attention() with the reference's signatures
(inferix/models/attention/flash_attention.py:42-56,153-167) on the tcgen05 kernel.

q [B, Lq, N, D], k/v [B, Lk, N, D] -> [B, Lq, N, D] in q's dtype.  Supported envelope = what the hot path uses:
bf16/fp16->bf16 compute, full (non-causal) attention, no dropout, no sliding window, Nq == Nk, D == 128.  Anything
else raises instead of silently taking another path (the reference would dispatch to FA3 / FA2 / SDPA).
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["flash_attention", "attention", "ifx_attn_forward", "collect_supported_attn", "CoreAttention"]


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
    # lengths are read once (one host sync for the whole batch, none when they are not given)
    q_len_list = None if q_lens is None else [int(x) for x in q_lens.tolist()]
    k_len_list = None if k_lens is None else [int(x) for x in k_lens.tolist()]
    for i in range(b):
        ql = lq if q_len_list is None else q_len_list[i]
        kl = k.size(1) if k_len_list is None else k_len_list[i]
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


# ----------------------------------------------------------------------------- distributed attention surface
def ifx_attn_forward(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                     alibi_slopes=None, return_softmax=False):
    """Attention backend with the calling convention of the reference's backend functions
    (inferix/models/attention/backends.py:58-72 `flash_attn_forward`): q [B, Lq, N, D], k / v [B, Lk, Nk, D] ->
    (out [B, Lq, N, D], lse [B, N, Lq] fp32).  The (out, lse) pair is what `update_out_and_lse`
    (distributed.py:27-46) merges across ring steps."""
    if causal or dropout_p != 0.0 or tuple(window_size) != (-1, -1) or softcap != 0.0 or alibi_slopes is not None:
        raise NotImplementedError("inferix_b200 attention backend: full, dropout-free attention only")
    b, lq, n, d = q.shape
    nk = k.shape[2]
    outs, lses = [], []
    for i in range(b):
        o, lse = ops.attention_lse(q[i].reshape(lq, n * d).to(torch.bfloat16).contiguous(),
                                   k[i].reshape(-1, nk * d).to(torch.bfloat16).contiguous(),
                                   v[i].reshape(-1, nk * d).to(torch.bfloat16).contiguous(), n, nk,
                                   softmax_scale=softmax_scale)
        outs.append(o.view(lq, n, d))
        lses.append(lse)
    return torch.stack(outs).to(q.dtype), torch.stack(lses)


def collect_supported_attn():
    """reference backends.py:154-166: name -> backend callable.  This library has exactly one backend."""
    return {"InferixB200": ifx_attn_forward}


class CoreAttention(torch.nn.Module):
    """Surface of the reference's `CoreAttention` (inferix/models/attention/distributed.py:53-330) for the strategies
    that survive the replicated-cache design:

      * "ulysses": all-to-all (scatter heads / gather sequence) over `ulysses_pg`, optional in-place write of the new
        K / V into caller-provided caches at `k_cache_offset`, attention over the cache prefix, inverse all-to-all —
        the arithmetic of distributed.py:175-268;
      * a group of size 1: plain attention over (cache prefix +) the given keys.

    The ring strategies ("pass_q" / "pass_kv": P - 1 P2P hops of the sharded cache with LSE merges, :564-712) are what
    the peer-memory exchange of `ifx_wan_block_forward_sp` replaces; asking for them with a ring group > 1 raises."""

    def __init__(self, scatter_idx: int = 2, gather_idx: int = 1, ring_impl_type: str = "basic",
                 use_pack_qkv: bool = False, attn_type=None, q_descale=None, k_descale=None, v_descale=None,
                 strategy: str = "auto", ulysses_pg=None, ring_pg=None):
        super().__init__()
        if q_descale is not None or k_descale is not None or v_descale is not None:
            raise NotImplementedError("descaled (FP8) attention inputs are not on the hot path")
        self.scatter_idx, self.gather_idx, self.use_pack_qkv = scatter_idx, gather_idx, use_pack_qkv
        self.strategy, self.attn_type = strategy, attn_type
        self.ulysses_pg, self.ring_pg = ulysses_pg, ring_pg
        self.supported_attn = collect_supported_attn()

    @staticmethod
    def _world(pg):
        import torch.distributed as dist
        return dist.get_world_size(pg) if (pg is not None and dist.is_initialized()) else 1

    def _select_strategy(self, query, key, value, k_cache=None, v_cache=None) -> str:
        if self.strategy != "auto":
            return self.strategy
        return "ulysses"        # the only distributed strategy kept; a size-1 group degenerates to local attention

    def _all_to_all(self, x, scatter_idx, gather_idx):
        """SeqAllToAll4D (yunchang): [B, S/P, N, D] -> [B, S, N/P, D] for (scatter 2, gather 1), and back."""
        import torch.distributed as dist
        p = self._world(self.ulysses_pg)
        if p == 1:
            return x
        b, s, n, d = x.shape
        if scatter_idx == 2 and gather_idx == 1:
            t = x.reshape(b, s, p, n // p, d).permute(2, 0, 1, 3, 4).contiguous()       # [P, B, S/P, N/P, D]
            out = torch.empty_like(t)
            dist.all_to_all_single(out, t, group=self.ulysses_pg)
            return out.permute(1, 0, 2, 3, 4).reshape(b, p * s, n // p, d)
        t = x.reshape(b, p, s // p, n, d).permute(1, 0, 2, 3, 4).contiguous()           # [P, B, S/P, N/P, D]
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t, group=self.ulysses_pg)
        return out.permute(1, 2, 0, 3, 4).reshape(b, s // p, p * n, d)

    def forward(self, query, key, value, k_cache=None, v_cache=None, k_cache_offset=0, v_cache_offset=0, *,
                dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), alibi_slopes=None,
                deterministic=False, return_attn_probs=False, custom_mask=None, q_ranges=None, k_ranges=None,
                attn_type_map=None, attn_backend=None):
        strategy = self._select_strategy(query, key, value, k_cache, v_cache)
        if strategy not in ("ulysses",) and self._world(self.ring_pg) > 1:
            raise NotImplementedError(
                f"CoreAttention strategy {strategy!r}: the ring pass-q / pass-kv exchange is replaced by the peer-memory "
                "exchange of the sequence-parallel block (ifx_wan_block_forward_sp); only 'ulysses' is built here")
        if custom_mask is not None or q_ranges is not None or isinstance(k_cache, list):
            raise NotImplementedError("masked / ranged / multi-cache CoreAttention calls are not on the Wan hot path")
        if query.shape[0] != 1:
            raise ValueError("CoreAttention: the batch size must be 1 (as in the reference)")
        q = self._all_to_all(query, self.scatter_idx, self.gather_idx)
        k = self._all_to_all(key, self.scatter_idx, self.gather_idx)
        v = self._all_to_all(value, self.scatter_idx, self.gather_idx)
        if k_cache is not None and v_cache is not None:             # distributed.py:197-203
            k_cache[0, k_cache_offset:k_cache_offset + k.shape[1]] = k[0]
            v_cache[0, v_cache_offset:v_cache_offset + v.shape[1]] = v[0]
            k = k_cache[:, :k_cache_offset + k.shape[1]]
            v = v_cache[:, :v_cache_offset + v.shape[1]]
        backend = self.supported_attn["InferixB200"]
        out, _lse = backend(q, k, v, dropout_p=dropout_p, softmax_scale=softmax_scale, causal=causal,
                            window_size=window_size)
        return self._all_to_all(out, self.gather_idx, self.scatter_idx)
