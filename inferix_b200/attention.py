"""attention() / flash_attention() with the reference's signatures
(inferix/models/attention/flash_attention.py:42-56,153-167) on the tcgen05 kernel.

q [B, Lq, N, D], k/v [B, Lk, N, D] -> [B, Lq, N, D] in q's dtype.  Supported envelope = what the hot path uses:
bf16/fp16->bf16 compute, full (non-causal) attention, no dropout, no sliding window, Nq == Nk, D == 128.  Anything
else raises instead of silently taking another path (the reference would dispatch to FA3 / FA2 / SDPA).
"""
from __future__ import annotations

from typing import Optional

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
    """Surface of the reference's `CoreAttention` (inferix/models/attention/distributed.py:53-712):

      * Ulysses over `ulysses_pg`: all-to-all (scatter heads / gather sequence), optional in-place write of the new
        K / V into caller-provided caches at `k_cache_offset`, attention, inverse all-to-all (:175-268);
      * ring over `ring_pg`, non-causal: "pass-kv" (K / V blocks travel P - 1 hops, partial outputs merged through
        their log-sum-exp, :564-712) and "pass-q" (queries travel, partials return by a ring reduce-scatter,
        :372-560), on the (out, lse) attention backend;
      * both groups of size 1: plain attention over (cache prefix +) the given keys.

    The sequence-parallel Wan block does NOT use the ring: its exchange is the peer-memory push inside the attention
    kernel (`ifx_wan_block_forward_sp`).  The ring strategies exist for callers of this class; they are
    host orchestration (torch.distributed P2P) around `ifx_attention_lse`."""

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
        """reference :92-130 (the size heuristic returns "pass_q" / "pass_kv" with an underscore, which the dispatch
        below — like the reference's, :209 — reads as pass-kv; only an explicit "pass-q" selects the query ring)."""
        if self.strategy != "auto":
            return self.strategy
        if k_cache is not None and v_cache is not None:
            return "ulysses"
        seq_len, num_heads = query.shape[1], query.shape[2]
        return "pass_q" if (seq_len > 2048 or (seq_len > 1024 and num_heads >= 16)) else "pass_kv"

    # ------------------------------------------------------------------ ring (reference :372-712)
    @staticmethod
    def _merge(out, lse, block_out, block_lse):
        """Fold one block's (out [B, L, N, D], lse [B, L, N, 1]) into the running pair: exact softmax merge written
        with sigmoid / logsigmoid (yunchang `update_out_and_lse`, the reference's `update_out_and_lse_pass_q` :29-46)."""
        block_out = block_out.to(torch.float32)
        if out is None:
            return block_out, block_lse
        out = out - torch.sigmoid(block_lse - lse) * (out - block_out)
        lse = lse - torch.nn.functional.logsigmoid(lse - block_lse)
        return out, lse

    @staticmethod
    def _ring_exchange(group, send: torch.Tensor, recv: Optional[torch.Tensor] = None):
        """Post send-to-next / receive-from-previous of one tensor; returns (recv buffer, requests)."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nxt = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
        prv = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
        if recv is None:
            recv = torch.empty(send.shape, dtype=send.dtype, device=send.device)
        ops_ = [dist.P2POp(dist.isend, send.contiguous(), nxt, group), dist.P2POp(dist.irecv, recv, prv, group)]
        if rank % 2:                                  # pair sends with receives in opposite order on odd ranks
            ops_.reverse()
        return recv, dist.batch_isend_irecv(ops_)

    def _backend(self, attn_backend):
        name = attn_backend if attn_backend is not None else "InferixB200"
        if name not in self.supported_attn:
            raise ValueError(f"Specified attention backend '{name}' is not available. "
                             f"Available backends: {list(self.supported_attn.keys())}")
        return self.supported_attn[name]

    @staticmethod
    def _check_ring_args(causal, dropout_p, custom_mask, q_ranges, k_ranges):
        if causal or dropout_p or custom_mask is not None or q_ranges is not None or k_ranges is not None:
            raise NotImplementedError("ring attention here is full, dropout-free attention (the Wan / MAGI hot path); "
                                      "causal / masked / ranged ring calls are not built")

    def ring_attention_forward_pass_kv(self, process_group, q, k, v, softmax_scale, k_cache_offset=0, v_cache_offset=0,
                                       dropout_p=0, deterministic=False, causal=False, window_size=(-1, -1),
                                       alibi_slopes=None, q_descale=None, k_descale=None, v_descale=None,
                                       custom_mask=None, q_ranges=None, k_ranges=None, attn_type_map=None,
                                       attn_backend=None):
        """K / V blocks travel around the ring; every hop's transfer is in flight while the block at hand is attended
        (reference :564-712).  q [B, Lq, N, D] stays; returns (out in q's dtype, lse [B, N, Lq] fp32)."""
        import torch.distributed as dist
        self._check_ring_args(causal, dropout_p, custom_mask, q_ranges, k_ranges)
        attn = self._backend(attn_backend)
        world = dist.get_world_size(process_group)
        out = lse = None
        for step in range(world):
            last = step + 1 == world
            if not last:
                next_k, req_k = self._ring_exchange(process_group, k)
                next_v, req_v = self._ring_exchange(process_group, v)
            block_out, block_lse = attn(q, k, v, dropout_p=dropout_p, softmax_scale=softmax_scale, causal=False,
                                        window_size=window_size)
            out, lse = self._merge(out, lse, block_out, block_lse.transpose(-2, -1).unsqueeze(-1))
            if not last:
                for r in req_k + req_v:
                    r.wait()
                k, v = next_k, next_v
        return out.to(q.dtype), lse.squeeze(-1).transpose(1, 2)

    def ring_attention_forward_pass_q(self, process_group, q, key, value, softmax_scale, k_cache_offset=0,
                                      v_cache_offset=0, dropout_p=0, deterministic=False, causal=False,
                                      window_size=(-1, -1), alibi_slopes=None, q_descale=None, k_descale=None,
                                      v_descale=None, custom_mask=None, q_ranges=None, k_ranges=None, attn_type_map=None,
                                      attn_backend=None):
        """Queries travel, K / V stay (reference :372-560): after P hops every rank holds, for every rank's query
        block, the partial over its own keys; a ring reduce-scatter of those partials returns each block to its
        owner merged over all keys.  Returns (out, lse) of the local queries, both in q's dtype and the lse in the
        merge layout [B, Lq, N, 1], like the reference (:560)."""
        import torch.distributed as dist
        self._check_ring_args(causal, dropout_p, custom_mask, q_ranges, k_ranges)
        attn = self._backend(attn_backend)
        world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        dtype = q.dtype
        outs, lses = [None] * world, [None] * world
        for step in range(world):
            last = step + 1 == world
            if not last:
                next_q, req = self._ring_exchange(process_group, q)
            block_out, block_lse = attn(q, key, value, dropout_p=dropout_p, softmax_scale=softmax_scale, causal=False,
                                        window_size=window_size)
            owner = (rank - step) % world                                   # whose queries these are
            outs[owner] = block_out.to(torch.float32)
            lses[owner] = block_lse.transpose(-2, -1).unsqueeze(-1).contiguous()
            if not last:
                for r in req:
                    r.wait()
                q = next_q
        for i in range(world - 1):                                           # ring reduce-scatter of the partials
            give, take = (rank - i - 1) % world, (rank - i - 2) % world
            got_out, req_o = self._ring_exchange(process_group, outs[give], torch.empty(
                outs[take].shape, dtype=outs[take].dtype, device=outs[take].device))
            got_lse, req_l = self._ring_exchange(process_group, lses[give], torch.empty(
                lses[take].shape, dtype=lses[take].dtype, device=lses[take].device))
            for r in req_o + req_l:
                r.wait()
            outs[take], lses[take] = self._merge(outs[take], lses[take], got_out, got_lse)
        return outs[rank].to(dtype), lses[rank].to(dtype)

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
        if softmax_scale is None:
            softmax_scale = 1.0 / (q.shape[-1] ** 0.5)
        lse = None
        if self._world(self.ring_pg) > 1:                            # distributed.py:206-267
            ring = (self.ring_attention_forward_pass_q if strategy in ("pass-q", "ulysses-pass-q")
                    else self.ring_attention_forward_pass_kv)
            out, lse = ring(self.ring_pg, q, k, v, softmax_scale=softmax_scale, k_cache_offset=k_cache_offset,
                            v_cache_offset=v_cache_offset, dropout_p=dropout_p, causal=causal, window_size=window_size,
                            alibi_slopes=alibi_slopes, deterministic=deterministic, attn_backend=attn_backend)
        else:
            out, lse = self._backend(attn_backend)(q, k, v, dropout_p=dropout_p, softmax_scale=softmax_scale,
                                                   causal=causal, window_size=window_size)
        out = self._all_to_all(out, self.gather_idx, self.scatter_idx)
        return (out, lse) if return_attn_probs else out
