"""`UlyssesScheduler` with the reference's surface (inferix/distributed/parallelism/context_parallel.py:382-598): the
generic Ulysses attention pipeline driven by caller-supplied closures — sequence-sharded q / k / v in, all-to-all to
head shards, KV-cache hook, core attention in `overlap_degree` query-head chunks whose output all-to-all overlaps the
next chunk's attention, cross-attention under the last exchange, sequence-sharded output.

The native MAGI layer does not go through this class: `magi_layer.py` writes q / k / v directly in the send layout
(`ifx_magi_qkv_post`) and receives K / V straight into the cache rows (`magi_cp.py`).  This class is the reference's
interface for callers that bring their own closures (any device; `core_attn_func` may be `ops.attention_gqa`).
Tensors are [seq, heads, head_dim] like the reference's.  Pinned bit-exact to the reference class on a 2-rank gloo run
(tests/golden/ulysses_sched.pt, oracle/make_golden_ulysses.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import magi_cp
from .magi_cp import FakeHandle


def _heads_to_ranks(t: torch.Tensor, cp: int) -> torch.Tensor:
    """[seq, cp * hn, hd] -> [cp * seq, hn, hd]: block c = the heads destined for rank c ("seq (cp hn) hd -> (cp seq) hn hd")."""
    seq, heads, hd = t.shape
    return t.reshape(seq, cp, heads // cp, hd).transpose(0, 1).reshape(cp * seq, heads // cp, hd).contiguous()


def _repeat_kv_heads(t: torch.Tensor, cp: int) -> torch.Tensor:
    heads = t.shape[1]
    if cp % heads == 0 and cp != heads:                       # fewer KV heads than ranks: every rank gets a copy (:396)
        t = torch.repeat_interleave(t, repeats=cp // heads, dim=1)
    return t


def all_to_all_input_split(tensor: torch.Tensor, cp_split_sizes: Optional[Sequence[int]]):
    """Scatter heads, gather sequence (:382-402): [seq_local, cp * hn, hd] -> ([sum(split), hn, hd], handle)."""
    cp = magi_cp.get_cp_world_size()
    if cp == 1:
        return tensor, FakeHandle()
    assert cp_split_sizes is not None
    send = _heads_to_ranks(_repeat_kv_heads(tensor, cp).contiguous(), cp)
    out = torch.empty((sum(cp_split_sizes),) + tuple(send.shape[1:]), device=send.device, dtype=send.dtype)
    handle = dist.all_to_all_single(out, send, output_split_sizes=list(cp_split_sizes), group=magi_cp.get_cp_group(),
                                    async_op=True)
    return out, handle


def all_to_all_output_split(tensor: torch.Tensor, cp_split_sizes: Optional[Sequence[int]]):
    """Scatter sequence, gather heads (:405-424): [sum(split), hn, hd] -> ([cp * seq_local, hn, hd], handle), block c of
    the result = the heads rank c computed."""
    cp = magi_cp.get_cp_world_size()
    if cp == 1:
        return tensor, FakeHandle()
    assert cp_split_sizes is not None and tensor.is_contiguous()
    local = cp_split_sizes[magi_cp.get_cp_rank()]
    out = torch.empty((local * cp,) + tuple(tensor.shape[1:]), device=tensor.device, dtype=tensor.dtype)
    handle = dist.all_to_all_single(out, tensor, input_split_sizes=list(cp_split_sizes), group=magi_cp.get_cp_group(),
                                    async_op=True)
    return out, handle


def fused_qkv_communication(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cp_split_sizes: Optional[Sequence[int]]):
    """One blocking all-to-all for q | k | v together (:427-450): fewer launches when the sequence is short."""
    cp = magi_cp.get_cp_world_size()
    if cp == 1:
        return q, k, v
    assert cp_split_sizes is not None
    parts = [_heads_to_ranks(q.contiguous(), cp), _heads_to_ranks(_repeat_kv_heads(k, cp).contiguous(), cp),
             _heads_to_ranks(_repeat_kv_heads(v, cp).contiguous(), cp)]
    widths = [p.shape[1] for p in parts]
    send = torch.cat(parts, dim=1).contiguous()
    out = torch.empty((sum(cp_split_sizes),) + tuple(send.shape[1:]), device=send.device, dtype=send.dtype)
    dist.all_to_all_single(out, send, output_split_sizes=list(cp_split_sizes), group=magi_cp.get_cp_group(),
                           async_op=False)
    return torch.split(out, widths, dim=1)


class UlyssesScheduler:
    """reference :453-598.  All methods are static, like the reference's."""

    @staticmethod
    def get_attn_and_xattn_with_comm_overlap(get_q_func: Callable, get_k_func: Callable, get_v_func: Callable,
                                             kv_cache_func: Callable, core_attn_func: Callable, cross_attn_func: Callable,
                                             overlap_degree: int, batch_size: int, cp_size: int,
                                             cp_split_sizes: List[int] = None):
        """v, k, q are produced in that order and each exchange starts as soon as its tensor exists, so the k and q
        computations (and the cache update) hide the v, k and q transfers (:468-499)."""
        value, wait_v = all_to_all_input_split(get_v_func(), cp_split_sizes)
        key, wait_k = all_to_all_input_split(get_k_func(), cp_split_sizes)
        query, wait_q = all_to_all_input_split(get_q_func(), cp_split_sizes)
        wait_v.wait()
        wait_k.wait()
        key, value = kv_cache_func(torch.concat([key, value], dim=-1))
        wait_q.wait()
        return UlyssesScheduler.get_attn_and_xattn_base(query, key, value, core_attn_func, cross_attn_func,
                                                        overlap_degree, batch_size, cp_size, cp_split_sizes)

    @staticmethod
    def get_attn_and_xattn_with_fused_kv_comm(get_q_func: Callable, get_kv_func: Callable, kv_cache_func: Callable,
                                              core_attn_func: Callable, cross_attn_func: Callable, overlap_degree: int,
                                              batch_size: int, cp_size: int, cp_split_sizes: List[int] = None):
        """K | V travel as one tensor (:501-526)."""
        kv, wait_kv = all_to_all_input_split(get_kv_func(), cp_split_sizes)
        query, wait_q = all_to_all_input_split(get_q_func(), cp_split_sizes)
        wait_kv.wait()
        key, value = kv_cache_func(kv)
        wait_q.wait()
        return UlyssesScheduler.get_attn_and_xattn_base(query, key, value, core_attn_func, cross_attn_func,
                                                        overlap_degree, batch_size, cp_size, cp_split_sizes)

    @staticmethod
    def get_attn_and_xattn_with_fused_qkv_comm(get_qkv_func: Callable, kv_cache_func: Callable, core_attn_func: Callable,
                                               cross_attn_func: Callable, overlap_degree: int, batch_size: int,
                                               cp_size: int, cp_split_sizes: List[int] = None):
        """q | k | v in one blocking exchange (:528-547)."""
        q, k, v = fused_qkv_communication(*get_qkv_func(), cp_split_sizes)
        k, v = kv_cache_func(torch.cat([k, v], dim=-1))
        return UlyssesScheduler.get_attn_and_xattn_base(q, k, v, core_attn_func, cross_attn_func, overlap_degree,
                                                        batch_size, cp_size, cp_split_sizes)

    @staticmethod
    def get_attn_and_xattn_base(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, core_attn_func: Callable,
                                cross_attn_func: Callable, overlap_degree: int, batch_size: int, cp_size: int,
                                cp_split_sizes: List[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Core attention in `overlap_degree` chunks of query heads (every chunk keeps all KV heads, so GQA groups stay
        whole); the output exchange of chunk i runs under the attention of chunk i + 1, the last one under the
        cross-attention (:549-598).  Returns ([seq_local, batch, cp * heads * head_dim], cross-attention output)."""
        q_seq, q_heads, hd = query.shape
        kv_heads = key.shape[1]
        if overlap_degree == -1:
            overlap_degree = q_heads // kv_heads
        else:
            assert overlap_degree <= q_heads
        if overlap_degree == 1:
            chunks = [query]
        elif kv_heads == 1:
            chunks = list(query.chunk(overlap_degree, dim=1))
        else:
            assert q_heads % (overlap_degree * kv_heads) == 0
            grouped = query.reshape(q_seq, kv_heads, -1, hd)
            chunks = [c.reshape(q_seq, -1, hd) for c in grouped.chunk(overlap_degree, dim=2)]

        done, in_flight, wait = [], None, None
        for chunk in chunks:
            fresh = core_attn_func(chunk, key, value)
            if wait is not None:
                wait.wait()
                done.append(in_flight)
            in_flight, wait = all_to_all_output_split(fresh, cp_split_sizes)
        xattn_out = cross_attn_func()
        wait.wait()
        done.append(in_flight)
        out = torch.cat(done, dim=1)                                   # [(cp sq b), heads_local, hd], chunk-major heads
        heads_local = out.shape[1]
        sq = out.shape[0] // (cp_size * batch_size)
        out = out.reshape(cp_size, sq, batch_size, heads_local, hd).permute(1, 2, 0, 3, 4)
        return out.reshape(sq, batch_size, cp_size * heads_local * hd), xattn_out
