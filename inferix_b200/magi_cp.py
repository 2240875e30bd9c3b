"""MAGI-1 context parallelism, Ulysses flavour — the reference's surface
(inferix/distributed/parallelism/context_parallel.py:30-87, 135-255, 309-376, 382-424 and the cp getters of
inferix/distributed/parallel_state.py:498-634) on `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).

Layout contract (what the native layer relies on):
  * the sequence is split contiguously, rank r owning `cp_split_sizes[r]` tokens (uneven allowed, :241-243);
  * "input split" = scatter heads, gather sequence: every rank ends up with ALL tokens of `heads / cp` heads, tokens in
    global order; "output split" is the inverse;
  * cross-attention stays local to the sequence shard; its (q, kv) ranges are re-cut per rank (:135-225).
The reference rearranges "seq (cp hn) hd -> (cp seq) hn hd" before the all-to-all (:397); here `ifx_magi_qkv_post`
writes that send layout directly and the receive side of K / V is the layer's KV-cache rows, so neither side of the
exchange costs an extra pass.  `cp_shuffle_overlap` (the pre-Hopper strategy, :258-306) is not offered.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

_CP_GROUP = None
_CP_SIZE = 1
_CP_RANK = 0


def init_context_parallel(group=None, size: Optional[int] = None, rank: Optional[int] = None) -> None:
    """parallel_state.initialize_model_parallel for the one axis this path shards (cp); tp = pp = 1."""
    global _CP_GROUP, _CP_SIZE, _CP_RANK
    _CP_GROUP = group
    _CP_SIZE = size if size is not None else dist.get_world_size(group)
    _CP_RANK = rank if rank is not None else dist.get_rank(group)


def destroy_context_parallel() -> None:
    global _CP_GROUP, _CP_SIZE, _CP_RANK
    _CP_GROUP, _CP_SIZE, _CP_RANK = None, 1, 0


def get_cp_group():
    return _CP_GROUP


def get_cp_world_size() -> int:
    return _CP_SIZE


def get_cp_rank() -> int:
    return _CP_RANK


class FakeHandle:
    def wait(self):
        pass


# ----------------------------------------------------------------------------- split / gather (:30-87)
def ulysses_split_sizes(seq_len: int, cp_size: int) -> List[int]:
    """cp_ulysses_process part 1 (:241-243)."""
    sizes = [seq_len // cp_size] * cp_size
    for i in range(seq_len % cp_size):
        sizes[i] += 1
    return sizes


def scatter_to_context_parallel_region(input_: torch.Tensor, cp_split_sizes: Sequence[int]) -> torch.Tensor:
    if get_cp_world_size() == 1:
        return input_
    off = sum(cp_split_sizes[:get_cp_rank()])
    return input_[off:off + cp_split_sizes[get_cp_rank()]].contiguous()


def gather_from_context_parallel_region(input_: torch.Tensor, cp_split_sizes: Sequence[int]) -> torch.Tensor:
    if get_cp_world_size() == 1:
        return input_
    input_ = input_.contiguous()
    out = torch.empty((sum(cp_split_sizes),) + tuple(input_.shape[1:]), dtype=input_.dtype, device=input_.device)
    dist.all_gather(list(torch.split(out, list(cp_split_sizes), dim=0)), input_, group=get_cp_group())
    return out


# ----------------------------------------------------------------------------- cross-attention ranges (:135-225)
def cp_update_cross_attn_qkv_range(cu_seqlens_q: Sequence[int], cu_seqlens_kv: Sequence[int],
                                   cp_split_sizes: Sequence[int], cp_rank: Optional[int] = None
                                   ) -> Tuple[List[List[int]], List[List[int]]]:
    """Batch 1, no shuffle / padding: intersect every query segment with this rank's token interval and re-base to
    the rank's first token.  Returns host lists (q_ranges, kv_ranges) of [start, end) — the integers the reference
    puts into PackedCrossAttnParams.q_ranges / kv_ranges."""
    r = get_cp_rank() if cp_rank is None else cp_rank
    lo = sum(cp_split_sizes[:r])
    hi = lo + cp_split_sizes[r]
    q_ranges, k_ranges = [], []
    for i in range(len(cu_seqlens_q) - 1):
        s, e = max(lo, int(cu_seqlens_q[i])), min(hi, int(cu_seqlens_q[i + 1]))
        if s < e:
            q_ranges.append([s - lo, e - lo])
            k_ranges.append([int(cu_seqlens_kv[i]), int(cu_seqlens_kv[i + 1])])
    if q_ranges:
        off = min(r_[0] for r_ in q_ranges)
        q_ranges = [[s - off, e - off] for s, e in q_ranges]
    return q_ranges, k_ranges


def cp_ulysses_process(cp_size: int, x: torch.Tensor, condition_map: torch.Tensor, rope: torch.Tensor,
                       cu_seqlens_q: Sequence[int], cu_seqlens_kv: Sequence[int]):
    """:228-255.  x [S, N, D]; returns this rank's shards, the split sizes and its cross-attention ranges."""
    seq_len = x.shape[0]
    assert seq_len == rope.size(0) and condition_map.size(0) == seq_len
    split = ulysses_split_sizes(seq_len, cp_size)
    x = scatter_to_context_parallel_region(x, split)
    condition_map = scatter_to_context_parallel_region(condition_map, split)
    rope = scatter_to_context_parallel_region(rope, split)
    q_ranges, k_ranges = cp_update_cross_attn_qkv_range(cu_seqlens_q, cu_seqlens_kv, split)
    return x, condition_map, rope, split, (q_ranges, k_ranges)


def cp_post_process(cp_size: int, cp_strategy: str, x: torch.Tensor, cp_split_sizes) -> torch.Tensor:
    """:365-376."""
    if cp_size == 1:
        return x
    if cp_strategy != "cp_ulysses":
        raise ValueError(f"Invalid CP strategy: {cp_strategy}, expected cp_ulysses")
    return gather_from_context_parallel_region(x, cp_split_sizes)


# ----------------------------------------------------------------------------- all-to-all (:382-424)
def all_to_all_input_split(send: torch.Tensor, cp_split_sizes: Sequence[int], out: Optional[torch.Tensor] = None,
                           async_op: bool = True):
    """Scatter heads, gather sequence.  `send` is already in the "(cp seq) hn*hd" layout: [cp, seq_local, W]
    (W = heads-per-rank * head_dim).  Returns ([sum(split), W] in global token order, work handle).
    `out` may be any contiguous [sum(split), W] view — the layer passes its KV-cache rows."""
    cp = get_cp_world_size()
    if cp == 1:
        return send.reshape(-1, send.shape[-1]), FakeHandle()
    assert send.is_contiguous() and send.dim() == 3 and send.shape[0] == cp
    total = sum(cp_split_sizes)
    if out is None:
        out = torch.empty((total, send.shape[-1]), dtype=send.dtype, device=send.device)
    assert out.is_contiguous() and out.shape == (total, send.shape[-1])
    handle = dist.all_to_all_single(out, send.view(-1, send.shape[-1]), output_split_sizes=list(cp_split_sizes),
                                    input_split_sizes=[send.shape[1]] * cp, group=get_cp_group(), async_op=async_op)
    return out, (handle if async_op else FakeHandle())


def all_to_all_output_split(full: torch.Tensor, cp_split_sizes: Sequence[int], async_op: bool = True):
    """Scatter sequence, gather heads: [sum(split), W] -> [cp, seq_local, W] (source-rank-major = head-group-major)."""
    cp = get_cp_world_size()
    if cp == 1:
        return full[None], FakeHandle()
    assert full.is_contiguous()
    local = cp_split_sizes[get_cp_rank()]
    out = torch.empty((cp, local, full.shape[-1]), dtype=full.dtype, device=full.device)
    handle = dist.all_to_all_single(out.view(-1, full.shape[-1]), full, output_split_sizes=[local] * cp,
                                    input_split_sizes=list(cp_split_sizes), group=get_cp_group(), async_op=async_op)
    return out, (handle if async_op else FakeHandle())
