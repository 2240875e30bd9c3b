"""Sequence-parallel plumbing: ParallelConfig (reference wan_base/utils/parallel_config.py:3-30) and the token
layout helpers of the spatial-token split (reference causal_model.py:939-942 scatter, :1008-1022 gather).

Partitioning is the reference's: every rank owns ``hw / P`` tokens of EVERY frame.  What differs is the exchange:
instead of a ring that ships the whole sharded cache every layer (models/attention/distributed.py:564-712) each rank
all-gathers only the block's new roped-K / V (one NCCL all-gather per layer over NVLink) into a replicated paged
cache and attends its local queries against it — no LSE merge, no per-step P2P.  The helpers below are pure index
logic and run on any backend (the CPU tests drive them over gloo).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

ATTN_BACKEND = "ifx_tcgen05"


class ParallelConfig:
    def __init__(self, ulysses_size=1, ring_size=1, local_rank=0, rank=0, world_size=1, ring_strategy="pass-kv",
                 attn_backend: Optional[str] = None, group=None):
        if ulysses_size * ring_size not in (1, world_size) and world_size > 1:
            raise ValueError("ulysses_size * ring_size must equal world_size")
        self.ulysses_size = ulysses_size
        self.ring_size = ring_size
        self.local_rank = local_rank
        self.rank = rank
        self.world_size = world_size
        self.ring_strategy = ring_strategy
        # the reference auto-selects among library backends; this build has exactly one
        if attn_backend not in (None, ATTN_BACKEND):
            raise ValueError(f"Specified attention backend '{attn_backend}' is not available. "
                             f"Available backends: ['{ATTN_BACKEND}']")
        self.attn_backend = ATTN_BACKEND
        self.group = group


def scatter_tokens(x: torch.Tensor, frames: int, world_size: int, rank: int) -> torch.Tensor:
    """[B, F*hw, C] -> this rank's [B, F*(hw/P), C]  (causal_model.py:939-942)."""
    if world_size == 1:
        return x
    b, s, c = x.shape
    hw = s // frames
    if hw % world_size:
        raise ValueError(f"tokens per frame ({hw}) must divide by world_size ({world_size})")
    chunk = hw // world_size
    return x.view(b, frames, hw, c)[:, :, rank * chunk:(rank + 1) * chunk].reshape(b, frames * chunk, c)


def interleave_gathered(gathered: torch.Tensor, frames: int, world_size: int) -> torch.Tensor:
    """[P, B, F*(hw/P), C] (rank-major, as all_gather returns) -> [B, F*hw, C] in single-process token order:
    'b (cp f hw) c -> b (f cp hw) c'  (causal_model.py:1016-1021)."""
    p, b, s, c = gathered.shape
    chunk = s // frames
    return gathered.view(p, b, frames, chunk, c).permute(1, 2, 0, 3, 4).reshape(b, frames * p * chunk, c)


def all_gather_tokens(x: torch.Tensor, frames: int, cfg: ParallelConfig) -> torch.Tensor:
    """Final all-gather of the head output + re-interleave (causal_model.py:1008-1022)."""
    if cfg.world_size == 1:
        return x
    x = x.contiguous()
    out = torch.empty((cfg.world_size,) + tuple(x.shape), dtype=x.dtype, device=x.device)
    # dim-0 concatenation form: accepted by both NCCL and gloo
    dist.all_gather_into_tensor(out.view((-1,) + tuple(x.shape[1:])), x, group=cfg.group)
    return interleave_gathered(out, frames, cfg.world_size)


def all_gather_rows(x: torch.Tensor, cfg: ParallelConfig, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[rows, C] per rank -> [P, rows, C] (rank-major)."""
    if out is None:
        out = torch.empty((cfg.world_size,) + tuple(x.shape), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out.view((-1,) + tuple(x.shape[1:])), x.contiguous(), group=cfg.group)
    return out
