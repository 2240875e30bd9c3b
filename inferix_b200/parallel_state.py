"""Process-group bookkeeping with the surface of the reference's `inferix/distributed/parallel_state.py:236-634`
(Megatron-style tp / cp / pp / dp groups), for callers that reach the MAGI context-parallel path through
`mpu.get_cp_group()` & co. instead of `magi_cp.init_context_parallel`.

Ranks form a mixed-radix grid whose axes are named by `order` (first name = fastest varying, default "tp-cp-pp-dp",
dist_utils.py:78); the group of a set of axes is every set of ranks that agree on all the other axes.  The native path
shards one axis only (cp, Ulysses); tp and pp sizes other than 1 are accepted for the group arithmetic (so that rank
lists can be compared with the reference's `RankGenerator`) but `initialize_model_parallel` refuses to build a
pipeline- or tensor-parallel run, which this library does not implement.
"""
from __future__ import annotations

from datetime import timedelta
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch.distributed as dist

_AXES = ("tp", "cp", "pp", "dp")
# the group kinds the reference creates (:320-423), in its creation order (new_group is collective: same order on all ranks)
_KINDS = ("dp", "dp-cp", "cp", "tp-pp", "tp", "tp-cp", "pp", "tp-cp-dp", "tp-dp")
_GLOO_TWINS = ("dp", "dp-cp")


def rank_groups(sizes: Dict[str, int], order: str, token: str) -> List[List[int]]:
    """All groups of the axes named in `token` ("cp", "tp-pp", ...): ranks ascending inside a group, groups ordered by
    their first rank.  Equals `RankGenerator(...).get_ranks(token)` of the reference (golden-tested)."""
    names = [n for n in order.lower().split("-") if n]
    for n in _AXES:
        if n not in names:
            if sizes.get(n, 1) != 1:
                raise RuntimeError(f"The size of ({n}) is ({sizes[n]}), but you haven't specified the order ({order}).")
            names.append(n)
    dims = [int(sizes.get(n, 1)) for n in names]                    # fastest axis first
    grid = np.arange(int(np.prod(dims))).reshape(dims[::-1])       # numpy: slowest axis first
    axis_of = {n: len(names) - 1 - i for i, n in enumerate(names)}
    inside = sorted(axis_of[t] for t in token.lower().split("-"))
    outside = [a for a in range(len(names)) if a not in inside]
    width = int(np.prod([grid.shape[a] for a in inside])) if inside else 1
    return grid.transpose(outside + inside).reshape(-1, width).tolist()


class _State:
    def __init__(self):
        self.groups: Dict[str, Tuple[object, List[int]]] = {}       # kind -> (process group, global ranks)
        self.gloo: Dict[str, object] = {}
        self.sizes: Dict[str, int] = {}


_STATE: Optional[_State] = None


def initialize_model_parallel(tp_size: int = 1, pp_size: int = 1, cp_size: int = 1,
                              nccl_communicator_config_path: Optional[str] = None,
                              distributed_timeout_minutes: int = 30, order: str = "tp-cp-pp-dp") -> None:
    """reference :236-423.  Builds the cp / dp / (size-1) tp / pp groups over the default process group and binds the
    cp group to the MAGI context-parallel path."""
    global _STATE
    assert dist.is_initialized()
    if _STATE is not None:
        raise AssertionError("model parallel groups are already initialized")
    if tp_size != 1 or pp_size != 1:
        raise NotImplementedError("inferix_b200 shards the sequence only (cp_ulysses); tp_size and pp_size must be 1")
    if nccl_communicator_config_path is not None:
        raise NotImplementedError("per-communicator NCCL options are not supported")
    world, rank = dist.get_world_size(), dist.get_rank()
    if world % (tp_size * pp_size * cp_size):
        raise RuntimeError(f"world_size ({world}) is not divisible by tp_size ({tp_size}) x pp_size ({pp_size}) "
                           f"x cp_size ({cp_size})")
    st = _State()
    st.sizes = {"tp": tp_size, "cp": cp_size, "pp": pp_size, "dp": world // (tp_size * pp_size * cp_size)}
    timeout = timedelta(minutes=distributed_timeout_minutes)
    for kind in _KINDS:
        for ranks in rank_groups(st.sizes, order, kind):
            group = dist.new_group(ranks, timeout=timeout)
            gloo = dist.new_group(ranks, timeout=timeout, backend="gloo") if kind in _GLOO_TWINS else None
            if rank in ranks:
                st.groups[kind] = (group, ranks)
                if gloo is not None:
                    st.gloo[kind] = gloo
    _STATE = st
    from . import magi_cp
    magi_cp.init_context_parallel(get_cp_group(), get_cp_world_size(), get_cp_rank())


def destroy_model_parallel() -> None:
    """reference :636-671 (drops the module state; the process groups die with the default group)."""
    global _STATE
    _STATE = None
    from . import magi_cp
    magi_cp.destroy_context_parallel()


def is_initialized() -> bool:
    return _STATE is not None


def is_unitialized() -> bool:          # (sic) reference :431
    return _STATE is None


def model_parallel_is_initialized() -> bool:
    return _STATE is not None


def _get(kind: str, check_initialized: bool = True):
    if _STATE is None:
        if check_initialized:
            raise AssertionError(f"{kind} parallel group is not initialized")
        return None, None
    return _STATE.groups[kind]


def _size(kind: str) -> int:
    return dist.get_world_size(group=_get(kind)[0])


def _rank(kind: str) -> int:
    return dist.get_rank(group=_get(kind)[0])


def get_model_parallel_group():
    return _get("tp-pp")[0]


def get_tp_group(check_initialized=True, with_context_parallel=False):
    return _get("tp-cp" if with_context_parallel else "tp", check_initialized)[0]


def get_pp_group():
    return _get("pp")[0]


def get_dp_group(with_context_parallel=False):
    return _get("dp-cp" if with_context_parallel else "dp")[0]


def get_dp_group_gloo(with_context_parallel=False):
    _get("dp")
    return _STATE.gloo["dp-cp" if with_context_parallel else "dp"]


def get_cp_group(check_initialized=True):
    return _get("cp", check_initialized)[0]


def get_tp_world_size(with_context_parallel=False) -> int:
    return _size("tp-cp" if with_context_parallel else "tp")


def get_pp_world_size() -> int:
    return _size("pp")


def get_tp_rank(with_context_parallel=False) -> int:
    return _rank("tp-cp" if with_context_parallel else "tp")


def get_pp_rank() -> int:
    return _rank("pp")


def is_pipeline_first_stage() -> bool:
    return get_pp_rank() == 0


def is_pipeline_last_stage() -> bool:
    return get_pp_rank() == get_pp_world_size() - 1


def get_tensor_model_parallel_ranks(with_context_parallel=False) -> Sequence[int]:
    return _get("tp-cp" if with_context_parallel else "tp")[1]


def get_tensor_model_parallel_src_rank(with_context_parallel=False) -> int:
    return get_tensor_model_parallel_ranks(with_context_parallel)[0]


def get_tensor_model_parallel_last_rank(with_context_parallel=False) -> int:
    return get_tensor_model_parallel_ranks(with_context_parallel)[-1]


def get_pipeline_model_parallel_first_rank() -> int:
    return _get("pp")[1][0]


def get_pipeline_model_parallel_last_rank() -> int:
    return _get("pp")[1][-1]


def get_pipeline_model_parallel_next_rank() -> int:
    ranks = _get("pp")[1]
    return ranks[(get_pp_rank() + 1) % len(ranks)]


def get_pipeline_model_parallel_prev_rank() -> int:
    ranks = _get("pp")[1]
    return ranks[(get_pp_rank() - 1) % len(ranks)]


def get_dp_world_size(with_context_parallel=False) -> int:
    return _size("dp-cp" if with_context_parallel else "dp")


def get_dp_rank(with_context_parallel=False) -> int:
    return _rank("dp-cp" if with_context_parallel else "dp")


def get_cp_world_size() -> int:
    return _size("cp") if _STATE is not None else 1


def get_cp_rank() -> int:
    return _rank("cp") if _STATE is not None else 0
