"""`dist_init` and the small rank helpers of the reference's `inferix/distributed/dist_utils.py:26-115`, on
`inferix_b200.parallel_state`.  `dist_init(config)` reads the same fields of the MAGI config
(`engine_config.distributed_backend`, `.distributed_timeout_minutes`, `.cp_size`, `.pp_size`)."""
from __future__ import annotations

import os
from datetime import timedelta

import torch
import torch.distributed as dist

from . import parallel_state as mpu


def print_rank_0(message) -> None:
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(message, flush=True)


def print_per_rank(message) -> None:
    if dist.is_initialized():
        print(f"[rank {dist.get_rank()}] {message}", flush=True)
    else:
        print(message, flush=True)


def dist_init(config) -> None:
    """reference :50-85: default process group from RANK / WORLD_SIZE (one process per GPU, device = rank modulo the
    visible devices), then the model-parallel groups; cp_size * pp_size must cover the world."""
    ec = config.engine_config
    n_dev = torch.cuda.device_count()
    if dist.is_initialized():
        print_rank_0("Torch distribution already initialized, skipping initialization ...")
    else:
        rank, world = int(os.getenv("RANK", "0")), int(os.getenv("WORLD_SIZE", "1"))
        if n_dev > 0:
            torch.cuda.set_device(rank % n_dev)
        dist.init_process_group(backend=ec.distributed_backend, world_size=world, rank=rank,
                                timeout=timedelta(minutes=ec.distributed_timeout_minutes))
    assert ec.cp_size * ec.pp_size == dist.get_world_size()
    if mpu.model_parallel_is_initialized():
        print_rank_0("Model parallel is already initialized")
    else:
        mpu.initialize_model_parallel(cp_size=ec.cp_size, pp_size=ec.pp_size, nccl_communicator_config_path=None,
                                      distributed_timeout_minutes=ec.distributed_timeout_minutes, order="tp-cp-pp-dp")
    print_rank_0("Initialize torch distribution and model parallel successfully")


def is_last_rank() -> bool:
    return dist.get_rank() == dist.get_world_size() - 1


def is_last_tp_cp_rank() -> bool:
    return mpu.get_tp_rank(with_context_parallel=True) == mpu.get_tp_world_size(with_context_parallel=True) - 1


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def get_device(local_rank=None) -> torch.device:
    """reference :104-115: the device class follows the default group's backend."""
    backend = dist.get_backend()
    if backend == "nccl":
        return torch.device("cuda") if local_rank is None else torch.device(f"cuda:{local_rank}")
    if backend == "gloo":
        return torch.device("cpu")
    raise RuntimeError(f"unsupported distributed backend {backend!r}")
