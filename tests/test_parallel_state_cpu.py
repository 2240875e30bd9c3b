"""inferix_b200.parallel_state / dist_utils (reference inferix/distributed/parallel_state.py:236-634,
dist_utils.py:50-115): rank-group arithmetic against the reference's own RankGenerator (tests/golden/
parallel_groups.json, oracle/make_golden_groups.py), and a world_size-2 gloo run of dist_init -> cp group -> the
MAGI context-parallel binding."""
import json
import os
import socket
import types
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inferix_b200 import parallel_state as mpu

GOLDEN = Path(__file__).parent / "golden" / "parallel_groups.json"


def test_rank_groups_match_reference_generator():
    cases = json.loads(GOLDEN.read_text())["cases"]
    assert len(cases) >= 100
    for case in cases:
        for kind, want in case["groups"].items():
            got = mpu.rank_groups(case["sizes"], case["order"], kind)
            assert got == want, (case["sizes"], case["order"], kind)


def test_rank_groups_properties():
    sizes = {"tp": 1, "cp": 8, "pp": 1, "dp": 1}                    # BASELINE config 4: cp over the 8 GPUs of a box
    assert mpu.rank_groups(sizes, "tp-cp-pp-dp", "cp") == [list(range(8))]
    assert mpu.rank_groups(sizes, "tp-cp-pp-dp", "dp") == [[r] for r in range(8)]
    with pytest.raises(RuntimeError):                               # a sharded axis missing from the order
        mpu.rank_groups({"tp": 1, "cp": 2, "pp": 1, "dp": 1}, "tp-pp-dp", "cp")


def test_uninitialised_accessors():
    assert not mpu.model_parallel_is_initialized() and mpu.is_unitialized()
    assert mpu.get_cp_world_size() == 1 and mpu.get_cp_rank() == 0
    assert mpu.get_cp_group(check_initialized=False) is None
    with pytest.raises(AssertionError):
        mpu.get_cp_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from inferix_b200 import dist_utils, magi_cp
    cfg = types.SimpleNamespace(engine_config=types.SimpleNamespace(
        distributed_backend="gloo", distributed_timeout_minutes=5, cp_size=world, pp_size=1))
    try:
        dist_utils.dist_init(cfg)
        ok = mpu.model_parallel_is_initialized() and mpu.get_cp_world_size() == world and mpu.get_cp_rank() == rank
        ok &= mpu.get_tp_world_size() == 1 and mpu.get_pp_world_size() == 1 and mpu.get_dp_world_size() == 1
        ok &= mpu.get_tp_world_size(with_context_parallel=True) == world
        ok &= mpu.is_pipeline_first_stage() and mpu.is_pipeline_last_stage()
        ok &= mpu.get_tensor_model_parallel_src_rank(with_context_parallel=True) == 0
        ok &= dist_utils.is_last_tp_cp_rank() == (rank == world - 1) and dist_utils.is_last_rank() == (rank == world - 1)
        ok &= dist_utils.get_world_size() == world and dist_utils.get_device().type == "cpu"
        # the cp group is the one the MAGI context-parallel code uses
        ok &= magi_cp.get_cp_group() is mpu.get_cp_group() and magi_cp.get_cp_world_size() == world
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, group=mpu.get_cp_group())
        ok &= t.item() == sum(range(1, world + 1))
        dist_utils.dist_init(cfg)                                   # idempotent, like the reference
        try:
            mpu.initialize_model_parallel(cp_size=world)
            ok = False
        except AssertionError:
            pass
        mpu.destroy_model_parallel()
        ok &= not mpu.model_parallel_is_initialized() and magi_cp.get_cp_world_size() == 1
        ret[rank] = bool(ok)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_dist_init_two_ranks_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_pipeline_parallel_is_refused():
    """pp_size > 1 is valid in the reference and outside this build: NotImplementedError, not a silent fallback."""
    port = _free_port()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        with pytest.raises(NotImplementedError):
            mpu.initialize_model_parallel(pp_size=2)
        assert not mpu.model_parallel_is_initialized()
    finally:
        dist.destroy_process_group()
