"""inferix_b200.ulysses_scheduler.UlyssesScheduler against the reference's own class (context_parallel.py:453-598):
both run the same closures under a 2-rank gloo group with an uneven sequence split; the reference's outputs are the
golden tests/golden/ulysses_sched.pt (oracle/make_golden_ulysses.py).  Bit-exact: the all-to-all layouts, the KV-head
repetition, the query-head chunking of `overlap_degree` and the final head order are pure data movement around the
same fp32 attention."""
import os
import socket
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import make_golden_ulysses as ref

GOLDEN = Path(__file__).parent / "golden" / "ulysses_sched.pt"


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from inferix_b200 import magi_cp
        from inferix_b200.ulysses_scheduler import UlyssesScheduler
        magi_cp.init_context_parallel(dist.group.WORLD, world, rank)
        ret[rank] = ref.run_variants(UlyssesScheduler, rank, world)
        magi_cp.destroy_context_parallel()
    finally:
        dist.destroy_process_group()


def test_scheduler_matches_reference_class_two_ranks():
    gold = torch.load(GOLDEN, weights_only=False)
    assert gold["split"] == ref.SPLIT
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        got = {r: ret[r] for r in range(2)}
    for rank in range(2):
        for case, outs in gold["ranks"][rank].items():
            for variant, want in outs.items():
                have = got[rank][case][variant]
                assert have.shape == want.shape, (rank, case, variant)
                assert torch.equal(have, want), (rank, case, variant)


def test_single_rank_is_plain_attention():
    """cp = 1: no communication; the result is the core attention over (history + new) keys in the caller's layout."""
    from inferix_b200.ulysses_scheduler import UlyssesScheduler
    g = torch.Generator().manual_seed(3)
    q, k, v = torch.randn(6, 4, 16, generator=g), torch.randn(6, 2, 16, generator=g), torch.randn(6, 2, 16, generator=g)
    out, xa = UlyssesScheduler.get_attn_and_xattn_with_comm_overlap(
        lambda: q, lambda: k, lambda: v, lambda kv: (kv[..., :16].contiguous(), kv[..., 16:].contiguous()),
        ref.core_attn, lambda: "x", 1, 1, 1, None)
    assert xa == "x" and torch.equal(out, ref.core_attn(q, k, v).reshape(6, 1, 64))
