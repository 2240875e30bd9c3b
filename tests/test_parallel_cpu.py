"""Host-side logic of the sequence-parallel path on CPU: token scatter / re-interleave helpers against the
reference's einops formulation, a world_size-2 gloo run of the all-gather + re-interleave, and the SP variant of the
oracle (sharded queries + gathered K/V into a replicated cache) against the single-process oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inferix_b200.parallel import ParallelConfig, all_gather_tokens, interleave_gathered, scatter_tokens


def test_scatter_and_interleave_roundtrip():
    b, f, hw, c, p = 2, 3, 8, 5, 4
    x = torch.arange(b * f * hw * c, dtype=torch.float32).view(b, f * hw, c)
    shards = [scatter_tokens(x, f, p, r) for r in range(p)]
    for r, s in enumerate(shards):          # causal_model.py:940-942: 'b (f hw) c' -> chunk hw -> 'b (f hw/p) c'
        want = x.view(b, f, hw, c).chunk(p, dim=2)[r].reshape(b, -1, c)
        assert torch.equal(s, want)
    back = interleave_gathered(torch.stack(shards), f, p)     # 'b (cp f hw) c -> b (f cp hw) c' (:1018)
    assert torch.equal(back, x)


def test_scatter_rejects_uneven_split():
    with pytest.raises(ValueError):
        scatter_tokens(torch.zeros(1, 3 * 10, 4), 3, 4, 0)


def test_parallel_config_validation():
    ParallelConfig(ring_size=2, world_size=2, rank=1)
    with pytest.raises(ValueError):
        ParallelConfig(attn_backend="FlashAttnV3")
    with pytest.raises(ValueError):
        ParallelConfig(ring_size=3, world_size=2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        f, hw, c = 3, 8, 4
        full = torch.arange(f * hw * c, dtype=torch.float32).view(1, f * hw, c)
        cfg = ParallelConfig(ring_size=world, world_size=world, rank=rank)
        mine = scatter_tokens(full, f, world, rank)
        back = all_gather_tokens(mine * 2, f, cfg)
        ret[rank] = bool(torch.equal(back, full * 2))
    finally:
        dist.destroy_process_group()


def test_all_gather_tokens_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_sp_oracle_equals_single_process_oracle():
    """The replicated-cache SP scheme is exact: each rank's output == its slice of the 1-rank output."""
    from inferix_b200.synthetic import TINY, synth_state_dict
    from oracle import wan_oracle as wo
    torch.set_num_threads(2)
    cfg = wo.WanConfig(**TINY, local_attn_size=6, sink_size=0)
    sd = synth_state_dict(TINY, seed=0)
    g = torch.Generator().manual_seed(5)
    frames, grid, C, P = 3, (3, 8, 8), TINY["dim"], 2
    fs = 64
    x = torch.randn(1, frames * fs, C, generator=g)
    e0 = torch.randn(1, frames, 6, C, generator=g) * 0.3
    ctx = torch.randn(1, 512, C, generator=g) * 0.5
    freqs = wo.rope_freqs(128)
    single = wo.new_cache(cfg, 6 * fs, 1, torch.float32)[0]
    ref = wo.block_forward(sd, 0, cfg, x, e0, grid, freqs, ctx, single, dict(is_init=False), 0)

    # emulate P ranks in one process: first pass collects every rank's new K/V, second pass attends
    shards = [scatter_tokens(x, frames, P, r) for r in range(P)]
    new_kv = {}

    def collect(rank):
        def hook(k, v):
            new_kv[rank] = (k, v)
            raise StopIteration
        return hook
    for r in range(P):
        try:
            wo.block_forward(sd, 0, cfg, shards[r], e0, grid, freqs, ctx, wo.new_cache(cfg, 6 * fs, 1, torch.float32)[0],
                             dict(is_init=False), 0, world_size=P, rank=r, peer_kv=collect(r))
        except StopIteration:
            pass

    def gathered(_k, _v):
        ks = torch.stack([new_kv[r][0] for r in range(P)])       # [P, 1, F*chunk, H, D]
        vs = torch.stack([new_kv[r][1] for r in range(P)])
        h, d = ks.shape[-2:]
        k = interleave_gathered(ks.flatten(3), frames, P).unflatten(2, (h, d))
        v = interleave_gathered(vs.flatten(3), frames, P).unflatten(2, (h, d))
        return k, v
    for r in range(P):
        cache = wo.new_cache(cfg, 6 * fs, 1, torch.float32)[0]
        out = wo.block_forward(sd, 0, cfg, shards[r], e0, grid, freqs, ctx, cache, dict(is_init=False), 0,
                               world_size=P, rank=r, peer_kv=gathered)
        want = scatter_tokens(ref, frames, P, r)
        assert torch.allclose(out, want, rtol=1e-4, atol=1e-4)
        assert torch.allclose(cache.k[:, :cache.local_end], single.k[:, :single.local_end], rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------------- peer-memory exchange: collective-safe setup
def _peer_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        import warnings
        from inferix_b200 import peer
        # CPU tensors cannot be exported through CUDA IPC: every rank must still take part in the handle exchange and
        # the vote, agree on "no peer path", leave the stores untouched and return None (-> NCCL all-gather path)
        stores = [types.SimpleNamespace(k=torch.zeros(8, 256, dtype=torch.bfloat16),
                                        v=torch.zeros(8, 256, dtype=torch.bfloat16)) for _ in range(3)]
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            pg = peer.try_setup(stores, world, rank, None, torch.device("cpu"))
        ok = pg is None and all(s.peer is None and s.peer_group is None for s in stores)
        ok = ok and (rank != 0 or any("peer-memory KV exchange unavailable" in str(x.message) for x in w))
        # and the helper used by the pipelines is a no-op for a single rank / when disabled
        pc1 = types.SimpleNamespace(world_size=1, rank=0, group=None)
        ok = ok and peer.setup_for_pipeline(None, None, [], pc1) is None
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_peer_setup_falls_back_collectively_without_ipc():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_peer_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_peer_group_rejects_more_than_eight_ranks():
    from inferix_b200 import peer
    with pytest.raises(ValueError):
        peer.PeerGroup(9, 0, None, "cpu")
