"""CoreAttention's ring strategies (reference inferix/models/attention/distributed.py:372-712) on CPU: 2- and 3-rank
gloo groups, an fp32 (out, lse) backend injected in place of the CUDA kernel.  pass-kv and pass-q must both equal the
attention of the local queries over the keys of ALL ranks (softmax attention is a sum over keys, merged exactly
through the log-sum-exp), and the returned lse must be the global one."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def cpu_attn(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), **_):
    """(out [B, Lq, N, D], lse [B, N, Lq]) like the reference's flash_attn_forward backend (backends.py:58-72)."""
    assert not causal and dropout_p == 0
    scale = softmax_scale if softmax_scale is not None else q.shape[-1] ** -0.5
    s = torch.einsum("blnd,bmnd->bnlm", q.float(), k.float()) * scale
    lse = torch.logsumexp(s, dim=-1)
    out = torch.einsum("bnlm,bmnd->blnd", torch.softmax(s, dim=-1), v.float())
    return out.to(q.dtype), lse


def _inputs(world, lq=5, lk=7, n=3, d=8):
    g = torch.Generator().manual_seed(17)
    return (torch.randn(world, 1, lq, n, d, generator=g), torch.randn(world, 1, lk, n, d, generator=g),
            torch.randn(world, 1, lk, n, d, generator=g))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from inferix_b200.attention import CoreAttention
        q, k, v = _inputs(world)
        want, want_lse = cpu_attn(q[rank], torch.cat(list(k), dim=1), torch.cat(list(v), dim=1))
        ca = CoreAttention(strategy="pass-kv", ring_pg=dist.group.WORLD)
        ca.supported_attn["cpu"] = cpu_attn
        scale = q.shape[-1] ** -0.5
        ok = True
        out, lse = ca.ring_attention_forward_pass_kv(dist.group.WORLD, q[rank], k[rank], v[rank], scale, attn_backend="cpu")
        ok &= torch.allclose(out, want, atol=1e-5) and torch.allclose(lse, want_lse, atol=1e-5)
        out, lse = ca.ring_attention_forward_pass_q(dist.group.WORLD, q[rank], k[rank], v[rank], scale, attn_backend="cpu")
        ok &= torch.allclose(out, want, atol=1e-5)
        ok &= torch.allclose(lse.squeeze(-1).transpose(1, 2), want_lse, atol=1e-5)       # merge layout [B, Lq, N, 1]
        for strategy in ("pass-kv", "pass-q", "auto"):                                    # through forward()
            ca.strategy = strategy
            got, got_lse = ca(q[rank], k[rank], v[rank], attn_backend="cpu", return_attn_probs=True)
            ok &= torch.allclose(got, want, atol=1e-5) and got_lse is not None
            ok &= torch.allclose(ca(q[rank], k[rank], v[rank], attn_backend="cpu"), want, atol=1e-5)
        try:
            ca.ring_attention_forward_pass_kv(dist.group.WORLD, q[rank], k[rank], v[rank], scale, causal=True,
                                              attn_backend="cpu")
            ok = False
        except NotImplementedError:
            pass
        try:
            ca.ring_attention_forward_pass_kv(dist.group.WORLD, q[rank], k[rank], v[rank], scale, attn_backend="FlashAttnV3")
            ok = False
        except ValueError:
            pass
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_strategies_equal_attention_over_all_keys(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {r: True for r in range(world)}


def test_strategy_selection_follows_the_reference_heuristic():
    from inferix_b200.attention import CoreAttention
    ca = CoreAttention()
    q = torch.zeros(1, 4096, 4, 8)
    assert ca._select_strategy(q, q, q) == "pass_q"
    assert ca._select_strategy(q[:, :512], q, q) == "pass_kv"
    assert ca._select_strategy(q, q, q, k_cache=q, v_cache=q) == "ulysses"
    assert CoreAttention(strategy="pass-q")._select_strategy(q, q, q) == "pass-q"
