"""Reference-surface shims on the native kernels (SURVEY §8b signature list): CausalWanSelfAttention.forward,
the (out, lse) attention backend + collect_supported_attn, CoreAttention (Ulysses strategy / local)."""
import pytest
import torch

from inferix_b200 import ops
from inferix_b200.attention import CoreAttention, collect_supported_attn, ifx_attn_forward
from inferix_b200.synthetic import TINY, synth_state_dict
from inferix_b200.wan_model import CausalWanModel
from oracle import wan_oracle as wo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_self_attention_forward_reference_signature():
    """CausalWanSelfAttention.forward(x, seq_lens, grid_sizes, freqs, block_mask, kv_cache_meta, current_start,
    cache_start) -> (y, k_view, v_view) on reference-layout cache tensors, vs oracle.self_attention: append, repeat,
    advance, evict (window 6 frames, sink 1)."""
    cfg = wo.WanConfig(**TINY, local_attn_size=6, sink_size=1)
    sd = {k: v.bfloat16() for k, v in synth_state_dict(TINY, seed=0).items()}
    model = CausalWanModel(**TINY, local_attn_size=6, sink_size=1)
    model.load_state_dict(synth_state_dict(TINY, seed=0))
    model = model.to(torch.bfloat16).to(DEV)
    sa = model.blocks[0].self_attn
    frames, fs, C, n, d = 3, 64, TINY["dim"], TINY["num_heads"], TINY["dim"] // TINY["num_heads"]
    g = torch.Generator().manual_seed(4)
    cache = wo.new_cache(cfg, 6 * fs, 1, torch.bfloat16)[0]
    meta = {"k": torch.zeros(1, 6 * fs, n, d, dtype=torch.bfloat16, device=DEV),
            "v": torch.zeros(1, 6 * fs, n, d, dtype=torch.bfloat16, device=DEV),
            "global_end_index": torch.zeros(1, dtype=torch.long, device=DEV),
            "local_end_index": torch.zeros(1, dtype=torch.long, device=DEV)}
    table = ops.rope_table(model.freqs, DEV)
    grid = (frames, 8, 8)
    for step, start in enumerate([0, 0, 3 * fs, 6 * fs, 9 * fs]):
        x = torch.randn(1, frames * fs, C, generator=g).bfloat16()
        ref = wo.self_attention(sd, "blocks.0.self_attn", cfg, x, grid, wo.rope_freqs(d), cache, start)
        y, kv, vv = sa(x.to(DEV), None, torch.tensor([grid]), table, None, meta, current_start=start)
        assert (int(meta["global_end_index"]), int(meta["local_end_index"])) == (cache.global_end, cache.local_end)
        assert kv.shape[1] == cache.local_end and vv.shape == kv.shape
        assert rel_l2(y, ref) <= 5e-3, f"step {step}"
        assert rel_l2(kv, cache.k[:, :cache.local_end]) <= 1e-3 and rel_l2(vv, cache.v[:, :cache.local_end]) <= 1e-3


@pytest.mark.parametrize("lq,lk,heads,kv_heads", [(300, 1000, 4, 4), (2000, 9000, 12, 12), (700, 3000, 8, 2)])
def test_attention_backend_returns_lse(lq, lk, heads, kv_heads):
    """(out, lse) as the reference's backend functions return them (backends.py:58-72): lse = log sum exp of the
    scaled scores, fp32 — with and without the key-split + combine path, and for grouped-query heads."""
    d = 128
    g = torch.Generator(device=DEV).manual_seed(0)
    q = torch.randn(1, lq, heads, d, device=DEV, generator=g).bfloat16()
    k = torch.randn(1, lk, kv_heads, d, device=DEV, generator=g).bfloat16()
    v = torch.randn(1, lk, kv_heads, d, device=DEV, generator=g).bfloat16()
    assert "InferixB200" in collect_supported_attn()
    out, lse = ifx_attn_forward(q, k, v)
    rep = heads // kv_heads
    kf = k[0].float().repeat_interleave(rep, dim=1).transpose(0, 1)           # [heads, lk, d]
    vf = v[0].float().repeat_interleave(rep, dim=1).transpose(0, 1)
    s = (q[0].float().transpose(0, 1) @ kf.transpose(1, 2)) / d ** 0.5
    ref_lse = torch.logsumexp(s, dim=-1)
    ref = (torch.softmax(s, dim=-1) @ vf).transpose(0, 1)
    assert lse.shape == (1, heads, lq) and lse.dtype == torch.float32
    assert (lse[0] - ref_lse).abs().max().item() <= 2e-3
    assert rel_l2(out[0], ref) <= 4e-3


def test_core_attention_local_with_cache():
    """CoreAttention.forward with k_cache / v_cache / offsets (distributed.py:197-203) on a size-1 group: the new keys
    are written into the caches at the offset and the queries attend the cache prefix."""
    heads, d, lq, off = 2, 128, 192, 384
    g = torch.Generator(device=DEV).manual_seed(1)
    q = torch.randn(1, lq, heads, d, device=DEV, generator=g).bfloat16()
    k = torch.randn(1, lq, heads, d, device=DEV, generator=g).bfloat16()
    v = torch.randn(1, lq, heads, d, device=DEV, generator=g).bfloat16()
    kc = torch.randn(1, 1024, heads, d, device=DEV, generator=g).bfloat16()
    vc = torch.randn(1, 1024, heads, d, device=DEV, generator=g).bfloat16()
    out = CoreAttention()(q, k, v, k_cache=kc, v_cache=vc, k_cache_offset=off, v_cache_offset=off)
    assert torch.equal(kc[0, off:off + lq], k[0]) and torch.equal(vc[0, off:off + lq], v[0])
    kk, vvv = kc[0, :off + lq].float().transpose(0, 1), vc[0, :off + lq].float().transpose(0, 1)
    ref = (torch.softmax(q[0].float().transpose(0, 1) @ kk.transpose(1, 2) / d ** 0.5, dim=-1) @ vvv).transpose(0, 1)
    assert rel_l2(out[0], ref) <= 4e-3
    with pytest.raises(NotImplementedError):
        CoreAttention()(q, k, v, custom_mask=torch.ones(1))


def test_block_forward_error_leaves_cache_untouched():
    """A block forward whose launches are refused after the append was planned (here: a misaligned scratch buffer,
    rejected by the QKV GEMM) must leave the native block table and end indices where kv_cache_meta says they are —
    the plan is host-side state that was already advanced when the error is found."""
    from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest
    from inferix_b200.wan_model import _Workspace
    model = CausalWanModel(**TINY, local_attn_size=6, sink_size=0)
    model.load_state_dict(synth_state_dict(TINY, seed=0))
    model = model.to(torch.bfloat16).to(DEV)
    blk = model.blocks[0]
    frames, fs, C = 3, 64, TINY["dim"]
    mgr, req = KVCacheManager(DEV), KVCacheRequest("rollback")
    blk.kv_cache_manager.allocate_kv_cache(mgr, req, 6 * fs, torch.bfloat16, page_tokens=fs)
    blk.kv_cache_manager.allocate_crossattn_cache(mgr, req, 512, torch.bfloat16)
    store = blk.kv_cache_manager.store(mgr, req)
    g = torch.Generator().manual_seed(9)
    table = ops.rope_table(model.freqs, DEV)
    meta = {"global_end_index": torch.zeros(1, dtype=torch.long, device=DEV),
            "local_end_index": torch.zeros(1, dtype=torch.long, device=DEV)}
    cmeta = {"is_init": False}
    ctx = (torch.randn(1, 512, C, generator=g) * 0.5).bfloat16().to(DEV)
    e0 = (torch.randn(1, frames, 6, C, generator=g) * 0.3).bfloat16().to(DEV)

    def run(start, ws=None):
        x = torch.randn(1, frames * fs, C, generator=g).bfloat16().to(DEV)
        return blk(x, e0, None, torch.tensor([(frames, 8, 8)]), table, ctx, None, None, meta, cmeta,
                   current_start=start, kv_cache_manager=mgr, kv_cache_requests=[req], workspace=ws)

    run(0)
    run(3 * fs)                                           # window full: the next append evicts (table rotation)
    before = store.state()
    bad = _Workspace(frames * fs, C, TINY["ffn_dim"], DEV)
    bad.qkv = torch.empty(frames * fs * 3 * C + 8, dtype=torch.bfloat16, device=DEV)[1:1 + frames * fs * 3 * C].view(
        frames * fs, 3 * C)                               # 2 bytes off a 16-byte boundary
    with pytest.raises((ValueError, RuntimeError, NotImplementedError)):
        run(6 * fs, bad)
    assert store.state() == before
    assert (int(meta["global_end_index"]), int(meta["local_end_index"])) == (6 * fs, 6 * fs)
    out = run(6 * fs)                                     # the same block, now with sound buffers, proceeds normally
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    after = store.state()
    assert after != before and int(meta["global_end_index"]) == 9 * fs


def test_reference_shaped_allocation_without_page_size():
    """allocate_kv_cache exactly as the reference calls it (no page_tokens: block_size=1,
    self_forcing_kv_cache_manager.py:33-60) must be usable: the first block forward re-cuts the empty cache into
    frame-sized pages.  Same outputs and indices as the explicitly paged allocation, incl. eviction; a cache that
    already holds tokens cannot be re-cut."""
    from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest
    model = CausalWanModel(**TINY, local_attn_size=6, sink_size=0)
    model.load_state_dict(synth_state_dict(TINY, seed=0))
    model = model.to(torch.bfloat16).to(DEV)
    blk = model.blocks[0]
    frames, fs, C = 3, 64, TINY["dim"]
    table = ops.rope_table(model.freqs, DEV)
    outs = []
    for paged in (True, False):
        mgr, req = KVCacheManager(DEV), KVCacheRequest(f"alloc{int(paged)}")
        kw = {"page_tokens": fs} if paged else {}
        blk.kv_cache_manager.allocate_kv_cache(mgr, req, 6 * fs, torch.bfloat16, **kw)
        blk.kv_cache_manager.allocate_crossattn_cache(mgr, req, 512, torch.bfloat16)
        g = torch.Generator().manual_seed(11)
        meta = {"global_end_index": torch.zeros(1, dtype=torch.long, device=DEV),
                "local_end_index": torch.zeros(1, dtype=torch.long, device=DEV)}
        cmeta = {"is_init": False}
        ctx = (torch.randn(1, 512, C, generator=g) * 0.5).bfloat16().to(DEV)
        e0 = (torch.randn(1, frames, 6, C, generator=g) * 0.3).bfloat16().to(DEV)
        res = []
        for start in (0, 3 * fs, 6 * fs):
            x = torch.randn(1, frames * fs, C, generator=g).bfloat16().to(DEV)
            res.append(blk(x, e0, None, torch.tensor([(frames, 8, 8)]), table, ctx, None, None, meta, cmeta,
                           current_start=start, kv_cache_manager=mgr, kv_cache_requests=[req]).clone())
        store = blk.kv_cache_manager.store(mgr, req)
        assert store.page_tokens == fs
        res.append(blk.kv_cache_manager.get_kv_cache(mgr, req).clone())      # reference-shaped read-out still works
        outs.append((res, store.state(), int(meta["global_end_index"]), int(meta["local_end_index"])))
        with pytest.raises(ValueError):
            store.repage(2 * fs)
        blk.kv_cache_manager.clear_cache(mgr, req)
    (ra, sa, ga, la), (rb, sb, gb, lb) = outs
    assert (sa, ga, la) == (sb, gb, lb)
    for a, b in zip(ra, rb):
        assert torch.equal(a, b)
