"""Parity at the BENCHMARKED shape (BASELINE.json configs[1]: Wan-1.3B widths, 720p block of 10 800 tokens, 86 400-token
window) — round 1 only proved parity at toy widths.  Covers what the small cases cannot: split-KV tail + combine,
multi-wave scheduling, 2-CTA 256x256 GEMM tiles, eviction at full size.

The oracle runs on a band of query rows.  It is given that band exactly as a sequence-parallel rank would own it
(`world_size=15`: 240 hw indices of every frame, the reference's own chunking, causal_model.py:64-100,939-942), so
oracle.block_forward is called unchanged: RoPE positions, per-frame modulation, cache append / eviction and the
attention over the whole window are the oracle's; only the K / V of the rows outside the band come from the native run
(handed over through the oracle's `peer_kv` hook, which is how its sequence-parallel form receives other ranks' rows).
"""
import pytest
import torch

from inferix_b200 import ops, synthetic
from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest
from inferix_b200.wan_model import CausalWanModel
from oracle import wan_oracle as wo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)

H_LAT, W_LAT, FRAMES, WINDOW_BLOCKS = 45, 80, 3, 8
FS = H_LAT * W_LAT                      # 3600 tokens per frame
S, L = FRAMES * FS, WINDOW_BLOCKS * FRAMES * FS
BAND_WORLD, BAND_RANK = 15, 7           # the band: hw indices [1680, 1920) of every frame = 720 query rows


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_attention_full_window_vs_fp32_rows():
    """ifx_attention, 12 heads x 10 800 queries x 86 400 keys (the roofline kernel's exact launch: 507 work items =
    3 whole waves + a split tail + combine) against fp32 softmax attention on sampled rows, plus rows-sum-to-one."""
    heads, d = 12, 128
    g = torch.Generator(device=DEV).manual_seed(0)
    q = torch.randn(S, heads * d, device=DEV, generator=g).bfloat16()
    k = torch.randn(L, heads * d, device=DEV, generator=g).bfloat16()
    v = torch.randn(L, heads * d, device=DEV, generator=g).bfloat16()
    out = ops.attention(q, k, v, heads)
    # rows from every region of the grid: first / last pair, the whole-wave / split-tail boundary, the ragged last tile
    rows = torch.cat([torch.arange(0, 64), torch.arange(3000, 3064), torch.arange(7400, 7464),
                      torch.arange(S - 120, S)]).to(DEV)
    qs = q[rows].float().view(-1, heads, d).transpose(0, 1)                       # [H, R, D]
    kh = k.float().view(L, heads, d).transpose(0, 1)
    vh = v.float().view(L, heads, d).transpose(0, 1)
    p = torch.softmax(qs @ kh.transpose(1, 2) / d ** 0.5, dim=-1)
    ref = (p @ vh).transpose(0, 1).reshape(len(rows), heads * d)
    err = rel_l2(out[rows], ref)
    print(f"full-window attention vs fp32 on {len(rows)} sampled rows: rel-L2 {err:.3e}")
    assert err <= 4e-3
    # size-independent property on ALL rows: with V = 1 the output is 1 (rows of softmax sum to one)
    ones = torch.ones_like(v)
    o1 = ops.attention(q, k, ones, heads)
    assert (o1.float() - 1.0).abs().max().item() <= 2e-2


@pytest.mark.parametrize("q_rows", [1350, 2700])
def test_attention_shard_shapes(q_rows):
    """The per-rank shapes of the 8- / 4-way sequence-parallel run (1350 / 2700 query rows x 12 heads against the
    86 400-key window: every item key-split in two + combine at 8 ranks, a ragged last row pair) against fp32 softmax
    attention on sampled rows (all rows for the sum-to-one property), dense and in extent mode."""
    heads, d = 12, 128
    g = torch.Generator(device=DEV).manual_seed(q_rows)
    q = torch.randn(q_rows, heads * d, device=DEV, generator=g).bfloat16()
    k = torch.randn(L, heads * d, device=DEV, generator=g).bfloat16()
    v = torch.randn(L, heads * d, device=DEV, generator=g).bfloat16()
    out = ops.attention(q, k, v, heads)
    rows = torch.cat([torch.arange(0, 96), torch.arange(q_rows // 2, q_rows // 2 + 96),
                      torch.arange(q_rows - 96, q_rows)]).to(DEV)
    qs = q[rows].float().view(-1, heads, d).transpose(0, 1)
    kh = k.float().view(L, heads, d).transpose(0, 1)
    vh = v.float().view(L, heads, d).transpose(0, 1)
    ref = (torch.softmax(qs @ kh.transpose(1, 2) / d ** 0.5, dim=-1) @ vh).transpose(0, 1).reshape(len(rows), heads * d)
    err = rel_l2(out[rows], ref)
    print(f"shard shape, {q_rows} rows: rel-L2 vs fp32 {err:.3e}")
    assert err <= 4e-3
    assert torch.isfinite(out.float()).all()
    o1 = ops.attention(q, k, torch.ones_like(v), heads)          # every row of softmax sums to one, on ALL rows
    assert (o1.float() - 1.0).abs().max().item() <= 2e-2
    # same keys as two extents with a ragged boundary (extent mode + V tail fix-up), unmapped rows poisoned
    cut = 40000 + 72
    k2 = torch.full((L + 3000, heads * d), float("nan"), dtype=torch.bfloat16, device=DEV)
    v2 = torch.full_like(k2, float("nan"))
    k2[:cut], v2[:cut] = k[:cut], v[:cut]
    k2[cut + 3000:], v2[cut + 3000:] = k[cut:], v[cut:]
    out2 = ops.attention_extents(q, k2, v2, [(0, cut), (cut + 3000, L - cut)], heads)
    assert torch.isfinite(out2.float()).all()
    assert rel_l2(out2[rows], ref) <= 4e-3


def _setup_layer(seed=0):
    cfg_d = dict(synthetic.WAN_1_3B, num_layers=1)
    sd32 = synthetic.synth_state_dict(cfg_d, seed=seed)
    model = CausalWanModel(**cfg_d, local_attn_size=WINDOW_BLOCKS * FRAMES, sink_size=0)
    model.load_state_dict(sd32)
    model = model.to(torch.bfloat16).to(DEV)
    sd = {k: v.bfloat16() for k, v in sd32.items()}
    cfg = wo.WanConfig(**cfg_d, local_attn_size=WINDOW_BLOCKS * FRAMES, sink_size=0)
    return model, sd, cfg


def test_block_forward_full_shape_vs_oracle_band():
    """One DiT layer through ifx_wan_block_forward at C=1536, 12 heads, S=10 800, window 86 400: (1) the block that
    fills the window, (2) the next block, which evicts three frames.  The oracle recomputes a 720-row band."""
    model, sd, cfg = _setup_layer()
    blk = model.blocks[0]
    C, heads, hd = cfg.dim, cfg.num_heads, cfg.head_dim
    mgr, req = KVCacheManager(DEV), KVCacheRequest("full")
    blk.kv_cache_manager.allocate_kv_cache(mgr, req, L, torch.bfloat16, page_tokens=FS)
    blk.kv_cache_manager.allocate_crossattn_cache(mgr, req, 512, torch.bfloat16)
    store = blk.kv_cache_manager.store(mgr, req)
    g = torch.Generator().manual_seed(5)
    # seven blocks of cached history (random K / V through the real append path), identical in the oracle's cache
    hist_k = torch.randn(L - S, C, generator=g).bfloat16()
    hist_v = torch.randn(L - S, C, generator=g).bfloat16()
    for b in range(WINDOW_BLOCKS - 1):
        plan = store.plan_append(b * S, S, 0, True)
        store.append(plan, hist_k[b * S:(b + 1) * S].to(DEV), hist_v[b * S:(b + 1) * S].to(DEV))
    cache = wo.LayerCache(torch.zeros(1, L, heads, hd, dtype=torch.bfloat16), torch.zeros(1, L, heads, hd, dtype=torch.bfloat16),
                          global_end=L - S, local_end=L - S)
    cache.k[0, :L - S] = hist_k.view(L - S, heads, hd)
    cache.v[0, :L - S] = hist_v.view(L - S, heads, hd)
    ctx = (torch.randn(1, 512, C, generator=g) * 0.5).bfloat16()
    cross = dict(is_init=False)
    meta = {"global_end_index": torch.full((1,), L - S, dtype=torch.long, device=DEV),
            "local_end_index": torch.full((1,), L - S, dtype=torch.long, device=DEV)}
    cmeta = {"is_init": False}
    table = ops.rope_table(model.freqs, DEV)
    freqs = wo.rope_freqs(hd)
    chunk = FS // BAND_WORLD
    band = torch.cat([torch.arange(f * FS + BAND_RANK * chunk, f * FS + (BAND_RANK + 1) * chunk) for f in range(FRAMES)])

    for step, start in enumerate([L - S, L]):                      # fill the window, then evict
        x = torch.randn(1, S, C, generator=g).bfloat16()
        e0 = (torch.randn(1, FRAMES, 6, C, generator=g) * 0.3).bfloat16()
        out = blk(x.clone().to(DEV), e0.to(DEV), None, torch.tensor([(FRAMES, H_LAT, W_LAT)]), table, ctx.to(DEV), None,
                  None, meta, cmeta, current_start=start, kv_cache_manager=mgr, kv_cache_requests=[req])
        torch.cuda.synchronize()
        assert int(meta["local_end_index"]) == L and int(meta["global_end_index"]) == start + S
        # the block's new K / V as the native run produced them (logical order = the reference's rolled tensor)
        k_new, v_new = store.export(L - S, S)
        k_new, v_new = k_new.cpu().view(1, S, heads, hd), v_new.cpu().view(1, S, heads, hd)
        seen = {}

        def peer_kv(k_band, v_band):
            seen["k"], seen["v"] = k_band, v_band
            return k_new, v_new
        ref = wo.block_forward(sd, 0, cfg, x[:, band], e0, (FRAMES, H_LAT, W_LAT), freqs, ctx, cache, cross, start,
                               world_size=BAND_WORLD, rank=BAND_RANK, peer_kv=peer_kv)
        # the oracle's own K / V for the band vs the rows the native kernel wrote into the cache
        ek, ev = rel_l2(k_new[0, band], seen["k"][0]), rel_l2(v_new[0, band], seen["v"][0])
        err = rel_l2(out[0, band.to(DEV)], ref[0])
        print(f"full-shape layer step {step} (start {start}): band rel-L2 {err:.3e}; new K {ek:.2e}, V {ev:.2e}")
        assert ek <= 1e-3 and ev <= 1e-3
        assert err <= 3e-3
        assert cache.trace[-1][:3] == (L - S, L, start + S)
        assert cache.trace[-1][3] == (S if step == 1 else 0)
        # whole cache, logical order, against the oracle's rolled tensor (rows outside the band are the native rows the
        # oracle was handed, so this checks the rotation: which page sits where after the eviction)
        kc, vc = store.export(0, L)
        assert torch.equal(kc.cpu().view(L, heads, hd), cache.k[0])
        assert torch.equal(vc.cpu().view(L, heads, hd), cache.v[0])


@pytest.mark.parametrize("site,n,k,epi", [("qkv", 4608, 1536, "bias"), ("ffn1", 8960, 1536, "gelu"),
                                          ("ffn2", 1536, 8960, "gate_res")])
@pytest.mark.parametrize("m", [10800, 1350])
def test_gemm_full_shape_vs_oracle_rows(site, n, k, epi, m):
    """The block's GEMM shapes at M = 10 800 (one GPU) and M = 1350 (the 8-way sequence-parallel shard: the 128-wide
    tile path) with their fused epilogues, against the oracle's F.linear (+GELU / gate+residual) on sampled rows."""
    g = torch.Generator(device=DEV).manual_seed(11)
    a = (torch.randn(m, k, device=DEV, generator=g) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=DEV, generator=g) * k ** -0.5).bfloat16()
    b = (torch.randn(n, device=DEV, generator=g) * 0.1).bfloat16()
    fs = m // 3
    rows = torch.cat([torch.arange(0, 96), torch.arange(m // 2, m // 2 + 96), torch.arange(m - 96, m)])
    lin = torch.nn.functional.linear(a[rows.to(DEV)].cpu(), w.cpu(), b.cpu())     # bf16 F.linear, as the oracle's _lin
    if epi == "bias":
        out, ref = ops.gemm(a, w, b), lin
    elif epi == "gelu":
        out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GELU)
        ref = torch.nn.functional.gelu(lin, approximate="tanh")
    else:
        res = torch.randn(m, n, device=DEV, generator=g).bfloat16()
        gate = (torch.randn(3, n, device=DEV, generator=g) * 0.3).bfloat16()
        out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GATE_RES, residual=res, gate=gate, tokens_per_frame=fs)
        gr = gate.cpu()[(rows // fs)]
        ref = res.cpu()[rows] + lin * gr
    err = rel_l2(out[rows.to(DEV)], ref)
    print(f"gemm {site} M={m}: rel-L2 {err:.2e}")
    assert err <= 1e-3
