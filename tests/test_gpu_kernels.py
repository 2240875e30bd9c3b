"""Parity of each sm_100a kernel, called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances: cache indexing and pure data movement are bit-exact; elementwise kernels reproduce every
reference rounding point, so they must agree with the oracle's bf16 result up to isolated 1-ulp flips
(bf16 ulp = 2^-8 relative): rel-L2 <= 1e-3 and >= 99.9 % of elements identical; GEMM / attention accumulate in
fp32 in a different order than the CPU kernels: rel-L2 <= 1e-3 for GEMM outputs (bf16-rounded the same way) and
<= 4e-3 for attention against an fp32 softmax (bf16 P and bf16 output rounding, as FlashAttention-2)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from inferix_b200 import ops                      # noqa: E402
from inferix_b200._lib import RopeGrid            # noqa: E402
from oracle import wan_oracle as wo               # noqa: E402

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def same_frac(a, b):
    return (a.cpu() == b.cpu()).float().mean().item()


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).bfloat16()


# ----------------------------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize("rows,cols,fs", [(192, 256, 64), (12, 1536, 4), (300, 1536, 100), (7, 8, 7)])
def test_ln_modulate(rows, cols, fs):
    x, frames = bf(rows, cols, seed=1), rows // fs
    sh, sc = bf(frames, cols, scale=0.2, seed=2), bf(frames, cols, scale=0.2, seed=3)
    ref = (wo.layer_norm(x[None], 1e-6).unflatten(1, (frames, fs)) * (1 + sc[None, :, None]) + sh[None, :, None]).flatten(1, 2)[0]
    out = ops.ln_modulate(x.to(DEV), shift=sh.to(DEV), scale=sc.to(DEV), tokens_per_frame=fs)
    assert rel_l2(out, ref) <= 1e-3 and same_frac(out, ref) >= 0.999
    # strided modulation views as the block passes them ([F, 6, C] -> [:, i])
    mod = bf(frames, 6, cols, scale=0.2, seed=4).to(DEV)
    out2 = ops.ln_modulate(x.to(DEV), shift=mod[:, 3], scale=mod[:, 4], tokens_per_frame=fs)
    ref2 = (wo.layer_norm(x[None], 1e-6).unflatten(1, (frames, fs)) * (1 + mod[:, 4].cpu()[None, :, None]) + mod[:, 3].cpu()[None, :, None]).flatten(1, 2)[0]
    assert rel_l2(out2, ref2) <= 1e-3


def test_ln_affine_and_rmsnorm():
    x = bf(130, 1536, seed=5)
    w, b = (1 + bf(1536, scale=0.1, seed=6).float()).bfloat16(), bf(1536, scale=0.1, seed=7)
    out = ops.ln_modulate(x.to(DEV), weight=w.to(DEV), bias=b.to(DEV))
    ref = wo.layer_norm(x, 1e-6, w, b)
    assert rel_l2(out, ref) <= 1e-3 and same_frac(out, ref) >= 0.999
    out = ops.rmsnorm(x.to(DEV), w.to(DEV))
    ref = wo.rms_norm(x, w, 1e-6)
    assert rel_l2(out, ref) <= 1e-3 and same_frac(out, ref) >= 0.999


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 1), (4, 2)])
def test_qk_norm_rope_append(world, rank):
    heads, hd, grid, start_frame = 2, 128, (3, 8, 8), 5
    C = heads * hd
    hw = grid[1] * grid[2]
    chunk = hw // world
    rows = grid[0] * chunk
    qkv = bf(rows, 3 * C, seed=8)
    wq, wk = (1 + bf(C, scale=0.1, seed=9).float()).bfloat16(), (1 + bf(C, scale=0.1, seed=10).float()).bfloat16()
    freqs = wo.rope_freqs(hd)
    q = wo.rms_norm(qkv[None, :, :C], wq, 1e-6).view(1, rows, heads, hd)
    k = wo.rms_norm(qkv[None, :, C:2 * C], wk, 1e-6).view(1, rows, heads, hd)
    q_ref = wo.causal_rope_apply(q, grid, freqs, start_frame, world, rank)[0].reshape(rows, C)
    k_ref = wo.causal_rope_apply(k, grid, freqs, start_frame, world, rank)[0].reshape(rows, C)
    g = RopeGrid(grid[0], grid[1], grid[2], start_frame, rank * chunk, chunk)
    q_out, k_out, v_out = ops.qk_norm_rope_append(qkv.to(DEV), wq.to(DEV), wk.to(DEV), ops.rope_table(freqs, DEV), g,
                                                  heads, hd)
    assert rel_l2(q_out, q_ref) <= 1e-3 and same_frac(q_out, q_ref) >= 0.999
    assert rel_l2(k_out, k_ref) <= 1e-3 and same_frac(k_out, k_ref) >= 0.999
    assert torch.equal(v_out.cpu(), qkv[:, 2 * C:])


# ----------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1, 8, 8), (300, 768, 256), (1000, 520, 1536), (257, 64, 8960),
                                   (4680, 1536, 1536)])
def test_gemm_bias(M, N, K):
    a, w, b = bf(M, K, seed=11), bf(N, K, scale=1 / math.sqrt(K), seed=12), bf(N, seed=13)
    ref = F.linear(a, w, b)                                    # CPU bf16 linear: fp32 accumulate, one rounding
    out = ops.gemm(a.to(DEV), w.to(DEV), b.to(DEV))
    assert rel_l2(out, ref) <= 1e-3 and same_frac(out, ref) >= 0.99


def test_gemm_epilogues():
    M, N, K, fs = 384, 512, 256, 128
    a, w, b = bf(M, K, seed=14), bf(N, K, scale=1 / math.sqrt(K), seed=15), bf(N, seed=16)
    res, gate = bf(M, N, seed=17), bf(M // fs, N, seed=18)
    t = F.linear(a, w, b)
    ref = res + (t.unflatten(0, (M // fs, fs)) * gate[:, None]).flatten(0, 1)
    out = ops.gemm(a.to(DEV), w.to(DEV), b.to(DEV), epilogue=ops.EPI_BIAS_GATE_RES, residual=res.to(DEV),
                   gate=gate.to(DEV), tokens_per_frame=fs)
    assert rel_l2(out, ref) <= 1e-3
    x = res.clone().to(DEV)                                     # in place on the residual stream, no gate
    ops.gemm(a.to(DEV), w.to(DEV), b.to(DEV), x, epilogue=ops.EPI_BIAS_GATE_RES, residual=x)
    assert rel_l2(x, res + t) <= 1e-3
    out = ops.gemm(a.to(DEV), w.to(DEV), b.to(DEV), epilogue=ops.EPI_BIAS_GELU)
    assert rel_l2(out, F.gelu(t, approximate="tanh")) <= 2e-3


def test_gemm_linearity_full_size():
    """Size-independent property at BASELINE config-2 shape: gemm(a1 + a2) == gemm(a1) + gemm(a2) (no bias)."""
    M, N, K = 10800, 1536, 1536
    a1, a2 = bf(M, K, seed=19).to(DEV), bf(M, K, seed=20).to(DEV)
    w = bf(N, K, scale=1 / math.sqrt(K), seed=21).to(DEV)
    s = ops.gemm(a1, w).float() + ops.gemm(a2, w).float()
    both = ops.gemm((a1.float() + a2.float()).bfloat16(), w)
    assert rel_l2(both, s) <= 6e-3     # inputs a1 + a2 re-rounded to bf16 + three output roundings


# ----------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("Lq,Lk,H", [(128, 128, 1), (1, 1, 1), (192, 64, 2), (300, 1000, 2), (257, 129, 3),
                                     (600, 5000, 2)])
def test_attention_vs_oracle(Lq, Lk, H):
    D = 128
    q, k, v = bf(Lq, H * D, seed=22), bf(Lk, H * D, seed=23), bf(Lk, H * D, seed=24)
    ref = wo.sdpa_attention(q.view(1, Lq, H, D).float(), k.view(1, Lk, H, D).float(), v.view(1, Lk, H, D).float())
    out = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), H)
    assert not torch.isnan(out).any()
    assert rel_l2(out, ref.reshape(Lq, H * D)) <= 4e-3


def test_attention_key_permutation_invariance_full_size():
    """Property the paged cache relies on: permuting (k, v) rows together does not change the output.
    Run at BASELINE config-2 shape (S = 10800, L = 86400, 12 heads)."""
    S, L, H, D = 10800, 86400, 12, 128
    g = torch.Generator(device=DEV).manual_seed(0)
    q = torch.randn(S, H * D, device=DEV, generator=g).bfloat16()
    k = torch.randn(L, H * D, device=DEV, generator=g).bfloat16()
    v = torch.randn(L, H * D, device=DEV, generator=g).bfloat16()
    o1 = ops.attention(q, k, v, H)
    perm = torch.randperm(L // 3600, device=DEV, generator=g)          # permute whole frames (pages)
    idx = (perm[:, None] * 3600 + torch.arange(3600, device=DEV)[None]).flatten()
    o2 = ops.attention(q, k[idx].contiguous(), v[idx].contiguous(), H)
    # each result is ~2.2e-3 from the fp32 answer (independent bf16 roundings of P and of the output), so two
    # differently-ordered runs sit sqrt(2) x 2.2e-3 = 3.1e-3 apart — the same distance as from FlashAttention-2.
    assert rel_l2(o2, o1) <= 5e-3
    # rows of the softmax sum to one: with v == 1 the output is exactly 1 up to bf16 rounding
    ones = torch.ones_like(v)
    o3 = ops.attention(q, k, ones, H)
    assert (o3.float() - 1).abs().max().item() <= 1e-2


def test_attention_reference_signature():
    from inferix_b200.attention import attention
    B, Lq, Lk, H, D = 2, 70, 200, 2, 128
    q, k, v = bf(B, Lq, H, D, seed=25), bf(B, Lk, H, D, seed=26), bf(B, Lk, H, D, seed=27)
    out = attention(q.to(DEV), k.to(DEV), v.to(DEV))
    ref = wo.sdpa_attention(q.float(), k.float(), v.float())
    assert out.shape == (B, Lq, H, D) and out.dtype == torch.bfloat16
    assert rel_l2(out, ref) <= 4e-3
    with pytest.raises(NotImplementedError):
        attention(q.to(DEV), k.to(DEV), v.to(DEV), causal=True)


# ----------------------------------------------------------------------------------------------- paged KV
@pytest.mark.parametrize("cache_frames,sink", [(6, 0), (6, 1), (7, 1)])
def test_paged_cache_matches_reference_roll(cache_frames, sink):
    """Contents in logical order == the reference's rolled tensor, bit for bit, over 6 blocks with eviction."""
    fs, block, heads, hd = 16, 3, 2, 128
    C = heads * hd
    kv = ops.PagedKV(cache_frames, fs, heads, hd, DEV)
    oc = wo.LayerCache(torch.zeros(1, cache_frames * fs, heads, hd, dtype=torch.bfloat16),
                       torch.zeros(1, cache_frames * fs, heads, hd, dtype=torch.bfloat16))
    for b in range(6):
        for rep in range(2):
            kn, vn = bf(block * fs, C, seed=100 + 2 * b + rep), bf(block * fs, C, seed=200 + 2 * b + rep)
            plan = kv.plan_append(b * block * fs, block * fs, sink * fs, True)
            kv.append(plan, kn.to(DEV), vn.to(DEV))
            ls, le = wo.cache_append(oc, kn.view(1, -1, heads, hd), vn.view(1, -1, heads, hd), b * block * fs,
                                     sink * fs, True)
            assert (plan.local_start, plan.local_end, plan.global_end) == (ls, le, oc.global_end)
            ke, ve = kv.export(0, le)
            assert torch.equal(ke.cpu(), oc.k[0, :le].reshape(le, C))
            assert torch.equal(ve.cpu(), oc.v[0, :le].reshape(le, C))
            # attention over the paged window == attention over the reference's contiguous window
            q = bf(40, C, seed=300 + b).to(DEV)
            o_paged = kv.attention(q)
            o_lin = ops.attention(q, oc.k[0, :le].reshape(le, C).to(DEV), oc.v[0, :le].reshape(le, C).to(DEV), heads)
            assert rel_l2(o_paged, o_lin) <= 5e-3
    kv.free()
    with pytest.raises(KeyError):
        kv.handle


def test_sp_append_layout():
    """all-gathered rank-major rows land in (frame, rank, hw) order == single-process token order."""
    world, frames, chunk, heads, hd = 4, 3, 8, 1, 128
    fs = world * chunk
    kv = ops.PagedKV(6, fs, heads, hd, DEV)
    full_k, full_v = bf(frames * fs, hd, seed=31), bf(frames * fs, hd, seed=32)
    shard = lambda t: torch.stack([t.view(frames, world, chunk, hd)[:, r].reshape(frames * chunk, hd) for r in range(world)])
    plan = kv.plan_append(0, frames * fs, 0, True)
    kv.append_sp(plan, shard(full_k).to(DEV).contiguous(), shard(full_v).to(DEV).contiguous(), frames)
    ke, ve = kv.export(0, frames * fs)
    assert torch.equal(ke.cpu(), full_k) and torch.equal(ve.cpu(), full_v)
    # K and V out of one [world, 2, rows, C] gather buffer (strided rank views)
    both = torch.stack([shard(full_v), shard(full_k)], dim=1).to(DEV).contiguous()
    kv.append_sp(kv.plan_append(frames * fs, frames * fs, 0, True), both[:, 0], both[:, 1], frames)
    ke, ve = kv.export(frames * fs, frames * fs)
    assert torch.equal(ke.cpu(), full_v) and torch.equal(ve.cpu(), full_k)


def test_kv_manager_reference_api():
    from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest, KVCacheRequestSpec, KVCacheSpec
    mgr = KVCacheManager(DEV)
    req = KVCacheRequest("req_0")
    spec = KVCacheRequestSpec(num_tokens=10, block_size=1, specs={
        "layer_0": KVCacheSpec(num_kv_heads=2, head_size=128, dtype=torch.bfloat16, kv_offload=False, use_mla=False)})
    mgr.allocate_slots(req, spec)
    with pytest.raises(ValueError, match="already exists"):
        mgr.allocate_slots(req, spec)
    assert list(mgr.layers(req)) == ["layer_0"] and mgr.layers(KVCacheRequest("nope")) == ()
    new = bf(2, 4, 1, 2, 128, seed=33).to(DEV)
    mgr.set(req, "layer_0", 3, 4, new)
    got = mgr.get(req, "layer_0")
    assert got.shape == (2, 10, 1, 2, 128) and torch.equal(got[:, 3:7], new)
    assert torch.equal(mgr.get_range(req, "layer_0", 4, 2), new[:, 1:3])
    assert torch.equal(mgr.select(req, "layer_0", [6, 3]), new[:, [3, 0]])
    assert mgr.layer_spec(req, "layer_0").num_blocks == 10
    mgr.free_layer(req, "layer_0")
    with pytest.raises(KeyError):
        mgr.get(req, "layer_0")
    mgr.free(req)
    with pytest.raises(KeyError):
        mgr.free(req)


# ----------------------------------------------------------------------------------------------- MAGI pieces
def test_gqa_range_attention():
    """Grouped-query attention over per-chunk key ranges (MAGI dit_module.py:1000-1014) vs an fp32 reference."""
    from inferix_b200 import magi_schedule as ms
    heads, kv_heads, D, clip = 6, 2, 128, 96
    kr = ms.noise2clean_kvrange(2, 3, clip, [3, 2], -1, [3, 1, 0], 4).tolist()      # 3 denoising chunks after 2 clean
    qr = [[i * clip, (i + 1) * clip] for i in range(3)]
    Sk = 5 * clip
    q, k, v = bf(3 * clip, heads * D, seed=40), bf(Sk, kv_heads * D, seed=41), bf(Sk, kv_heads * D, seed=42)
    out = ops.attention_ranges(q.to(DEV), k.to(DEV), v.to(DEV), qr, kr, heads, kv_heads)
    for (qs, qe), (ks, ke) in zip(qr, kr):
        qq = q[qs:qe].view(1, -1, heads, D).float()
        kk = k[ks:ke].view(1, -1, kv_heads, D).float().repeat_interleave(heads // kv_heads, dim=2)
        vv = v[ks:ke].view(1, -1, kv_heads, D).float().repeat_interleave(heads // kv_heads, dim=2)
        ref = wo.sdpa_attention(qq, kk, vv)[0].reshape(qe - qs, heads * D)
        assert rel_l2(out[qs:qe], ref) <= 4e-3


def test_magi_kv_cache_manager_matches_reference(golden_dir):
    """MagiKVCacheManager on the native cache vs the reference's adapter run on its CPU KVCacheManager."""
    from inferix_b200.kvcache_manager.model import InferenceParams, KVMetaArgs, MagiKVCacheManager
    gold = torch.load(golden_dir / "magi_kv.pt", weights_only=False)
    mgr = MagiKVCacheManager(3, gold["hn"], gold["d"])
    ip = InferenceParams(1, gold["clip"] * gold["chunks"], device=DEV)
    for st in gold["steps"]:
        ip.update_kv_cache = st["update"]
        k, v = mgr.adjust_key_and_value_for_inference(st["kv"].to(DEV), ip, KVMetaArgs(**st["meta"]))
        assert torch.equal(k.cpu(), st["k"]) and torch.equal(v.cpu(), st["v"])
    mapped = ip.kv_cache_manager.get_raw(ip.kv_cache_request, "layer_3")
    written = gold["chunks"] * gold["clip"]                  # the scenario stores all four clips
    assert torch.equal(mapped[:, :written].cpu(), gold["final_cache"][:, :written])
    mgr.clear_cache(ip)
    assert not mgr.is_cached(ip)


def test_two_phase_attention_equals_single_pass():
    """attention over [old extents] + [new extents] merged by the combine kernel == one pass over all keys
    (the sequence-parallel overlap path; extents are whole pages in arbitrary physical order)."""
    heads, D, pt, pages = 3, 128, 200, 7
    q = bf(700, heads * D, seed=50).to(DEV)
    k, v = bf(pages * pt, heads * D, seed=51).to(DEV), bf(pages * pt, heads * D, seed=52).to(DEV)
    ref = ops.attention(q, k, v, heads)
    new_ext = ops._coalesce_pages([5, 1, 2], pt)            # -> rows [200, 600) and [1000, 1200)
    old_ext = ops._coalesce_pages([0, 3, 4, 6], pt)
    assert new_ext == [(200, 400), (1000, 200)] and old_ext == [(0, 200), (600, 400), (1200, 200)]
    for n_old in (1, 2, 5):
        ws = torch.empty(ops.attention_workspace_bytes(700, heads, n_old + 1) // 4, dtype=torch.float32, device=DEV)
        ops.attention_partial(q, k, v, old_ext, heads, ws, n_old + 1, 0, n_old)
        ops.attention_partial(q, k, v, new_ext, heads, ws, n_old + 1, n_old, 1)
        out = ops.attention_combine(ws, n_old + 1, torch.empty_like(q), heads)
        assert rel_l2(out, ref) <= 5e-3, n_old
    # one phase only (first block of a video: nothing old yet)
    ws = torch.empty(ops.attention_workspace_bytes(700, heads, 1) // 4, dtype=torch.float32, device=DEV)
    ops.attention_partial(q, k, v, [(0, pages * pt)], heads, ws, 1, 0, 1)
    assert rel_l2(ops.attention_combine(ws, 1, torch.empty_like(q), heads), ref) <= 5e-3
