"""MAGI-1 layer on the GPU: each new row kernel against the oracle's torch statement, the zero-copy cache rows, and
the whole TransformerBlock (2 layers, 5-forward KV scenario) against the goldens of the reference's own module.

Tolerances: row kernels reproduce the reference's rounding points -> rel-L2 <= 1e-3 and >= 99.9 % identical elements
(99.5 % where fp32 rotary / LayerNorm order differs before the bf16 rounding); the block is judged the way SURVEY §8d
states for bf16: relL2(ours, oracle_fp32) <= 1.5 x relL2(reference_bf16, oracle_fp32), plus an absolute bar."""
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import fake_magi_ops as fk                         # noqa: E402  (CPU statements of the kernel contracts)
from inferix_b200 import magi_layer, ops           # noqa: E402
from magi_golden_util import meta_from_plain       # noqa: E402
from oracle import magi_oracle as mo               # noqa: E402

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def same_frac(a, b):
    return (a.cpu() == b.cpu()).float().mean().item()


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).bfloat16()


def f32(*shape, scale=1.0, seed=0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale + shift


# ----------------------------------------------------------------------------------------------- row kernels
@pytest.mark.parametrize("rows,hq,g,groups", [(200, 4, 2, 1), (77, 24, 8, 1), (96, 4, 2, 2), (50, 24, 8, 8)])
def test_magi_qkv_post(rows, hq, g, groups):
    d = 128
    qkvx = bf(rows, (2 * hq + 2 * g) * d, scale=1.5, seed=1)
    q_ln = (f32(d, scale=0.1, seed=2, shift=1.0), f32(d, scale=0.05, seed=3))
    k_ln = (f32(d, scale=0.1, seed=4, shift=1.0), f32(d, scale=0.05, seed=5))
    x_ln = (f32(d, scale=0.1, seed=6, shift=1.0).bfloat16(), f32(d, scale=0.05, seed=7).bfloat16())
    ang = f32(rows, 48, scale=2.0, seed=8)
    rope = torch.cat([ang.sin(), ang.cos()], dim=-1)
    shp_q = (rows, hq * d) if groups == 1 else (groups, rows, hq // groups * d)
    shp_k = (rows, g * d) if groups == 1 else (groups, rows, g // groups * d)
    ref = [torch.empty(shp_q, dtype=torch.bfloat16), torch.empty(shp_k, dtype=torch.bfloat16),
           torch.empty(shp_k, dtype=torch.bfloat16), torch.empty(rows, hq * d, dtype=torch.bfloat16)]
    fk.magi_qkv_post(qkvx, hq, g, q_ln, k_ln, x_ln, rope, ref[0], ref[1], ref[2], ref[3], groups=groups)
    if groups == 1:
        # K / V land in rows [5, 5 + rows) of a wider cache-like buffer
        cache_k = torch.zeros(rows + 9, g * d, dtype=torch.bfloat16, device=DEV)
        cache_v = torch.zeros_like(cache_k)
        k_dst, v_dst = cache_k[5:5 + rows], cache_v[5:5 + rows]
    else:
        k_dst = torch.empty(shp_k, dtype=torch.bfloat16, device=DEV)
        v_dst = torch.empty_like(k_dst)
    q_out = torch.empty(shp_q, dtype=torch.bfloat16, device=DEV)
    qx_out = torch.empty(rows, hq * d, dtype=torch.bfloat16, device=DEV)
    to = lambda t: tuple(u.to(DEV) for u in t)
    ops.magi_qkv_post(qkvx.to(DEV), hq, g, to(q_ln), to(k_ln), to(x_ln), rope.to(DEV), q_out, k_dst, v_dst, qx_out,
                      groups=groups)
    assert torch.equal(v_dst.cpu(), ref[2])                                   # raw copy: bit-exact
    for got, want, name in ((q_out, ref[0], "q"), (k_dst, ref[1], "k"), (qx_out, ref[3], "qx")):
        assert rel_l2(got, want) <= 1e-3 and same_frac(got, want) >= 0.995, (name, rel_l2(got, want), same_frac(got, want))
    if groups == 1:
        assert float(cache_k[:5].abs().max()) == 0 and float(cache_k[5 + rows:].abs().max()) == 0


def test_head_layernorm_in_place_on_strided_rows():
    rows, g, d = 38, 8, 128
    kvx = bf(rows, 2 * g * d, scale=2.0, seed=11)
    w, b = f32(d, scale=0.1, seed=12, shift=1.0).bfloat16(), f32(d, scale=0.05, seed=13).bfloat16()
    ref = F.layer_norm(kvx[:, :g * d].reshape(rows, g, d), (d,), w, b, 1e-6).reshape(rows, -1)
    dev = kvx.to(DEV)
    ops.head_layernorm(dev[:, :g * d], g, w.to(DEV), b.to(DEV))
    assert rel_l2(dev[:, :g * d], ref) <= 1e-3 and same_frac(dev[:, :g * d], ref) >= 0.999
    assert torch.equal(dev[:, g * d:].cpu(), kvx[:, g * d:])                  # V half untouched


@pytest.mark.parametrize("rows,cols,ranges", [(192, 256, 2), (100, 3072, 3), (64, 6144, 4), (33, 2048, 1)])
def test_gate_norm_residual(rows, cols, ranges):
    x, res = bf(rows, cols, scale=0.7, seed=21), bf(rows, cols, seed=22)
    gate = torch.tanh(f32(ranges, 2 * cols, seed=23)).bfloat16()
    rmap = (torch.arange(rows) * ranges // rows).to(torch.int32)
    nw, nb = f32(cols, scale=0.1, seed=24, shift=1.0), f32(cols, scale=0.05, seed=25)
    for half in (0, 1):
        gh = gate[:, half * cols:(half + 1) * cols]
        ref = torch.empty(rows, cols, dtype=torch.bfloat16)
        fk.gate_norm_residual(x, gh, rmap, nw, nb, res, ref)
        gd = gate.to(DEV)[:, half * cols:(half + 1) * cols]                    # strided view, as the layer passes it
        out = ops.gate_norm_residual(x.to(DEV), gd, rmap.to(DEV), nw.to(DEV), nb.to(DEV), res.to(DEV))
        assert rel_l2(out, ref) <= 1e-3 and same_frac(out, ref) >= 0.999, (rel_l2(out, ref), same_frac(out, ref))
    # fp32 x (the output projection's IFX_EPI_BIAS_F32 result)
    x32 = f32(rows, cols, scale=0.7, seed=26)
    fk.gate_norm_residual(x32, gate[:, :cols], rmap, nw, nb, res, ref)
    out = ops.gate_norm_residual(x32.to(DEV), gate.to(DEV)[:, :cols], rmap.to(DEV), nw.to(DEV), nb.to(DEV), res.to(DEV))
    assert rel_l2(out, ref) <= 1e-3 and same_frac(out, ref) >= 0.999
    # in place on the residual (second call of the layer)
    r2 = res.to(DEV).clone()
    ops.gate_norm_residual(x.to(DEV), gate.to(DEV)[:, :cols], rmap.to(DEV), nw.to(DEV), nb.to(DEV), r2, r2)
    fk.gate_norm_residual(x, gate[:, :cols], rmap, nw, nb, res, ref)
    assert rel_l2(r2, ref) <= 1e-3


def test_silu_mul_and_gelu_erf_epilogue():
    rows, f, k = 130, 512, 256
    x = bf(rows, 2 * f, scale=2.0, seed=31)
    out = ops.silu_mul(x.to(DEV))
    ref = mo.silu_and_mul(x)
    assert rel_l2(out, ref) <= 1e-3 and same_frac(out, ref) >= 0.995
    a, w = bf(rows, k, seed=32), bf(f, k, scale=k ** -0.5, seed=33)
    out = ops.gemm(a.to(DEV), w.to(DEV), None, epilogue=ops.EPI_BIAS_GELU_ERF)
    ref = F.gelu((a.float() @ w.float().t()).bfloat16())
    assert rel_l2(out, ref) <= 2e-3
    out = ops.gemm(a.to(DEV), w.to(DEV), None, epilogue=ops.EPI_BIAS_F32)      # fp32 result, no bf16 rounding
    assert out.dtype == torch.float32 and rel_l2(out, a.float() @ w.float().t()) <= 2e-5
    big = bf(700, k, seed=35)
    bias = bf(f, scale=0.1, seed=36)
    out = ops.gemm(big.to(DEV), w.to(DEV), bias.to(DEV), epilogue=ops.EPI_BIAS_F32)
    assert rel_l2(out, big.float() @ w.float().t() + bias.float()) <= 2e-5
    m = 700                                                                    # 2-CTA kernel path (M >= 256)
    a = bf(m, k, seed=34)
    out = ops.gemm(a.to(DEV), w.to(DEV), None, epilogue=ops.EPI_BIAS_GELU_ERF)
    assert rel_l2(out, F.gelu((a.float() @ w.float().t()).bfloat16())) <= 2e-3


def test_kv_map_rows_zero_copy_and_refusal_after_rotation():
    store = ops.PagedKV(64, 1, 2, 128, DEV)                                    # MAGI allocates block_size 1
    k, v = store.map_rows(40)
    data_k, data_v = bf(40, 256, seed=41).to(DEV), bf(40, 256, seed=42).to(DEV)
    k.copy_(data_k)
    v.copy_(data_v)
    ek, ev = store.export(8, 24)                                               # logical order == physical order
    assert torch.equal(ek, data_k[8:32]) and torch.equal(ev, data_v[8:32])
    with pytest.raises(IndexError):
        store.map_rows(65)
    frames = ops.PagedKV(4, 8, 2, 128, DEV)                                    # windowed Wan cache: rotate, then refuse
    for blk in range(3):
        frames.plan_append(blk * 16, 16, 0, True)
    with pytest.raises(NotImplementedError):
        frames.map_rows(8)


# ----------------------------------------------------------------------------------------------- the block
def build_block(cfg_dict, seed):
    mc = types.SimpleNamespace(layernorm_epsilon=1e-6, apply_layernorm_1p=False, cond_hidden_ratio=0.25,
                               cond_gating_ratio=1.0, xattn_cond_hidden_ratio=1.0, params_dtype=torch.bfloat16,
                               **cfg_dict)
    ec = types.SimpleNamespace(cp_size=1, cp_strategy="none", fp8_quant=False, kv_offload=False)
    block = magi_layer.TransformerBlock(mc, ec)
    sd = mo.synth_state_dict(mo.MagiConfig(**cfg_dict), seed=seed)
    block.load_state_dict(sd, strict=True)
    return block.to(DEV), sd


@pytest.mark.parametrize("name", ["gelu", "glu"])
def test_block_matches_reference_goldens(golden_dir, name):
    from inferix_b200.kvcache_manager.model import InferenceParams
    g = torch.load(golden_dir / f"magi_layer_{name}.pt")
    cfg = mo.MagiConfig(**g["cfg"])
    block, sd = build_block(g["cfg"], g["seed"])
    sd32 = {k: v.float() for k, v in sd.items()}
    ip = InferenceParams(1, g["max_seq"], device=DEV)
    c32 = mo.OracleMagiCache(g["max_seq"])
    for i, st in enumerate(g["steps"]):
        meta = meta_from_plain(st["meta"])
        ip.update_kv_cache = c32.update_kv_cache = st["update"]
        out = block(st["hidden"].to(DEV), st["condition"].to(DEV), st["condition_map"].to(DEV), st["y"].to(DEV),
                    st["rope"].to(DEV), ip, meta)
        torch.cuda.synchronize()
        assert out.dtype == torch.float32 and out.shape == st["out"].shape and bool(torch.isfinite(out).all())
        ref32 = mo.block_forward(sd32, cfg, st["hidden"].float(), st["condition"].float(), st["condition_map"],
                                 st["y"].float(), st["rope"], c32, meta)
        gap_ref, gap_ours, direct = rel_l2(st["out"], ref32), rel_l2(out, ref32), rel_l2(out, st["out"])
        print(f"{name} forward {i}: ours-vs-fp32 {gap_ours:.2e}  reference-bf16-vs-fp32 {gap_ref:.2e}  "
              f"ours-vs-reference {direct:.2e}")
        assert gap_ours <= 1.5 * gap_ref + 1e-3, (i, gap_ours, gap_ref)
        assert direct <= 3e-2, (i, direct)
    # cache rows the reference holds at the end (3 clips): K post-LN post-rotary, V raw
    for layer, ref in g["cache_prefix"].items():
        store = ip.kv_cache_manager.store(ip.kv_cache_request, f"layer_{layer}")
        n = ref.shape[1]
        k, v = store.export(0, n)
        assert rel_l2(k, ref[0, :, 0].reshape(n, -1)) <= 3e-2 and rel_l2(v, ref[1, :, 0].reshape(n, -1)) <= 3e-2
    block.clear_kv_cache(ip)
    assert not block.layers[0].self_attention.kv_cache_manager.is_cached(ip)


def test_single_layer_first_forward_is_tight(golden_dir):
    """Layer 0 alone, no cache: two bf16 implementations of one layer sit ~sqrt(2) x the bf16-vs-fp32 gap of one layer
    (4-5e-3) apart."""
    g = torch.load(golden_dir / "magi_layer_glu.pt")
    cfg = mo.MagiConfig(**g["cfg"])
    block, sd = build_block(g["cfg"], g["seed"])
    st = g["steps"][4]
    meta = meta_from_plain(st["meta"])
    out = block.layers[0](st["hidden"].to(DEV), st["condition"].to(DEV), st["condition_map"].to(DEV),
                          st["y"].to(DEV), st["rope"].to(DEV), None, meta)
    ref = mo.layer_forward(sd, 0, cfg, st["hidden"].clone(), st["condition"], st["condition_map"], st["y"], st["rope"],
                           None, meta)
    err = rel_l2(out, ref)
    print(f"single layer rel-L2 {err:.2e}")
    assert err <= 1.2e-2


def test_layer_at_magi_4p5b_width():
    """One layer at MAGI-1 4.5B widths (hidden 3072, 24 heads / 8 KV groups, ffn 12288, non-gated) on 2 x 384 tokens
    with one cached clip: exercises the multi-warp row kernels and the real GEMM shapes."""
    from inferix_b200.kvcache_manager.model import InferenceParams
    cfgd = dict(hidden_size=3072, ffn_hidden_size=12288, num_attention_heads=24, num_query_groups=8, kv_channels=128,
                num_layers=1, gated_linear_unit=False)
    cfg = mo.MagiConfig(**cfgd)
    block, sd = build_block(cfgd, seed=9)
    sd32 = {k: v.float() for k, v in sd.items()}
    clip = 384
    ip = InferenceParams(1, 4 * clip, device=DEV)
    cb, c32 = mo.OracleMagiCache(4 * clip), mo.OracleMagiCache(4 * clip)
    g = torch.Generator().manual_seed(10)
    plan = [(1, 0, [[0, clip]], [40], True, dict(extract_prefix_video_feature=True)),
            (2, 1, [[0, 2 * clip], [clip, 3 * clip]], [33, 50], False, {})]
    for ranges, sp, kr, ylens, update, flags in plan:
        s = ranges * clip
        hidden = torch.randn(s, 1, 3072, generator=g).bfloat16()
        cond = torch.randn(1, ranges, 768, generator=g).bfloat16()
        cmap = torch.arange(ranges).repeat_interleave(clip).reshape(-1, 1)
        y = torch.randn(sum(ylens), 3072, generator=g).bfloat16()
        ang = torch.randn(s, 48, generator=g) * 2
        rope = torch.cat([ang.sin(), ang.cos()], -1)
        cu_q = [i * clip for i in range(ranges + 1)]
        cu_k = [0]
        for n in ylens:
            cu_k.append(cu_k[-1] + n)
        meta = meta_from_plain(dict(slice_point=sp, denoising_range_num=ranges, clip_token_nums=clip,
                                    extract_prefix_video_feature=flags.get("extract_prefix_video_feature", False),
                                    fwd_extra_1st_chunk=False, distill_nearly_clean_chunk=False,
                                    q_range=[[cu_q[i], cu_q[i + 1]] for i in range(ranges)], k_range=kr,
                                    cu_seqlens_q=cu_q, cu_seqlens_kv=cu_k))
        ip.update_kv_cache = cb.update_kv_cache = c32.update_kv_cache = update
        out = block(hidden.to(DEV), cond.to(DEV), cmap.to(DEV), y.to(DEV), rope.to(DEV), ip, meta)
        ref = mo.block_forward(sd, cfg, hidden.clone(), cond, cmap, y, rope, cb, meta)
        ref32 = mo.block_forward(sd32, cfg, hidden.float(), cond.float(), cmap, y.float(), rope, c32, meta)
        gap_ref, gap_ours = rel_l2(ref, ref32), rel_l2(out, ref32)
        print(f"4.5B-width layer, {ranges} range(s): ours-vs-fp32 {gap_ours:.2e}  oracle-bf16-vs-fp32 {gap_ref:.2e}  "
              f"ours-vs-oracle {rel_l2(out, ref):.2e}")
        assert gap_ours <= 1.5 * gap_ref + 1e-3


@pytest.mark.parametrize("glu", [False, True])
def test_fp8_quant_block_matches_oracle(glu):
    """engine_config.fp8_quant: the middle layers run the reference's quantised linears (dit_module.py:410,525,538,865):
    PerTensorQuantizedFp8Linear for q / k / v / qx / fc1 (per-input-channel divisor, per-tensor GEMM scales) and
    PerChannelQuantizedFp8Linear for linear_proj / fc2 (smooth_scale divisor); first and last layer stay bf16.  Four
    layers at small widths, 2 forwards sharing the KV cache, against the oracle's restatement (magi_oracle.qlinear)
    on the SAME quantised parameters: the e4m3 codes are bit-exact per linear, so the distance is the bf16 one."""
    from inferix_b200.kvcache_manager.model import InferenceParams
    cfgd = dict(hidden_size=512, ffn_hidden_size=1024, num_attention_heads=4, num_query_groups=2, kv_channels=128,
                num_layers=4, gated_linear_unit=glu)
    cfg = mo.MagiConfig(**cfgd)
    mc = types.SimpleNamespace(layernorm_epsilon=1e-6, apply_layernorm_1p=False, cond_hidden_ratio=0.25,
                               cond_gating_ratio=1.0, xattn_cond_hidden_ratio=1.0, params_dtype=torch.bfloat16, **cfgd)
    ec = types.SimpleNamespace(cp_size=1, cp_strategy="none", fp8_quant=True, kv_offload=False)
    block = magi_layer.TransformerBlock(mc, ec)
    sd = mo.synth_state_dict(cfg, seed=21)                      # bf16 weights under the reference names
    sd_bf16 = dict(sd)
    g = torch.Generator().manual_seed(22)
    # quantise the middle layers' linears from the bf16 weights (what a quantised checkpoint would contain)
    for li in (1, 2):
        layer = block.layers[li]
        a, m = layer.self_attention, layer.mlp
        for name, lin, pfx in [(n, getattr(a.linear_qkv, n), f"layers.{li}.self_attention.linear_qkv.{n}") for n in ("q", "k", "v", "qx")] + \
                              [("proj", a.linear_proj, f"layers.{li}.self_attention.linear_proj"),
                               ("fc1", m.linear_fc1, f"layers.{li}.mlp.linear_fc1"), ("fc2", m.linear_fc2, f"layers.{li}.mlp.linear_fc2")]:
            w = sd.pop(pfx + ".weight")
            if isinstance(lin, magi_layer.PerChannelQuantizedFp8Linear):
                lin.quantize_from(w, input_amax=6.0, smooth=0.5 + torch.rand(w.shape[1], generator=g))
            else:
                lin.quantize_from(w, input_amax=6.0)
                lin.input_scale.mul_(1.0 + 0.25 * torch.rand(w.shape[1], generator=g))   # a genuine per-channel vector
            for pn, pv in lin.state_dict().items():
                sd[f"{pfx}.{pn}"] = pv.clone()
    block.load_state_dict(sd, strict=True)
    block = block.to(DEV)
    assert block.layers[1]._pack()["fp8"] and not block.layers[0]._pack()["fp8"]
    # each quantised linear type on its own first: codes bit-exact, product within the fp32 summation order
    for pfx, lin in (("layers.1.mlp.linear_fc2", block.layers[1].mlp.linear_fc2),
                     ("layers.2.self_attention.linear_qkv.k", block.layers[2].self_attention.linear_qkv.k)):
        xin = (torch.randn(300, lin.in_features, generator=g) * 2).bfloat16()
        div = sd[pfx + (".smooth_scale" if pfx.endswith("fc2") else ".input_scale")]
        codes = ops.quantize_fp8_cols(xin.to(DEV), lin.divisor().reshape(-1).contiguous())
        assert torch.equal(codes.cpu().view(torch.uint8), mo.div_clamp_to(xin, div).view(torch.uint8)), pfx
        assert rel_l2(lin(xin.to(DEV)), mo.qlinear(sd, pfx, xin)) <= 1e-3, pfx
    clip = 256
    ip = InferenceParams(1, 4 * clip, device=DEV)
    cb, cb16 = mo.OracleMagiCache(4 * clip), mo.OracleMagiCache(4 * clip)
    plan = [(1, 0, [[0, clip]], [40], True, dict(extract_prefix_video_feature=True)),
            (2, 1, [[0, 2 * clip], [clip, 3 * clip]], [33, 50], False, {})]
    for ranges, sp, kr, ylens, update, flags in plan:
        s = ranges * clip
        hidden = torch.randn(s, 1, 512, generator=g).bfloat16()
        cond = torch.randn(1, ranges, 128, generator=g).bfloat16()
        cmap = torch.arange(ranges).repeat_interleave(clip).reshape(-1, 1)
        y = torch.randn(sum(ylens), 512, generator=g).bfloat16()
        ang = torch.randn(s, 48, generator=g) * 2
        rope = torch.cat([ang.sin(), ang.cos()], -1)
        cu_q = [i * clip for i in range(ranges + 1)]
        cu_k = [0]
        for n in ylens:
            cu_k.append(cu_k[-1] + n)
        meta = meta_from_plain(dict(slice_point=sp, denoising_range_num=ranges, clip_token_nums=clip,
                                    extract_prefix_video_feature=flags.get("extract_prefix_video_feature", False),
                                    fwd_extra_1st_chunk=False, distill_nearly_clean_chunk=False,
                                    q_range=[[cu_q[i], cu_q[i + 1]] for i in range(ranges)], k_range=kr,
                                    cu_seqlens_q=cu_q, cu_seqlens_kv=cu_k))
        ip.update_kv_cache = cb.update_kv_cache = cb16.update_kv_cache = update
        out = block(hidden.to(DEV), cond.to(DEV), cmap.to(DEV), y.to(DEV), rope.to(DEV), ip, meta)
        ref = mo.block_forward(sd, cfg, hidden.clone(), cond, cmap, y, rope, cb, meta)
        ref16 = mo.block_forward(sd_bf16, cfg, hidden.clone(), cond, cmap, y, rope, cb16, meta)
        err, quant_effect = rel_l2(out, ref), rel_l2(ref, ref16)
        print(f"fp8_quant block ({'glu' if glu else 'gelu'}), {ranges} range(s): ours-vs-oracle {err:.2e}  "
              f"(quantisation itself moves the oracle by {quant_effect:.2e})")
        # Two bf16 executions of the same quantised stack differ where an activation lands on the other side of an e4m3
        # rounding boundary (3 mantissa bits: a flipped code moves that input by 6 %); the distance stays a fraction of
        # what quantisation itself does to the output.
        assert bool(torch.isfinite(out).all()) and err <= 4e-2 and err <= 0.6 * quant_effect


# ----------------------------------------------------------------------------------------------- the whole model
def test_video_dit_model_and_cfg_dispatcher_match_reference(golden_dir):
    """VideoDiTModel.forward (3 forwards sharing a cache) and forward_dispatcher (cfg_number 3 and 1, incl. the batched
    unconditional pass and the distill branch) on the real kernels vs the reference's own model run on CPU.  Two bf16
    implementations of the 2-layer stack sit ~7e-3 apart (see test_block_matches_reference_goldens); bar 2e-2."""
    from inferix_b200 import magi_model
    from inferix_b200.kvcache_manager.model import InferenceParams
    g = torch.load(golden_dir / "magi_model.pt")

    def build(cfg_number):
        mc = types.SimpleNamespace(model_name="tiny", params_dtype=torch.bfloat16, layernorm_epsilon=1e-6,
                                   apply_layernorm_1p=False, **g["model"])
        rc = types.SimpleNamespace(cfg_number=cfg_number, chunk_width=g["chunk_width"],
                                   cfg_t_range=[0, 0.0217, 0.1000, 0.3, 0.999], prev_chunk_scales=[1.5] * 5,
                                   text_scales=[7.5] * 5)
        ec = types.SimpleNamespace(cp_strategy="none", cp_size=1, fp8_quant=False, kv_offload=False, distill=False)
        m = magi_model.VideoDiTModel(types.SimpleNamespace(model_config=mc, runtime_config=rc, engine_config=ec))
        m.load_state_dict(mo.synth_model_state_dict(m, seed=g["seed"]), strict=True)
        return m.eval().to(DEV)

    dev = lambda t: t.to(DEV)
    model = build(1)
    ip = InferenceParams(1, 6 * g["clip"], device=DEV)
    for i, st in enumerate(g["forward"]):
        ip.update_kv_cache = st["update"]
        out = model(dev(st["x"]), dev(st["t"]), dev(st["y"]), caption_dropout_mask=torch.tensor([False], device=DEV),
                    xattn_mask=dev(st["mask"]), kv_range=dev(st["kv_range"]), inference_params=ip, **dict(st["kwargs"]))
        err = rel_l2(out, st["out"])
        print(f"model forward {i}: rel-L2 vs reference {err:.2e}")
        assert out.shape == st["out"].shape and err <= 2e-2
    for st in g["dispatch"]:
        model = build(st["cfg_number"])
        ip = InferenceParams(1, 6 * g["clip"], device=DEV)
        pf = st["prefix"]
        ip.update_kv_cache = True
        model(dev(pf["x"]), dev(pf["t"]), dev(pf["y"]), caption_dropout_mask=torch.tensor([False], device=DEV),
              xattn_mask=dev(pf["mask"]), kv_range=torch.tensor([[0, g["clip"]]], dtype=torch.int32, device=DEV),
              inference_params=ip, range_num=1, denoising_range_num=1, slice_point=0, chunk_width=g["chunk_width"],
              num_steps=12, distill_interval=4, extract_prefix_video_feature=True, fwd_extra_1st_chunk=False)
        out = model.forward_dispatcher(dev(st["x"]).clone(), dev(st["t"]), dev(st["y"]), dev(st["mask"]),
                                       dev(st["kv_range"]), ip, **dict(st["kwargs"]))
        err = rel_l2(out, st["out"])
        print(f"dispatcher cfg={st['cfg_number']} extra={st['kwargs']['fwd_extra_1st_chunk']}: rel-L2 {err:.2e}")
        assert out.shape == st["out"].shape and err <= 2.5e-2


def test_end_to_end_magi_job_on_gpu(golden_dir):
    """The whole MAGI-1 path on the real kernels: native SampleTransport + VideoDiTModel (3 chunks, window 2, 4 steps,
    3-way CFG = 24 model forwards sharing the native KV cache) vs the reference's scheduler + model run on CPU.
    The CFG combination (scales 1.5 / 7.5) amplifies the ~6e-3 bf16 distance of single forwards about tenfold."""
    from inferix_b200.kvcache_manager.model import InferenceParams
    from magi_e2e_util import run_native
    g = torch.load(golden_dir / "magi_e2e.pt")
    ip = InferenceParams(1, g["final_x"].shape[2] * (g["hw"] // 2) ** 2, device=DEV)
    chunks, final_x = run_native(g, torch.device(DEV), ip)
    assert [i for i, _ in chunks] == [i for i, _ in g["chunks"]]
    for (i, a), (_, b) in zip(chunks, g["chunks"]):
        err = rel_l2(a, b)
        print(f"e2e chunk {i}: rel-L2 vs reference stack {err:.2e}")
        assert bool(torch.isfinite(a).all()) and err <= 8e-2
