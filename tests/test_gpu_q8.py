"""Dynamic 8-bit linears (per-token activation x per-channel weight scales; e4m3 and int8) against the oracle's
restatement (oracle/wan_oracle.py: dynamic_q8_quantize / dynamic_q8_linear).  This is the qconfig the reference's
quantisation examples request from DAX (example/quantization/run_causvid_quantized.py:32-37); DAX is not vendored, so
the oracle restates the published scheme: PARITY UNPINNED (kernel == oracle is proven, oracle == DAX cannot be).

Tolerances: codes and scales are bit-exact (same fp32 arithmetic, RNE with saturation); the GEMM sums exact 8-bit
products (fp32 for e4m3, int32 for int8) in another order: rel-L2 <= 1e-3 on the bf16 output (int8: the integer sum is
exact, only the last rounding differs); a whole block differs through activations that land on the other side of a
rounding boundary after a 1e-3 perturbation: bounded at 2e-2 like the static FP8 block."""
import math

import pytest
import torch
import torch.nn.functional as F

from inferix_b200 import ops
from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest
from inferix_b200.synthetic import TINY, synth_state_dict
from inferix_b200.wan_model import CausalWanModel
from oracle import wan_oracle as wo

pytestmark = pytest.mark.gpu
DEV = "cuda"
KINDS = [("fp8", ops.Q8_E4M3), ("int8", ops.Q8_INT8)]


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).bfloat16()


def codes_as_float(t):
    return t.float() if t.dtype == torch.int8 else t.view(torch.float8_e4m3fn).float()


@pytest.mark.parametrize("name,kind", KINDS)
@pytest.mark.parametrize("rows,cols", [(300, 1536), (64, 8960), (37, 256)])
def test_quantize_rows_bit_exact(name, kind, rows, cols):
    x = bf(rows, cols, scale=3.0, seed=1)
    x[0, :8] = torch.tensor([1e4, -1e4, 448.0, -448.0, 0.0, 1e-6, 464.0, -0.017]).bfloat16()
    x[1] = 0                                                        # an all-zero token
    ref_codes, ref_s = wo.dynamic_q8_quantize(x, name)
    codes, s = ops.quantize_rows(x.to(DEV), kind)
    assert torch.equal(s.cpu(), ref_s.view(-1))
    assert torch.equal(codes_as_float(codes.cpu()), ref_codes)


@pytest.mark.parametrize("name,kind", KINDS)
def test_ln_modulate_quant_equals_quantised_ln_modulate(name, kind):
    rows, cols, fs = 192, 1536, 64
    x = bf(rows, cols, seed=2).to(DEV)
    sh, sc = bf(3, cols, scale=0.2, seed=3).to(DEV), bf(3, cols, scale=0.2, seed=4).to(DEV)
    plain = ops.ln_modulate(x, shift=sh, scale=sc, tokens_per_frame=fs)
    codes, s = ops.ln_modulate_quant(x, kind, shift=sh, scale=sc, tokens_per_frame=fs)
    codes2, s2 = ops.quantize_rows(plain, kind)
    assert torch.equal(s, s2) and torch.equal(codes.view(torch.uint8), codes2.view(torch.uint8))
    w, b = bf(cols, seed=5).to(DEV), bf(cols, scale=0.1, seed=6).to(DEV)
    codes, s = ops.ln_modulate_quant(x, kind, weight=w, bias=b)
    codes2, s2 = ops.quantize_rows(ops.ln_modulate(x, weight=w, bias=b), kind)
    assert torch.equal(s, s2) and torch.equal(codes.view(torch.uint8), codes2.view(torch.uint8))


@pytest.mark.parametrize("name,kind", KINDS)
@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (300, 768, 256), (1350, 1536, 1536), (257, 64, 8960)])
def test_gemm_q8_vs_oracle(name, kind, M, N, K):
    x, w, b = bf(M, K, seed=5), bf(N, K, scale=1 / math.sqrt(K), seed=6), bf(N, seed=7)
    w_codes, w_s = ops.quantize_weight_per_channel(w.to(DEV), kind)
    ref = wo.dynamic_q8_linear(x, codes_as_float(w_codes.cpu()), w_s.cpu(), name, b)
    a_codes, a_s = ops.quantize_rows(x.to(DEV), kind)
    out = ops.gemm_q8(a_codes, w_codes, a_s, w_s, kind, b.to(DEV))
    err = rel_l2(out, ref)
    print(f"gemm_q8 {name} M={M} N={N} K={K}: rel-L2 vs oracle {err:.2e}; vs bf16 linear {rel_l2(out, F.linear(x, w, b)):.2e}")
    assert err <= 1e-3
    assert rel_l2(out, F.linear(x, w, b)) <= (6e-2 if name == "fp8" else 3e-2)   # 8-bit accuracy of the product


@pytest.mark.parametrize("name,kind", KINDS)
def test_q8_block_matches_oracle(name, kind):
    cfg = wo.WanConfig(**TINY, local_attn_size=6, sink_size=0)
    sd = {k: v.bfloat16() for k, v in synth_state_dict(TINY, seed=0).items()}
    model = CausalWanModel(**TINY, local_attn_size=6, sink_size=0)
    model.load_state_dict(synth_state_dict(TINY, seed=0))
    model = model.to(torch.bfloat16).to(DEV)
    g = torch.Generator().manual_seed(3)
    frames, fs, C = 3, 64, TINY["dim"]
    x = torch.randn(1, frames * fs, C, generator=g).bfloat16()
    e0 = (torch.randn(1, frames, 6, C, generator=g) * 0.3).bfloat16()
    ctx = (torch.randn(1, 512, C, generator=g) * 0.5).bfloat16()
    grid = (frames, 8, 8)
    table = ops.rope_table(model.freqs, DEV)
    blk = model.blocks[1]

    def run(start_list):
        mgr, req = KVCacheManager(DEV), KVCacheRequest("r")
        blk.kv_cache_manager.allocate_kv_cache(mgr, req, 6 * fs, torch.bfloat16, page_tokens=fs)
        blk.kv_cache_manager.allocate_crossattn_cache(mgr, req, 512, torch.bfloat16)
        meta = {"global_end_index": torch.zeros(1, dtype=torch.long, device=DEV),
                "local_end_index": torch.zeros(1, dtype=torch.long, device=DEV)}
        cmeta = {"is_init": False}
        return [blk(x.clone().to(DEV), e0.to(DEV), None, torch.tensor([grid]), table, ctx.to(DEV), None, None, meta, cmeta,
                    current_start=start, kv_cache_manager=mgr, kv_cache_requests=[req]) for start in start_list]

    bf16_out = run([0])[0]
    model.quantize_dynamic(name)
    sd_q = dict(sd)
    sd_q["__q8__"] = {k: (codes_as_float(c.cpu()), s.cpu(), kn) for k, (c, s, kn) in blk.q8_state().items()}
    caches, cross = wo.new_cache(cfg, 6 * fs, 1, torch.bfloat16), [dict(is_init=False) for _ in range(2)]
    starts = [0, 0, 3 * fs, 6 * fs]
    outs = run(starts)
    for step, start in enumerate(starts):
        ref = wo.block_forward(sd_q, 1, cfg, x, e0, grid, wo.rope_freqs(128), ctx, caches[1], cross[1], start)
        err = rel_l2(outs[step], ref)
        print(f"{name} dynamic block step {step}: rel-L2 vs oracle {err:.3e}")
        assert err <= 2e-2
    print(f"{name} dynamic vs bf16 block output: {rel_l2(outs[0], bf16_out):.3e}")
    assert rel_l2(outs[0], bf16_out) <= 0.1
    model.disable_fp8()
