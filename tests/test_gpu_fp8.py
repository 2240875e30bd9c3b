"""FP8 (e4m3, per-tensor) linears: kernels against the oracle's restatement of MAGI's PerTensorQuantizedFp8Linear /
div_clamp_to (dit_module.py:367-387,434-459) — the only in-tree specification of the FP8 arithmetic (DAX, which the
Wan quantisation examples call, is not vendored: "DAX parity unpinned").

Tolerances: quantisation is bit-exact (same clamp -> bf16 -> e4m3 rounding chain); the FP8 GEMM sums exact fp8 x fp8
products in fp32 in a different order: rel-L2 <= 1e-3 on the bf16 output; a whole quantised block differs from the
oracle only through activations that land on the other side of an e4m3 rounding boundary (3 mantissa bits: a 1e-3
input perturbation flips ~2 % of the codes by 6 %): measured ~5e-3, bounded at 2e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from inferix_b200 import ops                                                 # noqa: E402
from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest      # noqa: E402
from inferix_b200.synthetic import TINY, synth_state_dict                    # noqa: E402
from inferix_b200.wan_model import CausalWanModel                            # noqa: E402
from oracle import wan_oracle as wo                                          # noqa: E402

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).bfloat16()


def test_quantize_bit_exact():
    x = bf(300, 1536, scale=3.0, seed=1)
    x[0, :8] = torch.tensor([1e4, -1e4, 448.0, -448.0, 0.0, 1e-6, 464.0, -0.017]).bfloat16()   # saturation / tiny values
    for scale in (0.05, 1.0, 0.0137):
        ref = wo.div_clamp_to_e4m3(x, scale)
        out = ops.quantize_fp8(x.to(DEV), scale)
        assert torch.equal(out.cpu().view(torch.uint8), ref.view(torch.uint8)), f"scale {scale}"


def test_ln_modulate_fp8_equals_quantised_ln_modulate():
    rows, cols, fs = 192, 256, 64
    x = bf(rows, cols, seed=2).to(DEV)
    sh, sc = bf(3, cols, scale=0.2, seed=3).to(DEV), bf(3, cols, scale=0.2, seed=4).to(DEV)
    plain = ops.ln_modulate(x, shift=sh, scale=sc, tokens_per_frame=fs)
    fused = ops.ln_modulate_fp8(x, 0.02, shift=sh, scale=sc, tokens_per_frame=fs)
    assert torch.equal(fused.view(torch.uint8), ops.quantize_fp8(plain, 0.02).view(torch.uint8))


@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (300, 768, 256), (1000, 520, 1536), (257, 64, 8960)])
def test_gemm_fp8_vs_oracle(M, N, K):
    x, w, b = bf(M, K, seed=5), bf(N, K, scale=1 / math.sqrt(K), seed=6), bf(N, seed=7)
    s_in, s_w = float(x.abs().max()) / 448, float(w.abs().max()) / 448
    w_q = wo.div_clamp_to_e4m3(w, s_w)
    ref = wo.fp8_linear(x, w_q, s_w, s_in, b)
    out = ops.gemm_fp8(ops.quantize_fp8(x.to(DEV), s_in), w_q.to(DEV), s_in * s_w, b.to(DEV))
    assert rel_l2(out, ref) <= 1e-3
    # and the quantised product tracks the un-quantised one at FP8 accuracy
    assert rel_l2(out, F.linear(x, w, b)) <= 6e-2


def test_fp8_block_matches_oracle():
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

    def run(start_list, mgr, req):
        meta = {"global_end_index": torch.zeros(1, dtype=torch.long, device=DEV),
                "local_end_index": torch.zeros(1, dtype=torch.long, device=DEV)}
        cmeta = {"is_init": False}
        outs = []
        for start in start_list:
            outs.append(blk(x.clone().to(DEV), e0.to(DEV), None, torch.tensor([grid]), table, ctx.to(DEV), None, None,
                            meta, cmeta, current_start=start, kv_cache_manager=mgr, kv_cache_requests=[req]))
        return outs

    def fresh():
        mgr, req = KVCacheManager(DEV), KVCacheRequest("r")
        blk.kv_cache_manager.allocate_kv_cache(mgr, req, 6 * fs, torch.bfloat16, page_tokens=fs)
        blk.kv_cache_manager.allocate_crossattn_cache(mgr, req, 512, torch.bfloat16)
        return mgr, req

    model.begin_fp8_calibration()
    bf16_out = run([0], *fresh())[0]
    for other in model.blocks:                     # only block 1 ran: give the others its statistics
        other._amax = dict(blk._amax)
    model.finish_fp8_calibration()
    assert blk._fp8 is not None and set(blk._fp8) == set(blk.FP8_SITES)

    sd_q = dict(sd)
    sd_q["__fp8__"] = {k: (w.cpu(), ws, si) for k, (w, ws, si) in blk.fp8_state().items()}
    caches, cross = wo.new_cache(cfg, 6 * fs, 1, torch.bfloat16), [dict(is_init=False) for _ in range(2)]
    starts = [0, 0, 3 * fs, 6 * fs]
    outs = run(starts, *fresh())
    for step, start in enumerate(starts):
        ref = wo.block_forward(sd_q, 1, cfg, x, e0, grid, wo.rope_freqs(128), ctx, caches[1], cross[1], start)
        err = rel_l2(outs[step], ref)
        print(f"fp8 block step {step}: rel-L2 vs fp8 oracle {err:.3e}")
        assert err <= 2e-2
    print(f"fp8 vs bf16 block output: {rel_l2(outs[0], bf16_out):.3e}")
    assert rel_l2(outs[0], bf16_out) <= 0.1
    model.disable_fp8()
