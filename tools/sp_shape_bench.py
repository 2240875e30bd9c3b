#!/usr/bin/env python
"""Per-kernel device times of ONE DiT layer at the per-rank shapes of the sequence-parallel run, on a single GPU:
M = 10800 / 5400 / 2700 / 1350 query rows (world 1 / 2 / 4 / 8) against the full 86 400-token window.  The compute a
rank does under SP is exactly these launches (only the K/V exchange needs the other GPUs), so this is where the
small-M losses are found without paying for a multi-GPU box.  Prints one JSON line per M.

    python tools/sp_shape_bench.py [--iters 10] [--rows 10800,5400,2700,1350]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from inferix_b200 import _lib, ops  # noqa: E402
from inferix_b200._lib import RopeGrid  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--rows", default="10800,5400,2700,1350")
    ap.add_argument("--window", type=int, default=86400)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    C, F, H, D, L = 1536, 8960, 12, 128, a.window
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, s=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * s).bfloat16()
    w_qkv, b_qkv = rnd(3 * C, C, s=C ** -0.5), rnd(3 * C, s=0.1)
    w_o, b_o = rnd(C, C, s=C ** -0.5), rnd(C, s=0.1)
    w_1, b_1 = rnd(F, C, s=C ** -0.5), rnd(F, s=0.1)
    w_2, b_2 = rnd(C, F, s=F ** -0.5), rnd(C, s=0.1)
    nq, nk = rnd(C).abs() + 0.5, rnd(C).abs() + 0.5
    k_cache, v_cache = rnd(L, C), rnd(L, C)
    k_txt, v_txt = rnd(512, C), rnd(512, C)
    freqs = torch.view_as_real(torch.polar(torch.ones(1024, 64, dtype=torch.float64),
                                           torch.randn(1024, 64, dtype=torch.float64))).contiguous().to(dev)
    for M in [int(r) for r in a.rows.split(",")]:
        fs = M // 3
        world = 10800 // M
        x, mod = rnd(M, C), rnd(3, 6, C, s=0.3)
        h, qkv, q, att, ffn = (torch.empty(M, n, dtype=torch.bfloat16, device=dev) for n in (C, 3 * C, C, C, F))
        kn, vn = torch.empty(M, C, dtype=torch.bfloat16, device=dev), torch.empty(M, C, dtype=torch.bfloat16, device=dev)
        grid = RopeGrid(3, 45, 80, 21, 0, fs)

        def layer():
            ops.ln_modulate(x, h, shift=mod[:, 0], scale=mod[:, 1], tokens_per_frame=fs)
            ops.gemm(h, w_qkv, b_qkv, qkv)
            ops.qk_norm_rope_append(qkv, nq, nk, freqs, grid, H, D, q_out=q, k_out=kn, v_out=vn)
            ops.attention(q, k_cache, v_cache, H, att)
            ops.gemm(att, w_o, b_o, x, epilogue=ops.EPI_BIAS_GATE_RES, residual=x, gate=mod[:, 2], tokens_per_frame=fs)
            ops.ln_modulate(x, h, weight=nq, bias=nk)
            ops.gemm(h, w_o, b_o, qkv[:, :C])
            ops.rmsnorm(qkv[:, :C], nq, q)
            ops.attention(q, k_txt, v_txt, H, att)
            ops.gemm(att, w_o, b_o, x, epilogue=ops.EPI_BIAS_GATE_RES, residual=x)
            ops.ln_modulate(x, h, shift=mod[:, 3], scale=mod[:, 4], tokens_per_frame=fs)
            ops.gemm(h, w_1, b_1, ffn, epilogue=ops.EPI_BIAS_GELU)
            ops.gemm(ffn, w_2, b_2, x, epilogue=ops.EPI_BIAS_GATE_RES, residual=x, gate=mod[:, 5], tokens_per_frame=fs)

        for _ in range(3):
            layer()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            layer()
        e1.record()
        torch.cuda.synchronize()
        layer_ms = e0.elapsed_time(e1) / a.iters
        _lib.prof_reset()
        _lib.prof_enable(True)
        for _ in range(a.iters):
            layer()
        torch.cuda.synchronize()
        _lib.prof_enable(False)
        per = {}
        for label in _lib.prof_labels():
            ms, n = _lib.prof_read(label)
            per[label] = round(1e3 * ms / n, 1)
        _lib.prof_reset()
        flops_attn = 4.0 * M * L * C
        attn_us = next(v for k, v in per.items() if k.startswith(f"attn_fwd_kernel[Lq={M},Lk={L}"))
        gemm_flops = 2.0 * M * (3 * C * C + 3 * C * C + 2 * C * F)
        gemm_us = sum(v for k, v in per.items() if k.startswith("gemm_"))
        print(json.dumps({"M": M, "world": world, "layer_us": round(1e3 * layer_ms, 1),
                          "ideal_layer_us_from_M10800": None,
                          "attn_self_us": attn_us, "attn_tflops": round(flops_attn / attn_us / 1e6, 1),
                          "gemm_us": round(gemm_us, 1), "gemm_tflops": round(gemm_flops / gemm_us / 1e6, 1),
                          "kernels_us": per}), flush=True)


if __name__ == "__main__":
    main()
