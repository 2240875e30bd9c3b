"""Developer probe for the GPU box: runs one named check, prints error statistics and timings.

Usage (under gpurun):  timeout 120 python tools/gpu_probe.py <check> [args]
Not a test and not part of the product; tests/ holds the parity tests proper.
"""
import math
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from inferix_b200 import ops  # noqa: E402
from inferix_b200._lib import RopeGrid  # noqa: E402

dev = "cuda"


def stats(name, got, ref):
    got, ref = got.float(), ref.float()
    diff = (got - ref).abs()
    rel = (got - ref).norm() / ref.norm().clamp_min(1e-30)
    print(f"  {name}: relL2={rel.item():.3e} max|d|={diff.max().item():.3e} ref|max|={ref.abs().max().item():.3e} "
          f"nan={torch.isnan(got).sum().item()} mismatch>{1e-2}: {(diff > 1e-2 * ref.abs().max()).float().mean().item():.4f}",
          flush=True)
    return rel.item()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def check_gemm():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 512), (300, 768, 256), (1000, 520, 1536)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        ref = (a.float() @ w.float().T + b.float()).bfloat16()
        out = ops.gemm(a, w, b)
        torch.cuda.synchronize()
        print(f"gemm bias M={M} N={N} K={K}")
        stats("out", out, ref)
    M, N, K, fs = 384, 512, 256, 128
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    res = torch.randn(M, N, device=dev).bfloat16()
    gate = torch.randn(M // fs, N, device=dev).bfloat16()
    t = (a.float() @ w.float().T + b.float()).bfloat16()
    ref = (res + (t.unflatten(0, (M // fs, fs)) * gate[:, None]).flatten(0, 1))
    out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GATE_RES, residual=res, gate=gate, tokens_per_frame=fs)
    print("gemm gate_res")
    stats("out", out, ref)
    ref = torch.nn.functional.gelu(t, approximate="tanh")
    out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GELU)
    print("gemm gelu")
    stats("out", out, ref)


def check_attn():
    torch.manual_seed(0)
    for (Lq, Lk, H) in [(128, 128, 1), (256, 128, 1), (256, 256, 2), (192, 384, 2), (300, 1000, 2), (600, 5000, 3)]:
        D = 128
        q = torch.randn(Lq, H * D, device=dev).bfloat16()
        k = torch.randn(Lk, H * D, device=dev).bfloat16()
        v = torch.randn(Lk, H * D, device=dev).bfloat16()
        ref = torch.nn.functional.scaled_dot_product_attention(
            q.view(Lq, H, D).transpose(0, 1).float(), k.view(Lk, H, D).transpose(0, 1).float(),
            v.view(Lk, H, D).transpose(0, 1).float()).transpose(0, 1).reshape(Lq, H * D)
        out = ops.attention(q, k, v, H)
        torch.cuda.synchronize()
        print(f"attn Lq={Lq} Lk={Lk} H={H}")
        stats("out", out, ref)


def check_elementwise():
    from oracle import wan_oracle as wo
    torch.manual_seed(0)
    rows, C, fs = 192, 256, 64
    x = torch.randn(rows, C).bfloat16()
    sh = (torch.randn(3, C) * 0.1).bfloat16()
    sc = (torch.randn(3, C) * 0.1).bfloat16()
    ref = (wo.layer_norm(x[None], 1e-6).unflatten(1, (3, fs)) * (1 + sc[None, :, None]) + sh[None, :, None]).flatten(1, 2)[0]
    out = ops.ln_modulate(x.to(dev), shift=sh.to(dev), scale=sc.to(dev), tokens_per_frame=fs)
    print("ln_modulate")
    stats("out", out.cpu(), ref)
    print("   exact-equal fraction:", (out.cpu() == ref).float().mean().item())
    w = (1 + 0.1 * torch.randn(C)).bfloat16()
    b = (0.1 * torch.randn(C)).bfloat16()
    ref = wo.layer_norm(x, 1e-6, w, b)
    out = ops.ln_modulate(x.to(dev), weight=w.to(dev), bias=b.to(dev))
    print("ln affine")
    stats("out", out.cpu(), ref)
    print("   exact-equal fraction:", (out.cpu() == ref).float().mean().item())
    ref = wo.rms_norm(x, w, 1e-6)
    out = ops.rmsnorm(x.to(dev), w.to(dev))
    print("rmsnorm")
    stats("out", out.cpu(), ref)
    print("   exact-equal fraction:", (out.cpu() == ref).float().mean().item())

    # qk norm + rope + append
    H, D = 2, 128
    grid = (3, 8, 8)
    qkv = torch.randn(rows, 3 * C).bfloat16()
    wq = (1 + 0.1 * torch.randn(C)).bfloat16()
    wk = (1 + 0.1 * torch.randn(C)).bfloat16()
    freqs = wo.rope_freqs(D)
    q = wo.rms_norm(qkv[None, :, :C], wq, 1e-6).view(1, rows, H, D)
    k = wo.rms_norm(qkv[None, :, C:2 * C], wk, 1e-6).view(1, rows, H, D)
    q_ref = wo.causal_rope_apply(q, grid, freqs, start_frame=3)[0].reshape(rows, C)
    k_ref = wo.causal_rope_apply(k, grid, freqs, start_frame=3)[0].reshape(rows, C)
    table = ops.rope_table(freqs, dev)
    kv = ops.PagedKV(6, fs, H, D, dev)
    plan0 = kv.plan_append(0, rows, 0, True)
    kv.append(plan0, torch.zeros(rows, C, device=dev).bfloat16(), torch.zeros(rows, C, device=dev).bfloat16())
    plan = kv.plan_append(rows, rows, 0, True)
    g = RopeGrid(3, 8, 8, 3, 0, 64)
    q_out, _, _ = ops.qk_norm_rope_append(qkv.to(dev), wq.to(dev), wk.to(dev), table, g, H, D, kv=kv, plan=plan)
    ke, ve = kv.export(rows, rows)
    print("qk_norm_rope_append plan pages", list(plan.pages[:plan.num_pages]), plan.local_start, plan.local_end)
    stats("q", q_out.cpu(), q_ref)
    print("   exact-equal fraction:", (q_out.cpu() == q_ref).float().mean().item())
    stats("k(cache)", ke.cpu(), k_ref)
    print("   exact-equal fraction:", (ke.cpu() == k_ref).float().mean().item())
    print("   v exact:", torch.equal(ve.cpu(), qkv[:, 2 * C:]))


def check_perf():
    torch.manual_seed(0)
    S, C, F, H = 10800, 1536, 8960, 12
    a = torch.randn(S, C, device=dev).bfloat16()
    for (N, K, name) in [(3 * C, C, "qkv"), (C, C, "o"), (F, C, "ffn1"), (C, F, "ffn2")]:
        x = torch.randn(S, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        out = torch.empty(S, N, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.gemm(x, w, b, out=out))
        ms_t = timeit(lambda: torch.nn.functional.linear(x, w, b))
        fl = 2.0 * S * N * K
        print(f"gemm {name}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s   (torch/cuBLAS {ms_t:.3f} ms {fl / ms_t / 1e9:.1f})",
              flush=True)
    for Lk in (10800, 43200, 86400):
        q = torch.randn(S, C, device=dev).bfloat16()
        k = torch.randn(Lk, C, device=dev).bfloat16()
        v = torch.randn(Lk, C, device=dev).bfloat16()
        out = torch.empty(S, C, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.attention(q, k, v, H, out=out), iters=5, warm=2)
        fl = 4.0 * S * Lk * C
        print(f"attn Lk={Lk}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        try:
            from flash_attn import flash_attn_func
            q4, k4, v4 = q.view(1, S, H, 128), k.view(1, Lk, H, 128), v.view(1, Lk, H, 128)
            ms2 = timeit(lambda: flash_attn_func(q4, k4, v4), iters=5, warm=2)
            print(f"   flash_attn 2 (reference's GPU kernel): {ms2:.3f} ms {fl / ms2 / 1e9:.1f} TFLOP/s", flush=True)
            ref = flash_attn_func(q4, k4, v4).view(S, C)
            stats("vs FA2", out, ref)
        except Exception as ex:  # noqa: BLE001
            print("   flash_attn unavailable:", ex)


def check_perf_sp8():
    """Per-rank shapes of the 8-way sequence-parallel run (S/P = 1350 rows)."""
    torch.manual_seed(0)
    S, C, F, H = 1350, 1536, 8960, 12
    for (N, K, name) in [(3 * C, C, "qkv"), (C, C, "o"), (F, C, "ffn1"), (C, F, "ffn2")]:
        x = torch.randn(S, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        out = torch.empty(S, N, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.gemm(x, w, b, out=out), iters=20)
        ms_t = timeit(lambda: torch.nn.functional.linear(x, w, b), iters=20)
        fl = 2.0 * S * N * K
        ref = torch.nn.functional.linear(x, w, b)
        print(f"gemm {name} M={S}: {ms * 1e3:.1f} us {fl / ms / 1e9:.1f} TFLOP/s   (cuBLAS {ms_t * 1e3:.1f} us {fl / ms_t / 1e9:.1f})",
              flush=True)
        stats("vs cuBLAS", out, ref)
    Lk = 86400
    q = torch.randn(S, C, device=dev).bfloat16()
    k = torch.randn(Lk, C, device=dev).bfloat16()
    v = torch.randn(Lk, C, device=dev).bfloat16()
    out = torch.empty(S, C, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(q, k, v, H, out=out), iters=10, warm=2)
    print(f"attn Lq={S} Lk={Lk}: {ms:.3f} ms {4.0 * S * Lk * C / ms / 1e9:.1f} TFLOP/s", flush=True)


def check_perf_fp8():
    torch.manual_seed(0)
    S, C, F = 10800, 1536, 8960
    for (N, K, name) in [(3 * C, C, "qkv"), (C, C, "o"), (F, C, "ffn1"), (C, F, "ffn2")]:
        x = torch.randn(S, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        s_in, s_w = float(x.abs().max()) / 448, float(w.abs().max()) / 448
        xq, wq = ops.quantize_fp8(x, s_in), ops.quantize_fp8(w, s_w)
        out = torch.empty(S, N, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.gemm_fp8(xq, wq, s_in * s_w, b, out=out))
        msq = timeit(lambda: ops.quantize_fp8(x, s_in, xq))
        ref = torch.nn.functional.linear(x, w, b)
        fl = 2.0 * S * N * K
        print(f"gemm_fp8 {name}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s; quantize {msq * 1e3:.1f} us "
              f"({S * K * 3 / msq / 1e6:.0f} GB/s)", flush=True)
        stats("vs bf16 linear", out, ref)


def check_attn_once():
    """Self-attention at the BASELINE config-2 shape, a few launches (target of `ncu --set full -k regex:attn_fwd`)."""
    S, C, H, Lk = 10800, 1536, 12, 86400
    q = torch.randn(S, C, device=dev).bfloat16()
    k = torch.randn(Lk, C, device=dev).bfloat16()
    v = torch.randn(Lk, C, device=dev).bfloat16()
    out = torch.empty(S, C, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.attention(q, k, v, H, out=out)


def check_gemm_once():
    """FFN1 GEMM (+GELU) at the config-2 shape (target of `ncu --set full -k regex:gemm_bf16`)."""
    S, C, F_ = 10800, 1536, 8960
    x = torch.randn(S, C, device=dev).bfloat16()
    w = (torch.randn(F_, C, device=dev) / math.sqrt(C)).bfloat16()
    b = torch.randn(F_, device=dev).bfloat16()
    out = torch.empty(S, F_, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.gemm(x, w, b, out=out, epilogue=ops.EPI_BIAS_GELU)


def check_gemm_fp8_once():
    """FFN1 GEMM with e4m3 operands at the config-2 shape (target of `ncu -k regex:gemm2_tn`)."""
    S, C, F_ = 10800, 1536, 8960
    x = torch.randn(S, C, device=dev).bfloat16()
    w = (torch.randn(F_, C, device=dev) / math.sqrt(C)).bfloat16()
    b = torch.randn(F_, device=dev).bfloat16()
    wq, ws = ops.quantize_weight_per_channel(w, ops.Q8_E4M3)
    out = torch.empty(S, F_, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        aq, a_s = ops.quantize_rows(x, ops.Q8_E4M3)
        ops.gemm_q8(aq, wq, a_s, ws, ops.Q8_E4M3, b, out, epilogue=ops.EPI_BIAS_GELU)


def check_gemm_small_once():
    """FFN2 GEMM at the 8-way sequence-parallel shard shape (M = 1350, K = 8960): the 128-wide cluster tile."""
    S, C, F_ = 1350, 1536, 8960
    x = torch.randn(S, F_, device=dev).bfloat16()
    w = (torch.randn(C, F_, device=dev) / math.sqrt(F_)).bfloat16()
    b = torch.randn(C, device=dev).bfloat16()
    res = torch.randn(S, C, device=dev).bfloat16()
    gate = torch.randn(3, C, device=dev).bfloat16()
    for _ in range(4):
        ops.gemm(x, w, b, res, epilogue=ops.EPI_BIAS_GATE_RES, residual=res, gate=gate, tokens_per_frame=S // 3)


def check_rows_once():
    """The HBM-bound row kernels at the config-2 shape: LN + modulate, and QK-RMSNorm + fp64 RoPE + paged append."""
    from inferix_b200._lib import RopeGrid
    S, C, H = 10800, 1536, 12
    x = torch.randn(S, C, device=dev).bfloat16()
    mod = (torch.randn(3, 6, C, device=dev) * 0.3).bfloat16()
    qkv = torch.randn(S, 3 * C, device=dev).bfloat16()
    nq, nk = torch.rand(C, device=dev).bfloat16() + 0.5, torch.rand(C, device=dev).bfloat16() + 0.5
    freqs = torch.view_as_real(torch.polar(torch.ones(1024, 64, dtype=torch.float64),
                                           torch.randn(1024, 64, dtype=torch.float64))).contiguous().to(dev)
    store = ops.PagedKV(24, 3600, H, 128, dev)
    h = torch.empty_like(x)
    q = torch.empty_like(x)
    for i in range(4):
        ops.ln_modulate(x, h, shift=mod[:, 0], scale=mod[:, 1], tokens_per_frame=3600)
        plan = store.plan_append(i * S, S, 0, True)
        ops.qk_norm_rope_append(qkv, nq, nk, freqs, RopeGrid(3, 45, 80, 3 * i, 0, 3600), H, 128, kv=store, plan=plan, q_out=q)


if __name__ == "__main__":
    name = sys.argv[1]
    t0 = time.time()
    globals()["check_" + name]()
    torch.cuda.synchronize()
    print(f"[{name}] done in {time.time() - t0:.1f}s", flush=True)
