#!/usr/bin/env python
"""A/B of attention-kernel build variants (inferix_b200/lib/variants/*.so, `make -C inferix_b200/csrc variant ...`).

For every library: parity of the full-shape launch against fp32 softmax on sampled rows, then a SUSTAINED timing —
the kernel back to back for ~3 s, so that the GPU settles at its power-capped clock like inside the denoising step
(a burst figure flatters every variant: 4.97 ms alone vs 5.3-5.5 ms in-step) — and the 1350-row shard shape.
Each library runs in its own process under a timeout (a mis-synchronised variant hangs instead of failing).

    python tools/attn_variants.py [name ...]        # default: the shipped library + every variant found
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def child():
    import torch
    sys.path.insert(0, str(ROOT))
    from inferix_b200 import ops
    dev = torch.device("cuda", 0)
    S, C, H, L, D = 10800, 1536, 12, 86400, 128
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(S, C, device=dev, generator=g).bfloat16()
    k = torch.randn(L, C, device=dev, generator=g).bfloat16()
    v = torch.randn(L, C, device=dev, generator=g).bfloat16()
    out = torch.empty_like(q)
    ops.attention(q, k, v, H, out=out)
    rows = torch.cat([torch.arange(0, 48), torch.arange(5000, 5048), torch.arange(S - 100, S)]).to(dev)
    qs = q[rows].float().view(-1, H, D).transpose(0, 1)
    ref = (torch.softmax(qs @ k.float().view(L, H, D).transpose(0, 1).transpose(1, 2) / D ** 0.5, dim=-1)
           @ v.float().view(L, H, D).transpose(0, 1)).transpose(0, 1).reshape(len(rows), C)
    err = ((out[rows].float() - ref).norm() / ref.norm()).item()
    o1 = torch.empty_like(q)
    ops.attention(q, k, torch.ones_like(v), H, out=o1)
    ones_err = (o1.float() - 1).abs().max().item()

    def sustained(qq, oo, seconds, window):
        evs = [torch.cuda.Event(enable_timing=True)]
        evs[0].record()
        n, t_est = 0, 0.0
        while t_est < seconds * 1e3:
            for _ in range(window):
                ops.attention(qq, k, v, H, out=oo)
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append(e)
            n += 1
            if n % 4 == 0:
                torch.cuda.synchronize()
                t_est = evs[0].elapsed_time(evs[-1])
        torch.cuda.synchronize()
        per = [evs[i].elapsed_time(evs[i + 1]) / window for i in range(len(evs) - 1)]
        tail = per[len(per) // 2:]
        return per[0], sum(tail) / len(tail), len(per) * window

    burst, sus, n = sustained(q, out, 3.0, 20)
    q8, o8 = q[:1350].contiguous(), torch.empty(1350, C, device=dev, dtype=torch.bfloat16)
    _, sus8, _ = sustained(q8, o8, 1.0, 50)
    fl = 4.0 * S * L * C
    print(json.dumps({"lib": os.environ.get("INFERIX_B200_LIB", "shipped"), "rel_l2_vs_fp32": round(err, 6),
                      "rows_sum_to_one_err": round(ones_err, 5), "burst_ms": round(burst, 4),
                      "sustained_ms": round(sus, 4), "sustained_tflops": round(fl / sus / 1e9, 1), "launches": n,
                      "shard1350_us": round(sus8 * 1e3, 1)}), flush=True)


def main():
    names = sys.argv[1:]
    libs = {"shipped": None}
    for pth in sorted((ROOT / "inferix_b200" / "lib" / "variants").glob("libinferix_b200_*.so")):
        libs[pth.stem.replace("libinferix_b200_", "")] = str(pth)
    for name, pth in libs.items():
        if names and name not in names:
            continue
        env = dict(os.environ)
        if pth:
            env["INFERIX_B200_LIB"] = pth
        try:
            r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True, timeout=150)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(json.dumps({"variant": name, **(json.loads(line[-1]) if line else {"error": (r.stderr or r.stdout)[-400:]})}),
                  flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"variant": name, "error": "timeout (hang)"}), flush=True)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
