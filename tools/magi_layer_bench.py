"""Per-kernel timing of ONE native MAGI-1 transformer layer at real widths on one B200 (BASELINE config 4 shapes,
reduced to one layer and one GPU; synthetic weights).  Not the headline bench: it reports where a MAGI layer's time
goes and the achieved TFLOP/s / GB/s of its kernels.

    python tools/magi_layer_bench.py [--model 4.5b|24b] [--clip-tokens 21600] [--ranges 4] [--history 4] [--reps 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/magi_layer_bench.py --model 24b
Under torchrun the layer runs with Ulysses context parallel over all ranks (BASELINE config 4: 8 KV groups <-> 8 GPUs);
rank 0 prints one JSON line (times = max over ranks).  The multi-rank mode was written after the round's GPU budget was
spent: verified at tiny widths on 2 GPUs through tools/magi_cp_check.py, not yet run at these sizes.
"""
import argparse
import json
import os
import sys
import types
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from inferix_b200 import _lib, magi_cp, magi_layer  # noqa: E402
from inferix_b200.kvcache_manager.model import InferenceParams  # noqa: E402

MODELS = {"4.5b": dict(hidden_size=3072, ffn_hidden_size=12288, num_attention_heads=24, num_query_groups=8,
                       gated_linear_unit=False),
          "24b": dict(hidden_size=6144, ffn_hidden_size=16384, num_attention_heads=48, num_query_groups=8,
                      gated_linear_unit=True)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="4.5b", choices=sorted(MODELS))
    ap.add_argument("--clip-tokens", type=int, default=21600)     # 720p: 6 latent frames x 45 x 80
    ap.add_argument("--ranges", type=int, default=4)
    ap.add_argument("--history", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--caption", type=int, default=120)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        magi_cp.init_context_parallel(None, world, rank)
    torch.set_grad_enabled(False)
    m = MODELS[a.model]
    mc = types.SimpleNamespace(layernorm_epsilon=1e-6, apply_layernorm_1p=False, cond_hidden_ratio=0.25,
                               cond_gating_ratio=1.0, xattn_cond_hidden_ratio=1.0, params_dtype=torch.bfloat16,
                               kv_channels=128, num_layers=1, **m)
    ec = types.SimpleNamespace(cp_size=world, cp_strategy="cp_ulysses" if world > 1 else "none", fp8_quant=False,
                               kv_offload=False)
    layer = magi_layer.TransformerLayer(mc, ec, 0)
    g = torch.Generator().manual_seed(0)
    for name, p in layer.named_parameters():
        if p.dim() == 2:
            p.data.copy_((torch.randn(p.shape, generator=g) * p.shape[1] ** -0.5).to(p.dtype))
        elif name.endswith("weight"):
            p.data.fill_(1.0)
    layer = layer.to(dev)
    clip, r, hist = a.clip_tokens, a.ranges, a.history
    s, h = clip * r, m["hidden_size"]
    ip = InferenceParams(1, (hist + r) * clip, device=dev)
    hidden = torch.randn(s, 1, h, generator=g).bfloat16().to(dev)
    cond = torch.randn(1, r, h // 4, generator=g).bfloat16().to(dev)
    cmap = torch.arange(r).repeat_interleave(clip).reshape(-1, 1).to(dev)
    y = torch.randn(a.caption * r, h, generator=g).bfloat16().to(dev)
    ang = torch.randn(s, 48, generator=g) * 2
    rope = torch.cat([ang.sin(), ang.cos()], -1).to(dev)
    cu_q = [i * clip for i in range(r + 1)]
    cu_k = [i * a.caption for i in range(r + 1)]
    # every denoising chunk sees the whole clean history and the chunks up to itself
    k_range = [[0, (hist + i + 1) * clip] for i in range(r)]
    core = types.SimpleNamespace(np_q_range=[[cu_q[i], cu_q[i + 1]] for i in range(r)], np_k_range=k_range)
    cross = types.SimpleNamespace(q_ranges=[[cu_q[i], cu_q[i + 1]] for i in range(r)],
                                  kv_ranges=[[cu_k[i], cu_k[i + 1]] for i in range(r)])
    meta = types.SimpleNamespace(slice_point=hist, denoising_range_num=r, clip_token_nums=clip,
                                 extract_prefix_video_feature=False, fwd_extra_1st_chunk=False,
                                 distill_nearly_clean_chunk=False, cp_split_sizes=None, core_attn_params=core,
                                 cross_attn_params=cross)
    if world > 1:                                   # this rank's contiguous shard of the sequence
        hidden, cmap, rope, split, (xq, xk) = magi_cp.cp_ulysses_process(world, hidden, cmap, rope, cu_q, cu_k)
        meta.cp_split_sizes = split
        meta.cross_attn_params = types.SimpleNamespace(q_ranges=xq, kv_ranges=xk)
        hidden = hidden.contiguous()
    # fill the history rows with something finite
    store = layer.self_attention.kv_cache_manager.native_store(ip)
    kk, vv = store.map_rows(hist * clip)
    kk.normal_()
    vv.normal_()
    for _ in range(2):
        layer(hidden, cond, cmap, y, rope, ip, meta)
    torch.cuda.synchronize()
    _lib.prof_reset()
    _lib.prof_enable(True)
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        layer(hidden, cond, cmap, y, rope, ip, meta)
    e1.record()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    ms = e0.elapsed_time(e1) / a.reps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        if rank != 0:
            dist.barrier()
            dist.destroy_process_group()
            return
    hq, gk, d, f = m["num_attention_heads"], m["num_query_groups"], 128, m["ffn_hidden_size"]
    flops_attn = sum(4.0 * clip * (ke - ks) * hq * d for ks, ke in k_range) + 4.0 * s * a.caption * hq * d
    flops_gemm = 2.0 * s * h * ((2 * hq + 2 * gk) * d) + 2.0 * s * (2 * hq * d) * h \
        + 2.0 * s * h * f * (3 if m["gated_linear_unit"] else 2)
    per = {}
    for label in _lib.prof_labels():
        t, n = _lib.prof_read(label)
        key = label.split("[")[0]
        per.setdefault(key, [0.0, 0])
        per[key][0] += t / a.reps
        per[key][1] += n // a.reps
    attn_ms = sum(v[0] for k, v in per.items() if k.startswith("attn"))
    gemm_ms = sum(v[0] for k, v in per.items() if k.startswith("gemm"))
    row_bytes = {  # algorithmic bytes of the row kernels (read + write), per launch
        "magi_qkv_post_kernel": 2.0 * s * (2 * hq + 2 * gk) * d * 2,
        "gate_norm_residual_kernel": 3.0 * s * h * 2,
        "ln_modulate_kernel": 2.0 * s * h * 2,
    }
    out = {"model": a.model, "cp_ulysses": world, "tokens": s, "ranges": r, "history_tokens": hist * clip, "layer_ms": ms,
           "launches_per_layer": _lib.launch_count() // a.reps,
           "attention": {"ms": attn_ms, "tflops": flops_attn / attn_ms / 1e9 if attn_ms else None},
           "gemm": {"ms": gemm_ms, "tflops": flops_gemm / gemm_ms / 1e9 if gemm_ms else None},
           "kernels_ms": {k: round(v[0], 4) for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])},
           "row_kernel_gbs": {k: round(b * per[k][1] / (per[k][0] * 1e6), 1) for k, b in row_bytes.items() if k in per and per[k][0] > 0}}
    if world > 1:                                   # per-rank kernels see 1/world of the heads or of the sequence
        out["note"] = "attention / GEMM TFLOP/s are whole-layer FLOPs over rank 0's kernel time x world"
        for k in ("attention", "gemm"):
            if out[k]["tflops"]:
                out[k]["tflops"] = out[k]["tflops"] / world
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
