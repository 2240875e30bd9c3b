#!/usr/bin/env python
"""BASELINE config 5 as specified: Self-Forcing 720p long horizon — 256 blocks (768 latent frames) through the 8-block
KV window with the shipped 4-step schedule, sequence-parallel over the launched ranks (4 x B200 in BASELINE.json).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/long_horizon.py

Checks, on every rank: the KV index trace of all 1280 forwards equals the oracle's arithmetic
(oracle/wan_oracle.py: plan_indices == causal_model.py:277-300), every block past the window evicts exactly one block,
allocated device memory does not grow once the window is full, per-block time stays flat (late vs early median), the
latents stay finite.  Rank 0 prints one JSON line (times = this rank's CUDA events).
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
from inferix_b200 import synthetic  # noqa: E402
from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest  # noqa: E402
from inferix_b200.parallel import ParallelConfig  # noqa: E402
from inferix_b200.pipeline import CausalInferencePipeline, DecodeMode  # noqa: E402
from inferix_b200.wan_model import CausalWanModel  # noqa: E402
from inferix_b200.wrapper import WanDiffusionWrapper  # noqa: E402
from oracle import wan_oracle as wo  # noqa: E402  (index arithmetic only: this is a checker, not a bench arm)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=256)
    ap.add_argument("--window-blocks", type=int, default=8)
    ap.add_argument("--sink-frames", type=int, default=0)
    ap.add_argument("--layers", type=int, default=30)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pc = ParallelConfig(ring_size=world, rank=rank, local_rank=local, world_size=world)
    torch.set_grad_enabled(False)
    cfg = dict(synthetic.WAN_1_3B, num_layers=a.layers)
    n, H, W = 3, 90, 160
    fs = (H // 2) * (W // 2)
    window = a.window_blocks * n
    model = CausalWanModel(**cfg, local_attn_size=window, sink_size=a.sink_frames, parallel_config=pc)
    with torch.device("cpu"):
        sd = synthetic.synth_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    model.load_state_dict(sd)
    del sd
    model = model.to(torch.bfloat16).to(dev)
    pargs = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                  num_frame_per_block=n, context_noise=0)
    pipe = CausalInferencePipeline(pargs, dev, generator=WanDiffusionWrapper(model=model, timestep_shift=5.0, parallel_config=pc),
                                   parallel_config=pc)
    g = torch.Generator(device=dev).manual_seed(3)
    noise = torch.randn(1, a.blocks * n, 16, H, W, device=dev, generator=g).bfloat16()      # 2.1 GB at 256 blocks
    context = torch.randn(1, 20, cfg["text_dim"], device=dev, generator=g).bfloat16()
    trace, mem = [], []
    hook = model.blocks[0].register_forward_hook(lambda m, i, o: trace.append(pipe.kv_cache_meta[0]["_ifx_plan"]))
    out = pipe.inference(noise=noise, text_prompts=context, kv_cache_manager=KVCacheManager(dev),
                         kv_cache_requests=[KVCacheRequest("long")], decode_mode=DecodeMode.NO_DECODE, profile=True,
                         block_callback=lambda lat, i: mem.append(torch.cuda.memory_allocated(dev)))
    hook.remove()
    torch.cuda.synchronize()
    want, ge, le = [], 0, 0
    for b in range(a.blocks):
        for _ in range(5):                                          # 4 noisy forwards + the clean re-run
            ls, le, ge, ev = wo.plan_indices(window * fs, ge, le, b * n * fs, n * fs, a.sink_frames * fs, True)
            want.append((ls, le, ge, ev))
    t = pipe.last_block_times_ms
    w = a.window_blocks
    early = sorted(t[w:w + 64])[32] if a.blocks >= w + 64 else sorted(t[w:])[len(t[w:]) // 2]
    late = sorted(t[-64:])[32] if a.blocks >= 64 else early
    res = {"workload": "self_forcing_720p_long_horizon", "world": world, "blocks": a.blocks, "forwards": len(trace),
           "window_blocks": w, "sink_frames": a.sink_frames, "layers": a.layers,
           "index_trace_equal_oracle": trace == want,
           "evicted_tokens_total": sum(x[3] for x in trace), "evicted_expected": max(0, a.blocks - w) * n * fs,
           "final_local_end": trace[-1][1], "final_global_end": trace[-1][2],
           "allocated_bytes_min_after_window": min(mem[w:]), "allocated_bytes_max_after_window": max(mem[w:]),
           "allocation_growth_bytes": max(mem[w:]) - min(mem[w:]),
           "block_ms_early_median": early, "block_ms_late_median": late, "block_ms_first_full": t[w],
           "block_ms_min": min(t[w:]), "block_ms_max": max(t[w:]), "late_over_early": late / early,
           "latent_frames_per_s_steady": n / (late / 1e3), "finite": bool(torch.isfinite(out.float()).all())}
    ok = (res["index_trace_equal_oracle"] and res["evicted_tokens_total"] == res["evicted_expected"]
          and res["allocation_growth_bytes"] == 0 and res["late_over_early"] <= 1.1 and res["finite"])
    res["ok"] = ok
    if world > 1:
        flags = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        res["ok_all_ranks"] = bool(flags.item())
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if a.out:
            Path(a.out).write_text(line + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
