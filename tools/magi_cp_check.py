"""Run under torchrun on N GPUs: the MAGI-1 block under Ulysses context parallel must reproduce the reference goldens
(tests/golden/magi_layer_glu.pt, produced by the reference's own single-rank TransformerBlock) as well as the
single-GPU native block does.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/magi_cp_check.py
Rank 0 prints one JSON line {"world":N,"worst_vs_golden":...,"worst_vs_single":...}.
"""
import json
import os
import sys
import types
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from inferix_b200 import magi_cp, magi_layer  # noqa: E402
from inferix_b200.kvcache_manager.model import InferenceParams  # noqa: E402
from magi_golden_util import meta_from_plain  # noqa: E402
from oracle import magi_oracle as mo  # noqa: E402  (synthetic weights only)


def build(cfgd, seed, cp, dev):
    mc = types.SimpleNamespace(layernorm_epsilon=1e-6, apply_layernorm_1p=False, cond_hidden_ratio=0.25,
                               cond_gating_ratio=1.0, xattn_cond_hidden_ratio=1.0, params_dtype=torch.bfloat16, **cfgd)
    ec = types.SimpleNamespace(cp_size=cp, cp_strategy="cp_ulysses" if cp > 1 else "none", fp8_quant=False,
                               kv_offload=False)
    block = magi_layer.TransformerBlock(mc, ec)
    block.load_state_dict(mo.synth_state_dict(mo.MagiConfig(**cfgd), seed=seed), strict=True)
    return block.to(dev)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    g = torch.load(ROOT / "tests/golden/magi_layer_glu.pt")
    outs = {}
    for cp in (world, 1):
        if cp == 1 and rank != 0:
            continue
        if cp > 1:
            magi_cp.init_context_parallel(None, world, rank)
        else:
            magi_cp.destroy_context_parallel()
        block = build(g["cfg"], g["seed"], cp, dev)
        ip = InferenceParams(1, g["max_seq"], device=dev)
        res = []
        for st in g["steps"]:
            meta = meta_from_plain(st["meta"])
            hidden, cmap, rope = st["hidden"].to(dev), st["condition_map"].to(dev), st["rope"].to(dev)
            split = None
            if cp > 1:
                hidden, cmap, rope, split, (xq, xk) = magi_cp.cp_ulysses_process(
                    cp, hidden, cmap, rope, st["meta"]["cu_seqlens_q"], st["meta"]["cu_seqlens_kv"])
                meta.cp_split_sizes = split
                meta.cross_attn_params = types.SimpleNamespace(q_ranges=xq, kv_ranges=xk)
            ip.update_kv_cache = st["update"]
            out = block(hidden.contiguous(), st["condition"].to(dev), cmap, st["y"].to(dev), rope, ip, meta)
            res.append(magi_cp.cp_post_process(cp, "cp_ulysses", out, split).cpu())
        outs[cp] = res
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        vs_gold = max(rel(o, st["out"]) for o, st in zip(outs[world], g["steps"]))
        vs_one = max(rel(a, b) for a, b in zip(outs[world], outs[1]))
        print(json.dumps({"world": world, "worst_vs_golden": vs_gold, "worst_vs_single": vs_one,
                          "single_vs_golden": max(rel(o, st["out"]) for o, st in zip(outs[1], g["steps"]))}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
