#!/usr/bin/env python
"""CoreAttention's ring strategies on real GPUs with the native (out, lse) kernel: every rank's pass-kv and pass-q
result against one ifx_attention over the all-gathered keys.  Launch under torch.distributed.run (>= 2 ranks);
rank 0 prints one JSON line."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from inferix_b200.attention import CoreAttention, ifx_attn_forward  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(3)
    lq, lk, n, d = 700, 1300, 4, 128
    q = torch.randn(world, 1, lq, n, d, generator=g).bfloat16().to(dev)
    k = torch.randn(world, 1, lk, n, d, generator=g).bfloat16().to(dev)
    v = torch.randn(world, 1, lk, n, d, generator=g).bfloat16().to(dev)
    want, want_lse = ifx_attn_forward(q[rank], k.transpose(0, 1).reshape(1, world * lk, n, d).contiguous(),
                                      v.transpose(0, 1).reshape(1, world * lk, n, d).contiguous())
    ca = CoreAttention(strategy="pass-kv", ring_pg=dist.group.WORLD)
    scale = d ** -0.5

    def rel(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm()).item()
    o_kv, l_kv = ca.ring_attention_forward_pass_kv(dist.group.WORLD, q[rank], k[rank], v[rank], scale)
    o_q, l_q = ca.ring_attention_forward_pass_q(dist.group.WORLD, q[rank], k[rank], v[rank], scale)
    ca.strategy = "pass-q"
    o_f = ca(q[rank], k[rank], v[rank])
    res = torch.tensor([rel(o_kv, want), rel(l_kv, want_lse), rel(o_q, want), rel(l_q.float().squeeze(-1).transpose(1, 2), want_lse),
                        rel(o_f, want)], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        names = ["pass_kv_out", "pass_kv_lse", "pass_q_out", "pass_q_lse_bf16", "forward_pass_q_out"]
        print(json.dumps({"world": world, **{nm: round(x, 6) for nm, x in zip(names, res.tolist())}}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
