"""Golden for inferix_b200.ulysses_scheduler: the reference's own UlyssesScheduler (context_parallel.py:382-598) run
under a 2-rank gloo group on CPU.  The class and its all-to-all helpers are lifted out of the reference source file
with `ast` and executed unmodified (the module itself imports half the framework); `mpu` is a 3-function stand-in
over the gloo group.  Test infrastructure, build container only:

    python oracle/make_golden_ulysses.py        # writes tests/golden/ulysses_sched.pt
"""
import ast
import os
import socket
import sys
import types
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF_FILE = Path("/root/reference/inferix/distributed/parallelism/context_parallel.py")
NAMES = {"FakeHandle", "all_to_all_input_split", "all_to_all_output_split", "fused_qkv_communication", "UlyssesScheduler"}
SPLIT = [5, 3]                      # uneven sequence split over the two ranks


def lift(rank, world):
    from typing import Callable, List, Tuple, Union
    from einops import rearrange
    tree = ast.parse(REF_FILE.read_text())
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in NAMES]
    mpu = types.SimpleNamespace(get_cp_world_size=lambda: world, get_cp_rank=lambda: rank,
                                get_cp_group=lambda: dist.group.WORLD)
    ns = {"torch": torch, "rearrange": rearrange, "mpu": mpu, "Callable": Callable, "List": List, "Tuple": Tuple,
          "Union": Union, "divide": lambda a, b: a // b}
    exec(compile(ast.Module(body=keep, type_ignores=[]), str(REF_FILE), "exec"), ns)
    return ns


def core_attn(q, k, v):
    """fp32 softmax attention with grouped KV heads: [S, hq, d], [Skv, hk, d] -> [S, hq, d] (deterministic on CPU)."""
    rep = q.shape[1] // k.shape[1]
    kf = k.repeat_interleave(rep, dim=1).transpose(0, 1)
    vf = v.repeat_interleave(rep, dim=1).transpose(0, 1)
    p = torch.softmax(q.transpose(0, 1) @ kf.transpose(1, 2) / q.shape[-1] ** 0.5, dim=-1)
    return (p @ vf).transpose(0, 1).contiguous()


def cases():
    # (name, q heads, kv heads (global, before the exchange), overlap_degree, cached history rows)
    return [("gqa_auto", 8, 4, -1, 0), ("gqa_od1", 8, 4, 1, 4), ("gqa_od2", 16, 4, 2, 2), ("mqa_repeat", 8, 1, 2, 0),
            ("mha", 4, 4, 1, 3)]


def make_inputs(rank, hq, hk, hist, hd=16):
    g = torch.Generator().manual_seed(100 * hq + 10 * hk + hist)
    total = sum(SPLIT)
    q = torch.randn(total, hq, hd, generator=g)
    k = torch.randn(total, hk, hd, generator=g)
    v = torch.randn(total, hk, hd, generator=g)
    lo = sum(SPLIT[:rank])
    sl = slice(lo, lo + SPLIT[rank])
    return q[sl].contiguous(), k[sl].contiguous(), v[sl].contiguous(), g


def run_variants(sched, rank, world):
    out = {}
    for name, hq, hk, od, hist in cases():
        q, k, v, g = make_inputs(rank, hq, hk, hist)
        hd = q.shape[-1]
        kv_heads_local = max(hk, world) // world
        hist_kv = torch.randn(hist, kv_heads_local, 2 * hd, generator=torch.Generator().manual_seed(7 + hist))

        def kv_cache(kv, hist_kv=hist_kv, hd=hd):
            full = torch.cat([hist_kv, kv], dim=0)
            return full[..., :hd].contiguous(), full[..., hd:].contiguous()
        xattn = lambda: torch.full((3,), float(rank))                                  # noqa: E731
        a, xa = sched.get_attn_and_xattn_with_comm_overlap(lambda: q, lambda: k, lambda: v, kv_cache, core_attn, xattn,
                                                           od, 1, world, SPLIT)
        b, _ = sched.get_attn_and_xattn_with_fused_kv_comm(lambda: q, lambda: torch.cat([k, v], dim=-1), kv_cache,
                                                           core_attn, xattn, od, 1, world, SPLIT)
        c, _ = sched.get_attn_and_xattn_with_fused_qkv_comm(lambda: (q, k, v), kv_cache, core_attn, xattn, od, 1, world,
                                                            SPLIT)
        out[name] = {"comm_overlap": a.clone(), "fused_kv": b.clone(), "fused_qkv": c.clone(), "xattn": xa.clone()}
    return out


def worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ns = lift(rank, world)
        ret[rank] = run_variants(ns["UlyssesScheduler"], rank, world)
    finally:
        dist.destroy_process_group()


def main():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(worker, args=(2, port, ret), nprocs=2, join=True)
        gold = {r: ret[r] for r in range(2)}
    out = ROOT / "tests" / "golden" / "ulysses_sched.pt"
    torch.save({"split": SPLIT, "ranks": gold, "source": "reference UlyssesScheduler, gloo world 2"}, out)
    print(out, out.stat().st_size, "bytes;", {k: tuple(v["comm_overlap"].shape) for k, v in gold[0].items()})


if __name__ == "__main__":
    main()
