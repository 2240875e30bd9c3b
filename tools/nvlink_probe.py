#!/usr/bin/env python
"""NVLink bytes of the sequence-parallel K/V exchange, measurable under ncu.

ncu must not wrap a multi-rank command, and the fused exchange blocks on flags that only the peers' concurrent launches
raise — a replayed kernel would spin into its trap.  This probe therefore runs rank 0 of an 8-rank group in ONE process
on a 2-GPU box: the seven peer caches (and their flag arrays) are plain allocations on cuda:1, addressed from cuda:0
through peer access over NVLink exactly like the CUDA-IPC mappings of the real run, and rank 0's own flag array is
pre-set far above any epoch so that every wait is satisfied immediately and every replay is idempotent.  The kernels,
their grids, the copy loops and the bytes that cross NVLink are those of rank 0 in the real 8-rank run
(ifx_wan_block_forward_sp, one 720p block of 3 frames: 1350 local rows, window 86 400).

    python tools/nvlink_probe.py --mode overlap|store [--iters 10]        # device times, one JSON line
    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,... -k regex:attn_fwd_kernel python tools/nvlink_probe.py --iters 1

Algorithmic bytes out of rank 0 per layer: 2 (K, V) x (P - 1) x S/P x C x 2 B = 2 * 7 * 1350 * 1536 * 2 = 58 060 800.
"""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path
from types import SimpleNamespace

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="overlap", choices=["overlap", "store"])
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--push-ctas", type=int, default=0)
a = ap.parse_args()
os.environ["IFX_SP_MODE"] = a.mode
if a.push_ctas:
    os.environ["IFX_SP_PUSH_CTAS"] = str(a.push_ctas)

import torch  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from inferix_b200 import _lib, ops, synthetic  # noqa: E402
from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest  # noqa: E402
from inferix_b200.parallel import ParallelConfig  # noqa: E402
from inferix_b200.wan_model import CausalWanModel  # noqa: E402

H_LAT, W_LAT, FRAMES, WINDOW_BLOCKS = 45, 80, 3, 8
FS = H_LAT * W_LAT
S, L = FRAMES * FS, WINDOW_BLOCKS * FRAMES * FS


def enable_peer_access(dev: int, peer: int) -> None:
    rt = C.CDLL("libcudart.so.12")
    rt.cudaSetDevice(dev)
    rc = rt.cudaDeviceEnablePeerAccess(peer, 0)
    if rc not in (0, 704):                                  # 704 = already enabled
        raise RuntimeError(f"cudaDeviceEnablePeerAccess({dev} -> {peer}) = {rc}")
    rt.cudaGetLastError()


def main():
    if torch.cuda.device_count() < 2:
        raise SystemExit("needs 2 GPUs (gpurun --gpus 2)")
    if not torch.cuda.can_device_access_peer(0, 1):
        raise SystemExit("no P2P between GPU 0 and GPU 1")
    d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
    torch.zeros(1, device=d1).to(d0)                        # contexts + torch's own peer enabling
    enable_peer_access(0, 1)
    torch.cuda.set_device(0)
    P = a.world
    rows = S // P

    cfg = dict(synthetic.WAN_1_3B, num_layers=1)
    model = CausalWanModel(**cfg, local_attn_size=WINDOW_BLOCKS * FRAMES, sink_size=0)
    model.load_state_dict(synthetic.synth_state_dict(cfg, seed=0))
    model = model.to(torch.bfloat16).to(d0)
    blk = model.blocks[0]
    blk.parallel_config = ParallelConfig(ring_size=P, rank=0, world_size=P)
    Cdim = cfg["dim"]
    mgr, req = KVCacheManager(d0), KVCacheRequest("probe")
    blk.kv_cache_manager.allocate_kv_cache(mgr, req, L, torch.bfloat16, page_tokens=FS)
    blk.kv_cache_manager.allocate_crossattn_cache(mgr, req, 512, torch.bfloat16)
    store = blk.kv_cache_manager.store(mgr, req)
    g = torch.Generator(device=d0).manual_seed(1)
    for b in range(WINDOW_BLOCKS - 1):                     # seven blocks of history through the real append path
        plan = store.plan_append(b * S, S, 0, True)
        store.append(plan, torch.randn(S, Cdim, device=d0, generator=g).bfloat16(),
                     torch.randn(S, Cdim, device=d0, generator=g).bfloat16())

    # the seven "peers": caches + flag arrays on GPU 1; rank 0's flags pre-satisfied
    peer_k = [torch.zeros_like(store.k, device=d1) for _ in range(P - 1)]
    peer_v = [torch.zeros_like(store.v, device=d1) for _ in range(P - 1)]
    peer_flags = [torch.zeros(P, dtype=torch.int64, device=d1) for _ in range(P - 1)]
    my_flags = torch.full((P,), 1 << 40, dtype=torch.int64, device=d0)
    dst = _lib.PeerDst()
    dst.world, dst.rank, dst.epoch, dst.local_only = P, 0, 0, 0
    dst.k[0], dst.v[0], dst.flags[0] = store.k.data_ptr(), store.v.data_ptr(), my_flags.data_ptr()
    for r in range(1, P):
        dst.k[r], dst.v[r], dst.flags[r] = peer_k[r - 1].data_ptr(), peer_v[r - 1].data_ptr(), peer_flags[r - 1].data_ptr()
    epoch = [0]

    def next_epoch():
        epoch[0] += 1
        return epoch[0]
    store.peer = dst
    store.peer_group = SimpleNamespace(next_epoch=next_epoch, flags=my_flags, world=P, rank=0)

    ctx = (torch.randn(1, 512, Cdim, device=d0, generator=g) * 0.5).bfloat16()
    table = ops.rope_table(model.freqs, d0)
    grid_sizes = torch.tensor([(FRAMES, H_LAT, W_LAT)])
    meta = {"global_end_index": torch.zeros(1, dtype=torch.long, device=d0),
            "local_end_index": torch.zeros(1, dtype=torch.long, device=d0)}
    cmeta = {"is_init": False}
    x = torch.randn(1, rows, Cdim, device=d0, generator=g).bfloat16()
    e0 = (torch.randn(1, FRAMES, 6, Cdim, device=d0, generator=g) * 0.3).bfloat16()
    start = [L - S]
    x_init = x.clone()

    def layer():
        x.copy_(x_init)
        # every call appends the next block: the first fills the window, the following ones evict (steady state)
        blk(x, e0, None, grid_sizes, table, ctx, None, None, meta, cmeta, current_start=start[0],
            kv_cache_manager=mgr, kv_cache_requests=[req])
        start[0] += S

    for _ in range(2):
        layer()
    torch.cuda.synchronize()
    e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_a.record()
    for _ in range(a.iters):
        layer()
    e_b.record()
    torch.cuda.synchronize()
    layer_us = 1e3 * e_a.elapsed_time(e_b) / a.iters
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

    # what arrived: rank 0's rows of the last appended block must equal the local cache on every "peer"
    torch.cuda.synchronize()
    k0 = store.k.to(d1)
    chunk = FS // P
    arrived = []
    for r in range(P - 1):
        pk = peer_k[r].view(-1, FS, Cdim)[:, :chunk]       # rank 0 owns the first `chunk` tokens of every page
        lk = k0.view(-1, FS, Cdim)[:, :chunk]
        written = pk.abs().sum(dim=(1, 2)) > 0
        arrived.append(bool(torch.equal(pk[written], lk[written])) and int(written.sum()) >= FRAMES)
    flags_seen = [int(f[0]) for f in peer_flags]
    print(json.dumps({"mode": a.mode, "world": P, "rows": rows, "window": L, "layer_us": round(layer_us, 1),
                      "kernels_us": per, "epoch": epoch[0], "peer_flags_rank0_slot": flags_seen,
                      "peer_rows_match_local_cache": arrived,
                      "algorithmic_nvlink_bytes_out_per_layer": 2 * (P - 1) * rows * Cdim * 2,
                      "note": "rank 0 of an emulated 8-rank group; all seven peer caches live on GPU 1"}), flush=True)


if __name__ == "__main__":
    main()
