"""Goldens for the MAGI host logic (scheduler index functions + MagiKVCacheManager), from the reference's own code.

ORACLE tooling (build container only).  `inferix/pipeline/magi/video_generate.py` cannot be imported here — its
package pulls timm / bs4 / a CUDA-only DiT — so the functions under test are lifted out of the reference SOURCE FILE
with `ast` (module-level functions + the SampleTransport class body) and executed unmodified in a namespace where
every unrelated name is a placeholder.  `MagiKVCacheManager` and `KVCacheManager` import as they are.
Outputs: tests/golden/magi_schedule.json, tests/golden/magi_kv.pt
"""
from __future__ import annotations

import ast
import builtins
import itertools
import json
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")


def lift(path: Path, names):
    tree = ast.parse(path.read_text())
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"torch": torch, "List": list, "Dict": dict, "Tuple": tuple, "Union": object, "Optional": object}
    for node in ast.walk(mod):
        if isinstance(node, ast.Name) and node.id not in ns and not hasattr(builtins, node.id):
            ns[node.id] = type(node.id, (), {})
    for n in keep:
        ns.pop(n.name, None)
    exec(compile(mod, str(path), "exec"), ns)
    return ns


def schedule_goldens():
    ns = lift(REF / "inferix/pipeline/magi/video_generate.py",
              {"generate_sequences", "init_t", "init_intervel", "SampleTransport"})
    out = {"generate_sequences": [], "init_t": [], "init_intervel": [], "kvrange": [], "status": [], "timestep": []}
    for chunk_num, window, offset in itertools.product([1, 3, 8, 13], [1, 4, 5], [0, 1, 2]):
        if offset >= chunk_num:
            continue
        out["generate_sequences"].append(dict(args=[chunk_num, window, offset],
                                              ret=[list(v) for v in ns["generate_sequences"](chunk_num, window, offset)]))
    for steps, cfg, sc in itertools.product([8, 12, 16, 64], [dict(tSchedulerFunc="sd3", shift=3.0), dict(tSchedulerFunc="square"),
                                                               dict(tSchedulerFunc="piecewise"), dict(tSchedulerFunc="id")],
                                            ["", "16,16,8"]):
        t = ns["init_t"](dict(cfg), steps, torch.device("cpu"), shortcut_mode=sc)
        out["init_t"].append(dict(steps=steps, cfg=cfg, shortcut=sc, ret=t.tolist()))
        out["init_intervel"].append(dict(steps=steps, shortcut=sc,
                                         ret=ns["init_intervel"](steps, torch.device("cpu"), shortcut_mode=sc).tolist()))

    ST = ns["SampleTransport"]
    for n2c, clean_kv, num_steps, window in [([], -1, 16, 4), ([5, 4, 3, 2], -1, 16, 4), ([5, 4, 3, 2], 6, 64, 4), ([3], 2, 12, 4)]:
        st = object.__new__(ST)
        st.device = torch.device("cpu")
        st.runtime_config = types.SimpleNamespace(clean_chunk_kvrange=clean_kv, noise2clean_kvrange=n2c, clean_t=0.9999,
                                                  chunk_width=6, window_size=window)
        st.model_config = types.SimpleNamespace(patch_size=2)
        st.chunk_width, st.window_size = 6, window
        chunk_num = 9
        st.transport_inputs = [types.SimpleNamespace(latent_size=[1, 16, 6 * chunk_num, 12, 20], num_steps=num_steps,
                                                     prefix_video=None, chunk_num=chunk_num)]
        _, ctn = st.get_batch_size_and_chunk_token_nums(0)
        for range_num in (1, 3, 7):
            out["kvrange"].append(dict(kind="prefix", n2c=n2c, clean_kv=clean_kv, range_num=range_num, ctn=ctn,
                                       ret=st.generate_kvrange_for_prefix_video(0, range_num).tolist()))
        per_stage = num_steps // window
        for step in range(0, per_stage * (chunk_num + window - 1), 3):
            (dps, stage, idx), (off, cs, ce, ts, te) = st.generate_denoise_status_and_sequences(0, step)
            out["status"].append(dict(step=step, num_steps=num_steps, chunk_num=chunk_num, window=window,
                                      ret=[[dps, stage, idx], [off, cs, ce, ts, te]]))
            for has_clean in (False, True):
                steps_each = st.get_denoise_step_of_each_chunk(0, dps, ts, te, idx, has_clean_t=has_clean)
                slice_point = max(0, cs - (1 if has_clean else 0))
                nrange = len(steps_each)
                kr = st.generate_kvrange_for_denoising_video(0, slice_point, nrange, steps_each)
                out["kvrange"].append(dict(kind="denoise", n2c=n2c, clean_kv=clean_kv, num_steps=num_steps, dps=dps,
                                           t_start=ts, t_end=te, idx=idx, has_clean=has_clean, slice_point=slice_point,
                                           steps_each=steps_each, ctn=ctn, ret=kr.tolist()))
        t_total = ns["init_t"](dict(tSchedulerFunc="sd3", shift=3.0), num_steps, torch.device("cpu"))
        for (s, e, i, hc) in [(0, 1, 0, False), (0, 4, 1, False), (1, 4, 2, True), (2, 4, per_stage - 1, False)]:
            out["timestep"].append(dict(num_steps=num_steps, dps=per_stage, start=s, end=e, idx=i, has_clean=hc,
                                        ret=st.get_timestep(t_total, per_stage, s, e, i, has_clean_t=hc).tolist()))
        out.setdefault("total_forward_step", []).append(dict(num_steps=num_steps, chunk_num=chunk_num, window=window,
                                                            ret=st.total_forward_step(0)))
    path = ROOT / "tests/golden/magi_schedule.json"
    path.write_text(json.dumps(out))
    print("wrote", path, {k: len(v) for k, v in out.items()}, f"{path.stat().st_size / 1e3:.0f} kB")


def kv_goldens():
    """MagiKVCacheManager._full_adjust_key_and_value / adjust_key_and_value_for_inference on the reference's CPU
    KVCacheManager: a 3-chunk history, then a forward that loads [0, slice_point) and appends one clip."""
    sys.path.insert(0, str(REF))
    from oracle.make_golden import install_shims
    install_shims()
    from inferix.kvcache_manager.kvcache_manager import KVCacheManager, KVCacheRequest
    from inferix.kvcache_manager.model.magi_kv_cache_manager import MagiKVCacheManager
    hn, d, clip, chunks = 2, 128, 8, 4
    g = torch.Generator().manual_seed(11)
    mgr = MagiKVCacheManager(3, hn, d, types.SimpleNamespace(kv_offload=False))
    ip = types.SimpleNamespace(max_sequence_length=clip * chunks, max_batch_size=1,
                               kv_cache_request=KVCacheRequest("magi"), kv_cache_manager=KVCacheManager("cpu"),
                               update_kv_cache=True)
    steps = []

    def meta(slice_point, extract=False, extra=False, distill=False):
        return types.SimpleNamespace(slice_point=slice_point, clip_token_nums=clip, extract_prefix_video_feature=extract,
                                     fwd_extra_1st_chunk=extra, distill_nearly_clean_chunk=distill)
    plan = [(meta(0, extract=True), 2 * clip, True), (meta(2), 2 * clip, True), (meta(2), 2 * clip, False),
            (meta(1, distill=True), 3 * clip, True), (meta(0), clip, True)]
    for m, ntok, update in plan:
        kv = torch.randn(ntok, hn, 2 * d, generator=g).bfloat16()
        ip.update_kv_cache = update
        k, v = mgr.adjust_key_and_value_for_inference(kv, ip, m)
        steps.append(dict(meta=vars(m), update=update, kv=kv, k=k.clone(), v=v.clone()))
    raw = ip.kv_cache_manager.get_raw(ip.kv_cache_request, "layer_3").clone()
    path = ROOT / "tests/golden/magi_kv.pt"
    torch.save(dict(hn=hn, d=d, clip=clip, chunks=chunks, steps=steps, final_cache=raw), path)
    print("wrote", path, f"{path.stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    schedule_goldens()
    kv_goldens()
