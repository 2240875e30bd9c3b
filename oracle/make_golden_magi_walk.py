"""Goldens for the MAGI-1 chunk scheduler loop (SampleTransport.walk / forward_velocity / integrate_velocity) from the
reference's own class.

ORACLE tooling (build container only).  `SampleTransport` is lifted out of the reference SOURCE FILE
(inferix/pipeline/magi/video_generate.py) with `ast` — the module cannot be imported here (timm / bs4 / CUDA-only
imports) — and executed unmodified around a stand-in model whose `forward_dispatcher` is a deterministic closed-form
function that records its arguments.  The trace of every model call (kwargs, timesteps, kv ranges, input checksums) and
every yielded clean chunk go to tests/golden/magi_walk.pt; tests/test_magi_pipeline_cpu.py replays the same stand-in
under inferix_b200.magi_pipeline.SampleTransport and compares bit-for-bit.
"""
from __future__ import annotations

import ast
import builtins
import sys
import types
from collections import Counter
from dataclasses import dataclass
from pathlib import Path
from queue import Queue

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/inferix/pipeline/magi/video_generate.py")


class StandInModel:
    """Has what SampleTransport reads from a VideoDiTModel; the velocity is a closed form of (x, timestep)."""

    def __init__(self, runtime, engine, patch_size=2):
        self.model_config = types.SimpleNamespace(patch_size=patch_size, in_channels=16, half_channel_vae=False)
        self.runtime_config, self.engine_config = runtime, engine
        self.y_embedder = object()
        self.calls = []

    def forward_dispatcher(self, x, timestep, y, mask, kv_range, inference_params, **kw):
        rec = {k: (float(v) if torch.is_tensor(v) else v) for k, v in kw.items()}
        self.calls.append(dict(kwargs=rec, timestep=timestep.clone(), kv_range=kv_range.tolist(), x_shape=tuple(x.shape),
                               x_sum=float(x.double().sum()), y_shape=tuple(y.shape), y_sum=float(y.double().sum()),
                               mask_sum=float(mask.double().sum())))
        cw = kw["chunk_width"]
        n, c, t, h, w = x.shape
        tt = timestep[:, :, None].expand(-1, -1, cw).reshape(n, 1, t, 1, 1)
        return torch.sin(x * 1.3) * 0.25 + (1.0 - tt) * 0.1 - x * 0.05


def lift_sample_transport():
    tree = ast.parse(REF.read_text())
    names = {"SampleTransport", "WorkStatus", "find_dit_model", "generate_sequences", "init_t", "init_intervel"}
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    mod = ast.Module(body=keep, type_ignores=[])
    tdist = types.SimpleNamespace(get_rank=lambda *a, **k: 0)
    torch_ns = types.SimpleNamespace(**{k: getattr(torch, k) for k in dir(torch) if not k.startswith("__")})
    torch_ns.distributed = tdist

    class FakeBar:
        def __init__(self, *a, **k):
            pass

        def update(self, n):
            pass

        def close(self):
            pass

    ns = {"torch": torch_ns, "List": list, "Dict": dict, "Tuple": tuple, "Union": object, "Optional": object,
          "Generator": object, "dataclass": dataclass, "Queue": Queue, "Counter": Counter, "tqdm": FakeBar,
          "print_rank_0": lambda *a, **k: None,
          "mpu": types.SimpleNamespace(get_pp_world_size=lambda: 1, is_pipeline_first_stage=lambda: True,
                                       is_pipeline_last_stage=lambda: True),
          "event_path_timer": lambda: types.SimpleNamespace(synced_record=lambda *a, **k: None),
          "InferenceParams": lambda max_batch_size, max_sequence_length: types.SimpleNamespace(
              max_batch_size=max_batch_size, max_sequence_length=max_sequence_length, update_kv_cache=False)}
    for node in ast.walk(mod):
        if isinstance(node, ast.Name) and node.id not in ns and not hasattr(builtins, node.id):
            ns[node.id] = type(node.id, (), {})
    for n in keep:
        ns.pop(n.name, None)
    exec(compile(mod, str(REF), "exec"), ns)
    return ns


CASES = {
    "t2v": dict(chunk_num=5, window=4, num_steps=16, n2c=[5, 4, 3, 2], clean_kv=-1, prefix_frames=0, shortcut="8,16,16"),
    "v2v_prefix2": dict(chunk_num=5, window=4, num_steps=12, n2c=[], clean_kv=-1, prefix_frames=6, shortcut="16,16,8"),
    "i2v": dict(chunk_num=3, window=4, num_steps=8, n2c=[3, 2], clean_kv=2, prefix_frames=1, shortcut=""),
}
CW, HW, L, CC = 3, 8, 6, 12


def case_inputs(c):
    g = torch.Generator().manual_seed(23)
    y = torch.randn(2, c["chunk_num"], L, CC, generator=g)
    masks = (torch.rand(2, c["chunk_num"], L, generator=g) > 0.3).float()
    prefix = torch.randn(1, 16, c["prefix_frames"], HW, HW, generator=g) if c["prefix_frames"] else None
    noise = torch.randn(1, 16, c["chunk_num"] * CW, HW, HW, generator=g)
    runtime = types.SimpleNamespace(chunk_width=CW, window_size=c["window"], clean_t=0.9999,
                                    noise2clean_kvrange=c["n2c"], clean_chunk_kvrange=c["clean_kv"])
    engine = types.SimpleNamespace(shortcut_mode=c["shortcut"], distill_nearly_clean_chunk_threshold=0.3)
    return y, masks, prefix, noise, runtime, engine


def main():
    ns = lift_sample_transport()
    out = {}
    for name, c in CASES.items():
        y, masks, prefix, noise, runtime, engine = case_inputs(c)
        model = StandInModel(runtime, engine)
        ti = types.SimpleNamespace(y=y, emb_masks=masks, prefix_video=prefix, latent_size=tuple(noise.shape),
                                   t_schedule_config=dict(tSchedulerFunc="sd3", shift=3.0), num_steps=c["num_steps"],
                                   chunk_num=c["chunk_num"], task_idx_list=[0], report_chunk_num_list=[c["chunk_num"]])
        orig_randn = torch.randn
        ns["torch"].randn = lambda *shape, device=None: noise.clone()      # SampleTransport draws its start latent (:309)
        st = ns["SampleTransport"](model=model, transport_inputs=[ti], device=torch.device("cpu"))
        ns["torch"].randn = orig_randn
        chunks = [(idx, chunk.clone()) for _, idx, chunk in st.walk()]
        out[name] = dict(case=c, calls=model.calls, chunks=chunks, final_x=st.xs[0].clone())
        print(name, "model calls", len(model.calls), "chunks yielded", [i for i, _ in chunks])
    path = ROOT / "tests/golden/magi_walk.pt"
    torch.save(out, path)
    print("wrote", path, f"{path.stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
