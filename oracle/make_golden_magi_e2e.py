"""End-to-end MAGI-1 golden: the reference's own SampleTransport (lifted, see make_golden_magi_walk.py) driving the
reference's own VideoDiTModel (CPU build of make_golden_magi_model.py) on a tiny text-to-video job with 3-way CFG.

ORACLE tooling (build container only).  Output: tests/golden/magi_e2e.pt — inputs, the clean chunks in the order the
reference yields them, and the final latent.  tests/ replays the job through inferix_b200.magi_pipeline +
inferix_b200.magi_model (kernel doubles on CPU, real kernels on the GPU).
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.make_golden_magi_model import CW, HW, MODEL, build_reference, new_ip  # noqa: E402
from oracle.make_golden_magi_walk import lift_sample_transport  # noqa: E402

JOB = dict(chunk_num=3, window=2, num_steps=4, cfg_number=3, n2c=[], clean_kv=-1)


def job_inputs():
    g = torch.Generator().manual_seed(31)
    L, C = MODEL["caption_max_length"], MODEL["caption_channels"]
    y = torch.randn(2, JOB["chunk_num"], L, C, generator=g)
    masks = torch.zeros(2, JOB["chunk_num"], L)
    for i, n in enumerate([9, 14, 5]):
        masks[0, i, :n] = 1
    masks[1, :, :1] = 1                      # the null caption keeps one token (as the reference's null mask does)
    noise = torch.randn(1, 16, JOB["chunk_num"] * CW, HW, HW, generator=g)
    return y, masks, noise


def main():
    model = build_reference(cfg_number=JOB["cfg_number"])
    model.runtime_config.window_size = JOB["window"]
    model.runtime_config.clean_t = 0.9999
    model.runtime_config.noise2clean_kvrange = JOB["n2c"]
    model.runtime_config.clean_chunk_kvrange = JOB["clean_kv"]
    model.engine_config.shortcut_mode = ""
    ns = lift_sample_transport()
    y, masks, noise = job_inputs()
    max_seq = noise.shape[2] * (HW // 2) ** 2
    ns["InferenceParams"] = lambda max_batch_size, max_sequence_length: new_ip(max_sequence_length)
    ti = types.SimpleNamespace(y=y, emb_masks=masks, prefix_video=None, latent_size=tuple(noise.shape),
                               t_schedule_config=dict(tSchedulerFunc="sd3", shift=3.0), num_steps=JOB["num_steps"],
                               chunk_num=JOB["chunk_num"], task_idx_list=[0], report_chunk_num_list=[JOB["chunk_num"]])
    ns["torch"].randn = lambda *shape, device=None: noise.clone()
    st = ns["SampleTransport"](model=model, transport_inputs=[ti], device=torch.device("cpu"))
    assert st.inference_params[0].max_sequence_length == max_seq
    with torch.no_grad():
        chunks = [(idx, c.clone()) for _, idx, c in st.walk()]
    path = ROOT / "tests/golden/magi_e2e.pt"
    torch.save(dict(job=JOB, model=MODEL, chunk_width=CW, hw=HW, seed=1, chunks=chunks, final_x=st.xs[0].clone()), path)
    print("wrote", path, f"{path.stat().st_size / 1e3:.0f} kB; chunks", [i for i, _ in chunks],
          "finite", all(bool(torch.isfinite(c).all()) for _, c in chunks))


if __name__ == "__main__":
    main()
