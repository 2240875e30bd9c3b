"""Shared by the CPU and GPU end-to-end MAGI tests: the tiny job of oracle/make_golden_magi_e2e.py on the native stack."""
import types

import torch

from oracle import magi_oracle as mo


def job_inputs(g):
    """Same draws as oracle/make_golden_magi_e2e.job_inputs (kept here so the GPU box needs no reference)."""
    gen = torch.Generator().manual_seed(31)
    m, job = g["model"], g["job"]
    L, C = m["caption_max_length"], m["caption_channels"]
    y = torch.randn(2, job["chunk_num"], L, C, generator=gen)
    masks = torch.zeros(2, job["chunk_num"], L)
    for i, n in enumerate([9, 14, 5]):
        masks[0, i, :n] = 1
    masks[1, :, :1] = 1
    noise = torch.randn(1, 16, job["chunk_num"] * g["chunk_width"], g["hw"], g["hw"], generator=gen)
    return y, masks, noise


def run_native(g, device, inference_params):
    from inferix_b200 import magi_model, magi_pipeline
    job = g["job"]
    mc = types.SimpleNamespace(model_name="tiny", params_dtype=torch.bfloat16, layernorm_epsilon=1e-6,
                               apply_layernorm_1p=False, **g["model"])
    rc = types.SimpleNamespace(cfg_number=job["cfg_number"], chunk_width=g["chunk_width"], window_size=job["window"],
                               clean_t=0.9999, noise2clean_kvrange=job["n2c"], clean_chunk_kvrange=job["clean_kv"],
                               cfg_t_range=[0, 0.0217, 0.1000, 0.3, 0.999], prev_chunk_scales=[1.5] * 5,
                               text_scales=[7.5] * 5)
    ec = types.SimpleNamespace(cp_strategy="none", cp_size=1, fp8_quant=False, kv_offload=False, distill=False,
                               shortcut_mode="", distill_nearly_clean_chunk_threshold=0.3)
    model = magi_model.VideoDiTModel(types.SimpleNamespace(model_config=mc, runtime_config=rc, engine_config=ec))
    model.load_state_dict(mo.synth_model_state_dict(model, seed=g["seed"]), strict=True)
    model = model.eval().to(device)
    y, masks, noise = job_inputs(g)
    ti = magi_pipeline.InferenceInput(y=y.to(device), emb_masks=masks.to(device), prefix_video=None,
                                      latent_size=tuple(noise.shape), num_steps=job["num_steps"],
                                      chunk_num=job["chunk_num"], t_schedule_config=dict(tSchedulerFunc="sd3", shift=3.0))
    st = magi_pipeline.SampleTransport(model, [ti], device, inference_params=inference_params, noise=noise.to(device))
    with torch.no_grad():
        chunks = [(idx, c.float().cpu().clone()) for _, idx, c in st.walk()]
    return chunks, st.xs[0].float().cpu()
