"""Goldens for the MAGI-1 VideoDiTModel (embedding prologue, layer stack, epilogue, CFG dispatcher) from the reference.

ORACLE tooling (build container only; reads /root/reference).  Builds the reference's own
`inferix.models.magi.dit.dit_model.VideoDiTModel` on CPU with the substitutions of make_golden_magi_layer.py (third-
party CUDA kernels -> torch statements, CUDA autocast(float32) emulated) plus three device-only shims: the learnable
rotary bands are created on CPU, `Tensor.cuda()` is the identity, and `generate_kv_range_for_uncondition` (whose body
builds the string "cuda:<current_device>") is re-stated with the input's device.  Runs:
  * three `forward` calls sharing a KV cache (prefix extraction, denoising with history, storing forward);
  * `forward_dispatcher` with cfg_number = 3 (with and without the extra clean chunk) and cfg_number = 1 (plain and
    distill_nearly_clean_chunk).
Output: tests/golden/magi_model.pt (inputs, kwargs and outputs; weights are regenerated from names + seed).
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import magi_oracle as mo  # noqa: E402
from oracle.make_golden_magi_layer import import_reference  # noqa: E402

MODEL = dict(num_layers=2, hidden_size=512, ffn_hidden_size=1024, num_attention_heads=4, num_query_groups=2,
             kv_channels=128, patch_size=2, t_patch_size=1, in_channels=16, out_channels=16, caption_channels=64,
             caption_max_length=16, cond_hidden_ratio=0.25, xattn_cond_hidden_ratio=1.0, cond_gating_ratio=1.0,
             gated_linear_unit=True, x_rescale_factor=1.0, half_channel_vae=False)
CW, HW = 3, 12                     # chunk_width (latent frames per chunk), latent height = width
CLIP = CW * (HW // 2) ** 2          # tokens per chunk


def default_kv_range(slice_point, ranges):
    return torch.tensor([[0, (slice_point + i + 1) * CLIP] for i in range(ranges)], dtype=torch.int32)


def build_reference(cfg_number):
    import_reference()
    import inferix.models.magi.dit.dit_model as dmodel
    import inferix.models.magi.dit.dit_module as dm
    from inferix.core.config import EngineConfig, ModelConfig, RuntimeConfig
    dm.torch.cuda = types.SimpleNamespace(current_device=lambda: "cpu", get_device_capability=lambda *a: (8, 0),
                                          nvtx=torch.cuda.nvtx)
    torch.Tensor.cuda = lambda self, *a, **k: self

    class CpuVideoDiT(dmodel.VideoDiTModel):
        def generate_kv_range_for_uncondition(self, uncond_x):      # dit_model.py:92-100 with the input's device
            B, C, T, H, W = uncond_x.shape
            n = (T // self.model_config.t_patch_size) * (H // self.model_config.patch_size) * (W // self.model_config.patch_size)
            s = torch.linspace(0, (B - 1) * n, steps=B).reshape((B, 1))
            e = torch.linspace(n, B * n, steps=B).reshape((B, 1))
            return torch.concat([s, e], dim=1).to(torch.int32).to(uncond_x.device)

    mc = ModelConfig(model_name="tiny", params_dtype=torch.bfloat16, **MODEL)
    rc = RuntimeConfig(cfg_number=cfg_number, chunk_width=CW)
    ec = EngineConfig(cp_strategy="none", cp_size=1, fp8_quant=False, kv_offload=False, distill=False)
    model = CpuVideoDiT(types.SimpleNamespace(model_config=mc, runtime_config=rc, engine_config=ec))
    dmodel._high_precision_promoter(model)
    sd = mo.synth_model_state_dict(model, seed=1)
    model.load_state_dict(sd, strict=True)
    return model.eval()


def new_ip(max_seq):
    from inferix.kvcache_manager.kvcache_manager import KVCacheManager, KVCacheRequest
    return types.SimpleNamespace(max_sequence_length=max_seq, max_batch_size=1, sequence_len_offset=0,
                                 kv_cache_request=KVCacheRequest("magi"), kv_cache_manager=KVCacheManager("cpu"),
                                 key_value_memory_dict={}, update_kv_cache=False)


def inputs(g, n, ranges, ylens):
    x = torch.randn(n, 16, ranges * CW, HW, HW, generator=g)
    t = torch.rand(n, ranges, generator=g)
    y = torch.randn(n * ranges, 1, 16, 64, generator=g)
    mask = torch.zeros(n * ranges, 1, 16)
    for i in range(n * ranges):
        mask[i, 0, :ylens[i % len(ylens)]] = 1
    return x, t, y, mask


def main():
    g = torch.Generator().manual_seed(17)
    out = dict(model=MODEL, chunk_width=CW, hw=HW, clip=CLIP, seed=1, forward=[], dispatch=[])
    max_seq = 6 * CLIP

    # ---- plain forwards sharing a cache
    model = build_reference(cfg_number=1)
    out["state"] = {k: (str(v.dtype), tuple(v.shape)) for k, v in model.state_dict().items()}
    ip = new_ip(max_seq)
    plan = [(1, 0, True, dict(extract_prefix_video_feature=True, fwd_extra_1st_chunk=False), [9]),
            (2, 1, False, dict(fwd_extra_1st_chunk=False), [5, 12]),
            (3, 1, True, dict(fwd_extra_1st_chunk=True), [16, 3, 7])]
    for ranges, sp, update, flags, ylens in plan:
        x, t, y, mask = inputs(g, 1, ranges, ylens)
        kw = dict(range_num=sp + ranges, denoising_range_num=ranges, slice_point=sp, chunk_width=CW, num_steps=12,
                  distill_interval=4, **flags)
        kv_range = default_kv_range(sp, ranges)
        ip.update_kv_cache = update
        with torch.no_grad():
            o = model.forward(x, t, y, caption_dropout_mask=torch.tensor([False]), xattn_mask=mask, kv_range=kv_range,
                              inference_params=ip, **dict(kw))
        out["forward"].append(dict(x=x, t=t, y=y, mask=mask, kv_range=kv_range, kwargs=kw, update=update, out=o.clone()))

    # ---- CFG dispatcher
    for cfg_number, extra, distill in [(3, False, False), (3, True, False), (1, False, False), (1, True, True)]:
        model = build_reference(cfg_number=cfg_number)
        ip = new_ip(max_seq)
        # one stored clean chunk first, so the dispatcher's forwards read a history
        x0, t0, y0, m0 = inputs(g, 1, 1, [6])
        ip.update_kv_cache = True
        with torch.no_grad():
            model.forward(x0, t0, y0, caption_dropout_mask=torch.tensor([False]), xattn_mask=m0,
                          kv_range=default_kv_range(0, 1), inference_params=ip, range_num=1, denoising_range_num=1,
                          slice_point=0, chunk_width=CW, num_steps=12, distill_interval=4,
                          extract_prefix_video_feature=True, fwd_extra_1st_chunk=False)
        ranges = 3 if extra else 2
        sp = 0 if extra else 1          # with the extra clean chunk the forward starts one chunk earlier
        x, t, y, mask = inputs(g, 2, ranges, [8, 4, 11])
        t = t[0:1].repeat(2, 1)
        if cfg_number == 3:
            t = t * 0.9 + 0.05          # inside cfg_t_range
        kw = dict(range_num=sp + ranges, denoising_range_num=ranges, slice_point=sp, chunk_width=CW, num_steps=12,
                  distill_interval=4, fwd_extra_1st_chunk=extra, distill_nearly_clean_chunk=distill)
        kv_range = default_kv_range(sp, ranges)
        with torch.no_grad():
            o = model.forward_dispatcher(x.clone(), t, y, mask, kv_range, ip, **dict(kw))
        out["dispatch"].append(dict(cfg_number=cfg_number, prefix=dict(x=x0, t=t0, y=y0, mask=m0), x=x, t=t, y=y,
                                    mask=mask, kv_range=kv_range, kwargs=kw, out=o.clone()))
        print(f"dispatch cfg={cfg_number} extra={extra} distill={distill}: out {tuple(o.shape)} finite={bool(torch.isfinite(o).all())}")
    path = ROOT / "tests/golden/magi_model.pt"
    torch.save(out, path)
    print("wrote", path, f"{path.stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
