#!/usr/bin/env python
"""Times the UNMODIFIED reference (alibaba-damo-academy/Inferix, installed into baseline/_ref) on the GPU of this box:
its own CausalWanModel + CausalInferencePipeline + KVCacheManager + flash-attn call, on the benchmark's shape
(Self-Forcing 720p, Wan-1.3B widths, synthetic weights, block = 3 latent frames, 24-frame window) — "the number to
beat" of BASELINE.md §4.4 next to the CPU reference arm.  None of this repo's kernels run here.

    python tools/ref_gpu_bench.py [--offload 0|1] [--blocks 10] [--timesteps 30] [--out file.json]

Method: the reference pipeline generates `blocks` blocks with a SHORT step list (2 noisy forwards + the clean pass per
block) so that the window fills within seconds; every model forward is bracketed by CUDA events (forward hooks).  In
steady state (window full) the first forward of a block evicts + rolls the cache and the others rewrite the newest
rows, so one production block of T timesteps costs  t_first + T * t_rest  (T noisy + 1 clean = T + 1 forwards), which
is what is reported, together with the raw per-forward times.  The third-party packages the reference imports but does
not use on this path (diffusers, yunchang, xfuser, ftfy) are stubbed exactly as oracle/make_golden.py does; flash-attn
is the real installed one (the reference's FA2 branch, models/attention/flash_attention.py:117-147).
"""
from __future__ import annotations

import argparse
import json
import sys
import time
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def install_shims():
    class Permissive(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return type(item, (), {})

    def mod(name, **attrs):
        m = Permissive(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    class ConfigMixin:
        pass

    class ModelMixin(torch.nn.Module):
        pass

    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=lambda fn: fn)
    mod("diffusers.models")
    mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    mod("diffusers.schedulers")
    mod("diffusers.schedulers.scheduling_utils", KarrasDiffusionSchedulers=[])
    mod("diffusers.utils", deprecate=lambda *a, **k: None, is_scipy_available=lambda: True)
    mod("diffusers.utils.torch_utils", randn_tensor=None)
    mod("yunchang", LongContextAttention=object)
    mod("yunchang.ring")
    mod("yunchang.ring.utils", RingComm=object, update_out_and_lse=None)
    mod("yunchang.kernels", AttnType=types.SimpleNamespace(FA="fa", TORCH="torch"))
    mod("yunchang.comm")
    mod("yunchang.comm.all_to_all", SeqAllToAll4D=object)
    mod("yunchang.globals", PROCESS_GROUP=object)
    mod("xfuser")
    mod("xfuser.logger", init_logger=lambda *a, **k: None)
    mod("xfuser.core")
    mod("xfuser.core.distributed", get_sp_group=None, get_sequence_parallel_rank=None,
        get_sequence_parallel_world_size=None, init_distributed_environment=None, initialize_model_parallel=None,
        get_world_group=None)
    mod("xfuser.core.long_ctx_attention", xFuserLongContextAttention=object)
    mod("ftfy")
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0


def find_reference() -> str:
    for cand in (ROOT / "baseline" / "_ref", Path("/root/reference")):
        if (cand / "inferix" / "__init__.py").exists():
            return str(cand)
    raise SystemExit(json.dumps({"ref_gpu": "unavailable", "why": "baseline/_ref/inferix not found (pip install "
                                 "--no-deps --target baseline/_ref of the reference, see DESIGN.md)"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--offload", type=int, default=0, help="enable_kv_offload of the reference model (its default is 1)")
    ap.add_argument("--blocks", type=int, default=10)
    ap.add_argument("--timesteps", type=int, default=30, help="T of the production block being extrapolated to")
    ap.add_argument("--tiny-cpu", action="store_true", help="plumbing dry run: tiny widths, CPU, SDPA fallback")
    ap.add_argument("--parity", action="store_true",
                    help="tiny widths on the GPU: the reference's own GPU path (FA2 + cuBLAS) and the native pipeline on "
                         "the same weights / noise / re-noise stream; prints their distance instead of timings")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    from inferix_b200 import synthetic
    install_shims()
    sys.path.insert(0, find_reference())
    import inferix.models.attention  # noqa: F401
    fa_mod = sys.modules["inferix.models.attention.flash_attention"]
    from inferix.models.self_forcing import causal_model as cm
    if a.tiny_cpu:
        fa_mod.HAS_FLASH_ATTN = fa_mod.HAS_FLASH_ATTN_HOPPER = False
        ref_attention = fa_mod.attention
        import inferix.models.attention as att_pkg

        def sdpa(q, k, v, **kw):
            kw.pop("k_lens", None)
            return ref_attention(q, k, v, dtype=q.dtype)
        att_pkg.flash_attention = att_pkg.attention = cm.attention = sdpa
        dev, cfg, hw, window_frames = torch.device("cpu"), dict(synthetic.TINY), (16, 16), 6
    else:
        assert torch.cuda.is_available(), "the reference's GPU path needs a CUDA device"
        assert fa_mod.HAS_FLASH_ATTN, "flash-attn is not importable: the reference would fall back to SDPA"
        fa_mod.HAS_FLASH_ATTN_HOPPER = False        # FA3 (Hopper-only) is not installed / not applicable on sm_100
        dev, cfg, hw, window_frames = torch.device("cuda", 0), dict(synthetic.WAN_1_3B), (90, 160), 24
        if a.parity:
            cfg, hw, window_frames, a.blocks = dict(synthetic.TINY), (16, 16), 6, 4
    torch.set_grad_enabled(False)

    from inferix.core.types import DecodeMode
    from inferix.kvcache_manager.kvcache_manager import KVCacheManager, KVCacheRequest
    from inferix.models.schedulers.flow_match import FlowMatchScheduler
    from inferix.models.self_forcing import wrapper as wr
    from inferix.models.wan_base import ParallelConfig
    from inferix.pipeline.self_forcing.CausalInferencePipeline import CausalInferencePipeline

    pc = ParallelConfig.__new__(ParallelConfig)
    pc.ulysses_size = pc.ring_size = pc.world_size = 1
    pc.rank = pc.local_rank = 0
    pc.ring_strategy, pc.attn_backend = "pass-kv", "FlexAttention"
    model = cm.CausalWanModel(model_type="t2v", patch_size=(1, 2, 2), text_len=cfg["text_len"], in_dim=cfg["in_dim"],
                              dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"], text_dim=cfg["text_dim"],
                              out_dim=cfg["out_dim"], num_heads=cfg["num_heads"], num_layers=cfg["num_layers"],
                              local_attn_size=window_frames, sink_size=0, qk_norm=True, cross_attn_norm=True, eps=1e-6,
                              enable_kv_offload=bool(a.offload), parallel_config=pc)
    if a.tiny_cpu:
        for blk in model.blocks:
            blk.self_attn.attention = cm.attention
    model.load_state_dict(synthetic.synth_state_dict(cfg, seed=0), strict=True)
    model = model.to(torch.bfloat16).to(dev).eval()

    gen = wr.WanDiffusionWrapper.__new__(wr.WanDiffusionWrapper)
    torch.nn.Module.__init__(gen)
    gen.parallel_config, gen.enable_kv_offload, gen.model, gen.uniform_timestep = pc, bool(a.offload), model, False
    gen.scheduler = FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    gen.scheduler.set_timesteps(1000, training=True)
    gen.seq_len = 32760

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": text_prompts}

    fs = (hw[0] // 2) * (hw[1] // 2)
    pipe = CausalInferencePipeline.__new__(CausalInferencePipeline)
    torch.nn.Module.__init__(pipe)
    pipe.parallel_config, pipe._profiler = pc, None
    pipe.generator, pipe.text_encoder, pipe.vae = gen, Text(), None
    pipe.scheduler = gen.scheduler
    steps = [1000, 750, 500, 250] if a.parity else [1000, 500]
    sched_ts = torch.cat((gen.scheduler.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
    pipe.denoising_step_list = sched_ts[1000 - torch.tensor(steps, dtype=torch.long)]
    pipe.num_transformer_blocks = cfg["num_layers"]
    pipe.frame_seq_length = fs
    pipe.kv_cache_meta = pipe.crossattn_cache_meta = None
    pipe.args = types.SimpleNamespace(context_noise=0)
    pipe.num_frame_per_block = 3
    pipe.independent_first_frame = False
    pipe.local_attn_size = model.local_attn_size
    model.num_frame_per_block = 3

    # per-forward device time
    events, wall = [], []
    use_ev = dev.type == "cuda"

    def pre(_m, _a, _k=None):
        if use_ev:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events.append([e, None])
        else:
            wall.append([time.perf_counter(), None])

    def post(_m, _a, _o):
        if use_ev:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events[-1][1] = e
        else:
            wall[-1][1] = time.perf_counter()
    model.register_forward_pre_hook(pre)
    model.register_forward_hook(post)

    g = torch.Generator().manual_seed(1)
    frames = 3 * a.blocks
    noise = torch.randn(1, frames, 16, hw[0], hw[1], generator=g).bfloat16().to(dev)
    context = torch.randn(1, 20, cfg["text_dim"], generator=g).bfloat16().to(dev)
    mgr = KVCacheManager(dev)
    t0 = time.perf_counter()
    if a.parity:
        # same re-noise stream on both sides: torch.randn_like (reference :307, ours pipeline.renoise_fn) drawn from a
        # CPU generator with a fixed seed
        true_randn_like = torch.randn_like

        def seeded(seed):
            rg = torch.Generator().manual_seed(seed)
            return lambda x, **kw: torch.randn(x.shape, generator=rg, dtype=torch.float32).to(x.dtype).to(x.device)
        torch.randn_like = seeded(5)
    ref_out = pipe.inference(noise=noise, text_prompts=context, kv_cache_manager=mgr,
                             kv_cache_requests=[KVCacheRequest("ref")], free_cache_before_vae=False,
                             decode_mode=DecodeMode.NO_DECODE)
    if a.parity:
        torch.randn_like = true_randn_like
        ref_idx = (int(pipe.kv_cache_meta[0]["global_end_index"]), int(pipe.kv_cache_meta[0]["local_end_index"]))
        ref_lat = (ref_out[0] if isinstance(ref_out, (tuple, list)) else ref_out).float().cpu()
        from inferix_b200.kvcache_manager import KVCacheManager as NKV, KVCacheRequest as NReq
        from inferix_b200.pipeline import CausalInferencePipeline as NPipe, DecodeMode as NDecode
        from inferix_b200.wan_model import CausalWanModel as NModel
        from inferix_b200.wrapper import WanDiffusionWrapper as NWrap
        nm = NModel(**cfg, local_attn_size=window_frames, sink_size=0)
        nm.load_state_dict(synthetic.synth_state_dict(cfg, seed=0))
        nm = nm.to(torch.bfloat16).to(dev)
        nargs = types.SimpleNamespace(denoising_step_list=steps, warp_denoising_step=True, num_frame_per_block=3,
                                      context_noise=0)
        npipe = NPipe(nargs, dev, generator=NWrap(model=nm, timestep_shift=5.0))
        npipe.renoise_fn = seeded(5)
        ours = npipe.inference(noise=noise, text_prompts=context, kv_cache_manager=NKV(dev),
                               kv_cache_requests=[NReq("ours")], decode_mode=NDecode.NO_DECODE,
                               free_cache_before_vae=False)
        ours_lat = (ours[0] if isinstance(ours, (tuple, list)) else ours).float().cpu()
        our_idx = (int(npipe.kv_cache_meta[0]["global_end_index"]), int(npipe.kv_cache_meta[0]["local_end_index"]))
        res = {"impl": "reference_gpu_parity",
               "what": "unmodified reference (its FA2 + cuBLAS GPU path) vs the native pipeline, same synthetic weights, "
                       "noise and re-noise stream, on this GPU",
               "shape": f"tiny widths {cfg['dim']}/{cfg['num_heads']} heads, {a.blocks} blocks x {len(steps)} steps + clean "
                        f"pass, window {window_frames} frames (eviction from block 3)",
               "rel_l2": ((ours_lat - ref_lat).norm() / ref_lat.norm()).item(),
               "max_abs": (ours_lat - ref_lat).abs().max().item(),
               "end_indices_reference": ref_idx, "end_indices_native": our_idx, "index_equal": ref_idx == our_idx,
               "finite": bool(torch.isfinite(ours_lat).all()), "gpu": torch.cuda.get_device_name(0)}
        line = json.dumps(res)
        print(line, flush=True)
        if a.out:
            Path(a.out).write_text(line + "\n")
        return
    if use_ev:
        torch.cuda.synchronize()
        ms = [s.elapsed_time(e) for s, e in events]
    else:
        ms = [(e - s) * 1e3 for s, e in wall]
    total_wall = time.perf_counter() - t0
    per_block = len(steps) + 1
    assert len(ms) == a.blocks * per_block, (len(ms), a.blocks, per_block)
    window_blocks = window_frames // 3
    steady = [ms[b * per_block:(b + 1) * per_block] for b in range(window_blocks, a.blocks)]   # blocks that evict
    assert steady, "not enough blocks to reach the steady state"
    t_first = sum(b[0] for b in steady) / len(steady)
    rest = [t for b in steady for t in b[1:]]
    t_rest = sum(rest) / len(rest)
    block_ms = t_first + a.timesteps * t_rest
    res = {
        "impl": "reference_gpu", "what": "unmodified reference CausalWanModel + CausalInferencePipeline + KVCacheManager"
                                         " on this GPU, bf16, synthetic weights",
        "attention": "SDPA (cpu dry run)" if a.tiny_cpu else f"flash-attn {__import__('flash_attn').__version__} "
                                                             "(flash_attn_varlen_func, the reference's FA2 branch)",
        "enable_kv_offload": bool(a.offload), "workload": "tiny_cpu" if a.tiny_cpu else "self_forcing_720p",
        "window_frames": window_frames, "tokens_per_frame": fs, "blocks_run": a.blocks,
        "forwards_run": len(ms), "ms_first_forward_of_block": t_first, "ms_other_forwards": t_rest,
        "steady_blocks_measured": len(steady), "timesteps": a.timesteps,
        "ms_per_block": block_ms, "value": 3.0 / (block_ms / 1e3), "unit": "latent frames/s",
        "method": "t_first + T * t_rest from per-forward CUDA events in the steady state (window full, evicting)",
        "wall_s_total": total_wall,
        "gpu": torch.cuda.get_device_name(0) if use_ev else "cpu",
    }
    line = json.dumps(res)
    print(line, flush=True)
    if a.out:
        Path(a.out).write_text(line + "\n")


if __name__ == "__main__":
    main()
