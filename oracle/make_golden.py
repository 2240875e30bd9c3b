"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

ORACLE tooling — runs only in the build container (the GPU box has no /root/reference); the produced fixtures
are committed.  Usage:  python oracle/make_golden.py

The reference imports diffusers / yunchang / xfuser / ftfy, none of which is installed, and picks a CUDA-only
flash-attn path whenever flash_attn is importable.  Following SURVEY §8c we (1) stub those modules in sys.modules
— they contribute no arithmetic on this path — and (2) force the reference's own SDPA fallback
(models/attention/flash_attention.py:185-199).  Everything else is the reference's code, executed as is:
CausalWanModel._forward_inference, CausalWanAttentionBlock, KVCacheManager, FlowMatchScheduler,
WanDiffusionWrapper.forward and CausalInferencePipeline.inference.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = os.environ.get("INFERIX_REFERENCE", "/root/reference")


def install_shims():
    class Permissive(types.ModuleType):
        """Any attribute the reference imports but never uses on this path resolves to a placeholder class."""

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

    def register_to_config(fn):
        return fn

    class ModelMixin(torch.nn.Module):
        pass

    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
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
        torch.cuda.current_device = lambda: 0   # evaluated at class-definition time in wan_base/text_encoder/t5.py:480


def import_reference():
    install_shims()
    sys.path.insert(0, REF)
    import inferix.models.attention  # noqa: F401
    fa_mod = sys.modules["inferix.models.attention.flash_attention"]
    fa_mod.HAS_FLASH_ATTN = False
    fa_mod.HAS_FLASH_ATTN_HOPPER = False
    ref_attention = fa_mod.attention   # with the flags off this is the reference's SDPA fallback

    def attention_q_dtype(q, k, v, **kw):
        kw.pop("k_lens", None)
        return ref_attention(q, k, v, dtype=q.dtype)   # default dtype=bf16 would break the fp32 run (:192-199)

    import inferix.models.attention as att_pkg
    att_pkg.flash_attention = attention_q_dtype        # cross-attn imports it at call time (wan_base/model.py:94)
    att_pkg.attention = attention_q_dtype
    from inferix.models.self_forcing import causal_model as cm
    cm.attention = attention_q_dtype
    return cm


def build_reference_model(cm, cfg: dict, sd, dtype, local_attn_size, sink_size):
    from inferix.models.wan_base import ParallelConfig
    pc = ParallelConfig.__new__(ParallelConfig)
    pc.ulysses_size = pc.ring_size = pc.world_size = 1
    pc.rank = pc.local_rank = 0
    pc.ring_strategy, pc.attn_backend = "pass-kv", "FlexAttention"
    model = cm.CausalWanModel(model_type="t2v", patch_size=(1, 2, 2), text_len=cfg["text_len"], in_dim=cfg["in_dim"],
                              dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"],
                              text_dim=cfg["text_dim"], out_dim=cfg["out_dim"], num_heads=cfg["num_heads"],
                              num_layers=cfg["num_layers"], local_attn_size=local_attn_size, sink_size=sink_size,
                              qk_norm=True, cross_attn_norm=True, eps=1e-6, enable_kv_offload=False,
                              parallel_config=pc)
    for blk in model.blocks:
        blk.self_attn.attention = cm.attention
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return model.to(dtype).eval(), pc


def reference_pipeline(cm, model, pc, cfg, frame_seq_length, steps, num_frame_per_block, shift):
    """CausalInferencePipeline around an already-built generator, without touching disk."""
    from inferix.models.self_forcing import wrapper as wr
    from inferix.models.schedulers.flow_match import FlowMatchScheduler
    from inferix.pipeline.self_forcing.CausalInferencePipeline import CausalInferencePipeline

    gen = wr.WanDiffusionWrapper.__new__(wr.WanDiffusionWrapper)
    torch.nn.Module.__init__(gen)
    gen.parallel_config, gen.enable_kv_offload = pc, False
    gen.model = model
    gen.uniform_timestep = False
    gen.scheduler = FlowMatchScheduler(shift=shift, sigma_min=0.0, extra_one_step=True)
    gen.scheduler.set_timesteps(1000, training=True)
    gen.seq_len = 32760

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": text_prompts}     # the "prompts" handed in below are the embeddings

    pipe = CausalInferencePipeline.__new__(CausalInferencePipeline)
    torch.nn.Module.__init__(pipe)
    pipe.parallel_config, pipe._profiler = pc, None
    pipe.generator, pipe.text_encoder, pipe.vae = gen, Text(), None
    pipe.scheduler = gen.scheduler
    sched_ts = torch.cat((gen.scheduler.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
    pipe.denoising_step_list = sched_ts[1000 - torch.tensor(steps, dtype=torch.long)]     # :86-90
    pipe.num_transformer_blocks = cfg["num_layers"]
    pipe.frame_seq_length = frame_seq_length
    pipe.kv_cache_meta = pipe.crossattn_cache_meta = None
    pipe.args = types.SimpleNamespace(context_noise=0)
    pipe.num_frame_per_block = num_frame_per_block
    pipe.independent_first_frame = False
    pipe.local_attn_size = model.local_attn_size
    model.num_frame_per_block = num_frame_per_block
    return pipe


def run_case(cm, name, cfg, dtype, frames, latent_hw, local_attn_size, sink_size, steps, shift=5.0, seed=1):
    from inferix.kvcache_manager.kvcache_manager import KVCacheManager, KVCacheRequest
    from inferix.core.types import DecodeMode
    from inferix_b200.synthetic import synth_state_dict

    sd = synth_state_dict(cfg, seed=0)
    model, pc = build_reference_model(cm, cfg, sd, dtype, local_attn_size, sink_size)
    fs = (latent_hw // 2) ** 2
    pipe = reference_pipeline(cm, model, pc, cfg, fs, steps, 3, shift)

    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(1, frames, 16, latent_hw, latent_hw, generator=g).to(dtype)      # fp32 draw, then cast
    context = torch.randn(1, 20, cfg["text_dim"], generator=g).to(dtype)

    # --- trace the cache indices and per-block activations of the very first forward
    taps, trace = {}, []
    first = {"done": False}
    hooks = []
    for i, blk in enumerate(model.blocks):
        def hook(_m, _inp, out, i=i):
            if not first["done"]:
                taps[f"block{i}"] = out.detach().clone()
        hooks.append(blk.register_forward_hook(hook))

    def head_hook(_m, _inp, _out):
        first["done"] = True
    hooks.append(model.head.register_forward_hook(head_hook))

    def trace_hook(_m, _inp, _out):
        meta = pipe.kv_cache_meta[0]
        trace.append((int(meta["global_end_index"].item()), int(meta["local_end_index"].item())))
    hooks.append(model.blocks[0].register_forward_hook(trace_hook))

    mgr = KVCacheManager("cpu")
    reqs = [KVCacheRequest("req_0")]
    blocks_out = []
    # Re-noising (CausalInferencePipeline.py:307) calls torch.randn_like on the latents.  randn draws different
    # numbers for bf16 and fp32 tensors, which would make the two dtype runs incomparable, so for the duration of
    # the run randn_like is bound to "draw fp32 from a seeded generator, cast to the tensor's dtype".  This changes
    # where the noise comes from, not a single arithmetic operation of the reference.
    regen = torch.Generator().manual_seed(1234)
    real_randn_like = torch.randn_like
    torch.randn_like = lambda x, **kw: torch.randn(x.shape, generator=regen, dtype=torch.float32).to(x.dtype)
    try:
        with torch.no_grad():
            out = pipe.inference(noise=noise, text_prompts=context, kv_cache_manager=mgr, kv_cache_requests=reqs,
                                 free_cache_before_vae=False, decode_mode=DecodeMode.NO_DECODE,
                                 block_callback=lambda lat, idx: blocks_out.append((idx, lat.clone())))
    finally:
        torch.randn_like = real_randn_like
    for h in hooks:
        h.remove()
    # (2, N, 1, H, D) -> valid prefix only; the full tensor is kept for the bf16 runs, a checksum for fp32
    cache_l1 = mgr.get_raw(reqs[0], f"layer_{cfg['num_layers'] - 1}")[:, :trace[-1][1]].clone()
    cache_sum = cache_l1.double().abs().sum(dim=(1, 2, 3, 4))
    if dtype == torch.float32:
        cache_l1 = None
    gold = dict(name=name, cfg=cfg, dtype=str(dtype), frames=frames, latent_hw=latent_hw,
                local_attn_size=local_attn_size, sink_size=sink_size, steps=steps, shift=shift,
                noise=noise, context=context, renoise_seed=1234, latents=out, taps=taps, index_trace=trace,
                callback_blocks=[i for i, _ in blocks_out], last_layer_cache=cache_l1,
                last_layer_cache_abs_sum=cache_sum,
                torch_version=torch.__version__)
    path = ROOT / "tests" / "golden" / f"{name}.pt"
    path.parent.mkdir(parents=True, exist_ok=True)
    torch.save(gold, path)
    print(f"wrote {path}  latents |x|={out.float().norm():.4f}  trace={trace[:3]}..{trace[-1]}  "
          f"size={path.stat().st_size / 1e3:.0f} kB")


def run_causvid_case(name, cfg, dtype, frames, latent_hw, steps, shift=8.0, seed=2, start_frames=0):
    """CausVid: reference models/causvid + pipeline/causvid, constructed without touching disk."""
    import torch.distributed  # noqa: F401
    from inferix.kvcache_manager.kvcache_manager import KVCacheManager, KVCacheRequest
    from inferix.models.causvid import causal_model as cvm
    from inferix.models.causvid import wrapper as cvw
    from inferix.models.schedulers.flow_match import FlowMatchScheduler
    from inferix.models.wan_base import ParallelConfig
    from inferix.pipeline.causvid.CausalInferencePipeline import CausalInferencePipeline as CVPipe
    from inferix_b200.synthetic import synth_state_dict
    import inferix.models.attention as att_pkg

    cvm.attention = att_pkg.attention                      # SDPA fallback with dtype=q.dtype (see import_reference)
    pc = ParallelConfig.__new__(ParallelConfig)
    pc.ulysses_size = pc.ring_size = pc.world_size = 1
    pc.rank = pc.local_rank = 0
    pc.ring_strategy, pc.attn_backend = "pass-kv", "FlexAttention"
    model = cvm.CausalWanModel(model_type="t2v", patch_size=(1, 2, 2), text_len=cfg["text_len"], in_dim=cfg["in_dim"],
                               dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"],
                               text_dim=cfg["text_dim"], out_dim=cfg["out_dim"], num_heads=cfg["num_heads"],
                               num_layers=cfg["num_layers"], qk_norm=True, cross_attn_norm=True, eps=1e-6,
                               enable_kv_offload=False, parallel_config=pc)
    for blk in model.blocks:
        blk.self_attn.attention = cvm.attention
    model.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
    model = model.to(dtype).eval()

    gen = cvw.WanDiffusionWrapper.__new__(cvw.WanDiffusionWrapper)
    torch.nn.Module.__init__(gen)
    gen.parallel_config, gen.enable_kv_offload, gen.model, gen.uniform_timestep = pc, False, model, False
    gen.scheduler = FlowMatchScheduler(shift=shift, sigma_min=0.0, extra_one_step=True)
    gen.scheduler.set_timesteps(1000, training=True)
    gen.seq_len = 32760

    class Text(torch.nn.Module):
        def forward(self, text_prompts):
            return {"prompt_embeds": text_prompts}

    class NoVae:
        def decode_to_pixel(self, x, **kw):
            return x

    fs = (latent_hw // 2) ** 2
    pipe = CVPipe.__new__(CVPipe)
    torch.nn.Module.__init__(pipe)
    pipe.parallel_config, pipe.generator, pipe.text_encoder, pipe.vae = pc, gen, Text(), NoVae()
    pipe.scheduler = gen.scheduler
    pipe.denoising_step_list = torch.tensor(steps, dtype=torch.long)[:-1]          # :37
    pipe.num_transformer_blocks = cfg["num_layers"]
    pipe.frame_seq_length = pipe.per_rank_frame_seq_length = fs
    pipe.is_kv_cache_initialized = False
    pipe.args = types.SimpleNamespace()
    pipe.num_frame_per_block = 3
    model.num_frame_per_block = 3

    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(1, frames, 16, latent_hw, latent_hw, generator=g).to(dtype)
    context = torch.randn(1, 20, cfg["text_dim"], generator=g).to(dtype)
    start = torch.randn(1, start_frames, 16, latent_hw, latent_hw, generator=g).to(dtype) if start_frames else None
    regen = torch.Generator().manual_seed(4321)
    real = torch.randn_like
    torch.randn_like = lambda x, **kw: torch.randn(x.shape, generator=regen, dtype=torch.float32).to(x.dtype)
    try:
        with torch.no_grad():
            _, out = pipe.inference(noise=noise, text_prompts=context, start_latents=start, return_latents=True,
                                    kv_cache_manager=KVCacheManager("cpu"), kv_cache_requests=[KVCacheRequest("req_0")])
    finally:
        torch.randn_like = real
    gold = dict(name=name, cfg=cfg, dtype=str(dtype), frames=frames, latent_hw=latent_hw, steps=steps, shift=shift,
                noise=noise, context=context, start_latents=start, renoise_seed=4321, latents=out,
                torch_version=torch.__version__)
    path = ROOT / "tests" / "golden" / f"{name}.pt"
    torch.save(gold, path)
    print(f"wrote {path}  latents |x|={out.float().norm():.4f}  size={path.stat().st_size / 1e3:.0f} kB")


def main():
    from inferix_b200.synthetic import TINY
    cm = import_reference()
    torch.set_num_threads(1)   # bit-reproducible reductions
    # BASELINE config 1: 2 layers, 1 block of 3 frames, 16x16 latent, 4 timesteps (fp32 and bf16)
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        run_case(cm, f"sf_tiny_1block_{tag}", TINY, dt, frames=3, latent_hw=16, local_attn_size=-1, sink_size=0,
                 steps=[1000, 750, 500, 250])
    # eviction: 4 blocks through a 6-frame window with a 1-frame sink (rolls at blocks 2 and 3)
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        run_case(cm, f"sf_tiny_evict_{tag}", TINY, dt, frames=12, latent_hw=16, local_attn_size=6, sink_size=1,
                 steps=[1000, 500])
    # window that is not a multiple of the block (7 frames), no sink
    run_case(cm, "sf_tiny_evict7_bf16", TINY, torch.bfloat16, frames=12, latent_hw=16, local_attn_size=7,
             sink_size=0, steps=[1000, 500])
    # CausVid: shipped step list [1000, 757, 522, 0] (causvid/default_config.yaml), 2 blocks, one given as start latent
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        run_causvid_case(f"causvid_tiny_{tag}", TINY, dt, frames=6, latent_hw=16, steps=[1000, 757, 522, 0])
    run_causvid_case("causvid_tiny_start_bf16", TINY, torch.bfloat16, frames=9, latent_hw=16,
                     steps=[1000, 757, 522, 0], start_frames=3)


if __name__ == "__main__":
    main()
