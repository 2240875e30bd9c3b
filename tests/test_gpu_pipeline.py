"""End-to-end parity of the native path against the reference's own outputs (tests/golden, produced by running
/root/reference on CPU) and against the oracle, through the reference-shaped Python surface.

Tolerance statement.  The path is bf16.  KV indices: bit-exact.  Latents:
  (a) relL2(ours_bf16, ref_fp32) <= 1.5 x relL2(ref_bf16, ref_fp32)   [the reference's own bf16-vs-fp32 gap], and
  (b) relL2(ours_bf16, ref_bf16) <= 5e-3 over the whole multi-block pipeline (measured 2.4e-3 .. 2.8e-3: the same
      distance two FlashAttention-2 runs with different key order sit apart; per-op parity at <= 1e-3 is in
      tests/test_gpu_kernels.py).
"""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest     # noqa: E402
from inferix_b200.pipeline import CausalInferencePipeline, DecodeMode        # noqa: E402
from inferix_b200.synthetic import synth_state_dict                          # noqa: E402
from inferix_b200.wan_model import CausalWanModel                            # noqa: E402
from inferix_b200.wrapper import WanDiffusionWrapper                         # noqa: E402
from oracle import wan_oracle as wo                                          # noqa: E402

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(cfg_d, local_attn_size, sink_size, steps, shift):
    model = CausalWanModel(**cfg_d, local_attn_size=local_attn_size, sink_size=sink_size)
    model.load_state_dict(synth_state_dict(cfg_d, seed=0))
    model = model.to(torch.bfloat16).to(DEV)
    gen = WanDiffusionWrapper(model=model, timestep_shift=shift)
    args = types.SimpleNamespace(denoising_step_list=steps, warp_denoising_step=True, num_frame_per_block=3,
                                 context_noise=0)
    return CausalInferencePipeline(args, DEV, generator=gen)


def run_ours(gold):
    pipe = build(gold["cfg"], gold["local_attn_size"], gold["sink_size"], gold["steps"], gold["shift"])
    cpu_gen = torch.Generator().manual_seed(gold["renoise_seed"])   # same draws as oracle/make_golden.py
    pipe.renoise_fn = lambda x: torch.randn(x.shape, generator=cpu_gen, dtype=torch.float32).to(x.dtype).to(x.device)
    mgr, reqs = KVCacheManager(DEV), [KVCacheRequest("req_0")]
    trace, blocks = [], []
    blk0 = pipe.generator.model.blocks[0]
    hook = blk0.register_forward_hook(lambda m, i, o: trace.append(pipe.kv_cache_meta[0]["_ifx_plan"]))
    out = pipe.inference(noise=gold["noise"].to(torch.bfloat16).to(DEV), text_prompts=gold["context"].to(torch.bfloat16).to(DEV),
                         kv_cache_manager=mgr, kv_cache_requests=reqs, free_cache_before_vae=False,
                         decode_mode=DecodeMode.NO_DECODE, block_callback=lambda lat, i: blocks.append(i))
    hook.remove()
    return pipe, mgr, reqs, out, trace, blocks


@pytest.mark.parametrize("case", ["sf_tiny_1block", "sf_tiny_evict"])
def test_pipeline_matches_reference(case, golden_dir):
    g16 = torch.load(golden_dir / f"{case}_bf16.pt", weights_only=False)
    g32 = torch.load(golden_dir / f"{case}_fp32.pt", weights_only=False)
    pipe, mgr, reqs, out, trace, blocks = run_ours(g16)
    # --- indices: bit-exact against the reference's (global_end, local_end) after every forward
    assert [(g, l) for (_, l, g, _) in trace] == g16["index_trace"]
    assert blocks == g16["callback_blocks"]
    meta = pipe.kv_cache_meta[0]
    assert (int(meta["global_end_index"]), int(meta["local_end_index"])) == g16["index_trace"][-1]
    # --- latents
    ref_gap = rel_l2(g16["latents"], g32["latents"])
    ours_vs_fp32 = rel_l2(out, g32["latents"])
    ours_vs_bf16 = rel_l2(out, g16["latents"])
    print(f"{case}: ref bf16-vs-fp32 {ref_gap:.3e}; ours-vs-fp32 {ours_vs_fp32:.3e}; ours-vs-ref-bf16 {ours_vs_bf16:.3e}")
    assert ours_vs_fp32 <= 1.5 * ref_gap
    assert ours_vs_bf16 <= 5e-3
    # --- last layer's cache in the reference's logical order
    last = pipe.generator.model.blocks[-1].kv_cache_manager
    kv = last.get_kv_cache(mgr, reqs[0])                        # (2, N, H, D)
    le = g16["index_trace"][-1][1]
    ref_kv = g16["last_layer_cache"][:, :, 0]                   # (2, le, H, D)
    # K/V are bf16 re-roundings of activations that already carry the ~3e-3 bf16-path noise, in a 2-layer net
    # with an un-trained (xavier) second layer that amplifies it: measured 6e-3 .. 8e-3, bounded at 1.5e-2.
    kv_err = rel_l2(kv[:, :le], ref_kv)
    print(f"{case}: last-layer cache rel-L2 vs reference {kv_err:.3e}")
    assert kv_err <= 1.5e-2


def test_pipeline_window_not_multiple_of_block(golden_dir):
    g16 = torch.load(golden_dir / "sf_tiny_evict7_bf16.pt", weights_only=False)
    _, _, _, out, trace, _ = run_ours(g16)
    assert [(g, l) for (_, l, g, _) in trace] == g16["index_trace"]
    assert rel_l2(out, g16["latents"]) <= 5e-3


def test_block_forward_matches_oracle():
    """One DiT block (the ifx_wan_block_forward call) against oracle.block_forward on the same inputs."""
    from inferix_b200.synthetic import TINY
    from inferix_b200 import ops
    cfg = wo.WanConfig(**TINY, local_attn_size=6, sink_size=0)
    sd = {k: v.bfloat16() for k, v in synth_state_dict(TINY, seed=0).items()}
    model = CausalWanModel(**TINY, local_attn_size=6, sink_size=0)
    model.load_state_dict(synth_state_dict(TINY, seed=0))
    model = model.to(torch.bfloat16).to(DEV)
    g = torch.Generator().manual_seed(3)
    frames, fs, C = 3, 64, TINY["dim"]
    x = torch.randn(1, frames * fs, C, generator=g).bfloat16()
    e0 = (torch.randn(1, frames, 6, C, generator=g) * 0.3).bfloat16()
    ctx = (torch.randn(1, 512, C, generator=g) * 0.5).bfloat16()
    grid = (frames, 8, 8)
    caches, cross = wo.new_cache(cfg, 6 * fs, 1, torch.bfloat16), [dict(is_init=False) for _ in range(2)]
    mgr, req = KVCacheManager(DEV), KVCacheRequest("r")
    blk = model.blocks[1]
    blk.kv_cache_manager.allocate_kv_cache(mgr, req, 6 * fs, torch.bfloat16, page_tokens=fs)
    blk.kv_cache_manager.allocate_crossattn_cache(mgr, req, 512, torch.bfloat16)
    meta = {"global_end_index": torch.zeros(1, dtype=torch.long, device=DEV),
            "local_end_index": torch.zeros(1, dtype=torch.long, device=DEV)}
    cmeta = {"is_init": False}
    table = ops.rope_table(model.freqs, DEV)
    for step, start in enumerate([0, 0, 3 * fs, 6 * fs]):     # repeat, advance, evict
        ref = wo.block_forward(sd, 1, cfg, x, e0, grid, wo.rope_freqs(128), ctx, caches[1], cross[1], start)
        out = blk(x.clone().to(DEV), e0.to(DEV), None, torch.tensor([grid]), table, ctx.to(DEV), None, None, meta, cmeta,
                  current_start=start, kv_cache_manager=mgr, kv_cache_requests=[req])
        assert rel_l2(out, ref) <= 5e-3, f"step {step}"
        assert (int(meta["global_end_index"]), int(meta["local_end_index"])) == (caches[1].global_end, caches[1].local_end)


@pytest.mark.parametrize("case", ["causvid_tiny", "causvid_tiny_start"])
def test_causvid_pipeline_matches_reference(case, golden_dir):
    """CausVid surface (explicit kv_start / kv_end, x0-only wrapper, steps[:-1]) against the reference's own run."""
    from inferix_b200.causvid import CausVidCausalWanModel, CausVidDiffusionWrapper, CausVidInferencePipeline
    g16 = torch.load(golden_dir / f"{case}_bf16.pt", weights_only=False)
    model = CausVidCausalWanModel(**g16["cfg"])
    model.load_state_dict(synth_state_dict(g16["cfg"], seed=0))
    model = model.to(torch.bfloat16).to(DEV)
    args = types.SimpleNamespace(denoising_step_list=g16["steps"], warp_denoising_step=False, num_frame_per_block=3)
    pipe = CausVidInferencePipeline(args, DEV, generator=CausVidDiffusionWrapper(model=model, timestep_shift=g16["shift"]))
    cpu_gen = torch.Generator().manual_seed(g16["renoise_seed"])
    pipe.renoise_fn = lambda x: torch.randn(x.shape, generator=cpu_gen, dtype=torch.float32).to(x.dtype).to(x.device)
    blocks = []
    start = g16["start_latents"]
    mgr, reqs = KVCacheManager(DEV), [KVCacheRequest("prompt text as id")]      # causvid keys requests by prompt
    _, out = pipe.inference(noise=g16["noise"].to(DEV), text_prompts=g16["context"].to(DEV),
                            start_latents=None if start is None else start.to(DEV), return_latents=True,
                            kv_cache_manager=mgr, kv_cache_requests=reqs,
                            block_callback=lambda lat, i: blocks.append(i))
    err = rel_l2(out, g16["latents"])
    print(f"{case}: ours-vs-ref-bf16 {err:.3e}")
    assert err <= 5e-3
    if start is None:
        g32 = torch.load(golden_dir / f"{case}_fp32.pt", weights_only=False)
        assert rel_l2(out, g32["latents"]) <= 1.5 * rel_l2(g16["latents"], g32["latents"])
        assert blocks == [0, 1]
    else:
        assert blocks == [1, 2]                                    # block 0 came from start_latents
    # second segment on the same pipeline object: caches are reset, result reproduces
    cpu_gen.manual_seed(g16["renoise_seed"])
    _, out2 = pipe.inference(noise=g16["noise"].to(DEV), text_prompts=g16["context"].to(DEV),
                             start_latents=None if start is None else start.to(DEV), return_latents=True,
                             kv_cache_manager=mgr, kv_cache_requests=reqs)
    assert torch.equal(out2, out)


def _tiny_pipe(local_attn_size=6, steps=(1000, 500)):
    return build(dict(__import__("inferix_b200.synthetic", fromlist=["TINY"]).TINY), local_attn_size, 0, list(steps), 5.0)


def test_video_extension_and_cache_reuse():
    """initial_latent (video extension, reference :213-253): the given blocks are copied to the output and cached with
    a t=0 forward; a second inference() on the same pipeline resets the native block tables and reproduces."""
    pipe = _tiny_pipe()
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(1, 6, 16, 16, 16, generator=g).bfloat16().to(DEV)
    init = torch.randn(1, 3, 16, 16, 16, generator=g).bfloat16().to(DEV)
    ctx = torch.randn(1, 20, 64, generator=g).bfloat16().to(DEV)
    mgr, reqs = KVCacheManager(DEV), [KVCacheRequest("r0")]
    outs = []
    for _ in range(2):
        gen = torch.Generator().manual_seed(5)
        pipe.renoise_fn = lambda x: torch.randn(x.shape, generator=gen, dtype=torch.float32).to(x.dtype).to(x.device)
        out = pipe.inference(noise=noise, text_prompts=ctx, kv_cache_manager=mgr, kv_cache_requests=reqs,
                             initial_latent=init, free_cache_before_vae=False, decode_mode=DecodeMode.NO_DECODE)
        outs.append(out)
        assert out.shape == (1, 9, 16, 16, 16) and torch.equal(out[:, :3], init)
        assert int(pipe.kv_cache_meta[0]["global_end_index"]) == 9 * 64
        assert int(pipe.kv_cache_meta[0]["local_end_index"]) == 6 * 64           # 6-frame window
    assert torch.equal(outs[0], outs[1])
    # free_cache_before_vae=True releases every layer and the meta lists (reference :398-400, :494-502)
    pipe.inference(noise=noise, text_prompts=ctx, kv_cache_manager=mgr, kv_cache_requests=reqs,
                   decode_mode=DecodeMode.NO_DECODE)
    assert pipe.kv_cache_meta is None and list(mgr.layers(reqs[0])) == []


def test_two_requests_batch_equals_two_single_runs():
    """kv_cache_requests is the batch dimension (reference :422-429): B = 2 equals two B = 1 runs."""
    g = torch.Generator().manual_seed(10)
    noise = torch.randn(2, 3, 16, 16, 16, generator=g).bfloat16().to(DEV)
    ctx = torch.randn(2, 20, 64, generator=g).bfloat16().to(DEV)

    def run(nz, cx, names):
        pipe = _tiny_pipe(steps=(1000,))       # single step: no re-noising, so batching cannot change the RNG stream
        return pipe.inference(noise=nz, text_prompts=cx, kv_cache_manager=KVCacheManager(DEV),
                              kv_cache_requests=[KVCacheRequest(n) for n in names], decode_mode=DecodeMode.NO_DECODE)
    both = run(noise, ctx, ["a", "b"])
    one_a, one_b = run(noise[:1], ctx[:1], ["a"]), run(noise[1:], ctx[1:], ["b"])
    assert rel_l2(both[0], one_a[0]) <= 1e-3 and rel_l2(both[1], one_b[0]) <= 1e-3


def test_decode_requires_vae_and_reports_block_times():
    pipe = _tiny_pipe(steps=(1000,))
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(1, 3, 16, 16, 16, generator=g).bfloat16().to(DEV)
    ctx = torch.randn(1, 20, 64, generator=g).bfloat16().to(DEV)
    with pytest.raises(RuntimeError, match="VAE"):
        pipe.inference(noise=noise, text_prompts=ctx, kv_cache_manager=KVCacheManager(DEV),
                       kv_cache_requests=[KVCacheRequest("r")])
    pipe2 = _tiny_pipe(steps=(1000,))
    pipe2.inference(noise=noise, text_prompts=ctx, kv_cache_manager=KVCacheManager(DEV),
                    kv_cache_requests=[KVCacheRequest("r")], decode_mode=DecodeMode.NO_DECODE, profile=True)
    assert len(pipe2.last_block_times_ms) == 1 and pipe2.last_block_times_ms[0] > 0


def test_wrong_dtype_and_device_fail_loudly():
    from inferix_b200 import ops
    with pytest.raises(ValueError):
        ops.gemm(torch.zeros(8, 8), torch.zeros(8, 8))                               # CPU tensors
    with pytest.raises(ValueError):
        ops.gemm(torch.zeros(8, 8, device=DEV), torch.zeros(8, 8, device=DEV))       # fp32
    with pytest.raises(ValueError):
        KVCacheManager("cpu")
    with pytest.raises(NotImplementedError):
        CausalWanModel(dim=512, num_heads=8)                                         # head_dim 64


def test_long_horizon_256_blocks_eviction_steady_state():
    """BASELINE config 5 (long-horizon eviction stress) at tiny widths on one GPU: 256 blocks = 768 latent frames
    (< the 1024-row RoPE table) through an 8-block window with one sink frame and the shipped 4-step schedule.
    Steady state from block 8 on: index trace == the oracle's arithmetic (causal_model.py:277-300) for all 2560
    forwards, allocated memory does not grow, per-block time stays flat, latents stay finite."""
    nblk, fpb, fs, window, sink = 256, 3, 64, 24, 1
    from inferix_b200.synthetic import TINY
    pipe = build(dict(TINY), window, sink, [1000, 750, 500, 250], 5.0)
    g = torch.Generator().manual_seed(21)
    noise = torch.randn(1, nblk * fpb, 16, 16, 16, generator=g).bfloat16().to(DEV)
    ctx = torch.randn(1, 20, TINY["text_dim"], generator=g).bfloat16().to(DEV)
    trace, mem = [], []
    blk0 = pipe.generator.model.blocks[0]
    hook = blk0.register_forward_hook(lambda m, i, o: trace.append(pipe.kv_cache_meta[0]["_ifx_plan"]))
    out = pipe.inference(noise=noise, text_prompts=ctx, kv_cache_manager=KVCacheManager(DEV),
                         kv_cache_requests=[KVCacheRequest("long")], decode_mode=DecodeMode.NO_DECODE, profile=True,
                         free_cache_before_vae=False,
                         block_callback=lambda lat, i: mem.append(torch.cuda.memory_allocated()))
    hook.remove()
    assert out.shape == noise.shape and bool(torch.isfinite(out.float()).all())
    # --- indices, bit-exact against the oracle's restatement, every forward of every block
    want, ge, le = [], 0, 0
    for b in range(nblk):
        for _ in range(5):                                                     # 4 noisy + 1 clean forward
            ls, le, ge, ev = wo.plan_indices(window * fs, ge, le, b * fpb * fs, fpb * fs, sink * fs, True)
            want.append((ls, le, ge, ev))
    assert trace == want
    assert trace[-1][1] == window * fs and trace[-1][2] == nblk * fpb * fs
    assert sum(t[3] for t in trace) == (nblk - 8) * fpb * fs                   # every block past the window evicts once
    # --- no growth in allocated memory once the window is full
    assert max(mem[8:]) == min(mem[8:]), (min(mem[8:]), max(mem[8:]))
    # --- flat per-block time in the steady state (median of the last 64 blocks vs blocks 8..72)
    t = pipe.last_block_times_ms
    early, late = sorted(t[8:72])[32], sorted(t[-64:])[32]
    print(f"long horizon: block time early {early:.2f} ms, late {late:.2f} ms")
    assert late <= 1.5 * early            # medians of 64 blocks; the margin covers host-side noise on a shared box


def test_kv_offload_tier_matches_hbm_resident():
    """kv_offload honoured (KVCacheManager(offload_tier=True)): layer windows live in pinned host memory and are staged
    through two device slots (layer i + 1 copied in while layer i computes, written pages copied back).  The pipeline
    result is bit-identical to the HBM-resident run — 4 blocks through a 6-frame window, eviction included — and the
    device holds two layers' worth of cache instead of one per layer."""
    import warnings
    from inferix_b200.synthetic import TINY
    cfg = dict(TINY, num_layers=4)
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(1, 12, 16, 16, 16, generator=g).bfloat16().to(DEV)
    context = torch.randn(1, 20, cfg["text_dim"], generator=g).bfloat16().to(DEV)

    def run(offload_tier):
        model = CausalWanModel(**cfg, local_attn_size=6, sink_size=1, enable_kv_offload=True)
        model.load_state_dict(synth_state_dict(cfg, seed=0))
        model = model.to(torch.bfloat16).to(DEV)
        args = types.SimpleNamespace(denoising_step_list=[1000, 500], warp_denoising_step=True, num_frame_per_block=3,
                                     context_noise=0)
        pipe = CausalInferencePipeline(args, DEV, generator=WanDiffusionWrapper(model=model, timestep_shift=5.0))
        rg = torch.Generator().manual_seed(5)
        pipe.renoise_fn = lambda x: torch.randn(x.shape, generator=rg, dtype=torch.float32).to(x.dtype).to(x.device)
        mgr = KVCacheManager(DEV, offload_tier=offload_tier)
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            out = pipe.inference(noise=noise, text_prompts=context, kv_cache_manager=mgr,
                                 kv_cache_requests=[KVCacheRequest("r")], decode_mode=DecodeMode.NO_DECODE,
                                 free_cache_before_vae=False)
        stores = [blk.kv_cache_manager.store(mgr, KVCacheRequest("r")) for blk in model.blocks]
        return out, stores, wlist

    out_hbm, stores_hbm, wl = run(False)
    assert any("stays in HBM" in str(w.message) for w in wl), "an ignored kv_offload request must be announced"
    out_off, stores_off, _ = run(True)
    torch.cuda.synchronize()
    assert torch.equal(out_off, out_hbm)
    slots = stores_off[0].offload
    assert slots is not None and slots.count == 2 and all(s.offload is slots for s in stores_off)
    assert len({s.k.data_ptr() for s in stores_off}) <= 2                      # two device slots for four layers
    for a, b in zip(stores_off, stores_hbm):                                   # host tier == HBM-resident cache contents
        _, local_end, _ = a.state()
        ka, va = a.export(0, local_end)
        kb, vb = b.export(0, local_end)
        assert torch.equal(ka, kb) and torch.equal(va, vb)


def test_cfg_diffusion_pipeline_matches_oracle():
    """CausalDiffusionInferencePipeline (many-step sampler + classifier-free guidance, two cache sets): 2 blocks x 6
    UniPC steps x 2 branches + clean re-runs = 28 native forwards against the oracle's restatement of the reference
    loop with the same (reference-pinned) sampler.  CFG (scale 3) amplifies the single-forward bf16 distance."""
    from inferix_b200.diffusion_pipeline import CausalDiffusionInferencePipeline
    from inferix_b200.synthetic import TINY
    from inferix_b200.unipc import FlowUniPCMultistepScheduler
    cfg_d = dict(TINY)
    g = torch.Generator().manual_seed(31)
    noise = torch.randn(1, 6, 16, 16, 16, generator=g).bfloat16()
    context = torch.randn(1, 20, cfg_d["text_dim"], generator=g).bfloat16()
    neg = torch.randn(1, 12, cfg_d["text_dim"], generator=g).bfloat16()
    model = CausalWanModel(**cfg_d, local_attn_size=-1, sink_size=0)
    model.load_state_dict(synth_state_dict(cfg_d, seed=0))
    model = model.to(torch.bfloat16).to(DEV)
    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, num_frame_per_block=3, guidance_scale=3.0,
                                 negative_prompt_embeds=neg.to(DEV), sampling_steps=6, kv_cache_frames=6)
    pipe = CausalDiffusionInferencePipeline(args, DEV, generator=WanDiffusionWrapper(model=model, timestep_shift=5.0))
    _, out = pipe.inference(noise.to(DEV), context.to(DEV), KVCacheManager(DEV), KVCacheManager(DEV),
                            [KVCacheRequest("r")], return_latents=True)
    cfg = wo.WanConfig(**cfg_d)
    sd = {k: v.bfloat16() for k, v in synth_state_dict(cfg_d, seed=0).items()}
    sched = wo.FlowMatchSigmas(shift=5.0)

    def factory():
        s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        s.set_timesteps(6, device="cpu", shift=5.0)
        return s
    ref, _ = wo.cfg_pipeline_inference(sd, cfg, sched, noise, context, neg, factory, 3.0, 3, 64, 6 * 64)
    err = rel_l2(out, ref)
    print(f"CFG diffusion pipeline (2 blocks, 6 UniPC steps, guidance 3): rel-L2 vs oracle {err:.3e}")
    assert bool(torch.isfinite(out).all()) and err <= 2e-2
