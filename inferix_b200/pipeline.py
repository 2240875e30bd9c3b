"""Semi-autoregressive block scheduler: CausalInferencePipeline with the reference's surface
(inferix/pipeline/self_forcing/CausalInferencePipeline.py:57-502).

Loop semantics are the reference's: for each block of ``num_frame_per_block`` latent frames, T noisy forwards with
re-noising to the next timestep in between (:276-310), the block's x0 written to the output (:346), then one clean
forward at ``context_noise`` that rewrites the block's K/V (:352-361) and the ``block_callback`` hook (:390-393).

What changed underneath: ``frame_seq_length`` and the cache size follow the latent shape (the reference hard-codes
480x832 -> 1560 / 32760, :92-93,457); the cache is allocated with frame-sized pages; resets go to the native block
table; nothing in the loop synchronises with the device unless profiling is on.
"""
from __future__ import annotations

from contextlib import contextmanager
from enum import Enum
from typing import Callable, List, Optional, Union

import torch

from .kvcache_manager import KVCacheManager, KVCacheRequest
from .parallel import ParallelConfig
from .wrapper import WanDiffusionWrapper


class DecodeMode(Enum):
    """inferix/core/types/inference.py:11-15."""
    AFTER_ALL = "after_all"
    PER_BLOCK = "per_block"
    NO_DECODE = "no_decode"


class PerformanceProfiler:
    """CUDA-event stage timer (reference :13-54)."""

    def __init__(self, enabled: bool = False):
        self.enabled = enabled
        self.events = {}

    def __bool__(self):
        return self.enabled

    @contextmanager
    def stage(self, name: str):
        if not self.enabled:
            yield
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        try:
            yield
        finally:
            e.record()
            torch.cuda.synchronize()
            self.events[name] = s.elapsed_time(e)

    def record_block_time(self, block_index: int, time_ms: float):
        if self.enabled:
            self.events[f"block_{block_index}"] = time_ms

    def get_results(self):
        return self.events if self.enabled else {}


class _IdentityTextEncoder(torch.nn.Module):
    """Text encoding (umT5) is out of scope: prompts handed to ``inference`` are already embeddings."""

    def forward(self, text_prompts):
        return {"prompt_embeds": text_prompts}


class CausalInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator: Optional[WanDiffusionWrapper] = None, text_encoder=None, vae=None,
                 parallel_config: Optional[ParallelConfig] = None, profiler=None):
        super().__init__()
        self.parallel_config = parallel_config if parallel_config is not None else ParallelConfig()
        self._profiler = profiler
        self.device = torch.device(device)
        self.generator = WanDiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True,
                                             parallel_config=self.parallel_config) if generator is None else generator
        self.text_encoder = _IdentityTextEncoder() if text_encoder is None else text_encoder
        self.vae = vae

        self.scheduler = self.generator.get_scheduler()
        self.denoising_step_list = torch.tensor(args.denoising_step_list, dtype=torch.long)
        if getattr(args, "warp_denoising_step", False):      # reference :86-90
            timesteps = torch.cat((self.scheduler.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
            self.denoising_step_list = timesteps[1000 - self.denoising_step_list]

        self.num_transformer_blocks = self.generator.model.num_layers
        self.frame_seq_length = getattr(args, "frame_seq_length", None)   # None: derived from the noise shape
        self.kv_cache_meta = None
        self.crossattn_cache_meta = None
        self.args = args
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.independent_first_frame = getattr(args, "independent_first_frame", False)
        if self.independent_first_frame:
            raise NotImplementedError("independent_first_frame (I2V start) is outside the T2V hot path")
        self.local_attn_size = self.generator.model.local_attn_size
        if self.num_frame_per_block > 1:
            self.generator.model.num_frame_per_block = self.num_frame_per_block
        # re-noising source; tests inject a seeded CPU generator to follow the oracle draw for draw
        self.renoise_fn: Callable[[torch.Tensor], torch.Tensor] = torch.randn_like
        self.last_block_times_ms: List[float] = []

    # ------------------------------------------------------------------------------------------ inference
    @torch.no_grad()   # inference only; the reference disables autograd globally (self_forcing/pipeline.py:62)
    def inference(self, noise: torch.Tensor, text_prompts, kv_cache_manager: KVCacheManager,
                  kv_cache_requests: List[KVCacheRequest], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, profile: bool = False, low_memory: bool = False,
                  free_cache_before_vae: bool = True, decode_mode: DecodeMode = DecodeMode.AFTER_ALL,
                  vae_chunk_size: Optional[int] = None, block_callback: Optional[callable] = None,
                  vae_decode_context=None, callback_stream: Optional["torch.cuda.Stream"] = None
                  ) -> Union[torch.Tensor, tuple]:
        """Reference signature (:108-131) plus `callback_stream`: when given, `block_callback` — the
        PER_BLOCK VAE decode in the reference's streaming path (self_forcing/pipeline.py:677-699) — is issued on that
        stream after an event that marks the block's latent final, so the decode of block i overlaps the denoising
        of block i+1; the main stream re-joins it before `inference` returns."""
        perf = PerformanceProfiler(enabled=profile)
        batch_size, num_frames, num_channels, height, width = noise.shape
        assert num_frames % self.num_frame_per_block == 0
        num_blocks = num_frames // self.num_frame_per_block
        num_input_frames = initial_latent.shape[1] if initial_latent is not None else 0
        num_output_frames = num_frames + num_input_frames
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        ps = self.generator.model.patch_size
        fs = (height // ps[1]) * (width // ps[2])
        if self.frame_seq_length is None:
            self.frame_seq_length = fs
        elif self.frame_seq_length != fs:
            raise ValueError(f"frame_seq_length={self.frame_seq_length} does not match the latent shape ({fs})")

        output = torch.zeros([batch_size, num_output_frames, num_channels, height, width], device=noise.device,
                             dtype=noise.dtype)

        with perf.stage("initialization"):
            if self.kv_cache_meta is None:
                self._initialize_kv_cache(kv_cache_manager, kv_cache_requests, dtype=noise.dtype)
            else:
                self._reset_kv_cache(kv_cache_manager, kv_cache_requests)
            if self.crossattn_cache_meta is None:
                self._initialize_crossattn_cache(kv_cache_manager, kv_cache_requests, dtype=noise.dtype)
            else:
                for i in range(self.num_transformer_blocks):
                    self.crossattn_cache_meta[i]["is_init"] = False

            common = dict(conditional_dict=conditional_dict, kv_cache_meta=self.kv_cache_meta,
                          crossattn_cache_meta=self.crossattn_cache_meta, kv_cache_manager=kv_cache_manager,
                          kv_cache_requests=kv_cache_requests)
            current_start_frame = 0
            if initial_latent is not None:       # video extension: cache the given context blocks (:213-253)
                timestep = torch.zeros([batch_size, 1], device=noise.device, dtype=torch.int64)
                assert num_input_frames % self.num_frame_per_block == 0
                for _ in range(num_input_frames // self.num_frame_per_block):
                    ref = initial_latent[:, current_start_frame:current_start_frame + self.num_frame_per_block]
                    output[:, current_start_frame:current_start_frame + self.num_frame_per_block] = ref
                    self.generator(noisy_image_or_video=ref,
                                   timestep=(timestep * 0).expand(batch_size, self.num_frame_per_block),
                                   current_start=current_start_frame * self.frame_seq_length, **common)
                    current_start_frame += self.num_frame_per_block

        with perf.stage("diffusion_generation"):
            block_times = []
            for block_index in range(num_blocks):
                n = self.num_frame_per_block
                if profile:
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record()
                noisy_input = noise[:, current_start_frame - num_input_frames:current_start_frame + n - num_input_frames]
                denoised_pred = self.denoise_block(noisy_input, current_start_frame, common)
                output[:, current_start_frame:current_start_frame + n] = denoised_pred
                if profile:
                    t1.record()
                    torch.cuda.synchronize()
                    block_times.append(t0.elapsed_time(t1))
                    perf.record_block_time(block_index, block_times[-1])
                    if self._profiler is not None and hasattr(self._profiler, "record_block_computation"):
                        try:
                            self._profiler.record_block_computation(
                                block_index=block_index, block_size=n, computation_time_ms=block_times[-1],
                                memory_usage_mb=torch.cuda.max_memory_allocated() / (1024 * 1024))
                        except Exception:
                            pass
                current_start_frame += n
                if block_callback is not None:
                    block_latent = output[:, current_start_frame - n:current_start_frame]
                    if callback_stream is None:
                        block_callback(block_latent, block_index)
                    else:
                        final = torch.cuda.Event()
                        final.record()
                        with torch.cuda.stream(callback_stream):
                            callback_stream.wait_event(final)
                            block_callback(block_latent, block_index)
            if block_callback is not None and callback_stream is not None:
                torch.cuda.current_stream().wait_stream(callback_stream)
            self.last_block_times_ms = block_times

        if free_cache_before_vae:
            self.clear_cache(kv_cache_manager, kv_cache_requests)

        if decode_mode == DecodeMode.NO_DECODE:
            return (output, output) if return_latents else output
        if self.vae is None:
            raise RuntimeError("no VAE attached: VAE decode is out of scope of inferix_b200, pass "
                               "decode_mode=DecodeMode.NO_DECODE or construct the pipeline with vae=...")
        with perf.stage("vae_decoding"):
            chunk = vae_chunk_size if vae_chunk_size is not None else 2
            if vae_decode_context is not None:
                with vae_decode_context:
                    video = self.vae.decode_to_pixel(output, use_cache=True, chunk_size=chunk)
            else:
                video = self.vae.decode_to_pixel(output, use_cache=True, chunk_size=chunk)
            video = (video * 0.5 + 0.5).clamp(0, 1)
        return (video, output) if return_latents else video

    @torch.no_grad()   # inference only; the reference disables autograd globally (self_forcing/pipeline.py:62)
    def denoise_block(self, noisy_input: torch.Tensor, current_start_frame: int, common: dict) -> torch.Tensor:
        """One unit of the metric: T noisy forwards with re-noising in between (:276-310), then the clean re-run that
        rewrites the block's K/V (:352-361).  noisy_input [B, n, C, H, W] -> denoised x0 of the block."""
        batch_size, n = noisy_input.shape[:2]
        steps = self.denoising_step_list
        device = noisy_input.device
        start = current_start_frame * self.frame_seq_length
        denoised_pred = timestep = None
        for index, current_timestep in enumerate(steps):
            timestep = torch.ones([batch_size, n], device=device, dtype=torch.int64) * current_timestep
            _, denoised_pred = self.generator(noisy_image_or_video=noisy_input, timestep=timestep, current_start=start,
                                              **common)
            if index < len(steps) - 1:
                next_timestep = steps[index + 1]
                flat = denoised_pred.flatten(0, 1)
                noisy_input = self.scheduler.add_noise(
                    flat, self.renoise_fn(flat),
                    next_timestep * torch.ones([batch_size * n], device=device, dtype=torch.long)
                ).unflatten(0, denoised_pred.shape[:2])
            if self._profiler is not None and hasattr(self._profiler, "record_diffusion_step"):
                try:
                    self._profiler.record_diffusion_step(step=index, timestep=float(current_timestep) / 1000.0,
                                                         block_size=n, computation_time_ms=0.0,
                                                         guidance_scale=getattr(self.args, "guidance_scale", None))
                except Exception:
                    pass
        context_timestep = torch.ones_like(timestep) * getattr(self.args, "context_noise", 0)
        self.generator(noisy_image_or_video=denoised_pred, timestep=context_timestep, current_start=start, **common)
        return denoised_pred

    # ------------------------------------------------------------------------------------------ caches
    def _kv_cache_size(self) -> int:
        if self.local_attn_size != -1:
            return self.local_attn_size * self.frame_seq_length          # reference :453-455
        # reference default 32760 = 21 frames x 1560; same 21-frame horizon at any resolution
        return getattr(self.args, "kv_cache_frames", 21) * self.frame_seq_length

    def _initialize_kv_cache(self, kv_cache_manager, kv_cache_requests, dtype):
        """reference :444-472; one paged cache per layer x request, pages = latent frames."""
        kv_cache_size = self._kv_cache_size()
        ulysses = self.parallel_config.ulysses_size
        ring = self.parallel_config.ring_size
        for layer_idx in range(self.num_transformer_blocks):
            adapter = self.generator.model.blocks[layer_idx].kv_cache_manager
            for req in kv_cache_requests:
                adapter.allocate_kv_cache(kv_cache_manager=kv_cache_manager, kv_cache_request=req,
                                          sequence_length=kv_cache_size, dtype=dtype, ulysses_size=ulysses,
                                          ring_size=ring, page_tokens=self.frame_seq_length)
        from . import peer
        self._peer_group = peer.setup_for_pipeline(self.generator.model, kv_cache_manager, kv_cache_requests,
                                                   self.parallel_config)
        device = kv_cache_manager.device
        # all layers' end indices are views of one tensor
        idx = torch.zeros((self.num_transformer_blocks, 2), dtype=torch.long, device=device)
        self.kv_cache_meta = [{"global_end_index": idx[i, 0:1], "local_end_index": idx[i, 1:2], "_ifx_shared": idx}
                              for i in range(self.num_transformer_blocks)]

    def _reset_kv_cache(self, kv_cache_manager, kv_cache_requests):
        """reference :193-199 (assigns fresh zero tensors); here also resets the native block tables."""
        for layer_idx in range(self.num_transformer_blocks):
            adapter = self.generator.model.blocks[layer_idx].kv_cache_manager
            for req in kv_cache_requests:
                adapter.reset_kv_cache(kv_cache_manager, req, kv_cache_manager.device)
            self.kv_cache_meta[layer_idx]["global_end_index"].zero_()
            self.kv_cache_meta[layer_idx]["local_end_index"].zero_()

    def _initialize_crossattn_cache(self, kv_cache_manager, kv_cache_requests, dtype):
        """reference :474-492."""
        text_len = self.generator.model.text_len
        for layer_idx in range(self.num_transformer_blocks):
            adapter = self.generator.model.blocks[layer_idx].kv_cache_manager
            for req in kv_cache_requests:
                adapter.allocate_crossattn_cache(kv_cache_manager=kv_cache_manager, kv_cache_request=req,
                                                 crossattn_length=text_len, dtype=dtype)
        self.crossattn_cache_meta = [{"is_init": False} for _ in range(self.num_transformer_blocks)]

    def clear_cache(self, kv_cache_manager, kv_cache_requests):
        """reference :494-502."""
        if getattr(self, "_peer_group", None) is not None:      # unmap the other ranks' caches before anyone frees
            self._peer_group.release()
            torch.distributed.barrier(group=self.parallel_config.group)
            self._peer_group = None
        for layer_idx in range(self.num_transformer_blocks):
            for req in kv_cache_requests:
                self.generator.model.blocks[layer_idx].kv_cache_manager.clear_cache(
                    kv_cache_manager=kv_cache_manager, kv_cache_request=req)
        self.kv_cache_meta = None
        self.crossattn_cache_meta = None
