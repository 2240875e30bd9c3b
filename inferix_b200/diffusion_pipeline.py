"""CausalDiffusionInferencePipeline — the many-step classifier-free-guidance scheduler of the reference
(inferix/pipeline/self_forcing/CausalDiffusionInferencePipeline.py:10-385): per block, `sampling_steps` UniPC steps of
TWO DiT forwards each (conditional and unconditional prompt, each with its own KV / cross-attention caches), the CFG
combination  flow = uncond + guidance_scale * (cond - uncond),  a FlowUniPCMultistepScheduler step, then a clean
(timestep 0) re-run of both branches that rewrites the block's K / V.

Same constructor / `inference` signature as the reference (two KVCacheManagers, one per branch).  The forwards are
the native ones (ifx_wan_block_forward per layer); the scheduler arithmetic is `inferix_b200.unipc`, pinned bit-exact
to the reference class.  Text encoder / VAE are out of scope: `text_prompts` are prompt embeddings (as everywhere in
this package) and the negative prompt comes as `args.negative_prompt_embeds`; decoding needs a caller-supplied vae.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .kvcache_manager import KVCacheManager, KVCacheRequest
from .parallel import ParallelConfig
from .unipc import FlowUniPCMultistepScheduler
from .wrapper import WanDiffusionWrapper


class _EmbeddingPassthrough(torch.nn.Module):
    def forward(self, text_prompts):
        return {"prompt_embeds": text_prompts}


class CausalDiffusionInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator: Optional[WanDiffusionWrapper] = None, text_encoder=None, vae=None,
                 parallel_config: Optional[ParallelConfig] = None):
        super().__init__()
        if generator is None:
            raise ValueError("inferix_b200 does no checkpoint I/O: pass generator=WanDiffusionWrapper(model=...)")
        self.generator = generator
        self.text_encoder = text_encoder if text_encoder is not None else _EmbeddingPassthrough()
        self.vae = vae
        self.parallel_config = parallel_config if parallel_config is not None else ParallelConfig()
        self.num_train_timesteps = getattr(args, "num_train_timestep", 1000)
        self.sampling_steps = getattr(args, "sampling_steps", 50)                    # reference: fixed 50 (:33)
        self.sample_solver = "unipc"
        self.shift = args.timestep_shift
        self.num_transformer_blocks = self.generator.model.num_layers
        self.frame_seq_length = None                                                 # reference: hard-wired 1560 (:38)
        self.args = args
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.independent_first_frame = getattr(args, "independent_first_frame", False)
        if self.independent_first_frame:
            raise NotImplementedError("independent_first_frame (the [1, 4, 4, ...] i2v schedule) is not built")
        self.local_attn_size = self.generator.model.local_attn_size
        if self.num_frame_per_block > 1:
            self.generator.model.num_frame_per_block = self.num_frame_per_block
        self.kv_cache_meta_pos = self.kv_cache_meta_neg = None
        self.crossattn_cache_meta_pos = self.crossattn_cache_meta_neg = None

    # ------------------------------------------------------------------------------------------ caches (:298-362)
    def _kv_cache_size(self) -> int:
        if self.local_attn_size != -1:
            return self.local_attn_size * self.frame_seq_length
        return getattr(self.args, "kv_cache_frames", 21) * self.frame_seq_length     # reference default 32760 = 21 x 1560

    def _initialize_kv_cache(self, kv_cache_manager_pos, kv_cache_manager_neg, kv_cache_requests, dtype):
        size = self._kv_cache_size()
        metas = []
        for mgr in (kv_cache_manager_pos, kv_cache_manager_neg):
            for layer_idx in range(self.num_transformer_blocks):
                adapter = self.generator.model.blocks[layer_idx].kv_cache_manager
                for req in kv_cache_requests:
                    adapter.allocate_kv_cache(kv_cache_manager=mgr, kv_cache_request=req, sequence_length=size, dtype=dtype,
                                              ulysses_size=self.parallel_config.ulysses_size,
                                              ring_size=self.parallel_config.ring_size, page_tokens=self.frame_seq_length)
            idx = torch.zeros((self.num_transformer_blocks, 2), dtype=torch.long, device=mgr.device)
            metas.append([{"global_end_index": idx[i, 0:1], "local_end_index": idx[i, 1:2], "_ifx_shared": idx}
                          for i in range(self.num_transformer_blocks)])
        self.kv_cache_meta_pos, self.kv_cache_meta_neg = metas

    def _initialize_crossattn_cache(self, kv_cache_manager_pos, kv_cache_manager_neg, kv_cache_requests, dtype):
        text_len = self.generator.model.text_len
        for mgr in (kv_cache_manager_pos, kv_cache_manager_neg):
            for layer_idx in range(self.num_transformer_blocks):
                adapter = self.generator.model.blocks[layer_idx].kv_cache_manager
                for req in kv_cache_requests:
                    adapter.allocate_crossattn_cache(kv_cache_manager=mgr, kv_cache_request=req, crossattn_length=text_len,
                                                     dtype=dtype)
        self.crossattn_cache_meta_pos = [{"is_init": False} for _ in range(self.num_transformer_blocks)]
        self.crossattn_cache_meta_neg = [{"is_init": False} for _ in range(self.num_transformer_blocks)]

    def _reset_caches(self, kv_cache_manager_pos, kv_cache_manager_neg, kv_cache_requests):
        """:120-134 (fresh zero indices, cross-attention caches invalidated) + the native block tables."""
        for mgr, metas, cmetas in ((kv_cache_manager_pos, self.kv_cache_meta_pos, self.crossattn_cache_meta_pos),
                                   (kv_cache_manager_neg, self.kv_cache_meta_neg, self.crossattn_cache_meta_neg)):
            for i in range(self.num_transformer_blocks):
                cmetas[i]["is_init"] = False
                metas[i]["global_end_index"].zero_()
                metas[i]["local_end_index"].zero_()
                for req in kv_cache_requests:
                    self.generator.model.blocks[i].kv_cache_manager.reset_kv_cache(mgr, req, mgr.device)

    def clear_cache(self, kv_cache_manager_pos, kv_cache_manager_neg, kv_cache_requests):
        for mgr in (kv_cache_manager_pos, kv_cache_manager_neg):
            for i in range(self.num_transformer_blocks):
                for req in kv_cache_requests:
                    self.generator.model.blocks[i].kv_cache_manager.clear_cache(kv_cache_manager=mgr, kv_cache_request=req)
        self.kv_cache_meta_pos = self.kv_cache_meta_neg = None
        self.crossattn_cache_meta_pos = self.crossattn_cache_meta_neg = None

    def _initialize_sample_scheduler(self, noise):
        """:364-372 (unipc)."""
        s = FlowUniPCMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1, use_dynamic_shifting=False)
        s.set_timesteps(self.sampling_steps, device=noise.device, shift=self.shift)
        self.timesteps = s.timesteps
        return s

    # ------------------------------------------------------------------------------------------ inference (:50-296)
    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts, kv_cache_manager_pos: KVCacheManager,
                  kv_cache_manager_neg: KVCacheManager, kv_cache_requests: List[KVCacheRequest],
                  initial_latent: Optional[torch.Tensor] = None, return_latents: bool = False,
                  start_frame_index: Optional[int] = 0):
        batch_size, num_frames, num_channels, height, width = noise.shape
        assert num_frames % self.num_frame_per_block == 0
        num_blocks = num_frames // self.num_frame_per_block
        num_input_frames = initial_latent.shape[1] if initial_latent is not None else 0
        ps = self.generator.model.patch_size
        fs = (height // ps[1]) * (width // ps[2])
        if self.frame_seq_length is None:
            self.frame_seq_length = fs
        elif self.frame_seq_length != fs:
            raise ValueError(f"frame_seq_length={self.frame_seq_length} does not match the latent shape ({fs})")
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        neg = getattr(self.args, "negative_prompt_embeds", None)
        if neg is None:
            raise ValueError("args.negative_prompt_embeds (the encoded negative prompt, [B, L, text_dim]) is required: the "
                             "text encoder is out of scope")
        unconditional_dict = self.text_encoder(text_prompts=neg)
        output = torch.zeros([batch_size, num_frames + num_input_frames, num_channels, height, width], device=noise.device,
                             dtype=noise.dtype)
        if self.kv_cache_meta_pos is None:
            self._initialize_kv_cache(kv_cache_manager_pos, kv_cache_manager_neg, kv_cache_requests, dtype=noise.dtype)
            self._initialize_crossattn_cache(kv_cache_manager_pos, kv_cache_manager_neg, kv_cache_requests, dtype=noise.dtype)
        else:
            self._reset_caches(kv_cache_manager_pos, kv_cache_manager_neg, kv_cache_requests)
        branches = ((conditional_dict, self.kv_cache_meta_pos, self.crossattn_cache_meta_pos, kv_cache_manager_pos),
                    (unconditional_dict, self.kv_cache_meta_neg, self.crossattn_cache_meta_neg, kv_cache_manager_neg))

        def forward_both(x, timestep, current_start_frame, cache_start_frame):
            outs = []
            for cond, metas, cmetas, mgr in branches:
                flow, _ = self.generator(noisy_image_or_video=x, conditional_dict=cond, timestep=timestep, kv_cache_meta=metas,
                                         crossattn_cache_meta=cmetas, current_start=current_start_frame * self.frame_seq_length,
                                         cache_start=cache_start_frame * self.frame_seq_length, kv_cache_manager=mgr,
                                         kv_cache_requests=kv_cache_requests)
                outs.append(flow)
            return outs

        current_start_frame, cache_start_frame = start_frame_index, 0
        if initial_latent is not None:                                   # :139-200 video extension: cache the context
            assert num_input_frames % self.num_frame_per_block == 0
            zero_t = torch.zeros([batch_size, self.num_frame_per_block], device=noise.device, dtype=torch.int64)
            for _ in range(num_input_frames // self.num_frame_per_block):
                ref = initial_latent[:, cache_start_frame:cache_start_frame + self.num_frame_per_block]
                output[:, cache_start_frame:cache_start_frame + self.num_frame_per_block] = ref
                forward_both(ref, zero_t, current_start_frame, cache_start_frame)
                current_start_frame += self.num_frame_per_block
                cache_start_frame += self.num_frame_per_block

        for _ in range(num_blocks):                                      # :203-287
            n = self.num_frame_per_block
            latents = noise[:, cache_start_frame - num_input_frames:cache_start_frame + n - num_input_frames]
            sample_scheduler = self._initialize_sample_scheduler(noise)
            timestep = None
            for t in sample_scheduler.timesteps:
                timestep = t * torch.ones([batch_size, n], device=noise.device, dtype=torch.float32)
                flow_cond, flow_uncond = forward_both(latents, timestep, current_start_frame, cache_start_frame)
                flow_pred = flow_uncond + self.args.guidance_scale * (flow_cond - flow_uncond)
                latents = sample_scheduler.step(flow_pred, t, latents, return_dict=False)[0]
            output[:, cache_start_frame:cache_start_frame + n] = latents
            forward_both(latents, timestep * 0, current_start_frame, cache_start_frame)     # clean context (:256-279)
            current_start_frame += n
            cache_start_frame += n

        if self.vae is None:
            if not return_latents:
                raise RuntimeError("no VAE attached (out of scope): call with return_latents=True")
            return None, output
        video = (self.vae.decode_to_pixel(output) * 0.5 + 0.5).clamp(0, 1)
        return (video, output) if return_latents else video
