"""CausVid on the native path: model, wrapper and block scheduler with the reference's CausVid surface.

Reference: ``inferix/models/causvid/causal_model.py`` (CausalWanModel.forward :485-600, block :229-319, self-attention
:128-179), ``inferix/models/causvid/wrapper.py:269-305`` (returns x0 only) and
``inferix/pipeline/causvid/CausalInferencePipeline.py:94-257``.

CausVid differs from Self-Forcing only in bookkeeping: the caller passes explicit ``kv_start / kv_end`` (cache range
written this forward, per rank) and ``current_start / current_end`` (global token range, used for RoPE); there is no
window / eviction; attention runs over ``cache[0:kv_end]``; the cross-attention init flag lives on the block module.
Writing block b at ``kv_start = b * S`` is exactly the Self-Forcing index arithmetic with ``windowed = False``
(local_end advances by current_end - global_end), so the same native block (`ifx_wan_block_forward`) serves both.
Under sequence parallelism the reference splits the flattened sequence and shards the cache (:574-575); here CausVid
uses the same per-frame hw split + replicated cache as Self-Forcing, so ``kv_start / kv_end`` are only
validated against ``current_start / current_end``.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .kvcache_manager import KVCacheManager, KVCacheRequest
from .kvcache_manager.model import KVCacheManagerFactory
from .parallel import ParallelConfig
from .scheduler import FlowMatchScheduler
from .wan_model import CausalWanModel


class CausVidCausalWanModel(CausalWanModel):
    """``inferix.models.causvid.CausalWanModel``: no local window, explicit cache ranges."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, window_size=(-1, -1),
                 qk_norm=True, cross_attn_norm=True, eps=1e-6, enable_kv_offload=True,
                 parallel_config: Optional[ParallelConfig] = None):
        super().__init__(model_type=model_type, patch_size=patch_size, text_len=text_len, in_dim=in_dim, dim=dim,
                         ffn_dim=ffn_dim, freq_dim=freq_dim, text_dim=text_dim, out_dim=out_dim, num_heads=num_heads,
                         num_layers=num_layers, local_attn_size=-1, sink_size=0, qk_norm=qk_norm,
                         cross_attn_norm=cross_attn_norm, eps=eps, enable_kv_offload=enable_kv_offload,
                         parallel_config=parallel_config)
        self.window_size = window_size
        for i, blk in enumerate(self.blocks):
            blk.kv_cache_manager = KVCacheManagerFactory.create_manager(i, num_heads, dim // num_heads,
                                                                        enable_kv_offload)
            blk.is_cross_attn_init = False          # reference causvid/causal_model.py:227
        self._meta = None

    def _metas(self, device):
        if self._meta is None or self._meta[0]["global_end_index"].device != device:
            idx = torch.zeros((self.num_layers, 2), dtype=torch.long, device=device)
            self._meta = [{"global_end_index": idx[i, 0:1], "local_end_index": idx[i, 1:2], "_ifx_shared": idx}
                          for i in range(self.num_layers)]
        return self._meta

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, kv_start: int = 0, kv_end: int = 0,
                current_start: int = 0, current_end: int = 0, kv_cache_manager: Optional[KVCacheManager] = None,
                kv_cache_requests: Optional[list] = None):
        world = self.parallel_config.world_size
        if (kv_end - kv_start) * world != current_end - current_start or kv_start * world != current_start:
            raise ValueError(f"kv range [{kv_start}, {kv_end}) x{world} does not match the token range "
                             f"[{current_start}, {current_end})")
        cross = [{"is_init": blk.is_cross_attn_init} for blk in self.blocks]
        out = super().forward(x, t, context, seq_len, clip_fea=clip_fea, y=y, kv_cache_meta=self._metas(x[0].device),
                              crossattn_cache_meta=cross, current_start=current_start, cache_start=current_start,
                              kv_cache_manager=kv_cache_manager, kv_cache_requests=kv_cache_requests)
        for blk, c in zip(self.blocks, cross):
            blk.is_cross_attn_init = c["is_init"]
        return out


class CausVidDiffusionWrapper(torch.nn.Module):
    """``inferix.models.causvid.WanDiffusionWrapper`` (wrapper.py:172-305): forward returns pred_x0 only."""

    def __init__(self, model: Optional[CausVidCausalWanModel] = None, timestep_shift: float = 8.0,
                 enable_kv_offload: bool = True, parallel_config: Optional[ParallelConfig] = None,
                 model_kwargs: Optional[dict] = None):
        super().__init__()
        self.parallel_config = parallel_config if parallel_config is not None else ParallelConfig()
        if model is None:
            model = CausVidCausalWanModel(enable_kv_offload=enable_kv_offload, parallel_config=self.parallel_config,
                                          **(model_kwargs or {}))
        self.model = model.eval()
        self.uniform_timestep = False
        self.scheduler = FlowMatchScheduler(shift=timestep_shift, sigma_min=0.0, extra_one_step=True)
        self.scheduler.set_timesteps(1000, training=True)
        self.seq_len = 32760

    def _convert_flow_pred_to_x0(self, flow_pred, xt, timestep):
        original_dtype = flow_pred.dtype
        flow_pred, xt, sigmas, timesteps = map(lambda v: v.double().to(flow_pred.device),
                                               [flow_pred, xt, self.scheduler.sigmas, self.scheduler.timesteps])
        timestep_id = torch.argmin((timesteps.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)
        return (xt - sigmas[timestep_id].reshape(-1, 1, 1, 1) * flow_pred).to(original_dtype)

    def forward(self, noisy_image_or_video: torch.Tensor, conditional_dict: dict, timestep: torch.Tensor,
                kv_start: Optional[int] = None, kv_end: Optional[int] = None, current_start: Optional[int] = None,
                current_end: Optional[int] = None, kv_cache_manager: Optional[KVCacheManager] = None,
                kv_cache_requests: Optional[List] = None) -> torch.Tensor:
        flow_pred = self.model(noisy_image_or_video.permute(0, 2, 1, 3, 4), t=timestep,
                               context=conditional_dict["prompt_embeds"], seq_len=self.seq_len, kv_start=kv_start,
                               kv_end=kv_end, current_start=current_start, current_end=current_end,
                               kv_cache_manager=kv_cache_manager, kv_cache_requests=kv_cache_requests
                               ).permute(0, 2, 1, 3, 4)
        return self._convert_flow_pred_to_x0(flow_pred.flatten(0, 1), noisy_image_or_video.flatten(0, 1),
                                             timestep.flatten(0, 1)).unflatten(0, flow_pred.shape[:2])

    def get_scheduler(self):
        return self.scheduler


class CausVidInferencePipeline(torch.nn.Module):
    """``inferix.pipeline.causvid.CausalInferencePipeline`` (:9-257).  ``block_callback(block_latent, block_index)``
    is added so a streaming caller can decode per block (the reference's CausVid streaming hook raises
    NotImplementedError, causvid/pipeline.py:342-360)."""

    def __init__(self, args, device, generator: Optional[CausVidDiffusionWrapper] = None, text_encoder=None, vae=None,
                 enable_kv_offload: bool = True, parallel_config: Optional[ParallelConfig] = None):
        super().__init__()
        self.parallel_config = parallel_config if parallel_config is not None else ParallelConfig()
        self.generator = generator if generator is not None else CausVidDiffusionWrapper(
            enable_kv_offload=enable_kv_offload, parallel_config=self.parallel_config,
            model_kwargs=getattr(args, "model_kwargs", None))
        self.text_encoder = text_encoder if text_encoder is not None else (lambda text_prompts: {"prompt_embeds": text_prompts})
        self.vae = vae
        self.scheduler = self.generator.get_scheduler()
        self.denoising_step_list = torch.tensor(args.denoising_step_list, dtype=torch.long)[:-1]      # :37
        if getattr(args, "warp_denoising_step", False):
            ts = torch.cat((self.scheduler.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
            self.denoising_step_list = ts[1000 - self.denoising_step_list]
        self.num_transformer_blocks = self.generator.model.num_layers
        self.frame_seq_length = getattr(args, "frame_seq_length", None)
        self.kv_cache_frames = getattr(args, "kv_cache_frames", 21)          # reference: 32760 = 21 x 1560
        self.is_kv_cache_initialized = False
        self.args = args
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        if self.num_frame_per_block > 1:
            self.generator.model.num_frame_per_block = self.num_frame_per_block
        self.renoise_fn = torch.randn_like

    @property
    def per_rank_frame_seq_length(self):
        return self.frame_seq_length // self.parallel_config.world_size

    def _initialize_kv_cache(self, kv_cache_manager, kv_cache_requests, dtype):
        for layer_idx in range(self.num_transformer_blocks):
            adapter = self.generator.model.blocks[layer_idx].kv_cache_manager
            for req in kv_cache_requests:
                adapter.allocate_kv_cache(kv_cache_manager=kv_cache_manager, kv_cache_request=req,
                                          sequence_length=self.kv_cache_frames * self.frame_seq_length, dtype=dtype,
                                          ulysses_size=self.parallel_config.ulysses_size,
                                          ring_size=self.parallel_config.ring_size, page_tokens=self.frame_seq_length)
        from . import peer
        self._peer_group = peer.setup_for_pipeline(self.generator.model, kv_cache_manager, kv_cache_requests,
                                                   self.parallel_config)

    def _initialize_crossattn_cache(self, kv_cache_manager, kv_cache_requests, dtype):
        for layer_idx in range(self.num_transformer_blocks):
            adapter = self.generator.model.blocks[layer_idx].kv_cache_manager
            for req in kv_cache_requests:
                adapter.allocate_crossattn_cache(kv_cache_manager=kv_cache_manager, kv_cache_request=req,
                                                 crossattn_length=self.generator.model.text_len, dtype=dtype)

    def _reset_crossattn_cache(self):
        for blk in self.generator.model.blocks:
            blk.is_cross_attn_init = False

    def _reset_kv_cache(self, kv_cache_manager, kv_cache_requests):
        for blk in self.generator.model.blocks:
            for req in kv_cache_requests:
                blk.kv_cache_manager.reset_kv_cache(kv_cache_manager, req)
        self.generator.model._meta = None

    def clear_cache(self, kv_cache_manager, kv_cache_requests):
        if getattr(self, "_peer_group", None) is not None:      # unmap the other ranks' caches before anyone frees
            self._peer_group.release()
            torch.distributed.barrier(group=self.parallel_config.group)
            self._peer_group = None
        for blk in self.generator.model.blocks:
            for req in kv_cache_requests:
                blk.kv_cache_manager.clear_cache(kv_cache_manager=kv_cache_manager, kv_cache_request=req)
            blk.is_cross_attn_init = False
        self.generator.model._meta = None
        self.is_kv_cache_initialized = False

    @torch.no_grad()   # inference only; the reference disables autograd globally (self_forcing/pipeline.py:62)
    def inference(self, noise: torch.Tensor, text_prompts, start_latents: Optional[torch.Tensor],
                  return_latents: bool = True, kv_cache_manager: Optional[KVCacheManager] = None,
                  kv_cache_requests: Optional[List] = None, vae_chunk_size: Optional[int] = None,
                  block_callback=None, decode: bool = True):
        batch_size, num_frames, num_channels, height, width = noise.shape
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        ps = self.generator.model.patch_size
        fs = (height // ps[1]) * (width // ps[2])
        if self.frame_seq_length is None:
            self.frame_seq_length = fs
        elif self.frame_seq_length != fs:
            raise ValueError(f"frame_seq_length={self.frame_seq_length} does not match the latent shape ({fs})")
        output = torch.zeros([batch_size, num_frames, num_channels, height, width], device=noise.device,
                             dtype=noise.dtype)
        if not self.is_kv_cache_initialized:
            self._initialize_kv_cache(kv_cache_manager, kv_cache_requests, dtype=noise.dtype)
            self._initialize_crossattn_cache(kv_cache_manager, kv_cache_requests, dtype=noise.dtype)
            self.is_kv_cache_initialized = True
        else:
            self._reset_crossattn_cache()
            self._reset_kv_cache(kv_cache_manager, kv_cache_requests)   # positions restart at 0 for a new segment

        n = self.num_frame_per_block
        num_input_blocks = start_latents.shape[1] // n if start_latents is not None else 0
        for block_index in range(num_frames // n):
            rng = dict(current_start=block_index * n * self.frame_seq_length,
                       current_end=(block_index + 1) * n * self.frame_seq_length,
                       kv_start=block_index * n * self.per_rank_frame_seq_length,
                       kv_end=(block_index + 1) * n * self.per_rank_frame_seq_length,
                       kv_cache_manager=kv_cache_manager, kv_cache_requests=kv_cache_requests)
            noisy_input = noise[:, block_index * n:(block_index + 1) * n]
            zeros = torch.zeros([batch_size, n], device=noise.device, dtype=torch.int64)
            if start_latents is not None and block_index < num_input_blocks:
                ref = start_latents[:, block_index * n:(block_index + 1) * n]
                output[:, block_index * n:(block_index + 1) * n] = ref
                self.generator(noisy_image_or_video=ref, conditional_dict=conditional_dict, timestep=zeros, **rng)
                continue
            denoised_pred = timestep = None
            for index, current_timestep in enumerate(self.denoising_step_list):
                timestep = torch.ones([batch_size, n], device=noise.device, dtype=torch.int64) * current_timestep
                denoised_pred = self.generator(noisy_image_or_video=noisy_input, conditional_dict=conditional_dict,
                                               timestep=timestep, **rng)
                if index < len(self.denoising_step_list) - 1:
                    next_timestep = self.denoising_step_list[index + 1]
                    flat = denoised_pred.flatten(0, 1)
                    noisy_input = self.scheduler.add_noise(
                        flat, self.renoise_fn(flat),
                        next_timestep * torch.ones([batch_size], device=noise.device, dtype=torch.long)
                    ).view(denoised_pred.shape)
            if denoised_pred is None:
                raise RuntimeError(f"denoised_pred is None after the denoising loop for block {block_index}")
            output[:, block_index * n:(block_index + 1) * n] = denoised_pred
            self.generator(noisy_image_or_video=denoised_pred, conditional_dict=conditional_dict,
                           timestep=timestep * 0, **rng)
            if block_callback is not None:
                block_callback(output[:, block_index * n:(block_index + 1) * n], block_index)

        if not decode or self.vae is None:
            if decode and self.vae is None and not return_latents:
                raise RuntimeError("no VAE attached (out of scope): call with return_latents=True / decode=False")
            return (None, output) if return_latents else output
        chunk = vae_chunk_size if vae_chunk_size is not None else 2
        video = (self.vae.decode_to_pixel(output, use_cache=True, chunk_size=chunk) * 0.5 + 0.5).clamp(0, 1)
        return (video, output) if return_latents else video
