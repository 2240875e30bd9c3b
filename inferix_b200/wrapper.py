"""WanDiffusionWrapper (reference inferix/models/self_forcing/wrapper.py:172-383): flow -> x0 conversion around the
DiT, scheduler binding.  Only the causal KV-cached branch is built; text encoder / VAE wrappers are out of scope.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .kvcache_manager import KVCacheManager, KVCacheRequest
from .parallel import ParallelConfig
from .scheduler import FlowMatchScheduler
from .wan_model import CausalWanModel


class WanDiffusionWrapper(torch.nn.Module):
    def __init__(self, model: Optional[CausalWanModel] = None, model_path: Optional[str] = None,
                 model_name: str = "Wan2.1-T2V-1.3B", timestep_shift: float = 8.0, is_causal: bool = True,
                 local_attn_size: int = -1, sink_size: int = 0, enable_kv_offload: bool = True,
                 parallel_config: Optional[ParallelConfig] = None, model_kwargs: Optional[dict] = None):
        super().__init__()
        if not is_causal:
            raise NotImplementedError("the bidirectional teacher model is out of scope")
        self.parallel_config = parallel_config if parallel_config is not None else ParallelConfig()
        self.enable_kv_offload = enable_kv_offload
        if model is None:
            # no checkpoint I/O here (reference: CausalWanModel.from_pretrained(model_path)); the caller loads a
            # state_dict into .model — parameter names match the reference's.
            model = CausalWanModel(local_attn_size=local_attn_size, sink_size=sink_size,
                                   enable_kv_offload=enable_kv_offload, parallel_config=self.parallel_config,
                                   **(model_kwargs or {}))
        self.model = model.eval()
        self.uniform_timestep = False
        self.scheduler = FlowMatchScheduler(shift=timestep_shift, sigma_min=0.0, extra_one_step=True)
        self.scheduler.set_timesteps(1000, training=True)
        self.seq_len = 32760

    def _convert_flow_pred_to_x0(self, flow_pred: torch.Tensor, xt: torch.Tensor, timestep: torch.Tensor):
        """x0 = x_t - sigma_t * flow in fp64 (reference :259-283)."""
        original_dtype = flow_pred.dtype
        flow_pred, xt, sigmas, timesteps = map(lambda x: x.double().to(flow_pred.device),
                                               [flow_pred, xt, self.scheduler.sigmas, self.scheduler.timesteps])
        timestep_id = torch.argmin((timesteps.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)
        sigma_t = sigmas[timestep_id].reshape(-1, 1, 1, 1)
        return (xt - sigma_t * flow_pred).to(original_dtype)

    def forward(self, noisy_image_or_video: torch.Tensor, conditional_dict: dict, timestep: torch.Tensor,
                kv_cache_meta: Optional[List[dict]] = None, crossattn_cache_meta: Optional[List[dict]] = None,
                current_start: Optional[int] = None, classify_mode: Optional[bool] = False,
                concat_time_embeddings: Optional[bool] = False, clean_x: Optional[torch.Tensor] = None,
                aug_t: Optional[torch.Tensor] = None, cache_start: Optional[int] = None,
                kv_cache_manager: Optional[KVCacheManager] = None,
                kv_cache_requests: Optional[List[KVCacheRequest]] = None):
        """noisy_image_or_video [B, F, C, H, W] -> (flow_pred, pred_x0), reference :308-383."""
        if kv_cache_meta is None:
            raise NotImplementedError("only the KV-cached inference branch is built")
        prompt_embeds = conditional_dict["prompt_embeds"]
        x = noisy_image_or_video
        if x.is_cuda and x.dtype == torch.bfloat16 and getattr(self.model, "patch_size", (1, 2, 2))[0] == 1:
            # native epilogue: unpatchify + fp64 flow -> x0 in one kernel straight from the head's token output
            from . import ops
            tokens, _grid = self.model(
                x.permute(0, 2, 1, 3, 4), t=timestep, context=prompt_embeds, seq_len=self.seq_len,
                kv_cache_meta=kv_cache_meta, crossattn_cache_meta=crossattn_cache_meta, current_start=current_start,
                cache_start=cache_start, kv_cache_manager=kv_cache_manager, kv_cache_requests=kv_cache_requests,
                return_tokens=True)
            dev = x.device
            if self.scheduler.sigmas.device != dev or self.scheduler.timesteps.device != dev:
                self.scheduler.sigmas = self.scheduler.sigmas.to(dev)
                self.scheduler.timesteps = self.scheduler.timesteps.to(dev)
            tt = self.scheduler.timesteps.float().contiguous()
            ss = self.scheduler.sigmas.float().contiguous()
            ps = self.model.patch_size
            flows, x0s = [], []
            for b in range(x.shape[0]):
                fl, x0 = ops.unpatchify_x0(tokens[b].contiguous(), x[b], timestep[b].to(torch.float64).contiguous(), tt, ss,
                                           (ps[1], ps[2]))
                flows.append(fl)
                x0s.append(x0)
            return torch.stack(flows), torch.stack(x0s)
        flow_pred = self.model(
            noisy_image_or_video.permute(0, 2, 1, 3, 4), t=timestep, context=prompt_embeds, seq_len=self.seq_len,
            kv_cache_meta=kv_cache_meta, crossattn_cache_meta=crossattn_cache_meta, current_start=current_start,
            cache_start=cache_start, kv_cache_manager=kv_cache_manager, kv_cache_requests=kv_cache_requests,
        ).permute(0, 2, 1, 3, 4)
        pred_x0 = self._convert_flow_pred_to_x0(
            flow_pred=flow_pred.flatten(0, 1), xt=noisy_image_or_video.flatten(0, 1), timestep=timestep.flatten(0, 1)
        ).unflatten(0, flow_pred.shape[:2])
        return flow_pred, pred_x0

    def get_scheduler(self):
        return self.scheduler
