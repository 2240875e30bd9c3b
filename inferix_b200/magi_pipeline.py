"""MAGI-1 semi-autoregressive chunk scheduler with the reference's surface
(`SampleTransport`, inferix/pipeline/magi/video_generate.py:252-756) on top of
`inferix_b200.magi_schedule` (index / timestep arithmetic) and `inferix_b200.magi_model.VideoDiTModel`.

One `walk()` iteration = the reference's: `forward_velocity` builds the window of chunks being denoised (+ the extra
clean chunk whose K/V must be stored when a stage starts), their timesteps and kv ranges, and calls
`model.forward_dispatcher`; `integrate_velocity` takes the Euler step per chunk and yields a chunk once it has received
all its steps.  Single input, pipeline-parallel size 1 (the reference's PP work queue over several inputs is not
replicated).  Host control flow only — every FLOP is inside the model.
"""
from __future__ import annotations

from collections import Counter
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import torch

from . import magi_schedule as ms
from .kvcache_manager.model.magi_kv_cache_manager import InferenceParams


@dataclass(frozen=True)
class InferenceInput:
    """video_generate.py:35-47 (the fields the scheduler reads)."""
    y: torch.Tensor                      # [2, chunk_num, L, C]: (conditional, null) captions per chunk
    emb_masks: torch.Tensor              # [2, chunk_num, L]
    prefix_video: Optional[torch.Tensor]  # [1, C, T_prefix, H, W] latents or None
    latent_size: Tuple[int, ...]         # (N, C, T, H, W)
    t_schedule_config: Dict = field(default_factory=dict)
    num_steps: int = 64
    chunk_num: int = 1
    caption_embs: Optional[torch.Tensor] = None


def find_dit_model(model):
    """:245-250."""
    if hasattr(model, "y_embedder"):
        return model
    if hasattr(model, "module"):
        return find_dit_model(model.module)
    raise ValueError("Cannot find the real model")


class SampleTransport:
    def __init__(self, model, transport_inputs: List[InferenceInput], device, inference_params=None, noise=None):
        """`inference_params` / `noise` may be injected (tests, callers that own the cache); by default the cache is an
        `InferenceParams` over the native KVCacheManager sized like the reference's (:313-316) and the start latent
        is `torch.randn(latent_size)` duplicated for the two CFG copies (:309-311)."""
        assert len(transport_inputs) == 1, "Only support single input for PP=1"
        self.model, self.transport_inputs, self.device = model, transport_inputs, device
        dit = find_dit_model(model)
        self.model_config, self.runtime_config, self.engine_config = dit.model_config, dit.runtime_config, dit.engine_config
        self.chunk_width, self.window_size = self.runtime_config.chunk_width, self.runtime_config.window_size
        ti = transport_inputs[0]
        shortcut = getattr(self.engine_config, "shortcut_mode", "")
        self.chunk_denoise_count = [Counter()]
        self.ts = [ms.init_t(ti.t_schedule_config, ti.num_steps, device, shortcut_mode=shortcut)]
        self.time_interval = [ms.init_intervel(ti.num_steps, device, shortcut_mode=shortcut)]
        self.x_chunks: List[Optional[torch.Tensor]] = [None]
        self.velocities: List[Optional[torch.Tensor]] = [None]
        x = noise if noise is not None else torch.randn(*ti.latent_size, device=device)
        self.xs = [torch.cat([x, x], 0)]
        p = self.model_config.patch_size
        max_seq = x.shape[2] * (x.shape[3] // p) * (x.shape[4] // p)
        self.inference_params = [inference_params if inference_params is not None
                                 else InferenceParams(max_batch_size=1, max_sequence_length=max_seq, device=device)]

    # ------------------------------------------------------------------ const helpers (:320-585)
    def _chunk_offset(self, infer_idx: int) -> int:
        pv = self.transport_inputs[infer_idx].prefix_video
        return 0 if pv is None else pv.size(2) // self.chunk_width

    def get_batch_size_and_chunk_token_nums(self, infer_idx: int):
        ls, p = self.transport_inputs[infer_idx].latent_size, self.model_config.patch_size
        return 1, self.chunk_width * (ls[3] // p) * (ls[4] // p)

    def get_timestep(self, t_total, denoise_step_per_stage, start, end, denoise_idx, has_clean_t=False):
        return ms.get_timestep(t_total, denoise_step_per_stage, start, end, denoise_idx, has_clean_t,
                               clean_t=self.runtime_config.clean_t)

    def generate_denoise_status_and_sequences(self, infer_idx: int, cur_denoise_step: int):
        ti = self.transport_inputs[infer_idx]
        return ms.denoise_status_and_sequences(cur_denoise_step, ti.num_steps, ti.chunk_num, self.window_size,
                                               self._chunk_offset(infer_idx))

    def total_forward_step(self, infer_idx: int) -> int:
        ti = self.transport_inputs[infer_idx]
        return ms.total_forward_step(ti.num_steps, ti.chunk_num, self.window_size, self._chunk_offset(infer_idx))

    def generate_kvrange_for_prefix_video(self, infer_idx: int, range_num: int):
        _, ctn = self.get_batch_size_and_chunk_token_nums(infer_idx)
        rc = self.runtime_config
        return ms.kvrange_for_prefix_video(range_num, ctn, rc.clean_chunk_kvrange, rc.noise2clean_kvrange).to(self.device)

    def generate_kvrange_for_denoising_video(self, infer_idx, slice_point, denoising_range_num, denoise_step_of_each_chunk):
        _, ctn = self.get_batch_size_and_chunk_token_nums(infer_idx)
        rc = self.runtime_config
        return ms.kvrange_for_denoising_video(slice_point, denoising_range_num, ctn, denoise_step_of_each_chunk,
                                              self.transport_inputs[infer_idx].num_steps, rc.noise2clean_kvrange,
                                              rc.clean_chunk_kvrange).to(self.device)

    # ------------------------------------------------------------------ prefix video (:390-454)
    def extract_prefix_video_feature(self, infer_idx, prefix_video, y, chunk_offset, model_kwargs):
        """Forward of the clean prefix chunks with the null caption at t = clean_t, only to fill the KV cache."""
        ti = self.transport_inputs[infer_idx]
        x_chunk = prefix_video[:, :, :chunk_offset * self.chunk_width]
        x_chunk = torch.cat([x_chunk, x_chunk], 0)
        null_y = torch.cat([ti.y[1:2, :chunk_offset]] * 2, 0)
        mask = torch.cat([ti.emb_masks[1:2, :chunk_offset]] * 2, 0)
        t = (torch.ones(chunk_offset, device=self.device) * self.runtime_config.clean_t).unsqueeze(0).repeat(x_chunk.size(0), 1)
        kw = dict(model_kwargs)
        kw.update(slice_point=0, range_num=chunk_offset, denoising_range_num=chunk_offset, fwd_extra_1st_chunk=False,
                  extract_prefix_video_feature=True, distill_interval=self.time_interval[infer_idx][0])
        find_dit_model(self.model).forward_dispatcher(
            x=x_chunk, timestep=t, y=null_y.flatten(0, 1).unsqueeze(1), mask=mask.flatten(0, 1).unsqueeze(1),
            kv_range=self.generate_kvrange_for_prefix_video(infer_idx, chunk_offset),
            inference_params=self.inference_params[infer_idx], **kw)

    def try_pad_prefix_video(self, infer_idx, x_chunk, t, prefix_video_start):
        pv = self.transport_inputs[infer_idx].prefix_video
        prefix_length = pv.size(2)
        if prefix_length <= prefix_video_start:
            return x_chunk, t
        pad = min(prefix_length - prefix_video_start, x_chunk.size(2))
        ret = x_chunk.clone()
        ret[:, :, :pad] = pv[:, :, prefix_video_start:prefix_video_start + pad]
        num_clean_t = (prefix_length - prefix_video_start) // self.chunk_width
        if num_clean_t > 0:
            t[:, :num_clean_t] = 1.0
        return ret, t

    # ------------------------------------------------------------------ one step (:587-756)
    def forward_velocity(self, infer_idx: int, cur_denoise_step: int) -> torch.Tensor:
        x, ti = self.xs[infer_idx], self.transport_inputs[infer_idx]
        (dps, _stage, didx), (chunk_offset, cs, ce, t_start, t_end) = self.generate_denoise_status_and_sequences(
            infer_idx, cur_denoise_step)
        kw = dict(chunk_width=self.chunk_width, fwd_extra_1st_chunk=False, num_steps=ti.num_steps)
        if chunk_offset > 0 and cur_denoise_step == 0:
            self.extract_prefix_video_feature(infer_idx, ti.prefix_video, ti.y, chunk_offset, kw)
        cw = self.chunk_width
        x_chunk = x[:, :, cs * cw:ce * cw].clone()
        y_chunk, mask_chunk = ti.y[:, cs:ce], ti.emb_masks[:, cs:ce]
        kw.update(slice_point=cs, range_num=ce, denoising_range_num=ce - cs)
        # a new stage starts: run the chunk that just became clean once more (null caption) so its K/V is stored
        extra = cs > chunk_offset and didx == 0
        if extra:
            x_chunk = torch.cat([x[:, :, (cs - 1) * cw:cs * cw].clone(), x_chunk], dim=2)
            y_chunk = torch.cat([ti.y[1:2, 0:1].expand(y_chunk.size(0), -1, -1, -1), y_chunk], dim=1)
            mask_chunk = torch.cat([ti.emb_masks[1:2, 1:2].expand(mask_chunk.size(0), -1, -1), mask_chunk], dim=1)
            kw.update(slice_point=cs - 1, denoising_range_num=ce - cs + 1, fwd_extra_1st_chunk=True)
        steps_each = ms.get_denoise_step_of_each_chunk(ti.num_steps, dps, t_start, t_end, didx, has_clean_t=extra)
        t = self.get_timestep(self.ts[infer_idx], dps, t_start, t_end, didx, has_clean_t=extra)
        t = t.unsqueeze(0).repeat(x_chunk.size(0), 1)
        kv_range = self.generate_kvrange_for_denoising_video(infer_idx, kw["slice_point"], kw["denoising_range_num"],
                                                             steps_each)
        if ti.prefix_video is not None:
            x_chunk, t = self.try_pad_prefix_video(infer_idx, x_chunk, t, prefix_video_start=kw["slice_point"] * cw)
        nearly_clean_t = t[0, int(kw["fwd_extra_1st_chunk"])].item()
        kw["distill_nearly_clean_chunk"] = nearly_clean_t > self.engine_config.distill_nearly_clean_chunk_threshold
        kw["distill_interval"] = self.time_interval[infer_idx][didx]
        velocity = find_dit_model(self.model).forward_dispatcher(
            x=x_chunk, timestep=t, y=y_chunk.flatten(0, 1).unsqueeze(1), mask=mask_chunk.flatten(0, 1).unsqueeze(1),
            kv_range=kv_range, inference_params=self.inference_params[infer_idx], **kw)
        self.x_chunks[infer_idx], self.velocities[infer_idx] = x_chunk, velocity
        return velocity

    def integrate(self, x_chunk, velocity, t_total, denoise_step_per_stage, t_start, t_end, i):
        return ms.integrate(x_chunk, velocity, t_total, denoise_step_per_stage, t_start, t_end, i, self.chunk_width)

    def integrate_velocity(self, infer_idx: int, cur_denoise_step: int):
        ti = self.transport_inputs[infer_idx]
        x_chunk, velocity = self.x_chunks[infer_idx], self.velocities[infer_idx]
        count = self.chunk_denoise_count[infer_idx]
        (dps, _stage, didx), (chunk_offset, cs, ce, t_start, t_end) = self.generate_denoise_status_and_sequences(
            infer_idx, cur_denoise_step)
        cw = self.chunk_width
        if cs > chunk_offset and didx == 0:              # drop the extra clean chunk
            x_chunk, velocity = x_chunk[:, :, cw:], velocity[:, :, cw:]
        x_chunk = self.integrate(x_chunk, velocity, self.ts[infer_idx], dps, t_start, t_end, didx)
        for c in range(cs, ce):
            count[c] += 1
        self.xs[infer_idx][:, :, cs * cw:ce * cw] = x_chunk
        if count[cs] == ti.num_steps:                    # the oldest chunk of the window is clean: hand it out
            if ti.prefix_video is not None:
                plen = ti.prefix_video.size(2)
                if (cs + 1) * cw <= plen:
                    return None, None
                real_start = max(cs * cw, plen)
                if cs == 0 and plen == 1:                # I2V: keep the first frames
                    real_start = 0
                clean, _ = self.xs[infer_idx][:, :, real_start:(cs + 1) * cw].chunk(2, dim=0)
                return clean, cs - chunk_offset
            clean, _ = self.xs[infer_idx][:, :, cs * cw:(cs + 1) * cw].chunk(2, dim=0)
            return clean, cs - chunk_offset
        return None, None

    def walk(self) -> Iterator[Tuple[int, int, torch.Tensor]]:
        """:731-766 for one input: yields (infer_idx, chunk_idx, clean_chunk) in generation order."""
        step, total = 0, self.total_forward_step(0)
        self.forward_velocity(0, 0)
        while True:
            clean, idx = self.integrate_velocity(0, step)
            if clean is not None:
                yield 0, idx, clean
            if step + 1 == total:
                return
            step += 1
            self.forward_velocity(0, step)
