"""MAGI-1 VideoDiTModel with the reference's surface (inferix/models/magi/dit/dit_model.py:42-596) around the native
TransformerBlock (inferix_b200/magi_layer.py).

What runs where.  The 34 / 48 transformer layers — all of the model's FLOPs but ~0.1 % — are the native kernels.  The
prologue (patch embedding, timestep / caption embedders, rotary table, range bookkeeping: `get_embedding_and_meta`,
:113-260) and the epilogue (final fp32 linear, unpatchify, :338-360) are once-per-forward torch ops kept in fp32 like
the reference's `_high_precision_promoter` (:620-637); the CFG dispatcher (:399-596) is host control flow over three
forwards that share the KV cache.  Class / method names, arguments and parameter names follow the reference so its
checkpoints load and `SampleTransport` can call `forward_dispatcher` unchanged.

Restated, not copied: every method cites the lines it follows; tests pin them to the reference's own module run on CPU
(tests/golden/magi_model_*.pt).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import magi_cp
from .magi_layer import TransformerBlock


# ----------------------------------------------------------------------------- meta types (core/types/inference.py:52-85)
@dataclass(frozen=True)
class PackedCoreAttnParams:
    q_range: torch.Tensor
    k_range: torch.Tensor
    np_q_range: np.ndarray
    np_k_range: np.ndarray
    max_seqlen_q: int
    max_seqlen_k: int


@dataclass(frozen=True)
class PackedCrossAttnParams:
    q_ranges: object = None
    kv_ranges: object = None
    cu_seqlens_q: torch.Tensor = None
    cu_seqlens_kv: torch.Tensor = None
    max_seqlen_q: Optional[int] = None
    max_seqlen_kv: Optional[int] = None


@dataclass(frozen=True)
class ModelMetaArgs:
    H: int
    W: int
    cp_pad_size: Optional[int]
    cp_split_sizes: Optional[List[int]]
    slice_point: int
    denoising_range_num: int
    range_num: int
    extract_prefix_video_feature: bool
    fwd_extra_1st_chunk: bool
    distill_nearly_clean_chunk: bool
    clip_token_nums: int
    enable_cuda_graph: bool
    core_attn_params: PackedCoreAttnParams
    cross_attn_params: PackedCrossAttnParams


# ----------------------------------------------------------------------------- embedders (dit_module.py:53-177)
class TimestepEmbedder(nn.Module):
    """dit_module.py:53-105: sinusoid of 1000*t (cos | sin, 256 wide) -> Linear -> SiLU -> Linear."""

    def __init__(self, model_config, frequency_embedding_size: int = 256):
        super().__init__()
        width = int(model_config.hidden_size * model_config.cond_hidden_ratio)
        self.data_type = model_config.params_dtype
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, width, bias=True), nn.SiLU(),
                                 nn.Linear(width, width, bias=True))
        self.frequency_embedding_size = frequency_embedding_size
        self.timestep_rescale_factor = 1000

    @staticmethod
    def timestep_embedding(t, dim, max_period=10000, timestep_rescale_factor=1):
        half = dim // 2
        freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
        args = t[:, None].float() * freqs[None] * timestep_rescale_factor
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        if dim % 2:
            emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
        return emb

    def forward(self, t):
        t_freq = self.timestep_embedding(t.to(torch.float32), self.frequency_embedding_size,
                                         timestep_rescale_factor=self.timestep_rescale_factor)
        # the reference rounds the sinusoid to params_dtype and the fp32 MLP then runs under autocast(float32)
        # (dit_module.py:102, dit_model.py:278): bf16 rounding first, fp32 arithmetic after
        return self.mlp(t_freq.to(self.data_type).to(self.mlp[0].weight.dtype))


class CaptionEmbedder(nn.Module):
    """dit_module.py:109-157: caption -> (cross-attention stream, AdaLN stream); null-caption tokens for CFG."""

    def __init__(self, model_config):
        super().__init__()
        c, h = model_config.caption_channels, model_config.hidden_size
        self.y_proj_xattn = nn.Sequential(nn.Linear(c, int(h * model_config.xattn_cond_hidden_ratio), bias=True), nn.SiLU())
        self.y_proj_adaln = nn.Sequential(nn.Linear(c, int(h * model_config.cond_hidden_ratio), bias=True))
        self.null_caption_embedding = nn.Parameter(torch.zeros(model_config.caption_max_length, c))

    def caption_drop(self, caption, mask):
        return torch.where(mask[:, None, None, None], self.null_caption_embedding[None, None, :], caption)

    def caption_drop_single_token(self, mask):
        return torch.where(mask[:, None, None], self.null_caption_embedding[None, -1, :],
                           self.null_caption_embedding[None, -2, :])

    def forward(self, caption, train, caption_dropout_mask=None):
        if train and caption_dropout_mask is not None:
            caption = self.caption_drop(caption, caption_dropout_mask)
        caption_xattn = self.y_proj_xattn(caption)
        if caption_dropout_mask is not None:
            caption = self.caption_drop_single_token(caption_dropout_mask)
        return caption_xattn, self.y_proj_adaln(caption)


class FinalLinear(nn.Module):
    """dit_module.py:163-174."""

    def __init__(self, hidden_size, patch_size, t_patch_size, out_channels):
        super().__init__()
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * t_patch_size * out_channels, bias=False)

    def forward(self, x):
        return self.linear(x)


class LearnableRotaryEmbeddingCat(nn.Module):
    """dit_module.py:723-776 with in_pixels=False (the only mode the model uses, dit_model.py:76-78): learnable
    inverse-frequency bands (dim/8 of them) over a centred (t, h, w) grid, returned as [T*H*W, 3*dim/8 * 2] = sin | cos."""

    def __init__(self, dim, temperature=10000):
        super().__init__()
        self.dim, self.temperature = dim, temperature
        n = dim // 8
        self.bands = nn.Parameter(1.0 / (temperature ** (torch.arange(0, n, dtype=torch.int64).to(torch.float32) / n)))

    def get_embed(self, shape, ref_feat_shape=None):
        dev = self.bands.device
        axes = [torch.arange(s, device=dev, dtype=torch.int64).to(torch.float32) for s in shape]
        axes[1] = axes[1] - (shape[1] - 1) / 2                      # spatial centre at (0, 0)   (:648-650)
        axes[2] = axes[2] - (shape[2] - 1) / 2
        if ref_feat_shape is not None:                              # EVA-style rescale to the reference grid (:651-662)
            scaled = []
            for x, f, r in zip(axes, shape, ref_feat_shape):
                if f == 1:
                    assert r == 1, "ref_feat_shape must be 1 when feat_shape is 1"
                    scaled.append(x)
                else:
                    scaled.append(x / (f - 1) * (r - 1))
            axes = scaled
        grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).unsqueeze(-1)
        pos = grid * self.bands
        n_tok = int(np.prod(shape))
        return torch.cat([pos.sin().reshape(n_tok, -1), pos.cos().reshape(n_tok, -1)], dim=-1)


# ----------------------------------------------------------------------------- the model
class VideoDiTModel(nn.Module):
    """dit_model.py:42-596.  `config` needs .model_config / .runtime_config / .engine_config (MagiConfig)."""

    def __init__(self, config, pre_process: bool = True, post_process: bool = True):
        super().__init__()
        mc = self.model_config = config.model_config
        self.runtime_config, self.engine_config = config.runtime_config, config.engine_config
        self.pre_process, self.post_process = pre_process, post_process
        self.in_channels, self.out_channels = mc.in_channels, mc.out_channels
        self.patch_size, self.t_patch_size = mc.patch_size, mc.t_patch_size
        self.caption_max_length, self.num_heads = mc.caption_max_length, mc.num_attention_heads
        k = (mc.t_patch_size, mc.patch_size, mc.patch_size)
        self.x_embedder = nn.Conv3d(mc.in_channels, mc.hidden_size, kernel_size=k, stride=k, bias=False)
        self.t_embedder = TimestepEmbedder(mc)
        self.y_embedder = CaptionEmbedder(mc)
        self.rope = LearnableRotaryEmbeddingCat(mc.hidden_size // mc.num_attention_heads)
        self.videodit_blocks = TransformerBlock(mc, self.engine_config, pre_process=pre_process, post_process=post_process)
        self.final_linear = FinalLinear(mc.hidden_size, mc.patch_size, mc.t_patch_size, mc.out_channels)
        self.high_precision_promoter()

    def high_precision_promoter(self):
        """dit_model.py:620-637: embedders, rope and the final linear run in fp32 (the block's norms are created in
        their final dtypes by magi_layer)."""
        for m in (self.x_embedder, self.y_embedder, self.t_embedder, self.final_linear, self.rope):
            m.float()
        return self

    def _device(self):
        return self.x_embedder.weight.device

    def generate_kv_range_for_uncondition(self, uncond_x) -> torch.Tensor:
        """:92-100 — every chunk of the unconditional pass attends itself only."""
        b, _, t, h, w = uncond_x.shape
        n = (t // self.t_patch_size) * (h // self.patch_size) * (w // self.patch_size)
        start = torch.linspace(0, (b - 1) * n, steps=b).reshape(b, 1)
        end = torch.linspace(n, b * n, steps=b).reshape(b, 1)
        return torch.cat([start, end], dim=1).to(torch.int32).to(uncond_x.device)

    def unpatchify(self, x, H, W):
        """:102-111: '(T H W) N (pT pH pW C) -> N C (T pT) (H pH) (W pW)'."""
        pt, p = self.t_patch_size, self.patch_size
        thw, n, _ = x.shape
        t = thw // (H * W)
        x = x.view(t, H, W, n, pt, p, p, -1).permute(3, 7, 0, 4, 1, 5, 2, 6)
        return x.reshape(n, -1, t * pt, H * p, W * p).contiguous()

    @torch.no_grad()
    def get_embedding_and_meta(self, x, t, y, caption_dropout_mask, xattn_mask, kv_range, **kwargs):
        """:113-260 (single-card behaviour)."""
        x = self.x_embedder(x)
        batch, _, T, H, W = x.shape
        range_num, ranges = kwargs["range_num"], kwargs["denoising_range_num"]
        slice_point = kwargs.get("slice_point", 0)
        frames_per_range = T // ranges
        t_total = T + frames_per_range * slice_point
        # rotary table over history + current frames, the current frames are its tail (:157-163)
        rescale = math.sqrt((H * W) / (16 * 16))
        rope = self.rope.get_embed(shape=[t_total, H, W], ref_feat_shape=[t_total, H / rescale, W / rescale])
        rope = rope[-(T * H * W):]
        # timestep (+ distillation step-size) embedding (:165-183)
        assert t.shape[0] == batch and t.shape[1] == ranges
        t_flat = t.flatten()
        t_emb = self.t_embedder(t_flat)
        if getattr(self.engine_config, "distill", False):
            if kwargs["num_steps"] == 12:
                factor = 4 / kwargs["distill_interval"] * 2
            else:
                factor = kwargs["num_steps"] / 4 * 2
            t_emb = t_emb + self.t_embedder(torch.ones_like(t_flat) * factor)
        t_emb = t_emb.reshape(batch, ranges, -1)
        # caption streams (:185-207)
        y_xattn, y_adaln = self.y_embedder(y, self.training, caption_dropout_mask)
        assert xattn_mask is not None
        xattn_mask = xattn_mask.squeeze(1).squeeze(1)
        condition = t_emb + y_adaln.squeeze(1).unsqueeze(1)
        assert condition.shape[:2] == (batch, ranges)
        seqlen_per_chunk = (T * H * W) // ranges
        condition_map = torch.repeat_interleave(torch.arange(batch * ranges, device=x.device), seqlen_per_chunk)
        condition_map = condition_map.reshape(batch, -1).transpose(0, 1).contiguous()
        y_flat = torch.masked_select(y_xattn.squeeze(1), xattn_mask.unsqueeze(-1).bool()).reshape(-1, y_xattn.shape[-1])
        # packed cross-attention ranges (:209-241)
        xattn_mask = xattn_mask.reshape(xattn_mask.shape[0], -1)
        y_index = torch.sum(xattn_mask, dim=-1)
        clip_tokens = H * W * frames_per_range
        cu_q = torch.tensor([0] + [clip_tokens] * (ranges * batch), dtype=torch.int64, device=x.device).cumsum(-1).to(torch.int32)
        cu_k = torch.cat([y_index.new_tensor([0]), y_index]).to(torch.int64).to(x.device).cumsum(-1).to(torch.int32)
        assert cu_q.shape == cu_k.shape
        q_ranges = torch.stack([cu_q[:-1], cu_q[1:]], dim=1)
        cross = PackedCrossAttnParams(q_ranges=q_ranges, kv_ranges=torch.stack([cu_k[:-1], cu_k[1:]], dim=1),
                                      cu_seqlens_q=cu_q, cu_seqlens_kv=cu_k, max_seqlen_q=clip_tokens,
                                      max_seqlen_kv=self.caption_max_length)
        # core-attention ranges (:243-258)
        flat_kv = torch.unique(kv_range, sorted=True)
        ardf_meta = dict(clip_token_nums=clip_tokens, slice_point=slice_point, range_num=range_num,
                         denoising_range_num=ranges, q_range=q_ranges.clone(), k_range=kv_range,
                         max_seqlen_q=clip_tokens, max_seqlen_k=int(flat_kv[-1] - flat_kv[0]))
        return x, condition, condition_map, rope, y_flat, None, H, W, ardf_meta, cross

    @torch.no_grad()
    def forward_pre_process(self, x, t, y, caption_dropout_mask=None, xattn_mask=None, kv_range=None, **kwargs):
        """:262-335."""
        assert kv_range is not None, "Please ensure kv_range is provided"
        mc = self.model_config
        x = x * mc.x_rescale_factor
        if mc.half_channel_vae:
            assert x.shape[1] == 16
            x = torch.cat([x, x], dim=1)
        (x, condition, condition_map, rope, y_flat, _mask_cg, H, W, am, cross) = self.get_embedding_and_meta(
            x.float(), t.float(), y.float(), caption_dropout_mask, xattn_mask, kv_range, **kwargs)
        x = x.to(mc.params_dtype)
        n, c, T = x.shape[0], x.shape[1], x.shape[2]
        x = x.permute(2, 3, 4, 0, 1).reshape(T * H * W, n, c).contiguous()          # 'N C T H W -> (T H W) N C'
        condition, y_flat = condition.to(mc.params_dtype), y_flat.to(mc.params_dtype)
        core = PackedCoreAttnParams(q_range=am["q_range"], k_range=am["k_range"],
                                    np_q_range=am["q_range"].cpu().numpy(), np_k_range=am["k_range"].cpu().numpy(),
                                    max_seqlen_q=am["max_seqlen_q"], max_seqlen_k=am["max_seqlen_k"])
        cp_size = max(1, getattr(self.engine_config, "cp_size", 1))
        cp_pad, cp_split = None, None
        if cp_size > 1:                                                             # cp_pre_process (:309-363)
            if self.engine_config.cp_strategy != "cp_ulysses":
                raise ValueError(f"Invalid CP strategy: {self.engine_config.cp_strategy}, expected cp_ulysses")
            x, condition_map, rope, cp_split, (xq, xk) = magi_cp.cp_ulysses_process(
                cp_size, x, condition_map, rope, cross.cu_seqlens_q.tolist(), cross.cu_seqlens_kv.tolist())
            cp_pad = 0
            cross = PackedCrossAttnParams(q_ranges=xq, kv_ranges=xk, cu_seqlens_q=cross.cu_seqlens_q,
                                          cu_seqlens_kv=cross.cu_seqlens_kv, max_seqlen_q=cp_split[magi_cp.get_cp_rank()],
                                          max_seqlen_kv=cross.max_seqlen_kv)
        meta = ModelMetaArgs(H=H, W=W, cp_pad_size=cp_pad, cp_split_sizes=cp_split, slice_point=am["slice_point"],
                             denoising_range_num=am["denoising_range_num"], range_num=am["range_num"],
                             extract_prefix_video_feature=kwargs.get("extract_prefix_video_feature", False),
                             fwd_extra_1st_chunk=kwargs["fwd_extra_1st_chunk"],
                             distill_nearly_clean_chunk=kwargs.get("distill_nearly_clean_chunk", False),
                             clip_token_nums=am["clip_token_nums"], enable_cuda_graph=False, core_attn_params=core,
                             cross_attn_params=cross)
        return x, condition, condition_map, y_flat, rope, meta

    @torch.no_grad()
    def forward_post_process(self, x, meta_args: ModelMetaArgs) -> torch.Tensor:
        """:337-360."""
        x = self.final_linear(x.float())
        cp_size = max(1, getattr(self.engine_config, "cp_size", 1))
        x = magi_cp.cp_post_process(cp_size, getattr(self.engine_config, "cp_strategy", "none"), x,
                                    meta_args.cp_split_sizes)
        x = self.unpatchify(x, meta_args.H, meta_args.W)
        if self.model_config.half_channel_vae:
            assert x.shape[1] == 32
            x = x[:, :16]
        return x / self.model_config.x_rescale_factor

    @torch.no_grad()
    def forward(self, x, t, y, caption_dropout_mask=None, xattn_mask=None, kv_range=None, inference_params=None,
                **kwargs) -> torch.Tensor:
        """:362-397 (single pipeline stage)."""
        x, condition, condition_map, y_flat, rope, meta = self.forward_pre_process(
            x, t, y, caption_dropout_mask, xattn_mask, kv_range, **kwargs)
        s_len, n = x.shape[0], x.shape[1]
        block_meta = meta
        if n > 1:
            # Batch > 1 only occurs in the unconditional CFG pass (one denoising range per sample, no cache,
            # dit_module.py:1000-1007).  N samples with one range each are the same computation as ONE sequence of N
            # ranges in which every range attends itself, so the batch is folded into the range dimension: tokens
            # sample-major (the order get_xqkv uses for the cross-attention segments, :956), the per-sample rotary table
            # tiled, condition rows b*R + r kept (condition_map already numbers them that way, dit_model.py:203-205).
            assert meta.denoising_range_num == 1 and inference_params is None, \
                "batch > 1 is supported for the cache-less single-range pass only (reference asserts the same)"
            if max(1, getattr(self.engine_config, "cp_size", 1)) > 1:
                raise NotImplementedError("batch > 1 under context parallel")
            (q0, q1), (k0, k1) = meta.core_attn_params.np_q_range[0], meta.core_attn_params.np_k_range[0]
            qr = np.asarray([[j * s_len + int(q0), j * s_len + int(q1)] for j in range(n)])
            kr = np.asarray([[j * s_len + int(k0), j * s_len + int(k1)] for j in range(n)])
            core = PackedCoreAttnParams(q_range=torch.as_tensor(qr, dtype=torch.int32), np_q_range=qr,
                                        k_range=torch.as_tensor(kr, dtype=torch.int32), np_k_range=kr,
                                        max_seqlen_q=meta.core_attn_params.max_seqlen_q,
                                        max_seqlen_k=meta.core_attn_params.max_seqlen_k)
            block_meta = ModelMetaArgs(**{**meta.__dict__, "denoising_range_num": n, "core_attn_params": core})
            x = x.permute(1, 0, 2).reshape(n * s_len, 1, -1)
            condition = condition.reshape(1, -1, condition.shape[-1])
            condition_map = condition_map.transpose(0, 1).reshape(n * s_len, 1)
            rope = rope.repeat(n, 1)
        x = self.videodit_blocks(hidden_states=x.contiguous().clone(), condition=condition, condition_map=condition_map,
                                 y_xattn_flat=y_flat, rotary_pos_emb=rope, inference_params=inference_params,
                                 meta_args=block_meta)
        if n > 1:
            x = x.reshape(n, s_len, -1).permute(1, 0, 2).contiguous()
        return self.forward_post_process(x, meta)

    # ------------------------------------------------------------------ classifier-free guidance (:399-596)
    def forward_3cfg(self, x, timestep, y, mask, kv_range, inference_params, **kwargs):
        """:399-487 — (text + previous chunks), (previous chunks only, also the pass that stores K/V), and the
        unconditional pass in which every denoising chunk becomes its own batch entry with no cache."""
        assert x.shape[0] == 2 and mask.shape[0] % 2 == 0
        x = torch.cat([x[0:1], x[0:1]], dim=0)
        drop = torch.tensor([False, True], dtype=torch.bool, device=x.device)
        half = y.shape[0] // 2
        inference_params.update_kv_cache = False
        out_text = self.forward(x[0:1], timestep[0:1], y[:half], caption_dropout_mask=drop[0:1], xattn_mask=mask[:half],
                                kv_range=kv_range, inference_params=inference_params, **kwargs)
        inference_params.update_kv_cache = True
        out_prev = self.forward(x[1:2], timestep[1:2], y[half:], caption_dropout_mask=drop[1:2], xattn_mask=mask[half:],
                                kv_range=kv_range, inference_params=inference_params, **kwargs)
        saved = {k: kwargs[k] for k in ("range_num", "denoising_range_num", "slice_point", "fwd_extra_1st_chunk")}
        try:
            if kwargs.get("fwd_extra_1st_chunk", False):        # the extra clean chunk takes no part in the uncond pass
                kwargs["denoising_range_num"] -= 1
                kwargs["slice_point"] += 1
                kwargs["fwd_extra_1st_chunk"] = False
            r, cw = kwargs["denoising_range_num"], kwargs["chunk_width"]
            denoise_width = cw * r
            ux = x[0:1, :, -denoise_width:].squeeze(0)
            ux = ux.reshape(-1, r, cw, *ux.shape[2:]).transpose(0, 1)              # (ranges, C, chunk_width, h, w)
            ut = timestep[0:1, -r:].transpose(0, 1)
            kwargs["range_num"], kwargs["denoising_range_num"], kwargs["slice_point"] = 1, 1, 0
            out_uncond = self.forward(ux, ut, y[half:][-r:], caption_dropout_mask=torch.tensor([True], device=x.device),
                                      xattn_mask=mask[half:][-r:],
                                      kv_range=self.generate_kv_range_for_uncondition(ux), inference_params=None,
                                      **kwargs)
            out_uncond = out_uncond.transpose(0, 1)
            out_uncond = out_uncond.reshape(1, -1, r * cw, *out_uncond.shape[3:])
        finally:
            kwargs.update(saved)
        return out_text, out_prev, out_uncond, denoise_width

    def get_cfg_scale(self, t, cfg_t_range, prev_chunk_scale_s, text_scale_s):
        idx = torch.searchsorted(cfg_t_range - 1e-7, t) - 1
        assert idx.min() >= 0 and idx.max() < len(prev_chunk_scale_s)
        return prev_chunk_scale_s[idx], text_scale_s[idx]

    def forward_dispatcher(self, x, timestep, y, mask, kv_range, inference_params, **kwargs):
        """:494-596.  x [2N, C, T, H, W] (the two CFG copies); returns the velocity in the same layout."""
        rc = self.runtime_config
        dev = x.device
        if rc.cfg_number == 3:
            out_text, out_prev, out_uncond, denoise_width = self.forward_3cfg(x, timestep, y, mask, kv_range,
                                                                              inference_params, **kwargs)
            prev_s = torch.tensor(rc.prev_chunk_scales, device=dev)
            text_s = torch.tensor(rc.text_scales, device=dev)
            t_range = torch.tensor(rc.cfg_t_range, device=dev)
            assert len(prev_s) == len(t_range) and len(text_s) == len(t_range)
            n, cw = kwargs["denoising_range_num"], kwargs["chunk_width"]
            if kwargs["fwd_extra_1st_chunk"]:
                n -= 1
            cfg_t = timestep[0, -n:]
            parts = []
            for i in range(n):
                ps, ts = self.get_cfg_scale(cfg_t[i], t_range, prev_s, text_s)
                sl = slice(i * cw, (i + 1) * cw)
                parts.append((1 - ps) * out_uncond[:, :, sl] + (ps - ts) * out_prev[:, :, -denoise_width:][:, :, sl]
                             + ts * out_text[:, :, -denoise_width:][:, :, sl])
            out = torch.cat([x[0:1, :, :-denoise_width], torch.cat(parts, dim=2)], dim=2)
            return torch.cat([out, out], dim=0)
        if rc.cfg_number == 1:
            assert x.shape[0] == 2
            x = torch.cat([x[0:1], x[0:1]], dim=0)
            half = y.shape[0] // 2
            kwargs["caption_dropout_mask"] = torch.tensor([False], dtype=torch.bool, device=dev)
            inference_params.update_kv_cache = True
            cw = kwargs["chunk_width"]
            if kwargs.get("distill_nearly_clean_chunk", False):                 # :540-575
                scale = float(os.getenv("prev_chunks_scale", 0.7))
                s0 = 1 if kwargs["fwd_extra_1st_chunk"] else 0
                width = x.shape[2]
                extra = x[0:1, :, s0 * cw:(s0 + 1) * cw]
                kwargs["denoising_range_num"] += 1
                cat_kv = torch.cat([kv_range, self.generate_kv_range_for_uncondition(extra) + kv_range.max()], dim=0)
                cat_out = self.forward(torch.cat([x[0:1], extra], dim=2),
                                       torch.cat([timestep[0:1], timestep[0:1, s0:s0 + 1]], dim=1),
                                       torch.cat([y[:half], y[s0:s0 + 1]], dim=0),
                                       xattn_mask=torch.cat([mask[:half], mask[s0:s0 + 1]], dim=0), kv_range=cat_kv,
                                       inference_params=inference_params, **kwargs)
                with_prev = cat_out[:, :, s0 * cw:(s0 + 1) * cw]
                text_only = cat_out[:, :, width:]
                cat_out[:, :, s0 * cw:(s0 + 1) * cw] = with_prev * scale + text_only * (1 - scale)
                out = cat_out[:, :, :width]
            else:
                out = self.forward(x[0:1], timestep[0:1], y[:half], xattn_mask=mask[:half], kv_range=kv_range,
                                   inference_params=inference_params, **kwargs)
            denoise_width = cw * kwargs["denoising_range_num"]
            if kwargs["fwd_extra_1st_chunk"]:
                denoise_width -= cw
            x = torch.cat([x[0:1, :, :-denoise_width], out[:, :, -denoise_width:]], dim=2)
            return torch.cat([x[0:1], x[0:1]], dim=0)
        raise NotImplementedError(f"cfg_number {rc.cfg_number}")
