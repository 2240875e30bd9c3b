"""MAGI-1 transformer layer / block with the reference's module surface on the native kernels
(inferix/models/magi/dit/dit_module.py:180-201 AdaModulateLayer, :326-360 FusedLayerNorm, :393-431
CustomLayerNormLinear, :496-556 CustomMLP, :833-1198 FullyParallelAttention, :1201-1318 TransformerLayer,
:1322-1390 TransformerBlock).

Same class names, constructor arguments (`model_config`, `engine_config`, `layer_number`), parameter names (reference
checkpoints load with `load_state_dict`) and forward signatures.  One layer forward is 13-15 kernels of
libinferix_b200.so:

    LN(affine)                         ifx_ln_modulate                                   (:415)
    q | k | v | qx projection          ifx_gemm_bf16, ONE GEMM over the four concatenated weights   (:418-431)
    head-LN + rotary + KV write        ifx_magi_qkv_post  (K / V land in the layer's cache rows)    (:902-958, kv :76-151)
    core attention per denoising range ifx_attention_gqa  (grouped-query, keys = cache rows)        (:972-1015)
    caption K | V projection, k-LN     ifx_gemm_bf16 + ifx_head_layernorm                           (:960-968)
    cross attention per range          ifx_attention_gqa                                            (:1047-1085)
    output projection                  ifx_gemm_bf16 on [core | cross] (weight columns pre-permuted for :1287), fp32
                                       result (IFX_EPI_BIAS_F32) as under the reference's autocast(float32) (:1291-1293)
    gate * x -> LN -> + residual       ifx_gate_norm_residual                                       (:295-313)
    MLP: LN, fc1 (+GELU-erf epilogue | + ifx_silu_mul), fc2, gate/LN/residual                       (:545-556)

What stays on torch ops: the AdaModulate gate of `denoising_range_num` rows (SiLU -> Linear -> softcap, :196-198,
:1299-1303) and the final fp32 LayerNorm of the block (once per forward), both < 0.1 % of the layer's bytes.

Context parallel: `engine_config.cp_strategy == "cp_ulysses"` (see inferix_b200/magi_cp.py).  Not built: batch > 1,
cp_shuffle_overlap,
pipeline parallel offsets.  There is no CPU path: every op below needs libinferix_b200.so and CUDA tensors.
"""
from __future__ import annotations

import numbers
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import magi_cp
from . import ops as _ops
from .kvcache_manager.model.magi_kv_cache_manager import MagiKVCacheManager


def softcap(x: torch.Tensor, cap: float) -> torch.Tensor:
    """dit_module.py:363-364."""
    return (cap * torch.tanh(x.float() / cap)).to(x.dtype)


class FusedLayerNorm(nn.Module):
    """dit_module.py:326-360 (parameters only; the fused kernels read them)."""

    def __init__(self, model_config, hidden_size, dtype=None):
        super().__init__()
        self.zero_centered_gamma = model_config.apply_layernorm_1p
        if isinstance(hidden_size, numbers.Integral):
            hidden_size = (hidden_size,)
        self.hidden_size = torch.Size(hidden_size)
        self.eps = model_config.layernorm_epsilon
        dtype = dtype or model_config.params_dtype
        self.weight = nn.Parameter(torch.ones(*hidden_size, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(*hidden_size, dtype=dtype))

    def affine(self):
        w = self.weight + 1 if self.zero_centered_gamma else self.weight
        return w.detach().contiguous(), self.bias.detach().contiguous()

    def forward(self, x):
        w, b = self.affine()
        return F.layer_norm(x, self.hidden_size, w, b, self.eps)


class AdaModulateLayer(nn.Module):
    """dit_module.py:180-201."""

    def __init__(self, model_config):
        super().__init__()
        h = model_config.hidden_size
        self.gate_num_chunks = 2
        self.act = nn.SiLU()
        self.proj = nn.Sequential(nn.Linear(int(h * model_config.cond_hidden_ratio),
                                            int(h * model_config.cond_gating_ratio * self.gate_num_chunks), bias=True,
                                            dtype=model_config.params_dtype))

    def forward(self, c):
        return self.proj(self.act(c))


def _fp8_layer(engine_config, model_config, layer_number: int) -> bool:
    """Which layers carry FP8 linears: fp8_quant set and neither the first nor the last layer (dit_module.py:410)."""
    return bool(getattr(engine_config, "fp8_quant", False)) and layer_number not in (0, model_config.num_layers - 1)


class PerTensorQuantizedFp8Linear(nn.Module):
    """dit_module.py:434-459 — parameter names / shapes of the reference (a quantised checkpoint loads as is):
    weight e4m3 [1, out, in], weight_scale fp32 [1], input_scale fp32 [in].  forward = div_clamp_to(x, input_scale)
    (one divisor per input channel) followed by bmm_fp8 with per-tensor scales, for which cuBLASLt reads ONE float per
    operand: input_scale[0] and weight_scale[0]."""

    def __init__(self, in_features, out_features, bias=False, dtype=torch.bfloat16, device=None):
        super().__init__()
        self.in_features, self.out_features, self.output_dtype = in_features, out_features, dtype
        self.weight = nn.Parameter(torch.zeros((1, out_features, in_features), dtype=torch.float8_e4m3fn), requires_grad=False)
        self.weight_scale = nn.Parameter(torch.ones(1, dtype=torch.float32), requires_grad=False)
        self.input_scale = nn.Parameter(torch.ones(in_features, dtype=torch.float32), requires_grad=False)

    def divisor(self):
        return self.input_scale

    def alpha(self) -> float:
        return float(self.input_scale.reshape(-1)[0]) * float(self.weight_scale.reshape(-1)[0])

    @torch.no_grad()
    def quantize_from(self, weight: torch.Tensor, input_amax: float = 8.0):
        """Test / tooling helper: fill the parameters from a bf16 weight (per-tensor amax scaling)."""
        ws = float(weight.float().abs().max()) / 448.0
        self.weight.copy_(torch.clamp(weight.float() / ws, -448.0, 448.0).bfloat16().to(torch.float8_e4m3fn).unsqueeze(0))
        self.weight_scale.fill_(ws)
        self.input_scale.fill_(input_amax / 448.0)

    def forward(self, x):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        a = _ops.quantize_fp8_cols(x2, self.divisor().reshape(-1).contiguous())
        return _ops.gemm_fp8(a, self.weight[0], self.alpha()).view(*shp[:-1], self.out_features)


class PerChannelQuantizedFp8Linear(PerTensorQuantizedFp8Linear):
    """dit_module.py:465-490: x / smooth_scale [1, in] -> e4m3, bmm_fp8 with input_scale [1] and weight_scale [1]."""

    def __init__(self, in_features, out_features, bias=False, dtype=torch.bfloat16, device=None):
        super().__init__(in_features, out_features, bias, dtype, device)
        self.input_scale = nn.Parameter(torch.ones(1, dtype=torch.float32), requires_grad=False)
        self.smooth_scale = nn.Parameter(torch.ones((1, in_features), dtype=torch.float32), requires_grad=False)

    def divisor(self):
        return self.smooth_scale

    @torch.no_grad()
    def quantize_from(self, weight: torch.Tensor, input_amax: float = 8.0, smooth=None):
        """Smooth-quant style split: activations are divided by smooth_scale[k] * 1, the weight absorbs smooth / input."""
        k = weight.shape[1]
        sm = torch.ones(k) if smooth is None else smooth.float().cpu()
        in_s = input_amax / 448.0
        self.smooth_scale.copy_((sm * in_s).view(1, k))          # x / (smooth * input_scale) lands in the e4m3 range
        w_eff = weight.float().cpu() * sm.view(1, k)             # (x / smooth) @ (W * smooth)^T == x @ W^T
        ws = float(w_eff.abs().max()) / 448.0
        self.weight.copy_(torch.clamp(w_eff / ws, -448.0, 448.0).bfloat16().to(torch.float8_e4m3fn).unsqueeze(0))
        self.weight_scale.fill_(ws)
        self.input_scale.fill_(in_s)


def _is_fp8(lin) -> bool:
    return isinstance(lin, PerTensorQuantizedFp8Linear)


class CustomLayerNormLinear(nn.Module):
    """dit_module.py:393-431 (parameter container: layer_norm + q / qx / k / v)."""

    def __init__(self, input_size, output_size_q, output_size_kv, layer_number, model_config, engine_config):
        super().__init__()
        dt = model_config.params_dtype
        self.layer_norm = nn.LayerNorm(input_size, eps=model_config.layernorm_epsilon, dtype=dt)
        self.layer_number = layer_number
        fp8 = _fp8_layer(engine_config, model_config, layer_number)
        for name, out in {"q": output_size_q, "qx": output_size_q, "k": output_size_kv, "v": output_size_kv}.items():
            setattr(self, name, PerTensorQuantizedFp8Linear(input_size, out) if fp8
                    else nn.Linear(input_size, out, bias=False, dtype=dt))


class CustomMLP(nn.Module):
    """dit_module.py:496-556 (parameter container)."""

    def __init__(self, model_config, engine_config, layer_number, input_size=None):
        super().__init__()
        dt = model_config.params_dtype
        self.input_size = input_size if input_size is not None else model_config.hidden_size
        self.layer_norm = nn.LayerNorm(self.input_size, eps=model_config.layernorm_epsilon, dtype=dt)
        f = model_config.ffn_hidden_size
        fp8 = _fp8_layer(engine_config, model_config, layer_number)
        fc1_out = (2 if model_config.gated_linear_unit else 1) * f
        self.linear_fc1 = (PerTensorQuantizedFp8Linear(self.input_size, fc1_out) if fp8                      # :525-536
                           else nn.Linear(self.input_size, fc1_out, bias=False, dtype=dt))
        self.linear_fc2 = (PerChannelQuantizedFp8Linear(f, model_config.hidden_size) if fp8                   # :538-543
                           else nn.Linear(f, model_config.hidden_size, bias=False, dtype=dt))


class FullyParallelAttention(nn.Module):
    """dit_module.py:779-879 (parameter container + the per-layer KV adapter)."""

    def __init__(self, model_config, engine_config, layer_number):
        super().__init__()
        mc = model_config
        self.layer_number = layer_number
        self.hidden_size_per_attention_head = mc.kv_channels
        self.query_projection_size = mc.kv_channels * mc.num_attention_heads
        self.kv_projection_size = mc.kv_channels * mc.num_query_groups
        cp = max(1, getattr(engine_config, "cp_size", 1))
        if mc.num_query_groups % cp:
            raise NotImplementedError("num_query_groups must divide by cp_size (the reference repeats KV heads when "
                                      "cp_size > num_query_groups, context_parallel.py:394-395; not built)")
        self.num_query_groups_per_partition = mc.num_query_groups // cp
        self.kv_cache_manager = MagiKVCacheManager(layer_number=layer_number,
                                                   num_query_groups_per_partition=self.num_query_groups_per_partition,
                                                   hidden_size_per_attention_head=mc.kv_channels,
                                                   engine_config=engine_config)
        dt = mc.params_dtype
        self.linear_qkv = CustomLayerNormLinear(mc.hidden_size, self.query_projection_size, self.kv_projection_size,
                                                layer_number, mc, engine_config)
        self.linear_kv_xattn = nn.Linear(int(mc.hidden_size * mc.xattn_cond_hidden_ratio), 2 * self.kv_projection_size,
                                         dtype=dt, bias=False)
        self.adapt_linear_quant = _fp8_layer(engine_config, mc, layer_number)                               # :864-866
        self.linear_proj = (PerChannelQuantizedFp8Linear(2 * self.query_projection_size, mc.hidden_size)
                            if self.adapt_linear_quant
                            else nn.Linear(2 * self.query_projection_size, mc.hidden_size, dtype=dt, bias=False))
        # dtypes as left by _high_precision_promoter (dit_model.py:620-637): self-attention q/k norms in fp32
        self.q_layernorm = FusedLayerNorm(mc, mc.kv_channels, dtype=torch.float32)
        self.q_layernorm_xattn = FusedLayerNorm(mc, mc.kv_channels)
        self.k_layernorm = FusedLayerNorm(mc, mc.kv_channels, dtype=torch.float32)
        self.k_layernorm_xattn = FusedLayerNorm(mc, mc.kv_channels)


class _Scratch:
    """Named scratch buffers shared by the layers of a block (grown on demand, never shrunk)."""

    def __init__(self):
        self._buf = {}

    def get(self, name, shape, dtype, device):
        n = 1
        for s in shape:
            n *= int(s)
        t = self._buf.get(name)
        if t is None or t.numel() < n or t.dtype != dtype or t.device != device:
            t = self._buf[name] = torch.empty(max(n, 1), dtype=dtype, device=device)
        return t[:n].view(*shape)


class _FwdCtx:
    """Per-forward state shared by all layers: int32 row map, contiguous fp32 rope, host ranges, scratch."""

    def __init__(self, condition_map, rotary_pos_emb, meta_args, scratch: _Scratch):
        if condition_map.shape[1] != 1:
            raise NotImplementedError("batch size > 1 (the reference folds it into the sequence; not built)")
        self.row_map = condition_map[:, 0].to(torch.int32).contiguous()
        self.rope = rotary_pos_emb.to(torch.float32).contiguous()
        self.scratch = scratch
        cap = meta_args.core_attn_params
        self.q_ranges = [[int(a), int(b)] for a, b in cap.np_q_range]
        self.k_ranges = [[int(a), int(b)] for a, b in cap.np_k_range]
        xp = meta_args.cross_attn_params
        q_r, k_r = getattr(xp, "q_ranges", None), getattr(xp, "kv_ranges", None)
        if q_r is None or k_r is None:
            cu_q, cu_k = xp.cu_seqlens_q.tolist(), xp.cu_seqlens_kv.tolist()
            q_r, k_r = list(zip(cu_q[:-1], cu_q[1:])), list(zip(cu_k[:-1], cu_k[1:]))
        else:
            q_r = q_r.tolist() if hasattr(q_r, "tolist") else q_r
            k_r = k_r.tolist() if hasattr(k_r, "tolist") else k_r
        self.xq_ranges = [[int(a), int(b)] for a, b in q_r]
        self.xk_ranges = [[int(a), int(b)] for a, b in k_r]


class TransformerLayer(nn.Module):
    """dit_module.py:1201-1318."""

    def __init__(self, model_config, engine_config, layer_number: int = 1):
        super().__init__()
        if getattr(engine_config, "cp_strategy", "none") not in ("none", "cp_ulysses"):
            raise NotImplementedError("cp_strategy must be 'none' or 'cp_ulysses'")
        if model_config.kv_channels != 128:
            raise NotImplementedError("the native kernels are built for kv_channels = 128 (every MAGI-1 model)")
        if model_config.params_dtype != torch.bfloat16:
            raise ValueError("the native layer computes in bf16 (the reference's params_dtype for inference)")
        self.model_config, self.engine_config, self.layer_number = model_config, engine_config, layer_number
        self.ada_modulate_layer = AdaModulateLayer(model_config)
        self.self_attention = FullyParallelAttention(model_config, engine_config, layer_number)
        self.self_attn_post_norm = FusedLayerNorm(model_config, model_config.hidden_size, dtype=torch.float32)
        self.mlp = CustomMLP(model_config, engine_config, layer_number)
        self.mlp_post_norm = FusedLayerNorm(model_config, model_config.hidden_size, dtype=torch.float32)
        self._packed = None
        self._stored_end = 0       # rows [0, _stored_end) of the KV cache hold tokens stored by update_kv_cache forwards
        self._own_scratch = None
        self.register_load_state_dict_post_hook(lambda m, _k: m.invalidate_packed())

    def invalidate_packed(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._packed = None
        return out

    def _pack(self):
        """Launch-ready weights: q|k|v|qx in one matrix; caption K/V rows regrouped to K-all-heads | V-all-heads; output
        projection columns permuted from the reference's '(hn n hd)' input order (dit_module.py:1287) to
        '[core | cross]' so the concat + rearrange of :1285-1287 never materialises."""
        sa, mc = self.self_attention, self.model_config
        lq = sa.linear_qkv
        d, g = mc.kv_channels, mc.num_query_groups
        fp8 = _is_fp8(lq.q)

        def proj_perm(wp):
            """output-projection input columns: reference '(hn n hd)' order (dit_module.py:1287) -> '[core | cross]'"""
            hd = wp.shape[1] // 16
            return wp.view(wp.shape[0], 8, 2, hd).permute(0, 2, 1, 3).reshape(wp.shape[0], -1).contiguous()

        def q8(lin, perm=None):
            """(e4m3 weight [out, in], per-input-channel divisor fp32 [in], alpha) of a quantised linear"""
            wq = lin.weight.detach()[0]
            div = lin.divisor().detach().reshape(-1).float()
            if perm is not None:                     # permute input channels of weight and divisor alike
                wq = perm(wq.view(torch.uint8)).view(torch.float8_e4m3fn)
                div = perm(div.view(1, -1)).reshape(-1)
            return wq.contiguous(), div.contiguous(), lin.alpha()

        w_kvx = sa.linear_kv_xattn.weight.detach()
        w_kvx = w_kvx.view(g, 2, d, w_kvx.shape[1]).permute(1, 0, 2, 3).reshape(2 * g * d, -1).contiguous()
        pk = dict(
            w_kvx=w_kvx, fp8=fp8,
            ln1=(lq.layer_norm.weight.detach().contiguous(), lq.layer_norm.bias.detach().contiguous()),
            q_ln=sa.q_layernorm.affine(), k_ln=sa.k_layernorm.affine(),
            qx_ln=sa.q_layernorm_xattn.affine(), kx_ln=sa.k_layernorm_xattn.affine(),
            post1=self.self_attn_post_norm.affine(), post2=self.mlp_post_norm.affine(),
            ln2=(self.mlp.layer_norm.weight.detach().contiguous(), self.mlp.layer_norm.bias.detach().contiguous()),
            ada_w=self.ada_modulate_layer.proj[0].weight.detach(), ada_b=self.ada_modulate_layer.proj[0].bias.detach())
        if fp8:
            # four PerTensor linears with their own divisors / scales (:410-413), PerChannel proj and fc2, PerTensor fc1
            pk["qkvx8"] = [q8(getattr(lq, n)) for n in ("q", "k", "v", "qx")]
            pk["proj8"] = q8(sa.linear_proj, proj_perm)
            pk["fc1_8"], pk["fc2_8"] = q8(self.mlp.linear_fc1), q8(self.mlp.linear_fc2)
        else:
            pk["w_qkvx"] = torch.cat([lq.q.weight, lq.k.weight, lq.v.weight, lq.qx.weight], dim=0).detach().contiguous()
            pk["w_proj"] = proj_perm(sa.linear_proj.weight.detach())
            pk["fc1"] = self.mlp.linear_fc1.weight.detach().contiguous()
            pk["fc2"] = self.mlp.linear_fc2.weight.detach().contiguous()
        self._packed = pk
        return self._packed

    # ------------------------------------------------------------------ KV rows
    def reset_kv_state(self):
        self._stored_end = 0

    def _kv_rows(self, ctx, inference_params, meta_args, n_new, kv_width, device):
        """Where this forward's K / V rows go and which rows attention reads.  Returns
        (k_all, v_all, start, restore) with k_all / v_all [start + n_new, kv_width]; the new tokens occupy rows
        [start, start + n_new).  Mirrors magi_kv_cache_manager.py:153-187 / :76-151 without the get_range + cat copies;
        rows of earlier update_kv_cache forwards that the scratch write would cover are saved and put back."""
        uses_cache = inference_params is not None and (meta_args.extract_prefix_video_feature or
                                                       meta_args.fwd_extra_1st_chunk or meta_args.slice_point > 0)
        if not uses_cache:
            k = ctx.scratch.get("k_plain", (n_new, kv_width), torch.bfloat16, device)
            v = ctx.scratch.get("v_plain", (n_new, kv_width), torch.bfloat16, device)
            return k, v, 0, None
        store = self.self_attention.kv_cache_manager.native_store(inference_params)
        start = meta_args.slice_point * meta_args.clip_token_nums * inference_params.max_batch_size
        if start + n_new > inference_params.max_sequence_length:
            raise IndexError(f"KV rows [{start}, {start + n_new}) beyond max_sequence_length "
                             f"{inference_params.max_sequence_length}")
        k_all, v_all = store.map_rows(start + n_new)
        keep = 0
        if inference_params.update_kv_cache:                                    # :127-146
            keep = (n_new - meta_args.clip_token_nums * inference_params.max_batch_size
                    if meta_args.distill_nearly_clean_chunk else n_new)
        lo, hi = start + keep, min(self._stored_end, start + n_new)
        restore = None
        if hi > lo:
            restore = (lo, hi, k_all[lo:hi].clone(), v_all[lo:hi].clone())
        if keep > 0:
            self._stored_end = max(self._stored_end, start + keep)
        return k_all, v_all, start, restore

    # ------------------------------------------------------------------ forward
    def forward(self, hidden_states, condition, condition_map, y_xattn_flat, rotary_pos_emb, inference_params,
                meta_args, _ctx: Optional[_FwdCtx] = None):
        """hidden_states [s, 1, h] bf16 (this rank's tokens under cp_ulysses) -> same shape."""
        mc = self.model_config
        if hidden_states.dim() != 3 or hidden_states.shape[1] != 1:
            raise NotImplementedError("hidden_states must be [s, 1, h] (batch size 1)")
        if _ctx is None:
            if self._own_scratch is None:
                self._own_scratch = _Scratch()
            _ctx = _FwdCtx(condition_map, rotary_pos_emb, meta_args, self._own_scratch)
        ctx, pk = _ctx, (self._packed or self._pack())
        dev, bf = hidden_states.device, torch.bfloat16
        s_loc, h = hidden_states.shape[0], mc.hidden_size
        d, hq, g, eps = mc.kv_channels, mc.num_attention_heads, mc.num_query_groups, mc.layernorm_epsilon
        cp = magi_cp.get_cp_world_size() if getattr(self.engine_config, "cp_strategy", "none") == "cp_ulysses" else 1
        qg, kg = hq // cp, g // cp
        x = hidden_states.reshape(s_loc, h)
        if not x.is_contiguous():
            x = x.contiguous()
        sc = ctx.scratch

        # ---- LN -> fused q|k|v|qx projection (:1103, :418-431)
        hbuf = sc.get("h", (s_loc, h), bf, dev)
        _ops.ln_modulate(x, hbuf, weight=pk["ln1"][0], bias=pk["ln1"][1], eps=eps)
        qkvx = sc.get("qkvx", (s_loc, (2 * hq + 2 * g) * d), bf, dev)
        if pk["fp8"]:
            off = 0
            for wq, div, alpha in pk["qkvx8"]:                    # q | k | v | qx, each with its own divisor and scales
                a8 = _ops.quantize_fp8_cols(hbuf, div, sc.get("a8_h", (s_loc, h), torch.float8_e4m3fn, dev))
                _ops.gemm_fp8(a8, wq, alpha, None, qkvx[:, off:off + wq.shape[0]])
                off += wq.shape[0]
        else:
            _ops.gemm(hbuf, pk["w_qkvx"], None, qkvx)

        # ---- head-LN + rotary + KV placement (:902-958; cache :76-151)
        qx = sc.get("qx", (s_loc, hq * d), bf, dev)
        if cp == 1:
            s_tot, splits = s_loc, None
            k_all, v_all, start, restore = self._kv_rows(ctx, inference_params, meta_args, s_tot, g * d, dev)
            q = sc.get("q", (s_loc, hq * d), bf, dev)
            _ops.magi_qkv_post(qkvx, hq, g, pk["q_ln"], pk["k_ln"], pk["qx_ln"], ctx.rope, q,
                               k_all[start:start + s_tot], v_all[start:start + s_tot], qx, eps=eps)
            handles = ()
        else:
            splits = list(meta_args.cp_split_sizes)
            s_tot = sum(splits)
            k_all, v_all, start, restore = self._kv_rows(ctx, inference_params, meta_args, s_tot, kg * d, dev)
            q_send = sc.get("q_send", (cp, s_loc, qg * d), bf, dev)
            k_send = sc.get("k_send", (cp, s_loc, kg * d), bf, dev)
            v_send = sc.get("v_send", (cp, s_loc, kg * d), bf, dev)
            _ops.magi_qkv_post(qkvx, hq, g, pk["q_ln"], pk["k_ln"], pk["qx_ln"], ctx.rope, q_send, k_send, v_send, qx,
                               eps=eps, groups=cp)
            # scatter heads / gather sequence; K and V are received straight into the cache rows
            _, hk = magi_cp.all_to_all_input_split(k_send, splits, out=k_all[start:start + s_tot])
            _, hv = magi_cp.all_to_all_input_split(v_send, splits, out=v_all[start:start + s_tot])
            q, hq_ = magi_cp.all_to_all_input_split(q_send, splits, out=sc.get("q", (s_tot, qg * d), bf, dev))
            handles = (hk, hv, hq_)

        # ---- cross attention on the local tokens (overlaps the exchange above) (:954-970, :1047-1085)
        attn_cat = sc.get("attn_cat", (s_loc, 2 * hq * d), bf, dev)
        n_y = y_xattn_flat.shape[0]
        kvx = sc.get("kvx", (n_y, 2 * g * d), bf, dev)
        _ops.gemm(y_xattn_flat, pk["w_kvx"], None, kvx)
        _ops.head_layernorm(kvx[:, :g * d], g, pk["kx_ln"][0], pk["kx_ln"][1], eps=eps)
        for (qs, qe), (ks, ke) in zip(ctx.xq_ranges, ctx.xk_ranges):
            if qe > qs:
                _ops.attention_gqa(qx[qs:qe], kvx[ks:ke, :g * d], kvx[ks:ke, g * d:], hq, g,
                                   attn_cat[qs:qe, hq * d:])

        # ---- core attention over the denoising ranges (:972-1015)
        for hdl in handles:
            hdl.wait()
        core_out = attn_cat[:, :hq * d] if cp == 1 else sc.get("core_full", (s_tot, qg * d), bf, dev)
        for (qs, qe), (ks, ke) in zip(ctx.q_ranges, ctx.k_ranges):
            _ops.attention_gqa(q[qs:qe], k_all[ks:ke], v_all[ks:ke], qg, kg, core_out[qs:qe])
        if restore is not None:
            lo, hi, k_old, v_old = restore
            k_all[lo:hi].copy_(k_old)
            v_all[lo:hi].copy_(v_old)
        if cp > 1:
            back, hb = magi_cp.all_to_all_output_split(core_out, splits)        # [cp, s_loc, qg*d], head-group-major
            hb.wait()
            attn_cat[:, :hq * d].unflatten(1, (cp, qg * d)).copy_(back.transpose(0, 1))

        # ---- output projection, gate, post-norm, residual (:1281-1311)
        # the reference runs this projection under autocast(float32) (:1291-1293): its result is consumed in fp32
        if pk["fp8"]:
            # adapt_linear_quant (:1288-1289): PerChannelQuantizedFp8Linear, bf16 result (no fp32 autocast region)
            wq, div, alpha = pk["proj8"]
            a8 = _ops.quantize_fp8_cols(attn_cat, div, sc.get("a8_attn", (s_loc, 2 * hq * d), torch.float8_e4m3fn, dev))
            proj32 = _ops.gemm_fp8(a8, wq, alpha, None, sc.get("proj_bf", (s_loc, h), bf, dev))
        else:
            proj32 = sc.get("proj32", (s_loc, h), torch.float32, dev)
            _ops.gemm(attn_cat, pk["w_proj"], None, proj32, epilogue=_ops.EPI_BIAS_F32)
        gate = softcap(F.linear(F.silu(condition.reshape(-1, condition.shape[-1])), pk["ada_w"], pk["ada_b"]), 1.0)
        gate = gate.to(bf).contiguous()                                          # [ranges, 2h]: gate_msa | gate_mlp
        x1 = torch.empty_like(x)
        _ops.gate_norm_residual(proj32, gate[:, :h], ctx.row_map, pk["post1"][0], pk["post1"][1], x, x1, eps=eps)

        # ---- MLP (:545-556) + second gate / post-norm / residual (:1313-1317)
        f = mc.ffn_hidden_size
        _ops.ln_modulate(x1, hbuf, weight=pk["ln2"][0], bias=pk["ln2"][1], eps=eps)
        act = sc.get("act", (s_loc, f), bf, dev)
        proj = sc.get("proj", (s_loc, h), bf, dev)
        if pk["fp8"]:
            wq, div, alpha = pk["fc1_8"]
            a8 = _ops.quantize_fp8_cols(hbuf, div, sc.get("a8_h", (s_loc, h), torch.float8_e4m3fn, dev))
            if mc.gated_linear_unit:
                ffn = sc.get("ffn", (s_loc, 2 * f), bf, dev)
                _ops.gemm_fp8(a8, wq, alpha, None, ffn)
                _ops.silu_mul(ffn, act)
            else:
                _ops.gemm_fp8(a8, wq, alpha, None, act, epilogue=_ops.EPI_BIAS_GELU_ERF)
            wq, div, alpha = pk["fc2_8"]
            a8 = _ops.quantize_fp8_cols(act, div, sc.get("a8_act", (s_loc, f), torch.float8_e4m3fn, dev))
            _ops.gemm_fp8(a8, wq, alpha, None, proj)
        elif mc.gated_linear_unit:
            ffn = sc.get("ffn", (s_loc, 2 * f), bf, dev)
            _ops.gemm(hbuf, pk["fc1"], None, ffn)
            _ops.silu_mul(ffn, act)
            _ops.gemm(act, pk["fc2"], None, proj)
        else:
            _ops.gemm(hbuf, pk["fc1"], None, act, epilogue=_ops.EPI_BIAS_GELU_ERF)
            _ops.gemm(act, pk["fc2"], None, proj)
        _ops.gate_norm_residual(proj, gate[:, h:], ctx.row_map, pk["post2"][0], pk["post2"][1], x1, x1, eps=eps)
        return x1.view(s_loc, 1, h)


class TransformerBlock(nn.Module):
    """dit_module.py:1322-1390: the layer stack and the fp32 final LayerNorm."""

    def __init__(self, model_config, engine_config, pre_process: bool = True, post_process: bool = True):
        super().__init__()
        self.model_config, self.engine_config = model_config, engine_config
        self.pre_process, self.post_process = pre_process, post_process
        self.input_tensor = None
        self.layers = nn.ModuleList([TransformerLayer(model_config, engine_config, layer_number=i)
                                     for i in range(model_config.num_layers)])
        if post_process:
            self.final_layernorm = FusedLayerNorm(model_config, model_config.hidden_size, dtype=torch.float32)
        self._scratch = _Scratch()

    def set_input_tensor(self, input_tensor):
        self.input_tensor = input_tensor

    def clear_kv_cache(self, inference_params):
        """Frees every layer's cache rows (the reference calls MagiKVCacheManager.clear_cache per layer)."""
        for layer in self.layers:
            mgr = layer.self_attention.kv_cache_manager
            if mgr.is_cached(inference_params):
                mgr.clear_cache(inference_params)
            layer.reset_kv_state()

    @torch.no_grad()
    def forward(self, hidden_states, condition, condition_map, y_xattn_flat, rotary_pos_emb, inference_params,
                meta_args):
        if not self.pre_process:
            assert self.input_tensor is not None, "please call set_input_tensor for pp"
            hidden_states = self.input_tensor
        ctx = _FwdCtx(condition_map, rotary_pos_emb, meta_args, self._scratch)
        y = y_xattn_flat.to(torch.bfloat16).contiguous()
        cond = condition.to(torch.bfloat16)
        for layer in self.layers:
            hidden_states = layer(hidden_states, cond, condition_map, y, rotary_pos_emb, inference_params, meta_args,
                                  _ctx=ctx)
        if self.post_process:
            hidden_states = self.final_layernorm(hidden_states.float())
        return hidden_states
