"""ORACLE — test infrastructure, not product code.

CPU restatement (plain PyTorch, functional, state-dict driven) of the MAGI-1 transformer layer
(`models/magi/dit/dit_module.py:833-1390`, cp_strategy "none", batch 1) and of the Ulysses context-parallel index /
layout logic (`distributed/parallelism/context_parallel.py`).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / reference arm may import this file; nothing under ``inferix_b200/`` does.

Parity status: PINNED to the reference's own modules.  `oracle/make_golden_magi_layer.py` imports the real
`TransformerBlock` / `cp_*` functions from /root/reference, runs them on CPU and stores the results under
`tests/golden/magi_layer_*.pt` / `magi_cp.json`; `tests/test_magi_layer_cpu.py` checks this restatement against those
files bit-for-bit.  The reference layer calls five CUDA-only third-party kernels; the golden run replaces each by the
torch statement of its published algorithm, which is also what this file restates — so on the CPU host the pin for
these five is to the algorithm, not to the vendor's binary.  That boundary is closed on the GPU box, where the
libraries exist: `tests/test_gpu_zzz_thirdparty_pin.py` runs each restatement below against the library kernel itself
(flashinfer silu_and_mul / bmm_fp8 and the reference's Triton range_mod: bit-exact; flash_attn rotary: 99.9996 %
identical; flash_attn_func / _varlen_func: 2.2e-3 from this fp32 statement, the bf16-P kernel's own distance;
profiles/r02z_pytest_thirdparty_pin.log).  The golden run also
emulates CUDA autocast(dtype=float32) — inactive on a CPU-only host — by casting the operands of `F.linear` to fp32
inside the reference's `torch.autocast("cuda", dtype=torch.float32)` regions, which is what CUDA autocast does:
  flash_attn.flash_attn_func / flash_attn_varlen_func  -> softmax(q k^T / sqrt(d)) v with grouped KV heads, fp32 math
  flash_attn.layers.rotary.apply_rotary_emb             -> flash_attn's own `apply_rotary_emb_torch` (non-interleaved)
  range_mod_triton (in-tree Triton, dit_module.py:205-292) -> y[row] = x[row] * gatings[map[row]]
  flashinfer.activation.silu_and_mul                    -> silu(x[..., :d]) * x[..., d:]   (fp32 math, input dtype out)
Every function cites the reference lines it follows (paths relative to /root/reference/inferix).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class MagiConfig:
    """The ModelConfig fields the layer reads (core/config/model.py:20-50)."""
    hidden_size: int = 3072
    ffn_hidden_size: int = 12288
    num_attention_heads: int = 24
    num_query_groups: int = 8
    kv_channels: int = 128
    num_layers: int = 34
    layernorm_epsilon: float = 1e-6
    apply_layernorm_1p: bool = False
    cond_hidden_ratio: float = 0.25
    cond_gating_ratio: float = 1.0
    xattn_cond_hidden_ratio: float = 1.0
    gated_linear_unit: bool = False

    @property
    def q_size(self) -> int:
        return self.kv_channels * self.num_attention_heads

    @property
    def kv_size(self) -> int:
        return self.kv_channels * self.num_query_groups


# ----------------------------------------------------------------------------- third-party kernels, restated
def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """flash_attn.layers.rotary.apply_rotary_emb_torch, interleaved=False.  x [b, s, h, d]; cos/sin [s, rd/2]."""
    rd = cos.shape[-1] * 2
    c = torch.cat([cos, cos], dim=-1)[:, None, :]
    s = torch.cat([sin, sin], dim=-1)[:, None, :]
    return torch.cat([x[..., :rd] * c + rotate_half(x[..., :rd]) * s, x[..., rd:]], dim=-1)


def gqa_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """flash_attn_func semantics without mask / dropout: q [sq, hq, d], k/v [sk, hk, d] -> [sq, hq, d] in q's dtype.
    Query head h reads KV head h // (hq / hk); scores, softmax and the PV product in fp32."""
    sq, hq, d = q.shape
    rep = hq // k.shape[1]
    qf = q.float().transpose(0, 1)                                     # [hq, sq, d]
    kf = k.float().repeat_interleave(rep, dim=1).transpose(0, 1)
    vf = v.float().repeat_interleave(rep, dim=1).transpose(0, 1)
    p = torch.softmax(qf @ kf.transpose(1, 2) * (1.0 / math.sqrt(d)), dim=-1)
    return (p @ vf).transpose(0, 1).to(q.dtype)


def varlen_attention(q, k, v, cu_q: Sequence[int], cu_k: Sequence[int]) -> torch.Tensor:
    """flash_attn_varlen_func semantics: segment i of q attends segment i of k/v."""
    out = torch.empty_like(q)
    for i in range(len(cu_q) - 1):
        qs, qe, ks, ke = int(cu_q[i]), int(cu_q[i + 1]), int(cu_k[i]), int(cu_k[i + 1])
        if qe > qs:
            out[qs:qe] = gqa_attention(q[qs:qe], k[ks:ke], v[ks:ke])
    return out


def range_mod(x: torch.Tensor, c_mapping: torch.Tensor, gatings: torch.Tensor) -> torch.Tensor:
    """range_mod_triton (dit_module.py:241-292): x [s, b, h], c_mapping [s, b], gatings [b, ranges, h]."""
    s, b, h = x.shape
    xf = x.transpose(0, 1).flatten(0, 1)
    y = xf * gatings.flatten(0, 1)[c_mapping.transpose(0, 1).flatten(0, 1)]
    return y.reshape(b, s, h).transpose(0, 1)


def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    d = x.shape[-1] // 2
    return (F.silu(x[..., :d].float()) * x[..., d:].float()).to(x.dtype)


# ----------------------------------------------------------------------------- in-tree pieces
def softcap(x: torch.Tensor, cap: float) -> torch.Tensor:
    """dit_module.py:363-364."""
    return (cap * torch.tanh(x.float() / cap)).to(x.dtype)


def fused_layer_norm(x, weight, bias, eps, zero_centered_gamma=False):
    """FusedLayerNorm.forward, dit_module.py:358-360."""
    w = weight + 1 if zero_centered_gamma else weight
    return F.layer_norm(x, (x.shape[-1],), w, bias, eps)


def bias_modulate_add(x, residual, condition_map, gate, norm_w, norm_b, cfg: MagiConfig):
    """dit_module.py:295-313: fp32( LN( x * gate[map] ) + residual ) -> x's dtype."""
    dt = x.dtype
    y = range_mod(x.float(), condition_map, gate.float())
    y = fused_layer_norm(y, norm_w, norm_b, cfg.layernorm_epsilon, cfg.apply_layernorm_1p)
    return (y + residual.float()).to(dt)


class OracleMagiCache:
    """MagiKVCacheManager (kvcache_manager/model/magi_kv_cache_manager.py:76-187) on a plain tensor per layer."""

    def __init__(self, max_sequence_length: int, max_batch_size: int = 1):
        self.max_sequence_length, self.max_batch_size = max_sequence_length, max_batch_size
        self.update_kv_cache = False
        self.mem: Dict[int, torch.Tensor] = {}

    def adjust(self, layer: int, key_and_value: torch.Tensor, meta) -> Tuple[torch.Tensor, torch.Tensor]:
        d = key_and_value.shape[-1] // 2
        if not (meta.extract_prefix_video_feature or meta.fwd_extra_1st_chunk or meta.slice_point > 0):   # :176-187
            k, v = torch.chunk(key_and_value, 2, dim=-1)
            return k.contiguous(), v.contiguous()
        kv = torch.stack([key_and_value[..., :d], key_and_value[..., d:]])          # coef, tok, hn, d   (:96)
        if layer not in self.mem:
            self.mem[layer] = torch.zeros((2, self.max_sequence_length) + tuple(kv.shape[2:]), dtype=kv.dtype)
        mem = self.mem[layer]
        start = meta.slice_point * meta.clip_token_nums * self.max_batch_size      # :114-116
        hist = mem[:, :start].clone()
        if self.update_kv_cache:                                                    # :127-146
            clip = (kv.shape[1] - meta.clip_token_nums * self.max_batch_size if meta.distill_nearly_clean_chunk
                    else kv.shape[1])
            assert start + clip <= mem.shape[1]
            mem[:, start:start + clip] = kv[:, :clip]
        full = torch.cat([hist, kv], dim=1)
        return full[0].contiguous(), full[1].contiguous()


# ----------------------------------------------------------------------------- FP8 linears
def div_clamp_to(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """dit_module.py:367-387: clamp(x.float() / scale.float(), +-448) -> bf16 -> e4m3 (scale broadcasts over the last dim)."""
    return torch.clamp(x.float() / scale.float(), -448.0, 448.0).bfloat16().to(torch.float8_e4m3fn)


def bmm_fp8(a_q: torch.Tensor, w_q: torch.Tensor, a_scale: torch.Tensor, b_scale: torch.Tensor) -> torch.Tensor:
    """flashinfer.bmm_fp8 (cuBLASLt FP8 matmul with per-tensor scaling): fp32 accumulation of the exact e4m3 products,
    times ONE scale per operand (cuBLASLt reads a single float through each scale pointer: element 0 when the tensor
    has more, as PerTensorQuantizedFp8Linear.input_scale does), rounded to bf16."""
    return ((a_q.float() @ w_q.float().t()) * (float(a_scale.reshape(-1)[0]) * float(b_scale.reshape(-1)[0]))).to(torch.bfloat16)


def qlinear(sd, name: str, x: torch.Tensor, fp32_autocast: bool = False) -> torch.Tensor:
    """A MAGI linear by parameter prefix `name`: nn.Linear, PerTensorQuantizedFp8Linear (:434-459) or
    PerChannelQuantizedFp8Linear (:465-490), told apart by the parameters present, as the reference's module types are."""
    if name + ".weight_scale" not in sd:
        if fp32_autocast:
            return F.linear(x.float(), sd[name + ".weight"].float())
        return F.linear(x, sd[name + ".weight"])
    w_q = sd[name + ".weight"][0]
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    if name + ".smooth_scale" in sd:                                            # PerChannel
        a = div_clamp_to(x2, sd[name + ".smooth_scale"].to(torch.float32))
    else:                                                                       # PerTensor (input_scale is [in])
        a = div_clamp_to(x2, sd[name + ".input_scale"])
    return bmm_fp8(a, w_q, sd[name + ".input_scale"], sd[name + ".weight_scale"]).view(*shp[:-1], -1)


# ----------------------------------------------------------------------------- the layer
def attention_block(sd, p, cfg: MagiConfig, layer: int, hidden, y_xattn_flat, rope, cache: Optional[OracleMagiCache],
                    meta):
    """FullyParallelAttention.forward, cp_strategy == "none" (dit_module.py:1087-1121), batch 1.
    hidden [s, 1, h]; returns (core_attn_out, xattn_out), each [s, 1, hq*d]."""
    a = p + "self_attention."
    d, hq, hk = cfg.kv_channels, cfg.num_attention_heads, cfg.num_query_groups
    eps, g1p = cfg.layernorm_epsilon, cfg.apply_layernorm_1p
    sin_emb, cos_emb = rope.tensor_split(2, -1)                                                     # :1097
    mixed = F.layer_norm(hidden, (cfg.hidden_size,), sd[a + "linear_qkv.layer_norm.weight"],
                         sd[a + "linear_qkv.layer_norm.bias"], eps)                                 # :1103, :415

    def roped(name, ln):                                                                            # get_q / get_k :902-934
        t = qlinear(sd, a + f"linear_qkv.{name}", mixed)
        t = t.reshape(t.size(0), t.size(1), -1, d)
        dt = t.dtype
        t = fused_layer_norm(t.float(), sd[a + ln + ".weight"], sd[a + ln + ".bias"], eps, g1p)
        t = apply_rotary(t.transpose(0, 1).contiguous(), cos_emb, sin_emb).to(dt)                    # [1, s, hn, d]
        return t[0]                                                                                 # (sq b) hn hd, b = 1

    key = roped("k", "k_layernorm")
    value = qlinear(sd, a + "linear_qkv.v", mixed).reshape(-1, hk, d)                               # get_v :936-938
    query = roped("q", "q_layernorm")
    key_and_value = torch.cat([key, value], dim=-1)                                                 # get_kv :940-945
    if cache is None:
        key, value = torch.chunk(key_and_value, 2, dim=-1)
    else:
        key, value = cache.adjust(layer, key_and_value, meta)                                       # :1118
    outs = []                                                                                       # core_attention :972-1015
    for i in range(meta.denoising_range_num):
        qs, qe = (int(t) for t in meta.core_attn_params.np_q_range[i])
        ks, ke = (int(t) for t in meta.core_attn_params.np_k_range[i])
        outs.append(gqa_attention(query[qs:qe], key[ks:ke], value[ks:ke]))
    core = torch.cat(outs, dim=0).reshape(-1, 1, hq * d)                                             # :1120

    # cross attention (get_xqkv :954-970, cross_attention :1047-1085)
    qx = qlinear(sd, a + "linear_qkv.qx", mixed).reshape(-1, hq, d)                                  # (b sq) hn hd
    qx = fused_layer_norm(qx, sd[a + "q_layernorm_xattn.weight"], sd[a + "q_layernorm_xattn.bias"], eps, g1p)
    w = sd[a + "linear_kv_xattn.weight"]
    mixed_kv = torch.cat([torch.matmul(y_xattn_flat, wc.t()) for wc in torch.chunk(w, 8, dim=0)], dim=1)
    mixed_kv = mixed_kv.view(y_xattn_flat.shape[0], -1, 2 * d)
    kx, vx = torch.split(mixed_kv, d, dim=-1)
    kx = fused_layer_norm(kx, sd[a + "k_layernorm_xattn.weight"], sd[a + "k_layernorm_xattn.bias"], eps, g1p)
    cp_ = meta.cross_attn_params
    xo = varlen_attention(qx, kx, vx, cp_.cu_seqlens_q.tolist(), cp_.cu_seqlens_kv.tolist())
    return core, xo.reshape(-1, 1, hq * d)


def layer_forward(sd, layer: int, cfg: MagiConfig, hidden, condition, condition_map, y_xattn_flat, rope,
                  cache: Optional[OracleMagiCache], meta, prefix="layers."):
    """TransformerLayer.forward (dit_module.py:1243-1318)."""
    p = f"{prefix}{layer}."
    residual = hidden
    core, xattn = attention_block(sd, p, cfg, layer, hidden, y_xattn_flat, rope, cache, meta)
    attn = torch.cat([core, xattn], dim=2)                                                          # :1285
    s, b, _ = attn.shape
    hd = attn.shape[2] // 16
    attn = attn.reshape(s, b, 2, 8, hd).transpose(2, 3).reshape(s, b, -1)                           # (n hn hd)->(hn n hd) :1287
    # non-quantised projection runs under torch.autocast("cuda", dtype=float32) (:1291-1293): CUDA autocast casts the
    # operands of `linear` to fp32, so the result stays fp32 through the gate / post-norm and is rounded only after
    # the residual add (bias_modulate_add works in x's dtype, then `.to(params_dtype)`, :1306-1308)
    pdt = hidden.dtype
    # adapt_linear_quant (:1288-1289): the quantised projection runs outside the autocast region and returns bf16
    h = qlinear(sd, p + "self_attention.linear_proj", attn, fp32_autocast=True)
    gate = F.linear(F.silu(condition), sd[p + "ada_modulate_layer.proj.0.weight"],
                    sd[p + "ada_modulate_layer.proj.0.bias"])                                       # :196-198
    gate_msa, gate_mlp = softcap(gate, 1.0).chunk(2, dim=-1)                                        # :1300-1303
    h = bias_modulate_add(h, residual, condition_map, gate_msa, sd[p + "self_attn_post_norm.weight"],
                          sd[p + "self_attn_post_norm.bias"], cfg).to(pdt)
    residual = h
    m = F.layer_norm(h, (cfg.hidden_size,), sd[p + "mlp.layer_norm.weight"], sd[p + "mlp.layer_norm.bias"],
                     cfg.layernorm_epsilon)                                                         # CustomMLP :545-556
    m = qlinear(sd, p + "mlp.linear_fc1", m)
    m = silu_and_mul(m) if cfg.gated_linear_unit else F.gelu(m)
    m = qlinear(sd, p + "mlp.linear_fc2", m)
    return bias_modulate_add(m, residual, condition_map, gate_mlp, sd[p + "mlp_post_norm.weight"],
                             sd[p + "mlp_post_norm.bias"], cfg)


def block_forward(sd, cfg: MagiConfig, hidden, condition, condition_map, y_xattn_flat, rope,
                  cache: Optional[OracleMagiCache], meta, final_norm=True):
    """TransformerBlock.forward (dit_module.py:1361-1390): the layer stack + fp32 final LayerNorm."""
    for i in range(cfg.num_layers):
        hidden = layer_forward(sd, i, cfg, hidden, condition, condition_map, y_xattn_flat, rope, cache, meta)
    if final_norm:
        hidden = fused_layer_norm(hidden.float(), sd["final_layernorm.weight"], sd["final_layernorm.bias"],
                                  cfg.layernorm_epsilon, cfg.apply_layernorm_1p)
    return hidden


# ----------------------------------------------------------------------------- synthetic weights
def synth_state_dict(cfg: MagiConfig, seed: int = 0, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Seeded weights with the reference's parameter names and the dtypes left by `_high_precision_promoter`
    (dit_model.py:620-637): q/k layernorm (self-attention only), post norms and final_layernorm fp32, the rest
    `dtype`."""
    g = torch.Generator().manual_seed(seed)
    h, f, hc = cfg.hidden_size, cfg.ffn_hidden_size, int(cfg.hidden_size * cfg.cond_hidden_ratio)
    hx = int(cfg.hidden_size * cfg.xattn_cond_hidden_ratio)
    hg = int(cfg.hidden_size * cfg.cond_gating_ratio * 2)
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, bias=False, scale=None):
        sd[name + ".weight"] = (torch.randn(out_f, in_f, generator=g) * (scale or in_f ** -0.5)).to(dtype)
        if bias:
            sd[name + ".bias"] = (torch.randn(out_f, generator=g) * 0.1).to(dtype)

    def norm(name, n, dt):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(n, generator=g)).to(dt)
        sd[name + ".bias"] = (0.05 * torch.randn(n, generator=g)).to(dt)

    for i in range(cfg.num_layers):
        p = f"layers.{i}."
        a = p + "self_attention."
        lin(p + "ada_modulate_layer.proj.0", hg, hc, bias=True)
        norm(a + "linear_qkv.layer_norm", h, dtype)
        lin(a + "linear_qkv.q", cfg.q_size, h)
        lin(a + "linear_qkv.qx", cfg.q_size, h)
        lin(a + "linear_qkv.k", cfg.kv_size, h)
        lin(a + "linear_qkv.v", cfg.kv_size, h)
        lin(a + "linear_kv_xattn", 2 * cfg.kv_size, hx)
        lin(a + "linear_proj", h, 2 * cfg.q_size)
        norm(a + "q_layernorm", cfg.kv_channels, torch.float32)
        norm(a + "k_layernorm", cfg.kv_channels, torch.float32)
        norm(a + "q_layernorm_xattn", cfg.kv_channels, dtype)
        norm(a + "k_layernorm_xattn", cfg.kv_channels, dtype)
        norm(p + "self_attn_post_norm", h, torch.float32)
        norm(p + "mlp.layer_norm", h, dtype)
        lin(p + "mlp.linear_fc1", (2 if cfg.gated_linear_unit else 1) * f, h)
        lin(p + "mlp.linear_fc2", h, f)
        norm(p + "mlp_post_norm", h, torch.float32)
    norm("final_layernorm", h, torch.float32)
    return sd


def synth_model_state_dict(module, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded weights for a whole VideoDiTModel (reference or native: same parameter names), one generator per
    parameter NAME so the result does not depend on registration order; dtypes follow the module's."""
    import zlib
    sd = {}
    for name, p in module.state_dict().items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
        if name.endswith("rope.bands"):
            t = p.detach().float().clone()
        elif p.dim() >= 2 and "null_caption" not in name:
            fan_in = p[0].numel()
            t = torch.randn(p.shape, generator=g) * fan_in ** -0.5
        elif "null_caption" in name:
            t = torch.randn(p.shape, generator=g) * 0.5
        elif name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
        else:
            t = 0.05 * torch.randn(p.shape, generator=g)
        sd[name] = t.to(p.dtype)
    return sd


# ----------------------------------------------------------------------------- context parallel (Ulysses) index logic
def cp_split_sizes(seq_len: int, cp_size: int) -> List[int]:
    """cp_ulysses_process part 1 (context_parallel.py:241-243)."""
    sizes = [seq_len // cp_size] * cp_size
    for i in range(seq_len % cp_size):
        sizes[i] += 1
    return sizes


def cp_cross_attn_ranges(cu_seqlens_q: Sequence[int], cu_seqlens_k: Sequence[int], split_sizes: Sequence[int],
                         cp_rank: int):
    """cp_update_cross_attn_qkv_range (context_parallel.py:135-225) for batch 1, cp_shuffle_num 1, no padding:
    intersect every query segment with this rank's token interval; ranges are re-based to the rank's first token.
    Returns (q_ranges, k_ranges) as lists of [start, end)."""
    lo = sum(split_sizes[:cp_rank])
    hi = lo + split_sizes[cp_rank]
    q_ranges, k_ranges = [], []
    for i in range(len(cu_seqlens_q) - 1):
        s, e = max(lo, int(cu_seqlens_q[i])), min(hi, int(cu_seqlens_q[i + 1]))
        if s < e:
            q_ranges.append([s, e])
            k_ranges.append([int(cu_seqlens_k[i]), int(cu_seqlens_k[i + 1])])
    off = min(r[0] for r in q_ranges)
    return [[s - off, e - off] for s, e in q_ranges], k_ranges


def ulysses_input_split(full: torch.Tensor, cp_size: int) -> List[torch.Tensor]:
    """Result of all_to_all_input_split (context_parallel.py:382-402) seen from every rank, given the tensor the
    ranks hold together: full [seq_total, heads, d] -> rank r owns full[:, r*hn:(r+1)*hn] (all tokens, its heads).
    With fewer KV heads than ranks the reference repeats heads first (`repeat_interleave`, :394-395)."""
    heads = full.shape[1]
    if cp_size % heads == 0 and cp_size != heads:
        full = torch.repeat_interleave(full, cp_size // heads, dim=1)
        heads = cp_size
    hn = heads // cp_size
    return [full[:, r * hn:(r + 1) * hn].contiguous() for r in range(cp_size)]
