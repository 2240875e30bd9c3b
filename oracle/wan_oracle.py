"""ORACLE — test infrastructure, not product code.

CPU restatement (plain PyTorch, functional, state-dict driven) of the reference's Self-Forcing / CausVid
denoising hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this file; nothing under ``inferix_b200/`` does.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference from /root/reference (with the
import shims of SURVEY §8c), runs it on CPU on seeded inputs and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those files bit-for-bit (fp32 and bf16).
The reference itself ships no golden vectors or tests for this path.

Every function cites the reference lines it follows (paths relative to /root/reference/inferix).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- configuration
@dataclass
class WanConfig:
    """Subset of CausalWanModel.__init__ arguments (models/self_forcing/causal_model.py:530-548)."""
    dim: int = 1536
    ffn_dim: int = 8960
    num_heads: int = 12
    num_layers: int = 30
    in_dim: int = 16
    out_dim: int = 16
    freq_dim: int = 256
    text_dim: int = 4096
    text_len: int = 512
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    eps: float = 1e-6
    local_attn_size: int = -1   # frames; -1 = global window
    sink_size: int = 0          # frames

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads


# ----------------------------------------------------------------------------- small components
def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """models/wan_base/components.py:11-31 — fp64 outer product, [cos | sin]."""
    half = dim // 2
    pos = position.to(torch.float64)
    inv = torch.pow(10000, -torch.arange(half).to(pos).div(half))
    ang = torch.outer(pos, inv)
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


def rope_params(max_seq_len: int, dim: int, theta: float = 10000) -> torch.Tensor:
    """models/wan_base/components.py:34-52 — complex128 [max_seq_len, dim/2]."""
    ang = torch.outer(torch.arange(max_seq_len),
                      1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(ang), ang)


def rope_freqs(head_dim: int) -> torch.Tensor:
    """CausalWanModel.freqs, causal_model.py:634-641: tables for (t, h, w) of widths d-4(d//6), 2(d//6), 2(d//6)."""
    d = head_dim
    return torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                      rope_params(1024, 2 * (d // 6))], dim=1)


def causal_rope_apply(x: torch.Tensor, grid: Tuple[int, int, int], freqs: torch.Tensor, start_frame: int = 0,
                      world_size: int = 1, rank: int = 0) -> torch.Tensor:
    """causal_model.py:33-61 (world_size == 1) and :64-100 (_chunked: the rank owns an hw slice of each frame).

    x: [B, S, heads, head_dim]; rotation of interleaved (even, odd) pairs in complex128, cast back to x.dtype."""
    n, c = x.size(2), x.size(3) // 2
    f, h, w = grid
    parts = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    ang = torch.cat([
        parts[0][start_frame:start_frame + f].view(f, 1, 1, -1).expand(f, h, w, -1),
        parts[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
        parts[2][:w].view(1, 1, w, -1).expand(f, h, w, -1),
    ], dim=-1).reshape(f, h * w, 1, -1)
    chunk = (h * w) // world_size
    ang = ang[:, rank * chunk:(rank + 1) * chunk].reshape(f * chunk, 1, -1)
    seq = f * chunk
    out = []
    for i in range(x.size(0)):
        xi = torch.view_as_complex(x[i, :seq].to(torch.float64).reshape(seq, n, -1, 2))
        xi = torch.view_as_real(xi * ang).flatten(2)
        out.append(torch.cat([xi, x[i, seq:]]))
    return torch.stack(out).type_as(x)


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """WanRMSNorm, components.py:107-126: fp32 normalise over the FULL channel dim, cast, then x weight."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * weight


def layer_norm(x: torch.Tensor, eps: float, weight=None, bias=None) -> torch.Tensor:
    """WanLayerNorm, components.py:129-142."""
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps).type_as(x)


def sdpa_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, dtype=None) -> torch.Tensor:
    """attention() SDPA branch, models/attention/flash_attention.py:185-199: q,k,v [B, L, N, D] -> [B, Lq, N, D].

    The reference's flash-attn branch (:117-147) computes the same full softmax attention (no mask, scale
    1/sqrt(D)); on CPU only this branch can run, and it is what the goldens were produced with."""
    dtype = dtype or q.dtype
    o = F.scaled_dot_product_attention(q.transpose(1, 2).to(dtype), k.transpose(1, 2).to(dtype),
                                       v.transpose(1, 2).to(dtype), attn_mask=None, is_causal=False, dropout_p=0.0)
    return o.transpose(1, 2).contiguous()


# ----------------------------------------------------------------------------- KV cache (reference layout)
@dataclass
class LayerCache:
    """One layer's cache as the reference keeps it: contiguous [B, N, H, D] K and V plus the two end indices
    (kv_cache_meta, pipeline/self_forcing/CausalInferencePipeline.py:462-470)."""
    k: torch.Tensor
    v: torch.Tensor
    global_end: int = 0
    local_end: int = 0
    trace: List[Tuple[int, int, int, int]] = field(default_factory=list)  # (local_start, local_end, global_end, evicted)


def new_cache(cfg: WanConfig, cache_tokens: int, batch: int, dtype, heads: Optional[int] = None) -> List[LayerCache]:
    h = heads or cfg.num_heads
    return [LayerCache(torch.zeros(batch, cache_tokens, h, cfg.head_dim, dtype=dtype),
                       torch.zeros(batch, cache_tokens, h, cfg.head_dim, dtype=dtype)) for _ in range(cfg.num_layers)]


def cache_append(c: LayerCache, k_new: torch.Tensor, v_new: torch.Tensor, current_start: int, sink_tokens: int,
                 windowed: bool) -> Tuple[int, int]:
    """causal_model.py:277-304,328-329: evict (roll left past the sink) when the window would overflow, then write the
    new tokens at [local_start, local_end).  Returns (local_start, local_end)."""
    num_new = k_new.shape[1]
    cache_size = c.k.shape[1]
    current_end = current_start + num_new
    evicted = 0
    if windowed and current_end > c.global_end and num_new + c.local_end > cache_size:
        evicted = num_new + c.local_end - cache_size
        rolled = c.local_end - evicted - sink_tokens
        src = slice(sink_tokens + evicted, sink_tokens + evicted + rolled)
        dst = slice(sink_tokens, sink_tokens + rolled)
        c.k[:, dst] = c.k[:, src].clone()
        c.v[:, dst] = c.v[:, src].clone()
        local_end = c.local_end + current_end - c.global_end - evicted
    else:
        local_end = c.local_end + current_end - c.global_end
    local_start = local_end - num_new
    c.k[:, local_start:local_end] = k_new
    c.v[:, local_start:local_end] = v_new
    c.global_end, c.local_end = current_end, local_end
    c.trace.append((local_start, local_end, current_end, evicted))
    return local_start, local_end


def plan_indices(cache_size: int, global_end: int, local_end: int, current_start: int, num_new: int,
                 sink_tokens: int, windowed: bool) -> Tuple[int, int, int, int]:
    """Index-only form of cache_append: (local_start, local_end, global_end, evicted)."""
    current_end = current_start + num_new
    evicted = 0
    if windowed and current_end > global_end and num_new + local_end > cache_size:
        evicted = num_new + local_end - cache_size
        new_local_end = local_end + current_end - global_end - evicted
    else:
        new_local_end = local_end + current_end - global_end
    return new_local_end - num_new, new_local_end, current_end, evicted


# ----------------------------------------------------------------------------- the DiT block
def div_clamp_to_e4m3(x: torch.Tensor, scale: float) -> torch.Tensor:
    """models/magi/dit/dit_module.py:367-387 with a per-tensor scale: clamp in fp32 -> bf16 -> e4m3."""
    return torch.clamp(x.float() / scale, -448.0, 448.0).bfloat16().to(torch.float8_e4m3fn)


def fp8_linear(x: torch.Tensor, w_q: torch.Tensor, weight_scale: float, input_scale: float, bias=None) -> torch.Tensor:
    """PerTensorQuantizedFp8Linear.forward, dit_module.py:434-459: bmm_fp8(x_q, W_q^T, input_scale, weight_scale)
    = (x_q @ W_q^T) * input_scale * weight_scale with fp32 accumulation, cast to bf16.  The Wan linears carry a bias
    (MAGI's do not): it is added in fp32 before the cast.  fp8 x fp8 products are exact in fp32."""
    xq = div_clamp_to_e4m3(x, input_scale).float()
    y = (xq @ w_q.float().t()) * (input_scale * weight_scale)
    if bias is not None:
        y = y + bias.float()
    return y.to(torch.bfloat16)


def dynamic_q8_quantize(x: torch.Tensor, kind: str):
    """Per-row dynamic quantisation: s = max(|x|_row, 1e-12) / qmax; codes = RNE with saturation (e4m3: qmax 448,
    int8: qmax 127).  Returns (codes as float32 values, scales [rows, 1] fp32).  PARITY UNPINNED: this is the published
    "dynamic per-token activation x per-channel weight" scheme the reference's quantisation examples request from DAX
    (example/quantization/run_causvid_quantized.py:32-37); DAX (RiseAI-Sys/DAX, cloned by the example's README, no
    version pinned) is absent from /root/reference, so no reference output exists to check against."""
    qmax = 448.0 if kind == "fp8" else 127.0
    xf = x.float()
    s = xf.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12) / qmax
    q = xf * (1.0 / s)                     # the kernels multiply by the reciprocal
    if kind == "fp8":
        codes = q.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()
    else:
        codes = torch.round(q).clamp(-127, 127)
    return codes, s


def dynamic_q8_linear(x: torch.Tensor, w_codes: torch.Tensor, w_scale: torch.Tensor, kind: str, bias=None) -> torch.Tensor:
    """y = bf16( (x_q @ W_q^T) * s_a[m] * s_w[n] + bias ), exact integer / fp8 products accumulated in fp32 (int8: the
    products are exact in fp32 up to K * 127^2 < 2^24 for K <= 1040; beyond that fp64 keeps the sum exact)."""
    xq, sa = dynamic_q8_quantize(x, kind)
    acc = (xq.double() @ w_codes.double().t()).float() if kind == "int8" else xq @ w_codes.float().t()
    y = acc * sa * w_scale.float().view(1, -1)
    if bias is not None:
        y = y + bias.float()
    return y.to(torch.bfloat16)


def _lin(sd: Dict[str, torch.Tensor], name: str, x: torch.Tensor) -> torch.Tensor:
    q8 = sd.get("__q8__", {}).get(name)
    if q8 is not None:     # (weight codes, weight scales [N], kind): dynamically quantised 8-bit linear
        shp = x.shape
        return dynamic_q8_linear(x.reshape(-1, shp[-1]), q8[0], q8[1], q8[2], sd.get(name + ".bias")).view(*shp[:-1], -1)
    q = sd.get("__fp8__", {}).get(name)
    if q is not None:      # (weight_q, weight_scale, input_scale): this linear is FP8-quantised
        return fp8_linear(x, q[0], q[1], q[2], sd.get(name + ".bias"))
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def self_attention(sd, pfx: str, cfg: WanConfig, x, grid, freqs, cache: LayerCache, current_start: int,
                   world_size: int = 1, rank: int = 0, attn_dtype=None, peer_kv: Optional[Callable] = None):
    """CausalWanSelfAttention.forward inference branch, causal_model.py:147-178,254-334."""
    b, s, n, d = x.shape[0], x.shape[1], cfg.num_heads, cfg.head_dim
    q = rms_norm(_lin(sd, pfx + ".q", x), sd[pfx + ".norm_q.weight"], cfg.eps).view(b, s, n, d)
    k = rms_norm(_lin(sd, pfx + ".k", x), sd[pfx + ".norm_k.weight"], cfg.eps).view(b, s, n, d)
    v = _lin(sd, pfx + ".v", x).view(b, s, n, d)
    frame_seqlen = grid[1] * grid[2]
    start_frame = current_start // frame_seqlen
    q = causal_rope_apply(q, grid, freqs, start_frame, world_size, rank).type_as(v)
    k = causal_rope_apply(k, grid, freqs, start_frame, world_size, rank).type_as(v)
    if peer_kv is not None:
        # sequence parallel with a replicated cache (this build's layout): gather every rank's new K/V
        # in (frame, rank, hw) order == the single-process token order
        k, v = peer_kv(k, v)
    sink_tokens = cfg.sink_size * frame_seqlen
    _, local_end = cache_append(cache, k, v, current_start, sink_tokens, cfg.local_attn_size != -1)
    o = sdpa_attention(q, cache.k[:, :local_end], cache.v[:, :local_end], attn_dtype or q.dtype)
    return _lin(sd, pfx + ".o", o.flatten(2))


def cross_attention(sd, pfx: str, cfg: WanConfig, x, context, cross_cache: dict, attn_dtype=None):
    """WanT2VCrossAttention.forward, models/wan_base/model.py:66-100 (text K/V computed once per prompt)."""
    b, n, d = x.size(0), cfg.num_heads, cfg.head_dim
    q = rms_norm(_lin(sd, pfx + ".q", x), sd[pfx + ".norm_q.weight"], cfg.eps).view(b, -1, n, d)
    if not cross_cache.get("is_init", False):
        cross_cache["is_init"] = True
        cross_cache["k"] = rms_norm(_lin(sd, pfx + ".k", context), sd[pfx + ".norm_k.weight"], cfg.eps).view(b, -1, n, d)
        cross_cache["v"] = _lin(sd, pfx + ".v", context).view(b, -1, n, d)
    o = sdpa_attention(q, cross_cache["k"], cross_cache["v"], attn_dtype or q.dtype)
    return _lin(sd, pfx + ".o", o.flatten(2))


def block_forward(sd, i: int, cfg: WanConfig, x, e0, grid, freqs, context, cache: LayerCache, cross_cache: dict,
                  current_start: int, world_size: int = 1, rank: int = 0, attn_dtype=None, peer_kv=None):
    """CausalWanAttentionBlock.forward, causal_model.py:384-484.  x [B, S, C]; e0 [B, F, 6, C]."""
    pfx = f"blocks.{i}"
    num_frames, fs = e0.shape[1], x.shape[1] // e0.shape[1]
    e = (sd[pfx + ".modulation"].unsqueeze(1) + e0).chunk(6, dim=2)          # :412

    def per_frame(t):
        return t.unflatten(dim=1, sizes=(num_frames, fs))

    h = (per_frame(layer_norm(x, cfg.eps)) * (1 + e[1]) + e[0]).flatten(1, 2)  # :433
    y = self_attention(sd, pfx + ".self_attn", cfg, h, grid, freqs, cache, current_start, world_size, rank,
                       attn_dtype, peer_kv)
    x = x + (per_frame(y) * e[2]).flatten(1, 2)                                # :444
    n3 = layer_norm(x, cfg.eps, sd.get(pfx + ".norm3.weight"), sd.get(pfx + ".norm3.bias")) \
        if (pfx + ".norm3.weight") in sd else x
    x = x + cross_attention(sd, pfx + ".cross_attn", cfg, n3, context, cross_cache, attn_dtype)  # :448
    h = (per_frame(layer_norm(x, cfg.eps)) * (1 + e[4]) + e[3]).flatten(1, 2)  # :451-452
    y = _lin(sd, pfx + ".ffn.2", F.gelu(_lin(sd, pfx + ".ffn.0", h), approximate="tanh"))
    return x + (per_frame(y) * e[5]).flatten(1, 2)                             # :455-456


# ----------------------------------------------------------------------------- the model
def embed(sd, cfg: WanConfig, latents: torch.Tensor, t: torch.Tensor, context: torch.Tensor):
    """Prologue of _forward_inference, causal_model.py:916-953.  latents [B, C_in, F, H, W]; t [B, F];
    context [B, L<=text_len, text_dim].  Returns x [B, S, C], e [B*F, C], e0 [B, F, 6, C], ctx, grid."""
    x = F.conv3d(latents, sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=cfg.patch_size)
    grid = tuple(int(v) for v in x.shape[2:])
    x = x.flatten(2).transpose(1, 2)
    e = sinusoidal_embedding_1d(cfg.freq_dim, t.flatten()).type_as(x)
    e = _lin(sd, "time_embedding.2", F.silu(_lin(sd, "time_embedding.0", e)))
    e0 = _lin(sd, "time_projection.1", F.silu(e)).unflatten(1, (6, cfg.dim)).unflatten(dim=0, sizes=t.shape)
    ctx = torch.stack([torch.cat([u, u.new_zeros(cfg.text_len - u.size(0), u.size(1))]) for u in context])
    ctx = _lin(sd, "text_embedding.2", F.gelu(_lin(sd, "text_embedding.0", ctx), approximate="tanh"))
    return x, e, e0, ctx, grid


def head_unpatchify(sd, cfg: WanConfig, x, e, t_shape, grid):
    """CausalHead.forward causal_model.py:504-515 + unpatchify :1196-1219.  Returns [B, C_out, F, H, W]."""
    f, h, w = grid
    num_frames, fs = t_shape[1], x.shape[1] // t_shape[1]
    eh = e.unflatten(dim=0, sizes=t_shape).unsqueeze(2)                         # [B, F, 1, C]
    m = (sd["head.modulation"].unsqueeze(1) + eh).chunk(2, dim=2)
    y = _lin(sd, "head.head", layer_norm(x, cfg.eps).unflatten(dim=1, sizes=(num_frames, fs)) * (1 + m[1]) + m[0])
    y = y.flatten(1, 2)                                                          # 'b f hw c -> b (f hw) c'
    c = cfg.out_dim
    out = []
    for u in y:
        u = u[: f * h * w].view(f, h, w, *cfg.patch_size, c)
        u = torch.einsum("fhwpqrc->cfphqwr", u)
        out.append(u.reshape(c, f * cfg.patch_size[0], h * cfg.patch_size[1], w * cfg.patch_size[2]))
    return torch.stack(out)


def model_forward(sd, cfg: WanConfig, latents, t, context, caches: List[LayerCache], cross_caches: List[dict],
                  current_start: int, freqs: Optional[torch.Tensor] = None, attn_dtype=None,
                  taps: Optional[dict] = None) -> torch.Tensor:
    """CausalWanModel._forward_inference (single process), causal_model.py:866-1026 -> flow prediction."""
    freqs = rope_freqs(cfg.head_dim) if freqs is None else freqs
    x, e, e0, ctx, grid = embed(sd, cfg, latents, t, context)
    if taps is not None:
        taps["x_in"], taps["e0"], taps["ctx"] = x.clone(), e0.clone(), ctx.clone()
    for i in range(cfg.num_layers):
        x = block_forward(sd, i, cfg, x, e0, grid, freqs, ctx, caches[i], cross_caches[i], current_start,
                          attn_dtype=attn_dtype)
        if taps is not None:
            taps[f"block{i}"] = x.clone()
    return head_unpatchify(sd, cfg, x, e, t.shape, grid)


# ----------------------------------------------------------------------------- scheduler + wrapper
class FlowMatchSigmas:
    """FlowMatchScheduler(shift, sigma_min=0, extra_one_step=True).set_timesteps(1000),
    models/schedulers/flow_match.py:118-141; add_noise :159-176."""

    def __init__(self, shift: float = 5.0, num_steps: int = 1000, sigma_max: float = 1.0, sigma_min: float = 0.0):
        s = torch.linspace(sigma_min + (sigma_max - sigma_min), sigma_min, num_steps + 1)[:-1]
        self.sigmas = shift * s / (1 + (shift - 1) * s)
        self.timesteps = self.sigmas * 1000

    def index_of(self, timestep: torch.Tensor) -> torch.Tensor:
        return torch.argmin((self.timesteps.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)

    def add_noise(self, x0: torch.Tensor, noise: torch.Tensor, timestep: torch.Tensor) -> torch.Tensor:
        sigma = self.sigmas[self.index_of(timestep)].reshape(-1, 1, 1, 1)
        return ((1 - sigma) * x0 + sigma * noise).type_as(noise)

    def flow_to_x0(self, flow: torch.Tensor, xt: torch.Tensor, timestep: torch.Tensor) -> torch.Tensor:
        """WanDiffusionWrapper._convert_flow_pred_to_x0, models/self_forcing/wrapper.py:259-283 (fp64)."""
        idx = torch.argmin((self.timesteps.double().unsqueeze(0) - timestep.double().unsqueeze(1)).abs(), dim=1)
        sigma = self.sigmas.double()[idx].reshape(-1, 1, 1, 1)
        return (xt.double() - sigma * flow.double()).to(flow.dtype)


def warp_steps(sched: FlowMatchSigmas, steps: List[int]) -> torch.Tensor:
    """CausalInferencePipeline.__init__, pipeline/self_forcing/CausalInferencePipeline.py:86-90."""
    ts = torch.cat((sched.timesteps, torch.tensor([0], dtype=torch.float32)))
    return ts[1000 - torch.tensor(steps, dtype=torch.long)]


def generator_forward(sd, cfg, sched: FlowMatchSigmas, noisy, timestep, context, caches, cross_caches, current_start,
                      attn_dtype=None):
    """WanDiffusionWrapper.forward (kv-cache branch), wrapper.py:308-383.  noisy [B, F, C, H, W] -> (flow, x0)."""
    flow = model_forward(sd, cfg, noisy.permute(0, 2, 1, 3, 4), timestep, context, caches, cross_caches,
                         current_start, attn_dtype=attn_dtype).permute(0, 2, 1, 3, 4)
    x0 = sched.flow_to_x0(flow.flatten(0, 1), noisy.flatten(0, 1), timestep.flatten(0, 1)).unflatten(0, flow.shape[:2])
    return flow, x0


def pipeline_inference(sd, cfg: WanConfig, sched: FlowMatchSigmas, noise: torch.Tensor, context: torch.Tensor,
                       denoising_steps: torch.Tensor, num_frame_per_block: int, frame_seq_length: int,
                       cache_tokens: int, context_noise: int = 0, attn_dtype=None,
                       block_callback: Optional[Callable] = None, noise_fn: Optional[Callable] = None):
    """CausalInferencePipeline.inference (T2V, no initial latent, NO_DECODE),
    pipeline/self_forcing/CausalInferencePipeline.py:108-442: per block, T noisy forwards with re-noising in between,
    then one clean forward at `context_noise` that rewrites the block's K/V.  Returns (latents, caches)."""
    b, num_frames = noise.shape[:2]
    assert num_frames % num_frame_per_block == 0
    caches = new_cache(cfg, cache_tokens, b, noise.dtype)
    cross = [dict(is_init=False) for _ in range(cfg.num_layers)]
    out = torch.zeros_like(noise)
    noise_fn = noise_fn or torch.randn_like
    start = 0
    for blk in range(num_frames // num_frame_per_block):
        n = num_frame_per_block
        x = noise[:, start:start + n]
        x0 = ts = None
        for idx, cur in enumerate(denoising_steps):
            ts = torch.ones([b, n], dtype=torch.int64) * cur
            _, x0 = generator_forward(sd, cfg, sched, x, ts, context, caches, cross, start * frame_seq_length, attn_dtype)
            if idx < len(denoising_steps) - 1:
                nxt = denoising_steps[idx + 1] * torch.ones([b * n], dtype=torch.long)
                x = sched.add_noise(x0.flatten(0, 1), noise_fn(x0.flatten(0, 1)), nxt).unflatten(0, x0.shape[:2])
        out[:, start:start + n] = x0
        generator_forward(sd, cfg, sched, x0, torch.ones_like(ts) * context_noise, context, caches, cross,
                          start * frame_seq_length, attn_dtype)
        start += n
        if block_callback is not None:
            block_callback(out[:, start - n:start], blk)
    return out, caches


def cfg_pipeline_inference(sd, cfg: WanConfig, sched: FlowMatchSigmas, noise, context, neg_context, scheduler_factory,
                           guidance_scale: float, num_frame_per_block: int, frame_seq_length: int, cache_tokens: int,
                           attn_dtype=None):
    """CausalDiffusionInferencePipeline.inference (T2V, no initial latent),
    pipeline/self_forcing/CausalDiffusionInferencePipeline.py:50-296: per block, for every sampler timestep a conditional
    and an unconditional forward (separate caches), flow = uncond + g * (cond - uncond), one multistep-sampler step; then
    the clean re-run of both branches.  `scheduler_factory()` returns a fresh sampler with `.timesteps` and
    `.step(flow, t, latents, return_dict=False)` (the reference's FlowUniPCMultistepScheduler; its restatement in
    inferix_b200/unipc.py is pinned bit-exact to the reference class by tests/golden/unipc.pt)."""
    b, num_frames = noise.shape[:2]
    caches = [new_cache(cfg, cache_tokens, b, noise.dtype) for _ in range(2)]
    cross = [[dict(is_init=False) for _ in range(cfg.num_layers)] for _ in range(2)]
    ctxs = (context, neg_context)
    out = torch.zeros_like(noise)
    start = 0
    for _ in range(num_frames // num_frame_per_block):
        n = num_frame_per_block
        latents = noise[:, start:start + n]
        s = scheduler_factory()
        ts = None
        for t in s.timesteps:
            ts = t * torch.ones([b, n], dtype=torch.float32)
            flows = [generator_forward(sd, cfg, sched, latents, ts, ctxs[k], caches[k], cross[k], start * frame_seq_length,
                                       attn_dtype)[0] for k in range(2)]
            flow = flows[1] + guidance_scale * (flows[0] - flows[1])
            latents = s.step(flow, t, latents, return_dict=False)[0]
        out[:, start:start + n] = latents
        for k in range(2):
            generator_forward(sd, cfg, sched, latents, ts * 0, ctxs[k], caches[k], cross[k], start * frame_seq_length, attn_dtype)
        start += n
    return out, caches


# ----------------------------------------------------------------------------- CausVid
def causvid_pipeline_inference(sd, cfg: WanConfig, sched: FlowMatchSigmas, noise: torch.Tensor, context: torch.Tensor,
                               denoising_steps: torch.Tensor, num_frame_per_block: int, frame_seq_length: int,
                               cache_tokens: int, start_latents: Optional[torch.Tensor] = None, attn_dtype=None,
                               noise_fn: Optional[Callable] = None):
    """CausVid block scheduler, pipeline/causvid/CausalInferencePipeline.py:94-257 (latents only), over the CausVid
    model (models/causvid/causal_model.py): every forward writes cache[kv_start:kv_end] and attends cache[0:kv_end]
    (:169-175, :262) — the un-windowed case of cache_append — and the wrapper returns x0 only (wrapper.py:269-305).
    `denoising_steps` is the list AFTER the reference dropped its trailing entry (:37)."""
    assert cfg.local_attn_size == -1
    b, num_frames = noise.shape[:2]
    n = num_frame_per_block
    caches = new_cache(cfg, cache_tokens, b, noise.dtype)
    cross = [dict(is_init=False) for _ in range(cfg.num_layers)]
    out = torch.zeros_like(noise)
    noise_fn = noise_fn or torch.randn_like
    num_input_blocks = start_latents.shape[1] // n if start_latents is not None else 0
    for blk in range(num_frames // n):
        start = blk * n * frame_seq_length
        x = noise[:, blk * n:(blk + 1) * n]
        zeros = torch.ones([b, n], dtype=torch.int64) * 0
        if start_latents is not None and blk < num_input_blocks:
            ref = start_latents[:, blk * n:(blk + 1) * n]
            out[:, blk * n:(blk + 1) * n] = ref
            generator_forward(sd, cfg, sched, ref, zeros, context, caches, cross, start, attn_dtype)
            continue
        x0 = ts = None
        for idx, cur in enumerate(denoising_steps):
            ts = torch.ones([b, n], dtype=torch.int64) * cur
            _, x0 = generator_forward(sd, cfg, sched, x, ts, context, caches, cross, start, attn_dtype)
            if idx < len(denoising_steps) - 1:
                nxt = denoising_steps[idx + 1] * torch.ones([b], dtype=torch.long)
                x = sched.add_noise(x0.flatten(0, 1), noise_fn(x0.flatten(0, 1)), nxt).view(x0.shape)
        out[:, blk * n:(blk + 1) * n] = x0
        generator_forward(sd, cfg, sched, x0, ts * 0, context, caches, cross, start, attn_dtype)
    return out, caches
