"""Torch-tensor front end of the C ABI: validates dtype / layout, passes raw pointers and the current stream.

PyTorch is used for device memory and streams only; every function below runs a hand-written sm_100a kernel
from ``libinferix_b200.so`` and raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_BIAS_F32, EPI_BIAS_GATE_RES, EPI_BIAS_GELU, EPI_BIAS_GELU_ERF, KvPlan, RopeGrid

__all__ = [
    "ln_modulate", "gemm", "rmsnorm", "quantize_fp8", "quantize_fp8_cols", "quantize_rows", "ln_modulate_quant", "gemm_q8",
    "quantize_weight_per_channel", "Q8_E4M3", "Q8_INT8", "ln_modulate_fp8", "gemm_fp8", "attention", "attention_gqa", "attention_ranges", "attention_partial", "attention_combine",
    "attention_workspace_bytes", "attention_extents", "attention_lse", "qk_norm_rope_append", "PagedKV", "rope_table",
    "EPI_BIAS", "EPI_BIAS_GELU", "EPI_BIAS_GATE_RES", "EPI_BIAS_GELU_ERF", "EPI_BIAS_F32",
    "qk_norm_rope_append_peers", "peer_push", "patchify", "sinusoidal_embedding", "linear_small", "unpatchify_x0", "add_noise", "magi_qkv_post", "head_layernorm", "gate_norm_residual", "silu_mul",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _bf16_2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (inferix_b200 has no CPU path)")
    if t.dtype != torch.bfloat16:
        raise ValueError(f"{name}: expected bfloat16, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)}/{t.stride()}")
    return t


def _bf16_vec(t: Optional[torch.Tensor], n: int, name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.bfloat16 or t.numel() != n or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA bfloat16 vector of {n} elements")
    return t


def rope_table(freqs: torch.Tensor, device) -> torch.Tensor:
    """complex128 [1024, D/2] (CausalWanModel.freqs) -> float64 [1024, D/2, 2] (cos, sin) on `device`."""
    if freqs.dtype != torch.complex128:
        raise ValueError("freqs must be complex128 as built by rope_params")
    return torch.view_as_real(freqs).contiguous().to(device)


def ln_modulate(x, out=None, *, weight=None, bias=None, shift=None, scale=None, tokens_per_frame=0, eps=1e-6):
    """LayerNorm (+affine) (+AdaLN modulate).  shift/scale: [frames, C] views with a common frame stride."""
    x = _bf16_2d(x, "x")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous")
    rows, cols = x.shape
    out = torch.empty_like(x) if out is None else _bf16_2d(out, "out")
    stride = 0
    if scale is not None:
        if shift is None or scale.shape != shift.shape or scale.dim() != 2 or scale.shape[1] != cols:
            raise ValueError("shift/scale must both be [frames, C]")
        if scale.stride(1) != 1 or shift.stride(1) != 1 or scale.stride(0) != shift.stride(0):
            raise ValueError("shift/scale must share a frame stride and have unit inner stride")
        if rows != scale.shape[0] * tokens_per_frame:
            raise ValueError("rows != frames * tokens_per_frame")
        stride = scale.stride(0)
    _lib.check(_lib.load().ifx_ln_modulate(
        x.data_ptr(), out.data_ptr(), _ptr(_bf16_vec(weight, cols, "weight")), _ptr(_bf16_vec(bias, cols, "bias")),
        _ptr(shift), _ptr(scale), stride, rows, cols, tokens_per_frame, eps, _stream()))
    return out


def gemm(a, w, bias=None, out=None, *, epilogue=EPI_BIAS, residual=None, gate=None, tokens_per_frame=0):
    """out = epilogue(a @ w.T + bias); a [M,K], w [N,K] (nn.Linear layout), gate [frames, N] or None."""
    a = _bf16_2d(a, "a")
    w = _bf16_2d(w, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"gemm: a is [{M},{K}] but w is {tuple(w.shape)}")
    if epilogue == EPI_BIAS_F32:
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        if not out.is_cuda or out.dtype != torch.float32 or out.dim() != 2 or out.stride(1) != 1:
            raise ValueError("out: EPI_BIAS_F32 writes a 2-D CUDA float32 tensor with unit inner stride")
    else:
        if out is None:
            out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        out = _bf16_2d(out, "out")
    if out.shape != (M, N):
        raise ValueError(f"gemm: out must be [{M},{N}], got {tuple(out.shape)}")
    gstride = 0
    if gate is not None:
        if gate.dim() != 2 or gate.shape[1] != N or gate.stride(1) != 1:
            raise ValueError("gate must be [frames, N] with unit inner stride")
        gstride = gate.stride(0)
    if residual is not None:
        residual = _bf16_2d(residual, "residual")
    _lib.check(_lib.load().ifx_gemm_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(_bf16_vec(bias, N, "bias")), out.data_ptr(),
        out.stride(0), M, N, K, epilogue, _ptr(residual), residual.stride(0) if residual is not None else 0,
        _ptr(gate), gstride, tokens_per_frame, _stream()))
    return out


def _fp8_2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float8_e4m3fn or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D CUDA float8_e4m3fn tensor with unit inner stride")
    return t


def quantize_fp8(x, scale: float, out=None):
    """e4m3(bf16(clamp(x / scale, +-448))) — MAGI div_clamp_to (dit_module.py:367-387) with a per-tensor scale."""
    x = _bf16_2d(x, "x")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=torch.float8_e4m3fn, device=x.device) if out is None else _fp8_2d(out, "out")
    _lib.check(_lib.load().ifx_quantize_fp8(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols,
                                            float(scale), _stream()))
    return out


def quantize_fp8_cols(x, col_scale, out=None):
    """e4m3(bf16(clamp(x[:, k] / col_scale[k], +-448))): MAGI's quantised linears divide by a per-input-channel vector
    (PerTensor input_scale [in] / PerChannel smooth_scale [1, in], dit_module.py:434-490)."""
    x = _bf16_2d(x, "x")
    rows, cols = x.shape
    col_scale = col_scale.reshape(-1)
    if not col_scale.is_cuda or col_scale.dtype != torch.float32 or col_scale.numel() != cols or not col_scale.is_contiguous():
        raise ValueError(f"col_scale: expected a contiguous CUDA float32 vector of {cols} elements")
    out = torch.empty((rows, cols), dtype=torch.float8_e4m3fn, device=x.device) if out is None else _fp8_2d(out, "out")
    _lib.check(_lib.load().ifx_quantize_fp8_cols(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols,
                                                 col_scale.data_ptr(), _stream()))
    return out


def ln_modulate_fp8(x, out_scale: float, out=None, *, weight=None, bias=None, shift=None, scale=None,
                    tokens_per_frame=0, eps=1e-6):
    """ln_modulate whose bf16 result is quantised to e4m3 on the way out (same rounding as quantize_fp8 after it)."""
    x = _bf16_2d(x, "x")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=torch.float8_e4m3fn, device=x.device) if out is None else _fp8_2d(out, "out")
    if not out.is_contiguous():
        raise ValueError("out must be contiguous")
    stride = 0
    if scale is not None:
        if shift is None or scale.shape != shift.shape or scale.stride(0) != shift.stride(0):
            raise ValueError("shift/scale must both be [frames, C] with a common frame stride")
        stride = scale.stride(0)
    _lib.check(_lib.load().ifx_ln_modulate_fp8(
        x.data_ptr(), out.data_ptr(), _ptr(_bf16_vec(weight, cols, "weight")), _ptr(_bf16_vec(bias, cols, "bias")),
        _ptr(shift), _ptr(scale), stride, rows, cols, tokens_per_frame, eps, float(out_scale), _stream()))
    return out


def gemm_fp8(a_q, w_q, alpha: float, bias=None, out=None, *, epilogue=EPI_BIAS, residual=None, gate=None,
             tokens_per_frame=0):
    """out = epilogue((a_q @ w_q.T) * alpha + bias); a_q [M,K], w_q [N,K] e4m3; alpha = input_scale * weight_scale."""
    a_q, w_q = _fp8_2d(a_q, "a_q"), _fp8_2d(w_q, "w_q")
    M, K = a_q.shape
    N = w_q.shape[0]
    if w_q.shape[1] != K:
        raise ValueError(f"gemm_fp8: a is [{M},{K}] but w is {tuple(w_q.shape)}")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a_q.device)
    out = _bf16_2d(out, "out")
    gstride = 0
    if gate is not None:
        if gate.dim() != 2 or gate.shape[1] != N or gate.stride(1) != 1:
            raise ValueError("gate must be [frames, N] with unit inner stride")
        gstride = gate.stride(0)
    if residual is not None:
        residual = _bf16_2d(residual, "residual")
    _lib.check(_lib.load().ifx_gemm_fp8(
        a_q.data_ptr(), a_q.stride(0), w_q.data_ptr(), w_q.stride(0), float(alpha), _ptr(_bf16_vec(bias, N, "bias")),
        out.data_ptr(), out.stride(0), M, N, K, epilogue, _ptr(residual),
        residual.stride(0) if residual is not None else 0, _ptr(gate), gstride, tokens_per_frame, _stream()))
    return out


# ----------------------------------------------------------------------------- dynamic 8-bit linears
Q8_E4M3, Q8_INT8 = _lib.IFX_Q8_E4M3, _lib.IFX_Q8_INT8
_Q8_DTYPE = {Q8_E4M3: torch.float8_e4m3fn, Q8_INT8: torch.int8}


def _q8_2d(t, kind, name):
    if not t.is_cuda or t.dtype != _Q8_DTYPE[kind] or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D CUDA {_Q8_DTYPE[kind]} tensor with unit inner stride")
    return t


def quantize_rows(x, kind=Q8_E4M3, out=None, scales=None):
    """Per-token dynamic quantisation: x [rows, cols] bf16 -> (codes [rows, cols] e4m3 | int8, scales [rows] fp32)."""
    x = _bf16_2d(x, "x")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=_Q8_DTYPE[kind], device=x.device) if out is None else _q8_2d(out, kind, "out")
    scales = torch.empty(rows, dtype=torch.float32, device=x.device) if scales is None else scales
    _lib.check(_lib.load().ifx_quantize_rows(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), scales.data_ptr(),
                                             rows, cols, kind, _stream()))
    return out, scales


def ln_modulate_quant(x, kind=Q8_E4M3, out=None, scales=None, *, weight=None, bias=None, shift=None, scale=None,
                      tokens_per_frame=0, eps=1e-6):
    """ln_modulate whose bf16 result is quantised per token on the way out -> (codes, scales)."""
    x = _bf16_2d(x, "x")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=_Q8_DTYPE[kind], device=x.device) if out is None else _q8_2d(out, kind, "out")
    if not out.is_contiguous():
        raise ValueError("out must be contiguous")
    scales = torch.empty(rows, dtype=torch.float32, device=x.device) if scales is None else scales
    stride = 0
    if scale is not None:
        if shift is None or scale.shape != shift.shape or scale.stride(0) != shift.stride(0):
            raise ValueError("shift/scale must both be [frames, C] with a common frame stride")
        stride = scale.stride(0)
    _lib.check(_lib.load().ifx_ln_modulate_quant(
        x.data_ptr(), out.data_ptr(), scales.data_ptr(), _ptr(_bf16_vec(weight, cols, "weight")),
        _ptr(_bf16_vec(bias, cols, "bias")), _ptr(shift), _ptr(scale), stride, rows, cols, tokens_per_frame, eps, kind,
        _stream()))
    return out, scales


def gemm_q8(a_q, w_q, row_scale, col_scale, kind=Q8_E4M3, bias=None, out=None, *, epilogue=EPI_BIAS, residual=None,
            gate=None, tokens_per_frame=0):
    """out = epilogue((a_q @ w_q.T) * row_scale[:, None] * col_scale[None, :] + bias); 8-bit codes, fp32 scales."""
    a_q, w_q = _q8_2d(a_q, kind, "a_q"), _q8_2d(w_q, kind, "w_q")
    M, K = a_q.shape
    N = w_q.shape[0]
    if w_q.shape[1] != K:
        raise ValueError(f"gemm_q8: a is [{M},{K}] but w is {tuple(w_q.shape)}")
    for t, n, name in ((row_scale, M, "row_scale"), (col_scale, N, "col_scale")):
        if not t.is_cuda or t.dtype != torch.float32 or t.numel() != n or not t.is_contiguous():
            raise ValueError(f"{name}: expected a contiguous CUDA float32 vector of {n} elements")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a_q.device)
    out = _bf16_2d(out, "out")
    gstride = 0
    if gate is not None:
        if gate.dim() != 2 or gate.shape[1] != N or gate.stride(1) != 1:
            raise ValueError("gate must be [frames, N] with unit inner stride")
        gstride = gate.stride(0)
    if residual is not None:
        residual = _bf16_2d(residual, "residual")
    _lib.check(_lib.load().ifx_gemm_q8(
        a_q.data_ptr(), a_q.stride(0), w_q.data_ptr(), w_q.stride(0), row_scale.data_ptr(), col_scale.data_ptr(), kind,
        _ptr(_bf16_vec(bias, N, "bias")), out.data_ptr(), out.stride(0), M, N, K, epilogue, _ptr(residual),
        residual.stride(0) if residual is not None else 0, _ptr(gate), gstride, tokens_per_frame, _stream()))
    return out


def quantize_weight_per_channel(w: torch.Tensor, kind=Q8_E4M3):
    """W [N, K] -> (codes [N, K], scales [N] fp32): s_w[n] = max(|W[n, :]|, 1e-12) / qmax, round-to-nearest-even with
    saturation (host-side torch arithmetic, done once per checkpoint)."""
    qmax = 448.0 if kind == Q8_E4M3 else 127.0
    wf = w.detach().float()
    s = wf.abs().amax(dim=1).clamp_min(1e-12) / qmax
    q = wf / s[:, None]
    if kind == Q8_E4M3:
        codes = q.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    else:
        codes = torch.round(q).clamp(-127, 127).to(torch.int8)
    return codes.contiguous(), s.contiguous()


def rmsnorm(x, weight, out=None, *, eps=1e-6):
    x = _bf16_2d(x, "x")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device) if out is None else _bf16_2d(out, "out")
    _lib.check(_lib.load().ifx_rmsnorm(x.data_ptr(), x.stride(0), _bf16_vec(weight, cols, "weight").data_ptr(),
                                       out.data_ptr(), out.stride(0), rows, cols, eps, _stream()))
    return out


def attention(q, k, v, heads, out=None, *, softmax_scale=None):
    """q [Lq, H*D], k/v [Lk, H*D] (same row stride) -> [Lq, H*D]; full (non-causal) softmax attention."""
    q = _bf16_2d(q, "q")
    k = _bf16_2d(k, "k")
    v = _bf16_2d(v, "v")
    width = q.shape[1]
    if k.shape != v.shape or k.shape[1] != width or k.stride(0) != v.stride(0):
        raise ValueError("attention: k and v must have identical shape/stride and match q's width")
    head_dim = width // heads
    if out is None:
        out = torch.empty((q.shape[0], width), dtype=torch.bfloat16, device=q.device)
    scale = softmax_scale if softmax_scale is not None else head_dim ** -0.5
    _lib.check(_lib.load().ifx_attention(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                         out.data_ptr(), out.stride(0), q.shape[0], k.shape[0], heads, head_dim,
                                         scale, _stream()))
    return out


def attention_gqa(q, k, v, heads, kv_heads, out=None, *, softmax_scale=None):
    """Grouped-query attention: q [Lq, heads*D], k/v [Lk, kv_heads*D]."""
    q, k, v = _bf16_2d(q, "q"), _bf16_2d(k, "k"), _bf16_2d(v, "v")
    head_dim = q.shape[1] // heads
    if k.shape != v.shape or k.shape[1] != kv_heads * head_dim or k.stride(0) != v.stride(0):
        raise ValueError("attention_gqa: k / v must be [Lk, kv_heads*D] with identical strides")
    if out is None:
        out = torch.empty((q.shape[0], q.shape[1]), dtype=torch.bfloat16, device=q.device)
    scale = softmax_scale if softmax_scale is not None else head_dim ** -0.5
    _lib.check(_lib.load().ifx_attention_gqa(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                             out.data_ptr(), out.stride(0), q.shape[0], k.shape[0], heads, kv_heads,
                                             head_dim, scale, _stream()))
    return out


def attention_lse(q, k, v, heads, kv_heads=None, out=None, *, softmax_scale=None):
    """(out [Lq, heads*D] bf16, lse [heads, Lq] fp32): attention plus the log-sum-exp of the scaled scores, the pair the
    reference's attention backends return (backends.py:58-72)."""
    q, k, v = _bf16_2d(q, "q"), _bf16_2d(k, "k"), _bf16_2d(v, "v")
    kv_heads = kv_heads or heads
    head_dim = q.shape[1] // heads
    if k.shape != v.shape or k.shape[1] != kv_heads * head_dim or k.stride(0) != v.stride(0):
        raise ValueError("attention_lse: k / v must be [Lk, kv_heads*D] with identical strides")
    if out is None:
        out = torch.empty((q.shape[0], q.shape[1]), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((heads, q.shape[0]), dtype=torch.float32, device=q.device)
    scale = softmax_scale if softmax_scale is not None else head_dim ** -0.5
    _lib.check(_lib.load().ifx_attention_lse(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                             out.data_ptr(), out.stride(0), lse.data_ptr(), q.shape[0], k.shape[0], heads,
                                             kv_heads, head_dim, scale, _stream()))
    return out, lse


def attention_ranges(q, k, v, q_ranges, k_ranges, heads, kv_heads, *, softmax_scale=None):
    """MAGI range attention (dit_module.py:1000-1014): output rows q_ranges[i] attend keys k_ranges[i].
    q [Sq, heads*D], k/v [Sk, kv_heads*D]; ranges are host int pairs [start, end) in tokens (np_q_range / np_k_range)."""
    out = torch.empty_like(q)
    for (qs, qe), (ks, ke) in zip(q_ranges, k_ranges):
        attention_gqa(q[qs:qe], k[ks:ke], v[ks:ke], heads, kv_heads, out[qs:qe], softmax_scale=softmax_scale)
    return out


def _coalesce_pages(pages, page_tokens):
    """sorted physical pages -> list of (first_row, rows) runs of consecutive pages."""
    runs = []
    for pg in sorted(pages):
        if runs and runs[-1][0] + runs[-1][1] == pg * page_tokens:
            runs[-1][1] += page_tokens
        else:
            runs.append([pg * page_tokens, page_tokens])
    return [tuple(r) for r in runs]


def attention_partial(q, k, v, extents, heads, workspace, pieces_per_item, piece_first, piece_count, *,
                      kv_heads=None, softmax_scale=None):
    """Phase of a two-phase attention: q [Lq, heads*D] against the key-row extents [(row0, rows), ...] of k/v;
    un-normalised partials go to `workspace` (float32, see attention_workspace_bytes)."""
    q, k, v = _bf16_2d(q, "q"), _bf16_2d(k, "k"), _bf16_2d(v, "v")
    kv_heads = kv_heads or heads
    head_dim = q.shape[1] // heads
    ext = (C.c_int64 * (2 * len(extents)))(*[int(x) for e in extents for x in e])
    scale = softmax_scale if softmax_scale is not None else head_dim ** -0.5
    _lib.check(_lib.load().ifx_attention_partial(
        q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), k.shape[0], ext, len(extents), q.shape[0],
        heads, kv_heads, head_dim, scale, workspace.data_ptr(), workspace.numel() * workspace.element_size(),
        pieces_per_item, piece_first, piece_count, _stream()))


def attention_combine(workspace, pieces_per_item, out, heads):
    out = _bf16_2d(out, "out")
    _lib.check(_lib.load().ifx_attention_combine(workspace.data_ptr(), pieces_per_item, out.data_ptr(), out.stride(0),
                                                 out.shape[0], heads, out.shape[1] // heads, _stream()))
    return out


def attention_extents(q, k, v, extents, heads, out=None, *, kv_heads=None, softmax_scale=None):
    """q [Lq, heads*D] against the key-row extents [(row0, rows), ...] of k/v (runs of consecutive cache pages).  Rows
    that follow an extent in memory are never attended, whatever they hold."""
    q, k, v = _bf16_2d(q, "q"), _bf16_2d(k, "k"), _bf16_2d(v, "v")
    kv_heads = kv_heads or heads
    head_dim = q.shape[1] // heads
    if k.shape != v.shape or k.stride(0) != v.stride(0):
        raise ValueError("attention_extents: k and v must have identical shape / stride")
    if out is None:
        out = torch.empty((q.shape[0], q.shape[1]), dtype=torch.bfloat16, device=q.device)
    ext = (C.c_int64 * (2 * len(extents)))(*[int(x) for e in extents for x in e])
    scale = softmax_scale if softmax_scale is not None else head_dim ** -0.5
    _lib.check(_lib.load().ifx_attention_extents(
        q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), k.shape[0], ext, len(extents),
        out.data_ptr(), out.stride(0), q.shape[0], heads, kv_heads, head_dim, scale, _stream()))
    return out


def attention_workspace_bytes(q_rows: int, heads: int, pieces_per_item: int) -> int:
    return heads * ((q_rows + 255) // 256) * pieces_per_item * 256 * 130 * 4


class OffloadSlots:
    """Device staging slots of the KV offload tier: `count` (K, V) buffer pairs of one layer's geometry shared by all
    layers of a request group, a copy stream, and per-slot events.  Layers take slots round-robin: while layer i
    computes out of slot i % count, layer i + 1 is staged into the next slot on the copy stream."""

    def __init__(self, count: int, rows: int, width: int, device):
        self.count, self.device = count, torch.device(device)
        self.k = [torch.empty((rows, width), dtype=torch.bfloat16, device=device) for _ in range(count)]
        self.v = [torch.empty((rows, width), dtype=torch.bfloat16, device=device) for _ in range(count)]
        self.stream = torch.cuda.Stream(device=device)
        self.owner = [None] * count                  # PagedKV currently staged in the slot
        self.free_event = [None] * count             # recorded when the slot's last user finished (compute + write-back)
        self.next = 0

    def take(self):
        i = self.next
        self.next = (self.next + 1) % self.count
        return i


class PagedKV:
    """One layer's self-attention cache: two bf16 buffers + the native block table (ifx_kv).

    offload=OffloadSlots: the reference's kv_offload tier (kvcache_manager.py:222-244 pinned CPU tensors, get() copies
    them to the GPU).  The window lives in pinned host memory; `stage()` copies its valid prefix into a device slot on
    the copy stream and re-points the native handle at it, `write_back()` copies the pages a block forward wrote back
    to the host.  HBM then holds `slots.count` layers instead of all of them."""

    def __init__(self, num_pages: int, page_tokens: int, heads: int, head_dim: int, device, offload: "OffloadSlots" = None):
        self.num_pages, self.page_tokens, self.heads, self.head_dim = num_pages, page_tokens, heads, head_dim
        width = heads * head_dim
        self.offload, self.slot = offload, None
        if offload is None:
            # like the reference (torch.empty, kvcache_manager.py:232) the memory starts uninitialised
            self.k = torch.empty((num_pages * page_tokens, width), dtype=torch.bfloat16, device=device)
            self.v = torch.empty_like(self.k)
        else:
            self.host_k = torch.empty((num_pages * page_tokens, width), dtype=torch.bfloat16, pin_memory=True)
            self.host_v = torch.empty_like(self.host_k).pin_memory()
            self.k, self.v = offload.k[0], offload.v[0]      # re-pointed by stage()
            self._ready = None                               # event: staging copy done
        h = C.c_void_p()
        _lib.check(_lib.load().ifx_kv_create(C.byref(h), self.k.data_ptr(), self.v.data_ptr(), num_pages, page_tokens,
                                             heads, head_dim))
        self._h = h

    @property
    def handle(self) -> int:
        if self._h is None:
            raise KeyError("PagedKV has been freed")
        return self._h.value

    def free(self) -> None:
        if self._h is not None:
            _lib.check(_lib.load().ifx_kv_destroy(self._h))
            self._h = None
            self.k = self.v = None
            if self.offload is not None:
                if self.slot is not None and self.offload.owner[self.slot] is self:
                    self.offload.owner[self.slot] = None
                self.host_k = self.host_v = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def reset(self) -> None:
        _lib.check(_lib.load().ifx_kv_reset(self.handle))

    def repage(self, page_tokens: int) -> None:
        """Re-cut the (still empty) cache into pages of `page_tokens` rows.  The reference's allocation call knows no
        page size (block_size=1, self_forcing_kv_cache_manager.py:47); the block forward calls this with the frame
        size the first time it sees such a cache, so the reference-shaped call stays usable.  Same buffers, new table."""
        if page_tokens == self.page_tokens:
            return
        total = self.num_pages * self.page_tokens
        g, l, table = self.state()
        if g or l or table:
            raise ValueError(f"cannot re-page a cache that already holds {l} tokens; allocate it with "
                             f"page_tokens={page_tokens} (allocate_kv_cache(..., page_tokens=tokens per latent frame))")
        if self.offload is not None:
            raise ValueError("the KV offload tier needs page_tokens at allocation time")
        if page_tokens <= 0 or total % page_tokens:
            raise ValueError(f"cache of {total} tokens is not a whole number of {page_tokens}-token frames")
        lib = _lib.load()
        _lib.check(lib.ifx_kv_destroy(self._h))
        h = C.c_void_p()
        _lib.check(lib.ifx_kv_create(C.byref(h), self.k.data_ptr(), self.v.data_ptr(), total // page_tokens,
                                     page_tokens, self.heads, self.head_dim))
        self._h, self.num_pages, self.page_tokens = h, total // page_tokens, page_tokens

    # ------------------------------------------------------------------ offload tier
    def stage(self) -> None:
        """Make this layer's window resident in a device slot (no-op for HBM-resident caches and when already staged).
        The copy runs on the slots' stream; `wait_staged()` orders the caller's stream after it."""
        sl = self.offload
        if sl is None or (self.slot is not None and sl.owner[self.slot] is self):
            return
        i = sl.take()
        prev = sl.owner[i]
        if prev is not None:
            prev.slot = None
        with torch.cuda.stream(sl.stream):
            if sl.free_event[i] is not None:
                sl.stream.wait_event(sl.free_event[i])       # the slot's previous layer has computed and written back
            _, local_end, table = self.state()
            rows = len(table) * self.page_tokens            # mapped pages = physical prefix (allocator invariant)
            if rows:
                sl.k[i][:rows].copy_(self.host_k[:rows], non_blocking=True)
                sl.v[i][:rows].copy_(self.host_v[:rows], non_blocking=True)
            self._ready = torch.cuda.Event()
            self._ready.record(sl.stream)
        sl.owner[i], self.slot = self, i
        self.k, self.v = sl.k[i], sl.v[i]
        _lib.check(_lib.load().ifx_kv_rebind(self.handle, self.k.data_ptr(), self.v.data_ptr()))

    def wait_staged(self) -> None:
        if self.offload is not None and self._ready is not None:
            torch.cuda.current_stream().wait_event(self._ready)

    def write_back(self, pages) -> None:
        """Copy the given physical pages (those a block forward just wrote) from the device slot to the host tier and
        mark the slot reusable once that is done."""
        sl = self.offload
        if sl is None:
            return
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        pt = self.page_tokens
        with torch.cuda.stream(sl.stream):
            sl.stream.wait_event(done)
            for r0, n in _coalesce_pages(pages, pt):
                self.host_k[r0:r0 + n].copy_(self.k[r0:r0 + n], non_blocking=True)
                self.host_v[r0:r0 + n].copy_(self.v[r0:r0 + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(sl.stream)
        sl.free_event[self.slot] = ev

    def plan_append(self, current_start: int, num_new: int, sink_tokens: int = 0, windowed: bool = True) -> KvPlan:
        plan = KvPlan()
        _lib.check(_lib.load().ifx_kv_plan_append(self.handle, current_start, num_new, sink_tokens, int(windowed),
                                                  C.byref(plan)))
        return plan

    def state(self):
        g, l, n = C.c_int64(), C.c_int64(), C.c_int32()
        table = (C.c_int32 * self.num_pages)()
        _lib.check(_lib.load().ifx_kv_state(self.handle, C.byref(g), C.byref(l), C.byref(n), table, self.num_pages))
        return g.value, l.value, list(table[: n.value])

    def _resident(self) -> None:
        """offload tier: stage the window into a device slot and order the current stream after the copy"""
        if self.offload is not None:
            self.stage()
            self.wait_staged()

    def append(self, plan: KvPlan, k_rows: torch.Tensor, v_rows: torch.Tensor) -> None:
        k_rows, v_rows = _bf16_2d(k_rows, "k_rows"), _bf16_2d(v_rows, "v_rows")
        if k_rows.stride(0) != v_rows.stride(0) or k_rows.shape != v_rows.shape:
            raise ValueError("k_rows / v_rows must share shape and stride")
        self._resident()
        _lib.check(_lib.load().ifx_kv_append(self.handle, C.byref(plan), k_rows.data_ptr(), v_rows.data_ptr(),
                                             k_rows.stride(0), k_rows.shape[0], _stream()))
        self.write_back(list(plan.pages[: plan.num_pages]))

    def append_sp(self, plan: KvPlan, k_gathered: torch.Tensor, v_gathered: torch.Tensor, frames: int) -> None:
        """k/v_gathered: [world, frames*chunk, H*D] rank-major views (rows contiguous, any common rank stride — e.g.
        the two halves of one [world, 2, rows, H*D] all-gather output) -> (frame, rank, hw) token order."""
        ok = (k_gathered.dim() == 3 and k_gathered.shape == v_gathered.shape and k_gathered.dtype == torch.bfloat16
              and k_gathered.stride(2) == 1 and k_gathered.stride(1) == k_gathered.shape[2]
              and k_gathered.stride() == v_gathered.stride())
        if not ok:
            raise ValueError("append_sp: bf16 [world, frames*chunk, H*D] views with contiguous rows expected")
        world, rows, _ = k_gathered.shape
        _lib.check(_lib.load().ifx_kv_append_sp(self.handle, C.byref(plan), k_gathered.data_ptr(),
                                                v_gathered.data_ptr(), k_gathered.stride(0), world, frames,
                                                rows // frames, _stream()))

    def split_extents(self, plan: KvPlan):
        """(old, new): physical key-row extents of the valid pages NOT written by `plan` and of the pages it writes.
        Attention is invariant to key order, so the two groups can be attended separately and merged."""
        _, _, table = self.state()
        new_pages = list(plan.pages[: plan.num_pages])
        old_pages = [pg for pg in table if pg not in set(new_pages)]
        return _coalesce_pages(old_pages, self.page_tokens), _coalesce_pages(new_pages, self.page_tokens)

    def map_rows(self, tokens: int):
        """Map logical tokens [0, tokens) (identity order; never-rotated caches only) and return the row views
        (k, v), each [tokens, H*D], aliasing the cache memory."""
        _lib.check(_lib.load().ifx_kv_map(self.handle, tokens, None, None))
        return self.k[:tokens], self.v[:tokens]

    def export(self, start: int, length: int):
        """Tokens [start, start+length) in the reference's logical order -> (k, v) each [length, H*D]."""
        width = self.heads * self.head_dim
        k = torch.empty((length, width), dtype=torch.bfloat16, device=self.k.device)
        v = torch.empty_like(k)
        self._resident()
        _lib.check(_lib.load().ifx_kv_export(self.handle, k.data_ptr(), v.data_ptr(), start, length, _stream()))
        self.write_back([])            # nothing written; releases the slot after this read
        return k, v

    def import_(self, start: int, k_rows: torch.Tensor, v_rows: torch.Tensor) -> None:
        k_rows, v_rows = _bf16_2d(k_rows, "k_rows"), _bf16_2d(v_rows, "v_rows")
        if not (k_rows.is_contiguous() and v_rows.is_contiguous()) or k_rows.shape != v_rows.shape:
            raise ValueError("import_: contiguous [length, H*D] tensors expected")
        self._resident()
        _lib.check(_lib.load().ifx_kv_import(self.handle, k_rows.data_ptr(), v_rows.data_ptr(), start,
                                             k_rows.shape[0], _stream()))
        if self.offload is not None:
            _, _, table = self.state()
            pt = self.page_tokens
            self.write_back(table[start // pt:(start + k_rows.shape[0] + pt - 1) // pt])

    def attention(self, q: torch.Tensor, out=None, *, softmax_scale=None, fresh: Optional[KvPlan] = None,
                  flags: Optional[torch.Tensor] = None, epoch: int = 0, timeout_ms: int = 60000):
        """Attention of q over the cached window, read through the block table.  With `fresh` (the plan of the block
        being appended) and `flags` (int64 [world] epoch flags of the peer-memory exchange) the fresh pages are
        attended last, after every rank's flag reached `epoch` (ifx_attention_kv_wait)."""
        q = _bf16_2d(q, "q")
        if out is None:
            out = torch.empty((q.shape[0], self.heads * self.head_dim), dtype=torch.bfloat16, device=q.device)
        scale = softmax_scale if softmax_scale is not None else self.head_dim ** -0.5
        if fresh is None:
            _lib.check(_lib.load().ifx_attention_kv(q.data_ptr(), q.stride(0), self.handle, out.data_ptr(),
                                                    out.stride(0), q.shape[0], scale, _stream()))
        else:
            if flags is None or flags.dtype != torch.int64 or not flags.is_cuda or not flags.is_contiguous():
                raise ValueError("flags must be a contiguous CUDA int64 vector (one epoch flag per rank)")
            _lib.check(_lib.load().ifx_attention_kv_wait(
                q.data_ptr(), q.stride(0), self.handle, C.byref(fresh), flags.data_ptr(), flags.numel(), int(epoch),
                int(timeout_ms), out.data_ptr(), out.stride(0), q.shape[0], scale, _stream()))
        return out


def qk_norm_rope_append(qkv, norm_q_w, norm_k_w, freqs_table, grid: RopeGrid, heads, head_dim, *, kv: PagedKV = None,
                        plan: KvPlan = None, q_out=None, k_out=None, v_out=None, eps=1e-6):
    """Fused QK-RMSNorm + RoPE + append.  With kv/plan the K/V rows land in the cache pages; otherwise in
    k_out/v_out (contiguous staging for the sequence-parallel all-gather)."""
    qkv = _bf16_2d(qkv, "qkv")
    rows = qkv.shape[0]
    C_ = heads * head_dim
    if qkv.shape[1] != 3 * C_:
        raise ValueError("qkv must be [rows, 3*heads*head_dim]")
    if freqs_table.dtype != torch.float64 or not freqs_table.is_cuda or not freqs_table.is_contiguous():
        raise ValueError("freqs_table: use ops.rope_table(model.freqs, device)")
    if q_out is None:
        q_out = torch.empty((rows, C_), dtype=torch.bfloat16, device=qkv.device)
    if kv is None:
        if k_out is None:
            k_out = torch.empty((rows, C_), dtype=torch.bfloat16, device=qkv.device)
            v_out = torch.empty_like(k_out)
        if not (k_out.is_contiguous() and v_out.is_contiguous()):
            raise ValueError("k_out / v_out must be contiguous")
    _lib.check(_lib.load().ifx_qk_norm_rope_append(
        qkv.data_ptr(), qkv.stride(0), _bf16_vec(norm_q_w, C_, "norm_q").data_ptr(),
        _bf16_vec(norm_k_w, C_, "norm_k").data_ptr(), freqs_table.data_ptr(), C.byref(grid), q_out.data_ptr(),
        q_out.stride(0), kv.handle if kv is not None else None, C.byref(plan) if plan is not None else None,
        _ptr(k_out), _ptr(v_out), rows, heads, head_dim, eps, _stream()))
    return q_out, k_out, v_out


def qk_norm_rope_append_peers(qkv, norm_q_w, norm_k_w, freqs_table, grid: RopeGrid, heads, head_dim, kv: "PagedKV",
                              plan: KvPlan, peers, *, q_out=None, eps=1e-6):
    """Sequence-parallel form of qk_norm_rope_append: this rank's K / V rows are stored into every rank's replicated
    cache (peers: an `_lib.PeerDst` built by inferix_b200.peer) and the epoch flag is published."""
    qkv = _bf16_2d(qkv, "qkv")
    rows = qkv.shape[0]
    C_ = heads * head_dim
    if qkv.shape[1] != 3 * C_:
        raise ValueError("qkv must be [rows, 3*heads*head_dim]")
    if freqs_table.dtype != torch.float64 or not freqs_table.is_cuda or not freqs_table.is_contiguous():
        raise ValueError("freqs_table: use ops.rope_table(model.freqs, device)")
    if q_out is None:
        q_out = torch.empty((rows, C_), dtype=torch.bfloat16, device=qkv.device)
    _lib.check(_lib.load().ifx_qk_norm_rope_append_peers(
        qkv.data_ptr(), qkv.stride(0), _bf16_vec(norm_q_w, C_, "norm_q").data_ptr(),
        _bf16_vec(norm_k_w, C_, "norm_k").data_ptr(), freqs_table.data_ptr(), C.byref(grid), q_out.data_ptr(),
        q_out.stride(0), kv.handle, C.byref(plan), C.byref(peers), rows, heads, head_dim, eps, _stream()))
    return q_out


def peer_push(kv: "PagedKV", plan: KvPlan, peers, frames: int, chunk: int, *, ctas: int = 4):
    """Exchange half of qk_norm_rope_append_peers as its own small grid (see ifx_peer_push): copies this rank's rows of
    the block's new pages to every other rank's cache and publishes peers.epoch.  Runs on the current stream."""
    _lib.check(_lib.load().ifx_peer_push(kv.handle, C.byref(plan), C.byref(peers), frames, chunk, ctas, _stream()))


# ----------------------------------------------------------------------------- MAGI-1 layer row kernels
def _f32_vec(t: torch.Tensor, n: int, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float32 or t.numel() != n or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA float32 vector of {n} elements")
    return t


def magi_qkv_post(qkvx, q_heads, kv_heads, q_ln, k_ln, qx_ln, rope, q_out, k_dst, v_dst, qx_out, *, eps=1e-6,
                  groups=1):
    """dit_module.py:902-958 in one pass over the fused projection qkvx [rows, (2*q_heads + 2*kv_heads)*128]
    (q | k | v | qx).  q_ln / k_ln: (weight, bias) fp32 [128]; qx_ln: (weight, bias) bf16 [128]; rope fp32
    [rows, 2*half] = sin | cos.
    groups == 1: q_out [rows, q_heads*128]; k_dst / v_dst [rows, kv_heads*128] row views with a common stride
    (normally rows of the layer's KV cache).
    groups == cp > 1 (Ulysses send layout): q_out [cp, rows, q_heads/cp*128], k_dst / v_dst [cp, rows, kv_heads/cp*128],
    contiguous."""
    qkvx = _bf16_2d(qkvx, "qkvx")
    rows = qkvx.shape[0]
    d = 128
    if qkvx.shape[1] != (2 * q_heads + 2 * kv_heads) * d:
        raise ValueError("qkvx must be [rows, (2*q_heads + 2*kv_heads)*128]")
    if rope.dtype != torch.float32 or not rope.is_cuda or rope.dim() != 2 or rope.shape[0] != rows or rope.stride(1) != 1:
        raise ValueError("rope must be a CUDA float32 [rows, 2*half] tensor")
    qx_out = _bf16_2d(qx_out, "qx_out")
    if qx_out.shape != (rows, q_heads * d):
        raise ValueError("qx_out must be [rows, q_heads*128]")
    if q_heads % groups or kv_heads % groups:
        raise ValueError("groups must divide q_heads and kv_heads")
    qg, kg = q_heads // groups, kv_heads // groups
    if groups == 1:
        q_out, k_dst, v_dst = _bf16_2d(q_out, "q_out"), _bf16_2d(k_dst, "k_dst"), _bf16_2d(v_dst, "v_dst")
        if k_dst.shape != (rows, kv_heads * d) or v_dst.shape != k_dst.shape or k_dst.stride(0) != v_dst.stride(0):
            raise ValueError("k_dst / v_dst must be [rows, kv_heads*128] views with a common row stride")
        if q_out.shape != (rows, q_heads * d):
            raise ValueError("q_out must be [rows, q_heads*128]")
        ld_q, ld_kv, qgs, kgs = q_out.stride(0), k_dst.stride(0), 0, 0
    else:
        for t, w, name in ((q_out, qg, "q_out"), (k_dst, kg, "k_dst"), (v_dst, kg, "v_dst")):
            if (t.dtype != torch.bfloat16 or not t.is_cuda or t.shape != (groups, rows, w * d) or not t.is_contiguous()):
                raise ValueError(f"{name}: expected a contiguous CUDA bf16 [{groups}, {rows}, {w * d}] tensor")
        ld_q, ld_kv, qgs, kgs = qg * d, kg * d, rows * qg * d, rows * kg * d
    _lib.check(_lib.load().ifx_magi_qkv_post(
        qkvx.data_ptr(), qkvx.stride(0), rows, q_heads, kv_heads, d,
        _f32_vec(q_ln[0], d, "q_layernorm.weight").data_ptr(), _f32_vec(q_ln[1], d, "q_layernorm.bias").data_ptr(),
        _f32_vec(k_ln[0], d, "k_layernorm.weight").data_ptr(), _f32_vec(k_ln[1], d, "k_layernorm.bias").data_ptr(),
        _bf16_vec(qx_ln[0], d, "q_layernorm_xattn.weight").data_ptr(),
        _bf16_vec(qx_ln[1], d, "q_layernorm_xattn.bias").data_ptr(), rope.data_ptr(), rope.stride(0),
        rope.shape[1] // 2, eps, q_out.data_ptr(), ld_q, qg, qgs, k_dst.data_ptr(), v_dst.data_ptr(), ld_kv, kg, kgs,
        qx_out.data_ptr(), qx_out.stride(0), _stream()))


def head_layernorm(x, heads, weight, bias, out=None, *, eps=1e-6):
    """Per-head LayerNorm (bf16 affine) on x [rows, heads*128] (any row stride); in place when out is None."""
    x = _bf16_2d(x, "x")
    out = x if out is None else _bf16_2d(out, "out")
    if x.shape[1] != heads * 128 or out.shape != x.shape:
        raise ValueError("head_layernorm: x / out must be [rows, heads*128]")
    _lib.check(_lib.load().ifx_head_layernorm(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0],
                                              heads, 128, _bf16_vec(weight, 128, "weight").data_ptr(),
                                              _bf16_vec(bias, 128, "bias").data_ptr(), eps, _stream()))
    return out


def gate_norm_residual(x, gate, row_map, norm_w, norm_b, residual, out=None, *, eps=1e-6):
    """bias_modulate_add (dit_module.py:295-313): bf16(LN_fp32(x * gate[row_map]) * norm_w + norm_b + residual).
    x bf16 or fp32 [rows, C]; gate bf16 [ranges, C]; row_map int32 [rows] (condition_map); norm_w / norm_b fp32 [C]."""
    gate, residual = _bf16_2d(gate, "gate"), _bf16_2d(residual, "residual")
    x_f32 = x.dtype == torch.float32
    if x_f32:
        if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1:
            raise ValueError("x: expected a 2-D CUDA tensor with unit inner stride")
    else:
        x = _bf16_2d(x, "x")
    rows, cols = x.shape
    if gate.shape[1] != cols or residual.shape != x.shape:
        raise ValueError("gate_norm_residual: gate must be [ranges, C] and residual [rows, C]")
    if row_map.dtype != torch.int32 or not row_map.is_cuda or row_map.numel() != rows or not row_map.is_contiguous():
        raise ValueError("row_map must be a contiguous CUDA int32 vector with one entry per row")
    out = torch.empty_like(residual) if out is None else _bf16_2d(out, "out")
    _lib.check(_lib.load().ifx_gate_norm_residual(
        x.data_ptr(), x.stride(0), int(x_f32), gate.data_ptr(), gate.stride(0), gate.shape[0], row_map.data_ptr(),
        _f32_vec(norm_w, cols, "norm weight").data_ptr(), _f32_vec(norm_b, cols, "norm bias").data_ptr(),
        residual.data_ptr(), residual.stride(0), out.data_ptr(), out.stride(0), rows, cols, eps, _stream()))
    return out


def silu_mul(x, out=None):
    """flashinfer silu_and_mul: x [rows, 2F] -> bf16(silu(x[:, :F]) * x[:, F:])."""
    x = _bf16_2d(x, "x")
    rows, two_f = x.shape
    out = torch.empty((rows, two_f // 2), dtype=torch.bfloat16, device=x.device) if out is None else _bf16_2d(out, "out")
    _lib.check(_lib.load().ifx_silu_mul(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, two_f // 2,
                                        _stream()))
    return out


# ----------------------------------------------------------------------------- forward prologue / epilogue
def patchify(x, patch, hw_offset=0, hw_count=None, out=None):
    """x: latent [C_in, F, H, W] bf16 (any strides) -> this rank's token rows [F/pt * hw_count, C_in*pt*ph*pw] (the A
    operand of the patch-embedding GEMM; column order = Conv3d weight flattening)."""
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4:
        raise ValueError("patchify: expected a CUDA bfloat16 [C_in, F, H, W] tensor")
    c_in, f, hh, ww = x.shape
    pt, ph, pw = patch
    if f % pt or hh % ph or ww % pw:
        raise ValueError("patchify: the latent must be a whole number of patches")
    gh, gw = hh // ph, ww // pw
    hw_count = gh * gw if hw_count is None else hw_count
    rows, k = (f // pt) * hw_count, c_in * pt * ph * pw
    out = torch.empty((rows, k), dtype=torch.bfloat16, device=x.device) if out is None else out
    _lib.check(_lib.load().ifx_patchify(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), x.stride(3), c_in, pt, ph, pw,
                                        f // pt, gh, gw, hw_offset, hw_count, out.data_ptr(), _stream()))
    return out


def _f64_vec(t, name):
    if not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous() or t.dim() != 1:
        raise ValueError(f"{name}: expected a contiguous 1-D CUDA float64 tensor")
    return t


def sinusoidal_embedding(positions, dim):
    """wan_base/components.py:11-31: positions fp64 [n] -> bf16 [n, dim] = [cos | sin]."""
    positions = _f64_vec(positions, "positions")
    out = torch.empty((positions.numel(), dim), dtype=torch.bfloat16, device=positions.device)
    _lib.check(_lib.load().ifx_sinusoidal_embedding(positions.data_ptr(), positions.numel(), dim, out.data_ptr(), _stream()))
    return out


def linear_small(x, w, bias=None, *, silu_input=False, mod_table=None):
    """nn.Linear on <= 8 rows: bf16(SiLU?(x) @ w.T + bias).  mod_table [layers, N]: returns [layers, M, N] =
    bf16(mod_table[l] + that)."""
    x, w = _bf16_2d(x, "x"), _bf16_2d(w, "w")
    m, k = x.shape
    n = w.shape[0]
    if w.shape[1] != k:
        raise ValueError("linear_small: shape mismatch")
    if mod_table is None:
        out = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
        layers, ms, os_ = 0, 0, 0
    else:
        mod_table = _bf16_2d(mod_table, "mod_table")
        if mod_table.shape[1] != n or not mod_table.is_contiguous():
            raise ValueError("mod_table must be contiguous [layers, N]")
        layers = mod_table.shape[0]
        out = torch.empty((layers, m, n), dtype=torch.bfloat16, device=x.device)
        ms, os_ = n, m * n
    _lib.check(_lib.load().ifx_linear_small(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0),
                                            _ptr(_bf16_vec(bias, n, "bias")), out.data_ptr(), n, m, n, k, int(silu_input),
                                            _ptr(mod_table), layers, ms, os_, _stream()))
    return out


def _sigma_tables(timesteps, sigmas):
    for t, name in ((timesteps, "timesteps"), (sigmas, "sigmas")):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.dim() != 1:
            raise ValueError(f"{name}: expected a contiguous 1-D CUDA float32 tensor (the scheduler's table)")
    if timesteps.numel() != sigmas.numel():
        raise ValueError("timesteps / sigmas tables differ in length")
    return timesteps, sigmas


def unpatchify_x0(head_tokens, xt, timestep, timesteps_table, sigmas_table, patch_hw, want_flow=True):
    """head_tokens [F*gh*gw, ph*pw*C] bf16, xt [F, C, H, W] bf16 (any strides), timestep fp64 [F] ->
    (flow [F, C, H, W] or None, x0 [F, C, H, W]): unpatchify + x0 = x_t - sigma_t * flow in fp64."""
    head_tokens = _bf16_2d(head_tokens, "head_tokens")
    if not head_tokens.is_contiguous():
        raise ValueError("head_tokens must be contiguous")
    f, c, hh, ww = xt.shape
    ph, pw = patch_hw
    gh, gw = hh // ph, ww // pw
    if head_tokens.shape != (f * gh * gw, ph * pw * c) or xt.dtype != torch.bfloat16 or not xt.is_cuda:
        raise ValueError("unpatchify_x0: head_tokens / xt shapes do not match")
    timestep = _f64_vec(timestep, "timestep")
    tt, ss = _sigma_tables(timesteps_table, sigmas_table)
    x0 = torch.empty((f, c, hh, ww), dtype=torch.bfloat16, device=xt.device)
    flow = torch.empty_like(x0) if want_flow else None
    _lib.check(_lib.load().ifx_unpatchify_x0(head_tokens.data_ptr(), xt.data_ptr(), xt.stride(0), xt.stride(1), xt.stride(2),
                                             xt.stride(3), timestep.data_ptr(), tt.data_ptr(), ss.data_ptr(), tt.numel(), f, c,
                                             gh, gw, ph, pw, _ptr(flow), x0.data_ptr(), _stream()))
    return flow, x0


def add_noise(x0, noise, timestep, timesteps_table, sigmas_table):
    """FlowMatchScheduler.add_noise: bf16((1 - sigma_f) * x0 + sigma_f * noise), x0 / noise [F, ...] contiguous bf16."""
    if x0.shape != noise.shape or x0.dtype != torch.bfloat16 or noise.dtype != torch.bfloat16 or not x0.is_cuda:
        raise ValueError("add_noise: x0 / noise must be CUDA bfloat16 tensors of one shape")
    x0, noise = x0.contiguous(), noise.contiguous()
    timestep = _f64_vec(timestep, "timestep")
    tt, ss = _sigma_tables(timesteps_table, sigmas_table)
    out = torch.empty_like(noise)
    f = x0.shape[0]
    _lib.check(_lib.load().ifx_add_noise(x0.data_ptr(), noise.data_ptr(), timestep.data_ptr(), tt.data_ptr(), ss.data_ptr(),
                                         tt.numel(), f, x0.numel() // f, out.data_ptr(), _stream()))
    return out
