"""TEST DOUBLE for `inferix_b200.ops`, CPU only — lets the host orchestration of inferix_b200/magi_layer.py (weight
packing / permutations, KV row placement and save-restore, range bookkeeping, the Ulysses all-to-all layouts over
gloo) run without a GPU.  Each function has the signature of the real wrapper and restates the kernel's contract
with the oracle's torch ops.  The product never imports this file; the GPU tests exercise the real kernels."""
import types

import torch
import torch.nn.functional as F

from oracle import magi_oracle as mo

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_GATE_RES, EPI_BIAS_GELU_ERF, EPI_BIAS_F32 = 0, 1, 2, 3, 4
calls = []


def ln_modulate(x, out=None, *, weight=None, bias=None, shift=None, scale=None, tokens_per_frame=0, eps=1e-6):
    calls.append("ln_modulate")
    assert shift is None and scale is None
    out.copy_(F.layer_norm(x, (x.shape[1],), weight, bias, eps))
    return out


def gemm(a, w, bias=None, out=None, *, epilogue=EPI_BIAS, residual=None, gate=None, tokens_per_frame=0):
    calls.append("gemm")
    if epilogue == EPI_BIAS_F32:
        assert out.dtype == torch.float32
        y = F.linear(a.float(), w.float(), None if bias is None else bias.float())
    else:
        y = F.linear(a, w, bias)
        if epilogue == EPI_BIAS_GELU_ERF:
            y = F.gelu(y)
        else:
            assert epilogue == EPI_BIAS
    out.copy_(y)
    return out


def magi_qkv_post(qkvx, q_heads, kv_heads, q_ln, k_ln, qx_ln, rope, q_out, k_dst, v_dst, qx_out, *, eps=1e-6, groups=1):
    calls.append("magi_qkv_post")
    d = 128
    rows = qkvx.shape[0]
    sin, cos = rope.tensor_split(2, -1)
    q, k, v, qx = torch.split(qkvx, [q_heads * d, kv_heads * d, kv_heads * d, q_heads * d], dim=1)

    def roped(t, heads, ln):
        t = F.layer_norm(t.reshape(rows, heads, d).float(), (d,), ln[0], ln[1], eps)
        return mo.apply_rotary(t[None], cos, sin)[0].to(torch.bfloat16).reshape(rows, heads * d)

    qn, kn = roped(q, q_heads, q_ln), roped(k, kv_heads, k_ln)
    qx_out.copy_(F.layer_norm(qx.reshape(rows, q_heads, d), (d,), qx_ln[0], qx_ln[1], eps).reshape(rows, -1))
    if groups == 1:
        q_out.copy_(qn)
        k_dst.copy_(kn)
        v_dst.copy_(v)
    else:
        q_out.copy_(qn.view(rows, groups, -1).transpose(0, 1))
        k_dst.copy_(kn.view(rows, groups, -1).transpose(0, 1))
        v_dst.copy_(v.reshape(rows, groups, -1).transpose(0, 1))


def head_layernorm(x, heads, weight, bias, out=None, *, eps=1e-6):
    calls.append("head_layernorm")
    out = x if out is None else out
    out.copy_(F.layer_norm(x.reshape(x.shape[0], heads, 128), (128,), weight, bias, eps).reshape(x.shape[0], -1))
    return out


def attention_gqa(q, k, v, heads, kv_heads, out=None, *, softmax_scale=None):
    calls.append("attention_gqa")
    o = mo.gqa_attention(q.reshape(q.shape[0], heads, 128), k.reshape(k.shape[0], kv_heads, 128),
                         v.reshape(v.shape[0], kv_heads, 128))
    out.copy_(o.reshape(q.shape[0], -1))
    return out


def gate_norm_residual(x, gate, row_map, norm_w, norm_b, residual, out=None, *, eps=1e-6):
    calls.append("gate_norm_residual")
    assert row_map.dtype == torch.int32
    y = x.float() * gate.float()[row_map.long()]
    y = F.layer_norm(y, (x.shape[1],), norm_w, norm_b, eps) + residual.float()
    out.copy_(y.to(torch.bfloat16))
    return out


def silu_mul(x, out=None):
    calls.append("silu_mul")
    out.copy_(mo.silu_and_mul(x))
    return out


class FakeStore:
    """Stands in for ops.PagedKV: identity-mapped rows on plain CPU tensors."""

    def __init__(self, tokens, width):
        self.k = torch.full((tokens, width), float("nan"), dtype=torch.bfloat16)
        self.v = torch.full((tokens, width), float("nan"), dtype=torch.bfloat16)

    def map_rows(self, tokens):
        assert tokens <= self.k.shape[0]
        return self.k[:tokens], self.v[:tokens]


def install(monkeypatch, magi_layer):
    """Route magi_layer's kernel calls and cache allocation to the doubles above."""
    me = types.SimpleNamespace(**{n: globals()[n] for n in (
        "ln_modulate", "gemm", "magi_qkv_post", "head_layernorm", "attention_gqa", "gate_norm_residual", "silu_mul",
        "EPI_BIAS", "EPI_BIAS_GELU", "EPI_BIAS_GATE_RES", "EPI_BIAS_GELU_ERF", "EPI_BIAS_F32")})
    monkeypatch.setattr(magi_layer, "_ops", me)
    stores = {}

    def native_store(self, inference_params, dtype=torch.bfloat16):
        key = (id(inference_params), self.layer_number)
        if key not in stores:
            stores[key] = FakeStore(inference_params.max_sequence_length,
                                    self.num_query_groups_per_partition * self.hidden_size_per_attention_head)
        return stores[key]

    monkeypatch.setattr(magi_layer.MagiKVCacheManager, "native_store", native_store)
    return stores
