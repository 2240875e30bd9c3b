"""Causal Wan DiT (Self-Forcing / CausVid) on the native kernels, with the reference's module surface.

Mirrors ``inferix/models/self_forcing/causal_model.py`` (inference branch): same class names, constructor arguments,
parameter names (a reference checkpoint's state_dict loads as is) and forward signatures —
``CausalWanModel.forward`` (:866-880,1186-1194), ``CausalWanAttentionBlock.forward`` (:384-400).  The nn.Linear /
LayerNorm children are parameter containers only: a block forward is ONE call into the C ABI
(``ifx_wan_block_forward``: 13 hand-written sm_100a kernels), or the same kernels op by op when the block is sharded
over ranks (one NCCL all-gather of the new K/V in the middle).

Differences from the reference, all deliberate:
  * bf16 only (the production dtype, base_pipeline.py:351-352); other dtypes raise.
  * the KV window lives in a paged cache (frame-sized pages, block table) instead of a rolled tensor; end indices are
    host integers mirrored into ``kv_cache_meta`` with ``fill_`` (no ``.item()`` syncs, reference :280-300,328-329).
  * the text embedding MLP runs once per prompt, not once per forward (reference :948-953 recomputes it although only
    the first forward consumes it, via the cross-attention K/V cache).
  * training branches (_forward_train, flex-attention masks) are out of scope.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from ._lib import KvPlan, RopeGrid, WanBlockIO, WanBlockWeights
from .kvcache_manager import KVCacheManager, KVCacheRequest
from .kvcache_manager.model import SelfForcingKVCacheManagerFactory
from .parallel import ParallelConfig, all_gather_rows, all_gather_tokens, scatter_tokens


# ----------------------------------------------------------------------------- small components (wan_base/components.py)
def sinusoidal_embedding_1d(dim, position):
    """wan_base/components.py:11-31 (fp64)."""
    assert dim % 2 == 0
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_params(max_seq_len, dim, theta=10000):
    """wan_base/components.py:34-52 (complex128)."""
    assert dim % 2 == 0
    freqs = torch.outer(torch.arange(max_seq_len),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


class WanRMSNorm(nn.Module):
    """Parameter holder for components.py:107-126; the arithmetic is in ifx_rmsnorm / ifx_qk_norm_rope_append."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        shp = x.shape
        return ops.rmsnorm(x.reshape(-1, shp[-1]), self.weight, eps=self.eps).view(shp)


class WanLayerNorm(nn.LayerNorm):
    """components.py:129-142; forward runs ifx_ln_modulate (LN only)."""

    def __init__(self, dim, eps=1e-6, elementwise_affine=False):
        super().__init__(dim, elementwise_affine=elementwise_affine, eps=eps)

    def forward(self, x):
        shp = x.shape
        w, b = (self.weight, self.bias) if self.elementwise_affine else (None, None)
        return ops.ln_modulate(x.reshape(-1, shp[-1]).contiguous(), weight=w, bias=b, eps=self.eps).view(shp)


class _Attn(nn.Module):
    """q/k/v/o + norm_q/norm_k parameter layout shared by self- and cross-attention (causal_model.py:124-130)."""

    def __init__(self, dim, num_heads, eps):
        super().__init__()
        self.dim, self.num_heads, self.head_dim, self.eps = dim, num_heads, dim // num_heads, eps
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q = WanRMSNorm(dim, eps=eps)
        self.norm_k = WanRMSNorm(dim, eps=eps)


class CausalWanSelfAttention(_Attn):
    def __init__(self, dim, num_heads, local_attn_size=-1, sink_size=0, qk_norm=True, eps=1e-6,
                 parallel_config: Optional[ParallelConfig] = None):
        assert dim % num_heads == 0
        if not qk_norm:
            raise NotImplementedError("qk_norm=False is not built (every shipped Wan config uses qk_norm=True)")
        super().__init__(dim, num_heads, eps)
        self.local_attn_size, self.sink_size, self.qk_norm = local_attn_size, sink_size, qk_norm
        self.parallel_config = parallel_config
        self._qkv = None

    def forward(self, x, seq_lens, grid_sizes, freqs, block_mask, kv_cache_meta=None, current_start=0,
                cache_start=None):
        """The reference's stand-alone call (causal_model.py:147-334, KV-cached branch, single process): x [B, L, C]
        (already normed / modulated), kv_cache_meta = {"k", "v": [B, N, H, D] cache tensors, "global_end_index",
        "local_end_index": 1-element int64 tensors}, freqs = ops.rope_table(model.freqs) -> (y [B, L, C], k_view,
        v_view).  Same arithmetic on the native kernels: fused q|k|v GEMM, QK-RMSNorm + fp64 RoPE, the reference's own
        evict / roll / write on its tensor layout, tcgen05 attention over the cache prefix, output projection.
        The hot path does NOT go through here (the block calls ifx_wan_block_forward on the paged cache); this method
        exists so that code written against the reference's module surface keeps working."""
        if kv_cache_meta is None:
            raise NotImplementedError("only the KV-cached inference branch is built")
        pc = self.parallel_config
        if pc is not None and pc.world_size > 1:
            raise NotImplementedError("the stand-alone self-attention call is single-process; the sequence-parallel "
                                      "path is CausalWanAttentionBlock.forward")
        b, s, c = x.shape
        n, d = self.num_heads, self.head_dim
        if self._qkv is None or self._qkv[0].device != x.device:
            self._qkv = (torch.cat([self.q.weight, self.k.weight, self.v.weight]).detach().contiguous(),
                         torch.cat([self.q.bias, self.k.bias, self.v.bias]).detach().contiguous())
        f_, h_, w_ = (int(v) for v in grid_sizes[0])
        frame_seqlen = h_ * w_
        grid = RopeGrid(f_, h_, w_, current_start // frame_seqlen, 0, frame_seqlen)
        kc, vc = kv_cache_meta["k"], kv_cache_meta["v"]
        cache_size = kc.shape[1]
        outs = []
        for bi in range(b):
            qkv = ops.gemm(x[bi].contiguous(), self._qkv[0], self._qkv[1])
            q, k_new, v_new = ops.qk_norm_rope_append(qkv, self.norm_q.weight, self.norm_k.weight, freqs, grid, n, d,
                                                      eps=self.eps)
            # causal_model.py:277-304 on the reference's tensor layout
            current_end = current_start + s
            global_end = int(kv_cache_meta["global_end_index"].item())
            local_end = int(kv_cache_meta["local_end_index"].item())
            sink_tokens = self.sink_size * frame_seqlen
            if self.local_attn_size != -1 and current_end > global_end and s + local_end > cache_size:
                evicted = s + local_end - cache_size
                rolled = local_end - evicted - sink_tokens
                kc[bi, sink_tokens:sink_tokens + rolled] = kc[bi, sink_tokens + evicted:sink_tokens + evicted + rolled].clone()
                vc[bi, sink_tokens:sink_tokens + rolled] = vc[bi, sink_tokens + evicted:sink_tokens + evicted + rolled].clone()
                local_end_new = local_end + current_end - global_end - evicted
            else:
                local_end_new = local_end + current_end - global_end
            local_start = local_end_new - s
            kc[bi, local_start:local_end_new] = k_new.view(s, n, d)
            vc[bi, local_start:local_end_new] = v_new.view(s, n, d)
            o = ops.attention(q, kc[bi, :local_end_new].reshape(local_end_new, c), vc[bi, :local_end_new].reshape(local_end_new, c), n)
            outs.append(ops.gemm(o, self.o.weight, self.o.bias))
        kv_cache_meta["global_end_index"].fill_(current_end)
        kv_cache_meta["local_end_index"].fill_(local_end_new)
        return torch.stack(outs), kc[:, :local_end_new], vc[:, :local_end_new]


class WanT2VCrossAttention(_Attn):
    """wan_base/model.py:63-100."""

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6):
        super().__init__(dim, num_heads, eps)

    def text_kv(self, context: torch.Tensor):
        """K = RMSNorm(k(context)), V = v(context): computed once per prompt (wan_base/model.py:79-88)."""
        ctx = context.reshape(-1, context.shape[-1]).contiguous()
        k = ops.rmsnorm(ops.gemm(ctx, self.k.weight, self.k.bias), self.norm_k.weight, eps=self.eps)
        v = ops.gemm(ctx, self.v.weight, self.v.bias)
        return k, v


class CausalWanAttentionBlock(nn.Module):
    def __init__(self, cross_attn_type, dim, ffn_dim, num_heads, layer_idx, local_attn_size=-1, sink_size=0,
                 qk_norm=True, cross_attn_norm=False, eps=1e-6, enable_kv_offload=True,
                 parallel_config: Optional[ParallelConfig] = None):
        super().__init__()
        if cross_attn_type != "t2v_cross_attn":
            raise NotImplementedError("i2v cross-attention is outside the T2V hot path")
        if not cross_attn_norm:
            raise NotImplementedError("cross_attn_norm=False is not built (Wan T2V configs use True)")
        self.dim, self.ffn_dim, self.num_heads, self.layer_idx = dim, ffn_dim, num_heads, layer_idx
        self.local_attn_size, self.sink_size = local_attn_size, sink_size
        self.qk_norm, self.cross_attn_norm, self.eps = qk_norm, cross_attn_norm, eps
        self.enable_kv_offload = enable_kv_offload
        self.parallel_config = parallel_config
        self.kv_cache_manager = SelfForcingKVCacheManagerFactory.create_manager(
            layer_idx, num_heads, dim // num_heads, enable_kv_offload=enable_kv_offload)

        self.norm1 = WanLayerNorm(dim, eps)
        self.self_attn = CausalWanSelfAttention(dim, num_heads, local_attn_size, sink_size, qk_norm, eps,
                                                parallel_config=parallel_config)
        self.norm3 = WanLayerNorm(dim, eps, elementwise_affine=True)
        self.cross_attn = WanT2VCrossAttention(dim, num_heads, (-1, -1), qk_norm, eps)
        self.norm2 = WanLayerNorm(dim, eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)
        self._packed = None
        self._fp8 = None      # site -> quantised weight / scales once quantize_fp8() ran
        self._q8 = None       # site -> per-channel quantised weight once quantize_dynamic() ran
        self._amax = None     # dict while calibrating

    # ------------------------------------------------------------------ weight packing for the C ABI
    def _pack(self):
        sa, ca = self.self_attn, self.cross_attn
        ps = [p for p in self.parameters()]
        if any(p.dtype != torch.bfloat16 or not p.is_cuda for p in ps):
            raise ValueError("inferix_b200 blocks run in bfloat16 on CUDA: call model.to(torch.bfloat16).cuda() first")
        qkv_w = torch.cat([sa.q.weight, sa.k.weight, sa.v.weight], dim=0).contiguous()
        qkv_b = torch.cat([sa.q.bias, sa.k.bias, sa.v.bias], dim=0).contiguous()
        w = WanBlockWeights()
        w.dim, w.ffn_dim, w.heads, w.head_dim, w.eps = self.dim, self.ffn_dim, self.num_heads, self.dim // self.num_heads, self.eps
        keep = [qkv_w, qkv_b]

        def ptr(t):
            t = t.detach()
            if not t.is_contiguous():
                t = t.contiguous()
                keep.append(t)
            return t.data_ptr()

        w.qkv_w, w.qkv_b = qkv_w.data_ptr(), qkv_b.data_ptr()
        w.norm_q_w, w.norm_k_w = ptr(sa.norm_q.weight), ptr(sa.norm_k.weight)
        w.o_w, w.o_b = ptr(sa.o.weight), ptr(sa.o.bias)
        w.norm3_w, w.norm3_b = ptr(self.norm3.weight), ptr(self.norm3.bias)
        w.cq_w, w.cq_b, w.cnorm_q_w = ptr(ca.q.weight), ptr(ca.q.bias), ptr(ca.norm_q.weight)
        w.co_w, w.co_b = ptr(ca.o.weight), ptr(ca.o.bias)
        w.ffn1_w, w.ffn1_b = ptr(self.ffn[0].weight), ptr(self.ffn[0].bias)
        w.ffn2_w, w.ffn2_b = ptr(self.ffn[2].weight), ptr(self.ffn[2].bias)
        self._packed = (w, keep, qkv_w, qkv_b)
        return self._packed

    def invalidate_packed(self):
        self._packed = None
        self._fp8 = None
        self._q8 = None

    # ------------------------------------------------------------------ forward
    def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens, block_mask, kv_cache_meta=None,
                crossattn_cache_meta=None, current_start=0, cache_start=None,
                kv_cache_manager: Optional[KVCacheManager] = None,
                kv_cache_requests: Optional[List[KVCacheRequest]] = None, workspace=None, mod=None):
        """x [B, L, C]; e [B, F, 6, C]; freqs: float64 (cos, sin) table from ops.rope_table.  Updates x in place and
        returns it.  kv_cache_meta / crossattn_cache_meta are the reference's per-layer dicts (mutated in place).
        mod: optional precomputed `modulation + e` [B, F, 6, C] (the model adds all layers' tables in one launch)."""
        if kv_cache_meta is None:
            raise NotImplementedError("only the KV-cached inference branch is built")
        assert kv_cache_manager is not None and kv_cache_requests is not None
        w, _keep, qkv_w, qkv_b = self._packed or self._pack()
        b, rows, c = x.shape
        frames, fs = e.shape[1], rows // e.shape[1]
        pc = self.parallel_config
        world = pc.world_size if pc is not None else 1
        rank = pc.rank if pc is not None else 0
        f_, h_, w_ = (int(v) for v in grid_sizes[0])
        frame_seqlen = h_ * w_                                    # global tokens per frame (causal_model.py:255)
        if mod is None:
            mod = (self.modulation.unsqueeze(1) + e).contiguous() # [B, F, 6, C]   causal_model.py:412
        ws = workspace if workspace is not None else _Workspace(rows, c, self.ffn_dim, x.device)
        stream = torch.cuda.current_stream().cuda_stream
        lib = _lib.load()
        windowed = self.local_attn_size != -1
        sink_tokens = self.sink_size * frame_seqlen
        grid = RopeGrid(f_, h_, w_, current_start // frame_seqlen, rank * (frame_seqlen // world), frame_seqlen // world)

        for bi, req in enumerate(kv_cache_requests):
            store = self.kv_cache_manager.store(kv_cache_manager, req)
            cstore = self.kv_cache_manager.crossattn_store(kv_cache_manager, req)
            if crossattn_cache_meta is not None and not crossattn_cache_meta["is_init"]:
                k_txt, v_txt = self.cross_attn.text_kv(context[bi])
                cstore.import_(0, k_txt, v_txt)
            xb = x[bi]
            if not xb.is_contiguous():
                raise ValueError("x must be contiguous per sample")
            if store.page_tokens == 1 and frame_seqlen > 1:
                store.repage(frame_seqlen)          # reference-shaped allocation (no page size): pages = latent frames
            if store.offload is not None:           # offload tier: the window must sit in a device slot
                if world > 1:
                    raise NotImplementedError("the KV offload tier is single-GPU (peer-mapped caches cannot move)")
                store.stage()
                store.wait_staged()
            peer_dst = getattr(store, "peer", None) if world > 1 else None
            native_block = getattr(self, "_fp8", None) is None and self._amax is None and self._q8 is None
            if native_block and (world == 1 or (peer_dst is not None and sp_mode(world) in ("overlap", "store"))):
                io = WanBlockIO()
                io.x, io.rows, io.tokens_per_frame = xb.data_ptr(), rows, fs
                io.mod, io.freqs, io.grid = mod[bi].data_ptr(), freqs.data_ptr(), grid
                io.kv, io.current_start, io.sink_tokens, io.windowed = store.handle, current_start, sink_tokens, int(windowed)
                io.cross_k, io.cross_v, io.text_len = cstore.k.data_ptr(), cstore.v.data_ptr(), cstore.k.shape[0]
                io.ws_h, io.ws_qkv, io.ws_q = ws.h.data_ptr(), ws.qkv.data_ptr(), ws.q.data_ptr()
                io.ws_attn, io.ws_ffn = ws.attn.data_ptr(), ws.ffn.data_ptr()
                plan = KvPlan()
                if world == 1:
                    _lib.check(lib.ifx_wan_block_forward(C.byref(w), C.byref(io), C.byref(plan), stream))
                else:
                    # sequence parallel: ONE call per layer, the K/V exchange over peer memory inside it
                    from . import peer as _peer
                    peer_dst.epoch = store.peer_group.next_epoch()
                    mode = _lib.IFX_SP_OVERLAP if sp_mode(world) == "overlap" else _lib.IFX_SP_STORE
                    _lib.check(lib.ifx_wan_block_forward_sp(C.byref(w), C.byref(io), C.byref(peer_dst), mode,
                                                            _sp_push_ctas(world), _peer.WAIT_TIMEOUT_MS,
                                                            C.byref(plan), stream))
            else:
                plan = self._forward_ops(xb, mod[bi], fs, frames, grid, freqs, store, cstore, current_start,
                                         sink_tokens, windowed, ws, qkv_w, qkv_b, pc, amax=self._amax)
            # mirror of causal_model.py:328-329 (host ints -> device scalars, no sync).  When the pipeline keeps all
            # layers' indices in one tensor (kv_cache_meta["_ifx_shared"]), the model writes them once per forward
            # instead of two tiny launches per layer.
            if "_ifx_shared" not in kv_cache_meta:
                kv_cache_meta["global_end_index"].fill_(plan.global_end)
                kv_cache_meta["local_end_index"].fill_(plan.local_end)
            kv_cache_meta["_ifx_plan"] = (plan.local_start, plan.local_end, plan.global_end, plan.num_evicted)
            if store.offload is not None:
                store.write_back(list(plan.pages[: plan.num_pages]))
        if crossattn_cache_meta is not None:
            crossattn_cache_meta["is_init"] = True
        return x

    # ------------------------------------------------------------------ FP8 (e4m3, per-tensor static scales)
    FP8_SITES = ("qkv", "o", "cq", "co", "ffn1", "ffn2")

    def quantize_fp8(self, input_scales: dict):
        """Quantise the six block GEMMs' weights to e4m3 with one scale per launched weight matrix (q|k|v fused = one
        tensor) and fix the activation scales.  input_scales: site -> float (amax / 448 of that GEMM's input)."""
        _w, _keep, qkv_w, qkv_b = self._packed or self._pack()
        sa, ca = self.self_attn, self.cross_attn
        mats = {"qkv": (qkv_w, qkv_b), "o": (sa.o.weight, sa.o.bias), "cq": (ca.q.weight, ca.q.bias),
                "co": (ca.o.weight, ca.o.bias), "ffn1": (self.ffn[0].weight, self.ffn[0].bias),
                "ffn2": (self.ffn[2].weight, self.ffn[2].bias)}
        q = {}
        for site, (wt, bias) in mats.items():
            ws_ = float(wt.detach().abs().max().float()) / 448.0
            wq = torch.clamp(wt.detach().float() / ws_, -448.0, 448.0).bfloat16().to(torch.float8_e4m3fn).contiguous()
            q[site] = dict(w=wq, w_scale=ws_, in_scale=float(input_scales[site]), bias=bias.detach())
        self._fp8 = q

    def quantize_dynamic(self, kind: str = "fp8"):
        """Dynamic per-token-activation x per-channel-weight 8-bit linears for the six block GEMMs — the qconfig of the
        reference's quantisation examples (example/quantization/run_causvid_quantized.py:32-37).  kind: "fp8" (e4m3) or
        "int8".  Weights are quantised once per output channel; activation scales are computed per token inside the
        kernel that produces the activation and consumed by the GEMM epilogue (no calibration, no host round trip)."""
        k = {"fp8": ops.Q8_E4M3, "int8": ops.Q8_INT8}[kind]
        _w, _keep, qkv_w, qkv_b = self._packed or self._pack()
        sa, ca = self.self_attn, self.cross_attn
        mats = {"qkv": (qkv_w, qkv_b), "o": (sa.o.weight, sa.o.bias), "cq": (ca.q.weight, ca.q.bias),
                "co": (ca.o.weight, ca.o.bias), "ffn1": (self.ffn[0].weight, self.ffn[0].bias),
                "ffn2": (self.ffn[2].weight, self.ffn[2].bias)}
        q = {"kind": k, "kind_name": kind}
        for site, (wt, bias) in mats.items():
            codes, scales = ops.quantize_weight_per_channel(wt, k)
            q[site] = dict(w=codes, w_scale=scales, bias=bias.detach())
        self._fp8 = None
        self._q8 = q

    def q8_state(self):
        """Reference-named view of the dynamically quantised weights for the oracle: name -> (codes, scales, kind)."""
        c, q, p = self.dim, self._q8, f"blocks.{self.layer_idx}"
        out = {}
        for j, name in enumerate(("q", "k", "v")):
            out[f"{p}.self_attn.{name}"] = (q["qkv"]["w"][j * c:(j + 1) * c], q["qkv"]["w_scale"][j * c:(j + 1) * c],
                                            q["kind_name"])
        for site, name in (("o", "self_attn.o"), ("cq", "cross_attn.q"), ("co", "cross_attn.o"), ("ffn1", "ffn.0"),
                           ("ffn2", "ffn.2")):
            out[f"{p}.{name}"] = (q[site]["w"], q[site]["w_scale"], q["kind_name"])
        return out

    def fp8_state(self):
        """Reference-named view of the quantised weights for the oracle: name -> (weight_q, weight_scale, input_scale)."""
        c = self.dim
        q = self._fp8
        p = f"blocks.{self.layer_idx}"
        out = {}
        for j, name in enumerate(("q", "k", "v")):
            out[f"{p}.self_attn.{name}"] = (q["qkv"]["w"][j * c:(j + 1) * c], q["qkv"]["w_scale"], q["qkv"]["in_scale"])
        for site, name in (("o", "self_attn.o"), ("cq", "cross_attn.q"), ("co", "cross_attn.o"), ("ffn1", "ffn.0"),
                           ("ffn2", "ffn.2")):
            out[f"{p}.{name}"] = (q[site]["w"], q[site]["w_scale"], q[site]["in_scale"])
        return out

    def _forward_ops(self, x, mod, fs, frames, grid, freqs, store, cstore, current_start, sink_tokens, windowed,
                     ws, qkv_w, qkv_b, pc: Optional[ParallelConfig], amax: Optional[dict] = None):
        """The 13 kernels of the block, op by op.  Used (a) under sequence parallelism: one all-gather of this rank's
        new K/V between the QKV epilogue and the attention, x [S/P, C] holding this rank's hw slice of
        every frame; (b) for FP8 linears (quantisation fused into the LN kernels, dequantisation into the GEMM
        epilogues); (c) for calibration (`amax` collects the absolute maxima of the six GEMM inputs)."""
        rows, c = x.shape
        sa, ca = self.self_attn, self.cross_attn
        heads, hd = self.num_heads, self.dim // self.num_heads
        m = mod.view(frames, 6, c)
        world = pc.world_size if pc is not None else 1
        f8 = getattr(self, "_fp8", None) if amax is None else None
        q8 = self._q8 if (amax is None and f8 is None) else None

        def note(site, t):
            if amax is not None:
                amax[site] = max(amax.get(site, 0.0), float(t.abs().max().float()))

        def linear(site, a_bf16, a_fp8, w, b, out, **kw):
            if q8 is not None:
                # a_fp8 carries (codes, per-token scales) when the producer (LN) quantised on the way out
                codes, scales = a_fp8 if a_fp8 is not None else ops.quantize_rows(
                    a_bf16, q8["kind"], ws.q8(site, a_bf16.shape, q8["kind"]), ws.q8_scales(site, a_bf16.shape[0]))
                s = q8[site]
                return ops.gemm_q8(codes, s["w"], scales, s["w_scale"], q8["kind"], s["bias"], out, **kw)
            if f8 is None:
                note(site, a_bf16)
                return ops.gemm(a_bf16, w, b, out, **kw)
            q = f8[site]
            if a_fp8 is None:                    # producer had no fused quantisation: one extra elementwise pass
                a_fp8 = ops.quantize_fp8(a_bf16, q["in_scale"], ws.q8(site, a_bf16.shape))
            return ops.gemm_fp8(a_fp8, q["w"], q["in_scale"] * q["w_scale"], q["bias"], out, **kw)

        def ln(site, **kw):
            if q8 is not None:
                return None, ops.ln_modulate_quant(x, q8["kind"], ws.q8(site, x.shape, q8["kind"]),
                                                   ws.q8_scales(site, x.shape[0]), eps=self.eps, **kw)
            if f8 is None:
                ops.ln_modulate(x, ws.h, eps=self.eps, **kw)
                return ws.h, None
            return None, ops.ln_modulate_fp8(x, f8[site]["in_scale"], ws.q8(site, x.shape), eps=self.eps, **kw)

        plan = store.plan_append(current_start, rows * world, sink_tokens, windowed)
        h, h8 = ln("qkv", shift=m[:, 0], scale=m[:, 1], tokens_per_frame=fs)
        linear("qkv", h, h8, qkv_w, qkv_b, ws.qkv)
        peer_dst = getattr(store, "peer", None) if world > 1 else None
        if peer_dst is not None:
            # exchange fused into the producer: K / V rows go straight into every rank's cache over NVLink, the
            # attention is ordered after all ranks' epoch flags (inferix_b200/peer.py) — no collective in the layer
            pg = store.peer_group
            peer_dst.epoch = pg.next_epoch()
            ops.qk_norm_rope_append_peers(ws.qkv, sa.norm_q.weight, sa.norm_k.weight, freqs, grid, heads, hd,
                                          store, plan, peer_dst, q_out=ws.q, eps=self.eps)
            pg.wait(peer_dst.epoch)
            store.attention(ws.q, ws.attn)
        elif world > 1:
            ops.qk_norm_rope_append(ws.qkv, sa.norm_q.weight, sa.norm_k.weight, freqs, grid, heads, hd, q_out=ws.q,
                                    k_out=ws.kv_new[0], v_out=ws.kv_new[1], eps=self.eps)
            old_ext, new_ext = store.split_extents(plan)
            if _SP_OVERLAP and len(old_ext) <= 4 and len(new_ext) <= 4:
                # exchange on a side stream while the local queries attend the keys that are already in the cache
                main = torch.cuda.current_stream()
                side = ws.side_stream
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    kvg = all_gather_rows(ws.kv_new, pc, ws.kv_all)   # ONE all-gather: [P, 2, rows, C]
                    gathered = torch.cuda.Event()
                    gathered.record(side)
                # one piece per item unless the grid would not fill the SMs (8-way SP: 72 items on 148 SMs -> 2)
                items = heads * ((rows + 255) // 256)
                old_tiles = sum((n + 127) // 128 for _, n in old_ext)
                n_old = min(8, old_tiles, max(1, ws.sm_count // items)) if old_ext else 0
                part = ws.partials(rows, heads, n_old + 1)
                if old_ext:
                    ops.attention_partial(ws.q, store.k, store.v, old_ext, heads, part, n_old + 1, 0, n_old)
                main.wait_event(gathered)
                store.append_sp(plan, kvg[:, 0], kvg[:, 1], frames)
                ops.attention_partial(ws.q, store.k, store.v, new_ext, heads, part, n_old + 1, n_old, 1)
                ops.attention_combine(part, n_old + 1, ws.attn, heads)
            else:
                kvg = all_gather_rows(ws.kv_new, pc, ws.kv_all)
                store.append_sp(plan, kvg[:, 0], kvg[:, 1], frames)
                store.attention(ws.q, ws.attn)
        else:
            ops.qk_norm_rope_append(ws.qkv, sa.norm_q.weight, sa.norm_k.weight, freqs, grid, heads, hd, kv=store,
                                    plan=plan, q_out=ws.q, eps=self.eps)
            store.attention(ws.q, ws.attn)
        linear("o", ws.attn, None, sa.o.weight, sa.o.bias, x, epilogue=ops.EPI_BIAS_GATE_RES, residual=x,
               gate=m[:, 2], tokens_per_frame=fs)
        h, h8 = ln("cq", weight=self.norm3.weight, bias=self.norm3.bias)
        linear("cq", h, h8, ca.q.weight, ca.q.bias, ws.qkv[:, :c])
        ops.rmsnorm(ws.qkv[:, :c], ca.norm_q.weight, ws.q, eps=self.eps)
        ops.attention(ws.q, cstore.k, cstore.v, heads, ws.attn)
        linear("co", ws.attn, None, ca.o.weight, ca.o.bias, x, epilogue=ops.EPI_BIAS_GATE_RES, residual=x)
        h, h8 = ln("ffn1", shift=m[:, 3], scale=m[:, 4], tokens_per_frame=fs)
        linear("ffn1", h, h8, self.ffn[0].weight, self.ffn[0].bias, ws.ffn, epilogue=ops.EPI_BIAS_GELU)
        linear("ffn2", ws.ffn, None, self.ffn[2].weight, self.ffn[2].bias, x, epilogue=ops.EPI_BIAS_GATE_RES,
               residual=x, gate=m[:, 5], tokens_per_frame=fs)
        return plan


# IFX_SP_MODE selects how the sequence-parallel layer exchanges the block's new K / V over peer memory:
#   overlap (default) one C-ABI call per layer; the attention kernel itself ships the rows to the peers (an idle warp
#                     of its first CTAs) while it attends the cached window, and waits for the peers' flags in-kernel
#   store             one C-ABI call per layer; the norm+RoPE kernel stores into every rank's cache, then a wait kernel
#   ops               the round-1 op-by-op path below (also what FP8 / calibration / the NCCL fallback use)
# Measured on B200 boxes, back to back (profiles/r02*_sp*, r02m_*): 8 ranks 966 ms / block fused (32 copy CTAs) vs
# 1028 ms store + wait; 2 ranks 3242 fused vs 3249 store.  A third variant (a separate copy grid on a side stream next
# to the attention) reached only 106 GB/s on the SMs the attention leaves free and delayed its tail; it was dropped.
_SP_MODE = __import__("os").environ.get("IFX_SP_MODE", "overlap")
if _SP_MODE not in ("overlap", "store", "ops"):
    raise ValueError(f"IFX_SP_MODE={_SP_MODE!r}: expected overlap, store or ops")


def sp_mode(world: int) -> str:
    """The exchange variant a `world`-rank sequence-parallel group runs."""
    return _SP_MODE


def _sp_push_ctas(world: int) -> int:
    """Cap on the attention CTAs that share the fused K/V exchange (0 = the library default, 32)."""
    return max(0, int(__import__("os").environ.get("IFX_SP_PUSH_CTAS", "0")))


# IFX_SP_OVERLAP=1 turns on the exchange/compute overlap of the sequence-parallel path: local queries attend the pages
# already in the cache (phase 1) while the all-gather of the new K/V runs on a side stream, then the new pages
# (phase 2), merged by attn_combine_kernel.  Off by default: measured 4 % SLOWER at 2 GPUs (the gather is only
# ~50 us there and the second launch + partial traffic cost more); it targets 4-8 ranks where the gather is 10 % of
# the layer, which this round could not re-measure.
_SP_OVERLAP = __import__("os").environ.get("IFX_SP_OVERLAP", "0") == "1"


def _pieces_for(q_rows: int, heads: int, sms: int) -> int:
    """Key-range pieces per (head, 256-row pair) item so that items * pieces fills whole waves of SMs."""
    items = heads * ((q_rows + 255) // 256)
    best, best_cost = 1, float("inf")
    for s in range(1, 9):
        cost = -(-items * s // sms) / s
        if cost < best_cost - 1e-9:
            best, best_cost = s, cost
    return best


class _Workspace:
    """Scratch activations of one block forward, reused by every layer (all bf16)."""

    def __init__(self, rows, dim, ffn_dim, device, world=1):
        def buf(*shape):
            return torch.empty(shape, dtype=torch.bfloat16, device=device)
        self.rows = rows
        self.h, self.qkv, self.q, self.attn, self.ffn = buf(rows, dim), buf(rows, 3 * dim), buf(rows, dim), buf(rows, dim), buf(rows, ffn_dim)
        self._q8 = {}
        self._part = None
        if world > 1:
            self.side_stream = torch.cuda.Stream(device=device)
            self.sm_count = torch.cuda.get_device_properties(device).multi_processor_count
            self.kv_new = buf(2, rows, dim)              # this rank's new roped-K | V, one send buffer
            self.kv_all = buf(world, 2, rows, dim)       # all-gather destination

    def partials(self, q_rows, heads, pieces_per_item):
        """float32 scratch for two-phase attention partials (grown on demand)."""
        need = ops.attention_workspace_bytes(q_rows, heads, pieces_per_item) // 4
        if self._part is None or self._part.numel() < need:
            self._part = torch.empty(need, dtype=torch.float32, device=self.h.device)
        return self._part

    def q8(self, site, shape, kind=None):
        """8-bit staging buffer for a GEMM input (lazily allocated, keyed by shape and code type)."""
        dtype = torch.int8 if kind == ops.Q8_INT8 else torch.float8_e4m3fn
        key = (tuple(shape), dtype)
        t = self._q8.get(key)
        if t is None:
            t = self._q8[key] = torch.empty(tuple(shape), dtype=dtype, device=self.h.device)
        return t

    def q8_scales(self, site, rows):
        """per-token activation scales of one GEMM input (fp32 [rows]); one buffer per site: the scales of a producer
        must survive until its GEMM has run"""
        key = ("scales", site, rows)
        t = self._q8.get(key)
        if t is None:
            t = self._q8[key] = torch.empty(rows, dtype=torch.float32, device=self.h.device)
        return t


class CausalHead(nn.Module):
    """causal_model.py:487-515.  [S, C] -> [S, 64] on the native kernels: LayerNorm + per-frame modulation
    (ifx_ln_modulate) and the projection with its bias (ifx_gemm_bf16), two launches."""

    def __init__(self, dim, out_dim, patch_size, eps=1e-6):
        super().__init__()
        self.dim, self.out_dim, self.patch_size, self.eps = dim, out_dim, patch_size, eps
        self.norm = WanLayerNorm(dim, eps)
        self.head = nn.Linear(dim, math.prod(patch_size) * out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)

    def forward(self, x, e):
        """x [B, L1, C]; e [B, F, 1, C]."""
        num_frames, frame_seqlen = e.shape[1], x.shape[1] // e.shape[1]
        m = (self.modulation.unsqueeze(1) + e).contiguous()          # [B, F, 2, C]: shift, scale  (:510)
        if not (x.is_cuda and x.dtype == torch.bfloat16):
            raise ValueError("inferix_b200 head runs in bfloat16 on CUDA")
        outs = []
        for bi in range(x.shape[0]):
            h = ops.ln_modulate(x[bi].contiguous(), shift=m[bi, :, 0], scale=m[bi, :, 1], tokens_per_frame=frame_seqlen,
                                eps=self.eps)
            outs.append(ops.gemm(h, self.head.weight, self.head.bias))
        return torch.stack(outs).unflatten(dim=1, sizes=(num_frames, frame_seqlen))


class CausalWanModel(nn.Module):
    """causal_model.py:518-654 (constructor) and :866-1026 (_forward_inference)."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, local_attn_size=-1, sink_size=0,
                 qk_norm=True, cross_attn_norm=True, eps=1e-6, enable_kv_offload=True,
                 parallel_config: Optional[ParallelConfig] = None):
        super().__init__()
        if model_type != "t2v":
            raise NotImplementedError("only the t2v variant is on the hot path")
        assert (dim % num_heads) == 0 and (dim // num_heads) % 2 == 0
        if dim // num_heads != 128:
            raise NotImplementedError("the tcgen05 attention kernel is built for head_dim 128 (all Wan models)")
        self.model_type, self.patch_size, self.text_len = model_type, tuple(patch_size), text_len
        self.in_dim, self.dim, self.ffn_dim, self.freq_dim, self.text_dim = in_dim, dim, ffn_dim, freq_dim, text_dim
        self.out_dim, self.num_heads, self.num_layers = out_dim, num_heads, num_layers
        self.local_attn_size, self.sink_size = local_attn_size, sink_size
        self.qk_norm, self.cross_attn_norm, self.eps = qk_norm, cross_attn_norm, eps
        self.parallel_config = parallel_config if parallel_config is not None else ParallelConfig()

        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"), nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        self.blocks = nn.ModuleList([
            CausalWanAttentionBlock("t2v_cross_attn", dim, ffn_dim, num_heads, i, local_attn_size, sink_size, qk_norm,
                                    cross_attn_norm, eps, enable_kv_offload=enable_kv_offload,
                                    parallel_config=self.parallel_config) for i in range(num_layers)])
        self.head = CausalHead(dim, out_dim, self.patch_size, eps)

        d = dim // num_heads
        self.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                                rope_params(1024, 2 * (d // 6))], dim=1)          # complex128, causal_model.py:634-641
        self._freqs_table = None
        self._workspace = None
        self._mod_table = None        # [layers, 6 * C] stack of the blocks' modulation parameters
        self.block_mask = None
        self.num_frame_per_block = 1
        self.independent_first_frame = False

    def load_state_dict(self, state_dict, strict=True, assign=False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        for blk in self.blocks:
            blk.invalidate_packed()
        self._mod_table = None
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        for blk in self.blocks:
            blk.invalidate_packed()
        self._freqs_table = None
        self._workspace = None
        self._mod_table = None
        return out

    # ------------------------------------------------------------------ FP8
    def begin_fp8_calibration(self):
        """Subsequent forwards run op by op in bf16 and record the absolute maximum of every block-GEMM input."""
        for blk in self.blocks:
            blk._fp8, blk._amax = None, {}

    def finish_fp8_calibration(self, margin: float = 1.0):
        """Static per-tensor activation scales = margin * amax / 448, weights quantised per launched matrix.
        First and last layers could be kept in bf16 as MAGI does (dit_module.py:410); Wan quantises all blocks."""
        for blk in self.blocks:
            amax, blk._amax = blk._amax, None
            if not amax:
                raise RuntimeError("finish_fp8_calibration: no forward ran since begin_fp8_calibration")
            blk.quantize_fp8({k: max(v, 1e-6) * margin / 448.0 for k, v in amax.items()})

    def disable_fp8(self):
        for blk in self.blocks:
            blk._fp8 = blk._amax = blk._q8 = None

    def quantize_dynamic(self, kind: str = "fp8"):
        """reference: quantize_dynamic(transformer, {"": get_dynamic_fp8_per_token_act_per_channel_weight_qconfig(),
        "condition_embedder": None, "proj_out": None}) — every block linear, embeddings and head left in bf16
        (example/quantization/run_causvid_quantized.py:28-37).  kind "fp8" or "int8"."""
        for blk in self.blocks:
            blk.quantize_dynamic(kind)

    def _get_workspace(self, rows, device):
        ws = self._workspace
        if ws is None or ws.rows != rows or ws.h.device != device:
            ws = self._workspace = _Workspace(rows, self.dim, self.ffn_dim, device, self.parallel_config.world_size)
        return ws

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, kv_cache_meta: list = None,
                crossattn_cache_meta: list = None, current_start: int = 0, cache_start: int = 0,
                kv_cache_manager: Optional[KVCacheManager] = None, kv_cache_requests: Optional[list] = None,
                return_tokens: bool = False):
        """x: list of [C_in, F, H, W] (or a [B, C_in, F, H, W] tensor); t [B, F]; context: list of [L, text_dim]
        (or [B, L, text_dim]).  Returns the flow prediction [B, C_out, F, H, W]; with return_tokens (used by the
        wrapper's fused unpatchify + x0 epilogue) the head output in token order [B, F*hw, ph*pw*C_out] and grid_sizes."""
        if kv_cache_meta is None:
            raise NotImplementedError("training forward (_forward_train) is out of scope")
        pc = self.parallel_config
        device = self.patch_embedding.weight.device
        if self._freqs_table is None or self._freqs_table.device != device:
            self._freqs_table = ops.rope_table(self.freqs, device)

        # embeddings (causal_model.py:916-936)
        frames = x[0].shape[1] // self.patch_size[0]
        gh, gw = x[0].shape[2] // self.patch_size[1], x[0].shape[3] // self.patch_size[2]
        grid_sizes = torch.tensor([[frames, gh, gw]] * len(x), dtype=torch.long)
        assert frames * gh * gw <= seq_len
        fused = (len(x) * t.shape[1] <= 8 and all(u.is_cuda and u.dtype == torch.bfloat16 for u in x)
                 and self.patch_embedding.weight.dtype == torch.bfloat16)
        if fused:
            # native prologue: patch gather (this rank's hw slice only) + GEMM, fp64 sinusoid, three small linears; the
            # last one writes every layer's `modulation + e0` table (:412) directly
            chunk = (gh * gw) // pc.world_size
            pw_ = self.patch_embedding.weight.view(self.dim, -1)
            x = torch.stack([ops.gemm(ops.patchify(u, self.patch_size, pc.rank * chunk, chunk), pw_,
                                      self.patch_embedding.bias) for u in x])
            if self._mod_table is None:
                self._mod_table = torch.stack([blk.modulation.reshape(-1) for blk in self.blocks]).detach().contiguous()
            tp, te = self.time_projection[1], self.time_embedding
            sin = ops.sinusoidal_embedding(t.flatten().to(torch.float64), self.freq_dim)
            e = ops.linear_small(ops.linear_small(sin, te[0].weight, te[0].bias), te[2].weight, te[2].bias, silu_input=True)
            mods = ops.linear_small(e, tp.weight, tp.bias, silu_input=True, mod_table=self._mod_table)
            mods = mods.view(len(self.blocks), *t.shape, 6, self.dim)
            e0 = None
        else:
            xs = [self.patch_embedding(u.unsqueeze(0)) for u in x]
            xs = [u.flatten(2).transpose(1, 2) for u in xs]
            x = torch.cat(xs).contiguous()
            e = self.time_embedding(sinusoidal_embedding_1d(self.freq_dim, t.flatten()).type_as(x))
            e0 = self.time_projection(e).unflatten(1, (6, self.dim)).unflatten(dim=0, sizes=t.shape)
            x = scatter_tokens(x, frames, pc.world_size, pc.rank).contiguous()           # :939-942
            if self._mod_table is None:
                self._mod_table = torch.stack([blk.modulation.reshape(-1) for blk in self.blocks]).detach().contiguous()
            # every layer's `modulation + e0` (causal_model.py:412) in one launch: [layers, B, F, 6, C]
            mods = self._mod_table.view(len(self.blocks), 1, 1, 6, self.dim) + e0.unsqueeze(0)

        # text embedding only when some layer still has to build its cross-attention K/V (reference: every call)
        ctx = None
        if any(not m["is_init"] for m in crossattn_cache_meta):
            ctx = self.text_embedding(torch.stack(
                [torch.cat([u, u.new_zeros(self.text_len - u.size(0), u.size(1))]) for u in context]))

        ws = self._get_workspace(x.shape[1], device)
        e_blk = mods[0] if e0 is None else e0          # the blocks read only its shape when `mod` is given
        def stage_layer(li):                           # offload tier: copy a layer's window into a device slot
            for req in kv_cache_requests:
                st = self.blocks[li].kv_cache_manager.store(kv_cache_manager, req)
                if st.offload is not None:
                    st.stage()
        stage_layer(0)
        for i, block in enumerate(self.blocks):
            if i + 1 < len(self.blocks):                # ... the next layer's, while this one runs
                stage_layer(i + 1)
            x = block(x, e=e_blk, seq_lens=None, grid_sizes=grid_sizes, freqs=self._freqs_table, context=ctx,
                      context_lens=None, block_mask=None, kv_cache_meta=kv_cache_meta[i],
                      crossattn_cache_meta=crossattn_cache_meta[i], current_start=current_start,
                      cache_start=cache_start, kv_cache_manager=kv_cache_manager,
                      kv_cache_requests=kv_cache_requests, workspace=ws, mod=mods[i])
        shared = kv_cache_meta[0].get("_ifx_shared") if kv_cache_meta else None
        if shared is not None:
            # every layer appended the same block: one [layers, 2] = (global_end, local_end) write per forward
            _ls, local_end, global_end, _ev = kv_cache_meta[0]["_ifx_plan"]
            if all(m["_ifx_plan"][1:3] == (local_end, global_end) for m in kv_cache_meta):
                shared[:, 0].fill_(global_end)
                shared[:, 1].fill_(local_end)
            else:                                     # layers diverged (not produced by the shipped pipelines)
                for m in kv_cache_meta:
                    m["global_end_index"].fill_(m["_ifx_plan"][2])
                    m["local_end_index"].fill_(m["_ifx_plan"][1])

        x = self.head(x, e.unflatten(dim=0, sizes=t.shape).unsqueeze(2))             # [B, F, hw/P, 64]
        x = x.flatten(1, 2)
        x = all_gather_tokens(x, frames, pc)                                         # :1008-1022
        if return_tokens:
            return x, grid_sizes
        return torch.stack(self.unpatchify(x, grid_sizes))

    def unpatchify(self, x, grid_sizes):
        """causal_model.py:1196-1219."""
        c = self.out_dim
        out = []
        for u, v in zip(x, grid_sizes.tolist()):
            u = u[:math.prod(v)].view(*v, *self.patch_size, c)
            u = torch.einsum("fhwpqrc->cfphqwr", u)
            out.append(u.reshape(c, *[i * j for i, j in zip(v, self.patch_size)]))
        return out
