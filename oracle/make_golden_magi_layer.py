"""Goldens for the MAGI-1 transformer layer and the Ulysses context-parallel index logic, from the reference's own code.

ORACLE tooling (build container only; reads /root/reference, which does not exist on the GPU box).
`inferix.models.magi.dit.dit_module` imports here as it is.  Its `TransformerBlock` is built on CPU with
`parallel_state` answering "one rank" and the five CUDA-only third-party kernels the layer calls replaced by the torch
statement of their published algorithm (see the header of oracle/magi_oracle.py — the same functions the oracle
restates, flash_attn's rotary one being flash_attn's own `apply_rotary_emb_torch`).  Everything else — module
structure, parameter layout, rearranges, dtype flow, the KV-cache adapter on the reference's KVCacheManager — runs
unmodified.

Outputs: tests/golden/magi_layer_{gelu,glu}.pt  (inputs + outputs of a 4-forward sequence through 2 layers)
         tests/golden/magi_cp.json             (cp_ulysses split sizes / cross-attention ranges per rank)
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")

from oracle import magi_oracle as mo  # noqa: E402


def import_reference():
    from oracle.make_golden import install_shims
    install_shims()
    sys.path.insert(0, str(REF))
    import flashinfer
    import inferix.models.magi.dit.dit_module as dm
    from flash_attn.layers.rotary import apply_rotary_emb_torch
    from inferix.distributed import parallel_state as ps

    ps.get_tp_world_size = lambda with_context_parallel=False: 1
    ps.get_pp_world_size = lambda: 1
    ps.get_pp_rank = lambda: 0
    torch.cuda.get_device_capability = lambda *a: (8, 0)     # selects the flash_attn_func branch (:1000-1014)

    def flash_attn_func(q, k, v, deterministic=False):
        return torch.stack([mo.gqa_attention(q[b], k[b], v[b]) for b in range(q.shape[0])])

    def flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, deterministic=False):
        return mo.varlen_attention(q, k, v, cu_seqlens_q.tolist(), cu_seqlens_k.tolist())

    class Fp32Autocast:
        """What torch.autocast("cuda", dtype=torch.float32) does to `linear` on a GPU (it is a no-op on a CPU-only host):
        operands cast to fp32.  Used by dit_module.py:1291-1293 and dit_model.py:278,344."""

        def __init__(self, device_type=None, dtype=None, **kw):
            assert dtype == torch.float32
            self.orig = None

        def __enter__(self):
            self.orig = torch.nn.functional.linear
            orig = self.orig
            torch.nn.functional.linear = lambda x, w, b=None: orig(x.float(), w.float(), None if b is None else b.float())
            return self

        def __exit__(self, *exc):
            torch.nn.functional.linear = self.orig
            return False

    import inferix.models.magi.dit.dit_model as dmodel
    dm.torch = types.SimpleNamespace(**{k: getattr(torch, k) for k in dir(torch) if not k.startswith("__")})
    dm.torch.autocast = Fp32Autocast
    dmodel_torch = types.SimpleNamespace(**{k: getattr(torch, k) for k in dir(torch) if not k.startswith("__")})
    dmodel_torch.autocast = Fp32Autocast
    dmodel.torch = dmodel_torch
    dm.flash_attn_func = flash_attn_func
    dm.flash_attn_varlen_func = flash_attn_varlen_func
    dm.flash_apply_rotary_emb = lambda x, cos, sin: apply_rotary_emb_torch(x, cos, sin)
    dm.range_mod_triton = mo.range_mod
    flashinfer.activation.silu_and_mul = mo.silu_and_mul
    return dm


TINY = dict(hidden_size=256, ffn_hidden_size=512, num_attention_heads=4, num_query_groups=2, kv_channels=128,
            num_layers=2)


def make_meta(clip, ranges, slice_point, k_ranges, y_lens, extract=False, extra=False, distill=False):
    from inferix.core.types.inference import ModelMetaArgs, PackedCoreAttnParams, PackedCrossAttnParams
    cu_q = torch.tensor([0] + [clip] * ranges).cumsum(0).to(torch.int32)
    cu_k = torch.tensor([0] + list(y_lens)).cumsum(0).to(torch.int32)
    q_range = torch.stack([cu_q[:-1], cu_q[1:]], dim=1)
    k_range = torch.tensor(k_ranges, dtype=torch.int32)
    core = PackedCoreAttnParams(q_range=q_range, k_range=k_range, np_q_range=q_range.numpy(), np_k_range=k_range.numpy(),
                                max_seqlen_q=clip, max_seqlen_k=int(k_range.max() - k_range.min()))
    cross = PackedCrossAttnParams(q_ranges=q_range, kv_ranges=torch.stack([cu_k[:-1], cu_k[1:]], dim=1),
                                  cu_seqlens_q=cu_q, cu_seqlens_kv=cu_k, max_seqlen_q=clip, max_seqlen_kv=max(y_lens))
    return ModelMetaArgs(H=8, W=8, cp_pad_size=None, cp_split_sizes=None, slice_point=slice_point,
                         denoising_range_num=ranges, range_num=ranges + slice_point,
                         extract_prefix_video_feature=extract, fwd_extra_1st_chunk=extra,
                         distill_nearly_clean_chunk=distill, clip_token_nums=clip, enable_cuda_graph=False,
                         core_attn_params=core, cross_attn_params=cross)


def meta_to_plain(m):
    return dict(slice_point=m.slice_point, denoising_range_num=m.denoising_range_num, clip_token_nums=m.clip_token_nums,
                extract_prefix_video_feature=m.extract_prefix_video_feature, fwd_extra_1st_chunk=m.fwd_extra_1st_chunk,
                distill_nearly_clean_chunk=m.distill_nearly_clean_chunk,
                q_range=m.core_attn_params.np_q_range.tolist(), k_range=m.core_attn_params.np_k_range.tolist(),
                cu_seqlens_q=m.cross_attn_params.cu_seqlens_q.tolist(),
                cu_seqlens_kv=m.cross_attn_params.cu_seqlens_kv.tolist())


def layer_goldens(dm, name: str, gated: bool):
    from inferix.core.config import EngineConfig, ModelConfig
    from inferix.kvcache_manager.kvcache_manager import KVCacheManager, KVCacheRequest
    cfg = mo.MagiConfig(**TINY, gated_linear_unit=gated)
    mc = ModelConfig(model_name="tiny", params_dtype=torch.bfloat16, gated_linear_unit=gated, **TINY)
    ec = EngineConfig(cp_strategy="none", cp_size=1, fp8_quant=False, kv_offload=False)
    block = dm.TransformerBlock(mc, ec)
    sd = mo.synth_state_dict(cfg, seed=3)
    # dtypes as left by _high_precision_promoter (dit_model.py:620-637)
    for n, sub in block.named_modules():
        if "_xattn" in n:
            continue
        if any(t in n for t in ("q_layernorm", "k_layernorm", "self_attn_post_norm", "mlp_post_norm", "final_layernorm")):
            sub.float()
    missing = block.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k_, v_ in block.state_dict().items():
        assert v_.dtype == sd[k_].dtype, (k_, v_.dtype, sd[k_].dtype)
    block.eval()

    clip, max_seq = 96, 96 * 4
    ip = types.SimpleNamespace(max_sequence_length=max_seq, max_batch_size=1, sequence_len_offset=0,
                               kv_cache_request=KVCacheRequest("magi"), kv_cache_manager=KVCacheManager("cpu"),
                               key_value_memory_dict={}, update_kv_cache=False)
    ocache = mo.OracleMagiCache(max_seq)
    g = torch.Generator().manual_seed(5)
    hc = int(cfg.hidden_size * cfg.cond_hidden_ratio)
    # (ranges, slice_point, key ranges over cat(history, new), caption lengths, update_kv_cache, flags)
    plan = [
        (2, 0, [[0, 96], [0, 192]], [20, 13], True, dict(extract=True)),       # prefix: stores 2 clips
        (2, 2, [[96, 288], [0, 384]], [31, 7], False, {}),                       # reads history, stores nothing
        (2, 2, [[0, 288], [192, 384]], [9, 40], True, dict(distill=True)),       # stores all but the last clip
        (1, 3, [[96, 384]], [25], False, {}),                                    # history incl. the clip just stored
        (1, 0, [[0, 96]], [17], False, {}),                                      # no cache involvement (:186-187)
    ]
    steps = []
    for ranges, sp, kr, ylens, update, flags in plan:
        s = ranges * clip
        hidden = torch.randn(s, 1, cfg.hidden_size, generator=g).bfloat16()
        condition = torch.randn(1, ranges, hc, generator=g).bfloat16()
        cmap = torch.arange(ranges).repeat_interleave(clip).reshape(1, -1).transpose(0, 1).contiguous()
        y = torch.randn(sum(ylens), int(cfg.hidden_size * cfg.xattn_cond_hidden_ratio), generator=g).bfloat16()
        ang = torch.randn(s, 48, generator=g) * 2.0
        rope = torch.cat([ang.sin(), ang.cos()], dim=-1)                        # [s, 96] = sin | cos (:1097)
        meta = make_meta(clip, ranges, sp, kr, ylens, **flags)
        ip.update_kv_cache = ocache.update_kv_cache = update
        with torch.no_grad():
            out = block(hidden.clone(), condition, cmap, y, rope, ip, meta)
            mine = mo.block_forward(sd, cfg, hidden.clone(), condition, cmap, y, rope, ocache, meta)
        assert out.dtype == torch.float32 and torch.equal(out, mine), f"{name}: oracle != reference ({(out - mine).abs().max()})"
        steps.append(dict(meta=meta_to_plain(meta), update=update, hidden=hidden, condition=condition,
                          condition_map=cmap, y=y, rope=rope, out=out.clone()))
    caches = {}
    for i in range(cfg.num_layers):
        raw = ip.kv_cache_manager.get_raw(ip.kv_cache_request, f"layer_{i}")    # [2, tokens, 1, hn, d]
        caches[i] = raw[:, :3 * clip].clone()
        assert torch.equal(raw[:, :3 * clip, 0], ocache.mem[i][:, :3 * clip])
    path = ROOT / f"tests/golden/magi_layer_{name}.pt"
    torch.save(dict(cfg=dict(TINY, gated_linear_unit=gated), seed=3, clip=clip, max_seq=max_seq, steps=steps,
                    cache_prefix=caches), path)
    print("wrote", path, f"{path.stat().st_size / 1e3:.0f} kB; oracle == reference bit-for-bit on {len(steps)} forwards")


def cp_goldens():
    """cp_ulysses_process / cp_update_cross_attn_qkv_range with parallel_state answering (cp_size, rank)."""
    import inferix.distributed.parallelism.context_parallel as cpm
    from inferix.core.types.inference import PackedCrossAttnParams
    out = []
    for cp_size, clip, ranges, ylens in [(2, 96, 2, [20, 13]), (4, 50, 3, [5, 9, 2]), (8, 603, 4, [800, 1, 33, 64]),
                                         (3, 7, 5, [4, 4, 4, 4, 4]), (8, 6030, 1, [120])]:
        seq = clip * ranges
        cu_q = torch.tensor([0] + [clip] * ranges).cumsum(0).to(torch.int32)
        cu_k = torch.tensor([0] + ylens).cumsum(0).to(torch.int32)
        params = PackedCrossAttnParams(q_ranges=torch.stack([cu_q[:-1], cu_q[1:]], 1),
                                       kv_ranges=torch.stack([cu_k[:-1], cu_k[1:]], 1), cu_seqlens_q=cu_q,
                                       cu_seqlens_kv=cu_k, max_seqlen_q=clip, max_seqlen_kv=800)
        for rank in range(cp_size):
            cpm.mpu.get_cp_world_size = lambda n=cp_size: n
            cpm.mpu.get_cp_rank = lambda r=rank: r
            x = torch.arange(seq, dtype=torch.float32).reshape(seq, 1, 1)
            cmap = torch.arange(seq).reshape(seq, 1)
            rope = torch.arange(seq, dtype=torch.float32).reshape(seq, 1)
            xs, cm_, rp, split, cross = cpm.cp_ulysses_process(cp_size, x, cmap, rope, None, params)
            out.append(dict(cp_size=cp_size, rank=rank, clip=clip, ranges=ranges, ylens=ylens, split=split,
                            first_token=int(xs[0, 0, 0]), n_tokens=int(xs.shape[0]),
                            q_ranges=cross.q_ranges.tolist(), k_ranges=cross.kv_ranges.tolist(),
                            cu_q=cross.cu_seqlens_q.tolist(), cu_k=cross.cu_seqlens_kv.tolist(),
                            max_seqlen_q=int(cross.max_seqlen_q)))
            mine_split = mo.cp_split_sizes(seq, cp_size)
            mq, mk = mo.cp_cross_attn_ranges(cu_q.tolist(), cu_k.tolist(), mine_split, rank)
            assert mine_split == split and mq == cross.q_ranges.tolist() and mk == cross.kv_ranges.tolist()
    path = ROOT / "tests/golden/magi_cp.json"
    path.write_text(json.dumps(out))
    print("wrote", path, len(out), "cases; oracle == reference")


if __name__ == "__main__":
    dm = import_reference()
    layer_goldens(dm, "gelu", gated=False)
    layer_goldens(dm, "glu", gated=True)
    cp_goldens()
