"""Host orchestration of the native MAGI-1 layer (inferix_b200/magi_layer.py, magi_cp.py) on CPU.

The CUDA kernels are replaced by the test doubles of tests/fake_magi_ops.py (oracle torch ops behind the real op
signatures); everything else is the product code: weight packing and the output-projection column permutation, the
zero-copy KV row placement with save/restore, range bookkeeping, and — over gloo, world_size 2 — the Ulysses
all-to-all layouts.  Results are compared with the goldens of the reference's own TransformerBlock.  (The kernels
themselves are checked on the GPU: tests/test_gpu_magi_layer.py.)"""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fake_magi_ops
from magi_golden_util import meta_from_plain
from oracle import magi_oracle as mo


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def build_block(cfg_dict, seed, cp_size=1):
    from inferix_b200 import magi_layer
    mc = types.SimpleNamespace(layernorm_epsilon=1e-6, apply_layernorm_1p=False, cond_hidden_ratio=0.25,
                               cond_gating_ratio=1.0, xattn_cond_hidden_ratio=1.0, params_dtype=torch.bfloat16,
                               **cfg_dict)
    ec = types.SimpleNamespace(cp_size=cp_size, cp_strategy="cp_ulysses" if cp_size > 1 else "none", fp8_quant=False,
                               kv_offload=False)
    block = magi_layer.TransformerBlock(mc, ec)
    sd = mo.synth_state_dict(mo.MagiConfig(**cfg_dict), seed=seed)
    res = block.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in block.state_dict().items():
        assert v.dtype == sd[k].dtype, k
    return block


@pytest.mark.parametrize("name", ["gelu", "glu"])
def test_block_orchestration_matches_reference(golden_dir, monkeypatch, name):
    from inferix_b200 import magi_layer
    stores = fake_magi_ops.install(monkeypatch, magi_layer)
    g = torch.load(golden_dir / f"magi_layer_{name}.pt")
    block = build_block(g["cfg"], g["seed"])
    ip = types.SimpleNamespace(max_sequence_length=g["max_seq"], max_batch_size=1, update_kv_cache=False)
    for i, st in enumerate(g["steps"]):
        ip.update_kv_cache = st["update"]
        out = block(st["hidden"].clone(), st["condition"], st["condition_map"], st["y"], st["rope"], ip,
                    meta_from_plain(st["meta"]))
        assert out.dtype == torch.float32 and out.shape == st["out"].shape
        err = rel_l2(out, st["out"]); print(f"forward {i}: rel-L2 {err:.2e}"); assert err <= 6e-3, f"forward {i}: {err}"
    # the rows the reference's cache holds after the sequence (forward 1 scribbles over rows of forward 0's clips only
    # transiently: the save/restore must have put them back)
    for layer, ref in g["cache_prefix"].items():
        store = stores[(id(ip), layer)]
        n = ref.shape[1]
        assert rel_l2(store.k[:n], ref[0, :, 0].reshape(n, -1)) <= 6e-3
        assert rel_l2(store.v[:n], ref[1, :, 0].reshape(n, -1)) <= 6e-3


def test_kernel_sequence_and_unsupported_configs(golden_dir, monkeypatch):
    from inferix_b200 import magi_layer
    fake_magi_ops.install(monkeypatch, magi_layer)
    g = torch.load(golden_dir / "magi_layer_gelu.pt")
    block = build_block(g["cfg"], g["seed"])
    st = g["steps"][4]                                   # no cache involvement
    fake_magi_ops.calls.clear()
    block.layers[0](st["hidden"].clone(), st["condition"], st["condition_map"], st["y"], st["rope"], None,
                    meta_from_plain(st["meta"]))
    assert fake_magi_ops.calls == ["ln_modulate", "gemm", "magi_qkv_post", "gemm", "head_layernorm", "attention_gqa",
                                   "attention_gqa", "gemm", "gate_norm_residual", "ln_modulate", "gemm", "gemm",
                                   "gate_norm_residual"]
    mc, ec = block.model_config, block.engine_config
    # fp8_quant: first / last layers stay bf16, the middle ones carry the reference's quantised linear types
    ec8 = types.SimpleNamespace(cp_size=1, cp_strategy="none", fp8_quant=True)
    assert isinstance(magi_layer.TransformerLayer(mc, ec8, 0).mlp.linear_fc1, torch.nn.Linear)
    if mc.num_layers > 2:
        mid = magi_layer.TransformerLayer(mc, ec8, 1)
        assert isinstance(mid.mlp.linear_fc1, magi_layer.PerTensorQuantizedFp8Linear)
        assert isinstance(mid.mlp.linear_fc2, magi_layer.PerChannelQuantizedFp8Linear)
        assert isinstance(mid.self_attention.linear_proj, magi_layer.PerChannelQuantizedFp8Linear)
        assert isinstance(mid.self_attention.linear_qkv.q, magi_layer.PerTensorQuantizedFp8Linear)
    with pytest.raises(NotImplementedError):
        magi_layer.TransformerLayer(mc, types.SimpleNamespace(cp_size=2, cp_strategy="cp_shuffle_overlap",
                                                              fp8_quant=False), 0)
    with pytest.raises(NotImplementedError):             # batch 2
        block.layers[0](st["hidden"].repeat(1, 2, 1), st["condition"], st["condition_map"], st["y"], st["rope"], None,
                        meta_from_plain(st["meta"]))


def test_real_ops_refuse_cpu_tensors(golden_dir):
    """No CPU fallback: without the doubles the layer raises on CPU tensors."""
    g = torch.load(golden_dir / "magi_layer_gelu.pt")
    block = build_block(g["cfg"], g["seed"])
    st = g["steps"][4]
    with pytest.raises(ValueError, match="CUDA"):
        block(st["hidden"].clone(), st["condition"], st["condition_map"], st["y"], st["rope"], None,
              meta_from_plain(st["meta"]))


# ----------------------------------------------------------------------------- Ulysses context parallel over gloo
def _cp_worker(rank, world, port, golden_path, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pytest as _pt
        from inferix_b200 import magi_cp, magi_layer
        mpatch = _pt.MonkeyPatch()
        fake_magi_ops.install(mpatch, magi_layer)
        magi_cp.init_context_parallel(None, world, rank)
        g = torch.load(golden_path)
        block = build_block(g["cfg"], g["seed"], cp_size=world)
        ip = types.SimpleNamespace(max_sequence_length=g["max_seq"], max_batch_size=1, update_kv_cache=False)
        worst = 0.0
        for st in g["steps"]:
            meta = meta_from_plain(st["meta"])
            x, cmap, rope, split, (xq, xk) = magi_cp.cp_ulysses_process(
                world, st["hidden"], st["condition_map"], st["rope"], st["meta"]["cu_seqlens_q"],
                st["meta"]["cu_seqlens_kv"])
            meta.cp_split_sizes = split
            meta.cross_attn_params = types.SimpleNamespace(q_ranges=xq, kv_ranges=xk)
            ip.update_kv_cache = st["update"]
            out = block(x.clone(), st["condition"], cmap, st["y"], rope, ip, meta)
            full = magi_cp.cp_post_process(world, "cp_ulysses", out, split)
            worst = max(worst, rel_l2(full, st["out"]))
        if rank == 0:
            result.put(worst)
        mpatch.undo()
    finally:
        dist.destroy_process_group()


def test_ulysses_cp2_matches_single_rank_reference(golden_dir):
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_cp_worker, args=(r, 2, port, str(golden_dir / "magi_layer_glu.pt"), result))
             for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert result.get(timeout=5) <= 6e-3


def test_cp_ranges_match_reference_goldens(golden_dir):
    import json
    from inferix_b200 import magi_cp
    for c in json.loads((golden_dir / "magi_cp.json").read_text()):
        split = magi_cp.ulysses_split_sizes(c["clip"] * c["ranges"], c["cp_size"])
        assert split == c["split"]
        cu_q = [i * c["clip"] for i in range(c["ranges"] + 1)]
        cu_k = [0]
        for n in c["ylens"]:
            cu_k.append(cu_k[-1] + n)
        q, k = magi_cp.cp_update_cross_attn_qkv_range(cu_q, cu_k, split, cp_rank=c["rank"])
        assert q == c["q_ranges"] and k == c["k_ranges"], c
