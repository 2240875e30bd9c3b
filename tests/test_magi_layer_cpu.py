"""The MAGI-1 layer oracle (oracle/magi_oracle.py) against goldens produced by running the reference's own
TransformerBlock and cp_* functions on CPU (oracle/make_golden_magi_layer.py).  Bit-exact: the oracle performs the
same torch ops in the same dtypes."""
import json

import pytest
import torch

from oracle import magi_oracle as mo
from magi_golden_util import meta_from_plain


@pytest.mark.parametrize("name", ["gelu", "glu"])
def test_oracle_block_matches_reference(golden_dir, name):
    g = torch.load(golden_dir / f"magi_layer_{name}.pt")
    cfg = mo.MagiConfig(**g["cfg"])
    sd = mo.synth_state_dict(cfg, seed=g["seed"])
    cache = mo.OracleMagiCache(g["max_seq"])
    for i, st in enumerate(g["steps"]):
        cache.update_kv_cache = st["update"]
        out = mo.block_forward(sd, cfg, st["hidden"].clone(), st["condition"], st["condition_map"], st["y"],
                               st["rope"], cache, meta_from_plain(st["meta"]))
        assert out.dtype == torch.float32
        assert torch.equal(out, st["out"]), f"forward {i}: max |diff| {(out - st['out']).abs().max()}"
    # what the reference's KVCacheManager holds after the sequence (layout [2, tokens, 1, hn, d])
    for layer, ref in g["cache_prefix"].items():
        assert torch.equal(cache.mem[layer][:, :ref.shape[1]], ref[:, :, 0])


def test_rotary_is_partial_and_non_interleaved():
    """flash_attn apply_rotary_emb convention: dims [0, rd/2) pair with [rd/2, rd); dims >= rd pass through."""
    x = torch.randn(1, 5, 2, 128)
    ang = torch.randn(5, 48)
    y = mo.apply_rotary(x, ang.cos(), ang.sin())
    assert torch.equal(y[..., 96:], x[..., 96:])
    c, s = ang.cos()[:, None, :], ang.sin()[:, None, :]
    assert torch.allclose(y[0, :, :, :48], x[0, :, :, :48] * c - x[0, :, :, 48:96] * s, atol=1e-6)
    assert torch.allclose(y[0, :, :, 48:96], x[0, :, :, 48:96] * c + x[0, :, :, :48] * s, atol=1e-6)


def test_gqa_attention_head_grouping():
    q, k, v = torch.randn(7, 4, 128), torch.randn(9, 2, 128), torch.randn(9, 2, 128)
    out = mo.gqa_attention(q, k, v)
    for h in range(4):
        p = torch.softmax(q[:, h] @ k[:, h // 2].T / 128 ** 0.5, dim=-1)
        assert torch.allclose(out[:, h], p @ v[:, h // 2], atol=1e-5)


def test_cp_index_logic_matches_reference(golden_dir):
    cases = json.loads((golden_dir / "magi_cp.json").read_text())
    assert len(cases) >= 20
    for c in cases:
        split = mo.cp_split_sizes(c["clip"] * c["ranges"], c["cp_size"])
        assert split == c["split"]
        assert sum(split[:c["rank"]]) == c["first_token"] and split[c["rank"]] == c["n_tokens"]
        cu_q = [i * c["clip"] for i in range(c["ranges"] + 1)]
        cu_k = [0]
        for n in c["ylens"]:
            cu_k.append(cu_k[-1] + n)
        q, k = mo.cp_cross_attn_ranges(cu_q, cu_k, split, c["rank"])
        assert q == c["q_ranges"] and k == c["k_ranges"], c
