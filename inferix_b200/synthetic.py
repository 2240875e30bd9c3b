"""Deterministic synthetic weights / inputs with the reference's state-dict key names.

There are no checkpoints offline, so tests, goldens and the benchmark all use weights produced here.  Values
come from numpy's PCG64 bit generator seeded per tensor name (independent of creation order and of torch's RNG),
with the reference's initialisation scales (CausalWanModel.init_weights, causal_model.py:1221-1243): Xavier-uniform
linears, N(0, .02)-scale embeddings MLPs, modulation ~ 1/sqrt(dim).  Biases, norm weights and the output head get
small non-trivial values (the reference zero-initialises them) so every term of the block is exercised.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import numpy as np
import torch

__all__ = ["WAN_1_3B", "TINY", "wan_shapes", "synth_state_dict", "uniform_tensor"]

# wan_base/configs/wan_t2v_1_3B.py
WAN_1_3B = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, in_dim=16, out_dim=16, freq_dim=256,
                text_dim=4096, text_len=512)
# BASELINE.json configs[0]: 2 layers, head_dim 128 (needed for the production [22,21,21] RoPE split)
TINY = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, in_dim=16, out_dim=16, freq_dim=64, text_dim=64,
            text_len=512)


def uniform_tensor(name: str, shape: Tuple[int, ...], bound: float, seed: int = 0, center: float = 0.0) -> torch.Tensor:
    """U(center - bound, center + bound) float32 tensor, reproducible from (name, seed) alone."""
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
    n = int(np.prod(shape)) if len(shape) else 1
    a = rng.random(n, dtype=np.float32)
    a = (a * 2.0 - 1.0) * np.float32(bound) + np.float32(center)
    return torch.from_numpy(a.reshape(shape))


def wan_shapes(cfg: dict) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """name -> (shape, kind) for every parameter of CausalWanModel (causal_model.py:608-631, 365-382, 496-502)."""
    d, f = cfg["dim"], cfg["ffn_dim"]
    out: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def lin(name, o, i, kind="xavier"):
        out[name + ".weight"] = ((o, i), kind)
        out[name + ".bias"] = ((o,), "bias")

    out["patch_embedding.weight"] = ((d, cfg["in_dim"], 1, 2, 2), "xavier_conv")
    out["patch_embedding.bias"] = ((d,), "bias")
    lin("text_embedding.0", d, cfg["text_dim"], "normal02")
    lin("text_embedding.2", d, d, "normal02")
    lin("time_embedding.0", d, cfg["freq_dim"], "normal02")
    lin("time_embedding.2", d, d, "normal02")
    lin("time_projection.1", 6 * d, d)
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}"
        out[p + ".modulation"] = ((1, 6, d), "modulation")
        for attn in ("self_attn", "cross_attn"):
            for proj in ("q", "k", "v", "o"):
                lin(f"{p}.{attn}.{proj}", d, d)
            out[f"{p}.{attn}.norm_q.weight"] = ((d,), "norm")
            out[f"{p}.{attn}.norm_k.weight"] = ((d,), "norm")
        out[p + ".norm3.weight"] = ((d,), "norm")
        out[p + ".norm3.bias"] = ((d,), "bias")
        lin(p + ".ffn.0", f, d)
        lin(p + ".ffn.2", d, f)
    lin("head.head", cfg["out_dim"] * 4, d, "normal02")
    out["head.modulation"] = ((1, 2, d), "modulation")
    return out


def synth_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32, device="cpu") -> Dict[str, torch.Tensor]:
    sd = {}
    d = cfg["dim"]
    for name, (shape, kind) in wan_shapes(cfg).items():
        if kind == "xavier":
            t = uniform_tensor(name, shape, math.sqrt(6.0 / (shape[0] + shape[1])), seed)
        elif kind == "xavier_conv":
            fan_out, fan_in = shape[0], int(np.prod(shape[1:]))
            t = uniform_tensor(name, shape, math.sqrt(6.0 / (fan_in + fan_out)), seed)
        elif kind == "normal02":
            t = uniform_tensor(name, shape, 0.02 * math.sqrt(3.0), seed)       # same variance as N(0, .02)
        elif kind == "modulation":
            t = uniform_tensor(name, shape, math.sqrt(3.0 / d), seed)          # same variance as randn / sqrt(d)
        elif kind == "norm":
            t = uniform_tensor(name, shape, 0.1, seed, center=1.0)
        else:  # bias
            t = uniform_tensor(name, shape, 0.02, seed)
        sd[name] = t.to(dtype=dtype, device=device)
    return sd
