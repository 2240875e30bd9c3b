"""Pins the oracle (oracle/wan_oracle.py) to outputs of the real reference run on CPU (oracle/make_golden.py).

The reference ships no golden vectors for this path, so the fixtures under tests/golden/ are the
reference's own outputs on seeded inputs.  Bar: the restatement uses the same torch ops in the same order, so it
must reproduce them BIT-EXACTLY in fp32 and bf16 (single-threaded CPU), including the KV index trace.
"""
import pytest
import torch

from inferix_b200.synthetic import synth_state_dict
from oracle import wan_oracle as wo

CASES = ["sf_tiny_1block_fp32", "sf_tiny_1block_bf16", "sf_tiny_evict_fp32", "sf_tiny_evict_bf16",
         "sf_tiny_evict7_bf16"]


def run_oracle(gold):
    cfg_d = gold["cfg"]
    dtype = torch.float32 if "float32" in gold["dtype"] else torch.bfloat16
    cfg = wo.WanConfig(**cfg_d, local_attn_size=gold["local_attn_size"], sink_size=gold["sink_size"])
    sd = {k: v.to(dtype) for k, v in synth_state_dict(cfg_d, seed=0).items()}
    sched = wo.FlowMatchSigmas(shift=gold["shift"])
    steps = wo.warp_steps(sched, gold["steps"])
    fs = (gold["latent_hw"] // 2) ** 2
    cache_tokens = 32760 if gold["local_attn_size"] == -1 else gold["local_attn_size"] * fs
    regen = torch.Generator().manual_seed(gold["renoise_seed"])    # fp32 draws cast to the run dtype, as make_golden
    blocks = []
    out, caches = wo.pipeline_inference(
        sd, cfg, sched, gold["noise"], gold["context"], steps, 3, fs, cache_tokens,
        block_callback=lambda lat, i: blocks.append(i),
        noise_fn=lambda x: torch.randn(x.shape, generator=regen, dtype=torch.float32).to(x.dtype))
    return out, caches, blocks


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_bit_exact(case, golden_dir):
    torch.set_num_threads(1)
    gold = torch.load(golden_dir / f"{case}.pt", weights_only=False)
    out, caches, blocks = run_oracle(gold)
    assert out.dtype == gold["latents"].dtype
    assert torch.equal(out, gold["latents"]), f"max |diff| = {(out.float() - gold['latents'].float()).abs().max()}"
    assert blocks == gold["callback_blocks"]
    # KV index trace of layer 0: (global_end, local_end) after every forward
    trace = [(g, l) for (_, l, g, _) in caches[0].trace]
    assert trace == gold["index_trace"]
    # cache contents of the last layer (valid prefix)
    last = caches[-1]
    kv = torch.stack([last.k[0, :last.local_end], last.v[0, :last.local_end]]).unsqueeze(2)
    if gold["last_layer_cache"] is not None:
        assert torch.equal(kv, gold["last_layer_cache"])
    assert torch.equal(kv.double().abs().sum(dim=(1, 2, 3, 4)), gold["last_layer_cache_abs_sum"])


def test_index_kat_from_survey():
    """SURVEY §8c index KAT: 4 tokens/frame, 3-frame blocks, 2 noisy + 1 clean forward per block."""
    def run(cache_frames, sink, nblocks):
        g = l = 0
        res = []
        for b in range(nblocks):
            for _ in range(3):
                ls, le, ge, ev = wo.plan_indices(cache_frames * 4, g, l, b * 12, 12, sink * 4, True)
                g, l = ge, le
            res.append((g, l))
        return res
    assert run(6, 0, 4) == [(12, 12), (24, 24), (36, 24), (48, 24)]
    assert run(6, 1, 4) == [(12, 12), (24, 24), (36, 24), (48, 24)]
    assert run(9, 0, 5) == [(12, 12), (24, 24), (36, 36), (48, 36), (60, 36)]


def test_frame_provenance_kat_from_survey():
    """SURVEY §8c provenance KAT: which source frame sits in which cache slot after each block."""
    def run(cache_frames, sink, nblocks):
        c = wo.LayerCache(torch.full((1, cache_frames * 4, 1, 1), -1.0), torch.full((1, cache_frames * 4, 1, 1), -1.0))
        snaps = []
        for b in range(nblocks):
            tag = torch.arange(3 * b, 3 * b + 3).repeat_interleave(4).float().view(1, 12, 1, 1)
            for _ in range(2):
                wo.cache_append(c, tag, tag, b * 12, sink * 4, True)
            snaps.append([int(c.v[0, 4 * i, 0, 0]) for i in range(c.local_end // 4)])
        return snaps
    assert run(6, 0, 5) == [[0, 1, 2], [0, 1, 2, 3, 4, 5], [3, 4, 5, 6, 7, 8], [6, 7, 8, 9, 10, 11],
                            [9, 10, 11, 12, 13, 14]]
    assert run(6, 1, 5) == [[0, 1, 2], [0, 1, 2, 3, 4, 5], [0, 4, 5, 6, 7, 8], [0, 7, 8, 9, 10, 11],
                            [0, 10, 11, 12, 13, 14]]
    assert run(7, 1, 5) == [[0, 1, 2], [0, 1, 2, 3, 4, 5], [0, 3, 4, 5, 6, 7, 8], [0, 6, 7, 8, 9, 10, 11],
                            [0, 9, 10, 11, 12, 13, 14]]


@pytest.mark.parametrize("case", ["causvid_tiny_fp32", "causvid_tiny_bf16", "causvid_tiny_start_bf16"])
def test_causvid_oracle_matches_reference_bit_exact(case, golden_dir):
    torch.set_num_threads(1)
    gold = torch.load(golden_dir / f"{case}.pt", weights_only=False)
    dtype = torch.float32 if "float32" in gold["dtype"] else torch.bfloat16
    cfg = wo.WanConfig(**gold["cfg"])
    sd = {k: v.to(dtype) for k, v in synth_state_dict(gold["cfg"], seed=0).items()}
    sched = wo.FlowMatchSigmas(shift=gold["shift"])
    steps = torch.tensor(gold["steps"], dtype=torch.long)[:-1]          # causvid pipeline :37, no warping
    fs = (gold["latent_hw"] // 2) ** 2
    regen = torch.Generator().manual_seed(gold["renoise_seed"])
    out, _ = wo.causvid_pipeline_inference(
        sd, cfg, sched, gold["noise"], gold["context"], steps, 3, fs, 32760, start_latents=gold["start_latents"],
        noise_fn=lambda x: torch.randn(x.shape, generator=regen, dtype=torch.float32).to(x.dtype))
    assert torch.equal(out, gold["latents"]), f"max |diff| = {(out.float() - gold['latents'].float()).abs().max()}"
