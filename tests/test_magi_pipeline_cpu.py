"""inferix_b200.magi_pipeline.SampleTransport against the reference's own SampleTransport (lifted from
pipeline/magi/video_generate.py and run by oracle/make_golden_magi_walk.py) around the same closed-form stand-in model:
every model call (kwargs incl. slice_point / fwd_extra_1st_chunk / distill flags, timesteps, int32 kv ranges, input
checksums) and every yielded clean chunk must be identical.  Pure host logic, bit-exact."""
import types

import pytest
import torch

from oracle.make_golden_magi_walk import CASES, StandInModel, case_inputs


@pytest.mark.parametrize("name", sorted(CASES))
def test_walk_matches_reference(golden_dir, name):
    from inferix_b200.magi_pipeline import InferenceInput, SampleTransport
    gold = torch.load(golden_dir / "magi_walk.pt")[name]
    c = gold["case"]
    y, masks, prefix, noise, runtime, engine = case_inputs(c)
    model = StandInModel(runtime, engine)
    ti = InferenceInput(y=y, emb_masks=masks, prefix_video=prefix, latent_size=tuple(noise.shape),
                        t_schedule_config=dict(tSchedulerFunc="sd3", shift=3.0), num_steps=c["num_steps"],
                        chunk_num=c["chunk_num"])
    ip = types.SimpleNamespace(max_batch_size=1, max_sequence_length=0, update_kv_cache=False)
    st = SampleTransport(model, [ti], torch.device("cpu"), inference_params=ip, noise=noise.clone())
    chunks = [(idx, chunk.clone()) for _, idx, chunk in st.walk()]
    assert len(model.calls) == len(gold["calls"])
    for i, (mine, ref) in enumerate(zip(model.calls, gold["calls"])):
        assert mine["kwargs"] == ref["kwargs"], (i, mine["kwargs"], ref["kwargs"])
        assert mine["kv_range"] == ref["kv_range"], i
        assert torch.equal(mine["timestep"], ref["timestep"]), i
        assert mine["x_shape"] == ref["x_shape"] and mine["y_shape"] == ref["y_shape"], i
        assert mine["x_sum"] == ref["x_sum"] and mine["y_sum"] == ref["y_sum"] and mine["mask_sum"] == ref["mask_sum"], i
    assert [i for i, _ in chunks] == [i for i, _ in gold["chunks"]]
    for (_, a), (_, b) in zip(chunks, gold["chunks"]):
        assert torch.equal(a, b)
    assert torch.equal(st.xs[0], gold["final_x"])
    assert st.total_forward_step(0) == len([k for k in gold["calls"] if not k["kwargs"].get("extract_prefix_video_feature")])


def test_default_cache_is_sized_like_the_reference():
    """max_sequence_length = T * (H / patch) * (W / patch) of the whole video (video_generate.py:313-316)."""
    from inferix_b200 import magi_pipeline as mp_
    c = CASES["t2v"]
    y, masks, prefix, noise, runtime, engine = case_inputs(c)
    model = StandInModel(runtime, engine)
    ti = mp_.InferenceInput(y=y, emb_masks=masks, prefix_video=None, latent_size=tuple(noise.shape), num_steps=16,
                            chunk_num=5, t_schedule_config=dict(tSchedulerFunc="sd3", shift=3.0))
    made = {}

    class FakeIP:
        def __init__(self, max_batch_size, max_sequence_length, device=None):
            made.update(b=max_batch_size, n=max_sequence_length)
    orig, mp_.InferenceParams = mp_.InferenceParams, FakeIP
    try:
        mp_.SampleTransport(model, [ti], torch.device("cpu"), noise=noise)
    finally:
        mp_.InferenceParams = orig
    assert made == dict(b=1, n=15 * 4 * 4)


def test_end_to_end_job_matches_reference_stack(golden_dir, monkeypatch):
    """Native scheduler + native VideoDiTModel (kernel doubles) vs the reference's scheduler + model (3 chunks, window
    2, 4 steps, 3-way CFG: 24 model forwards incl. the batched unconditional passes and the extra clean-chunk forwards)."""
    import fake_magi_ops
    from inferix_b200 import magi_layer
    from magi_e2e_util import run_native
    fake_magi_ops.install(monkeypatch, magi_layer)
    g = torch.load(golden_dir / "magi_e2e.pt")
    ip = types.SimpleNamespace(max_batch_size=1, max_sequence_length=g["final_x"].shape[2] * (g["hw"] // 2) ** 2,
                               update_kv_cache=False)
    chunks, final_x = run_native(g, torch.device("cpu"), ip)
    assert [i for i, _ in chunks] == [i for i, _ in g["chunks"]]
    for (i, a), (_, b) in zip(chunks, g["chunks"]):
        err = ((a - b).norm() / b.norm()).item()
        print(f"chunk {i}: rel-L2 {err:.2e}")
        assert err <= 2e-2          # 3-way CFG (scales 1.5 / 7.5) amplifies the bf16 noise of the three passes ~10x
    assert ((final_x - g["final_x"]).norm() / g["final_x"].norm()).item() <= 2e-2
