"""MAGI-1 scheduler index / timestep logic (inferix_b200.magi_schedule) against goldens produced by executing the
reference's own functions (oracle/make_golden_magi.py lifts them out of pipeline/magi/video_generate.py).
kv ranges and stage tuples: bit-exact integers; timestep tables: exact float32 equality."""
import json

import torch

from inferix_b200 import magi_schedule as ms


def load(golden_dir):
    return json.loads((golden_dir / "magi_schedule.json").read_text())


def test_generate_sequences(golden_dir):
    for c in load(golden_dir)["generate_sequences"]:
        assert [list(v) for v in ms.generate_sequences(*c["args"])] == c["ret"], c["args"]


def test_timestep_tables(golden_dir):
    g = load(golden_dir)
    for c in g["init_t"]:
        assert ms.init_t(dict(c["cfg"]), c["steps"], shortcut_mode=c["shortcut"]).tolist() == c["ret"], c
    for c in g["init_intervel"]:
        assert ms.init_intervel(c["steps"], shortcut_mode=c["shortcut"]).tolist() == c["ret"], c
    for c in g["timestep"]:
        t_total = ms.init_t(dict(tSchedulerFunc="sd3", shift=3.0), c["num_steps"])
        got = ms.get_timestep(t_total, c["dps"], c["start"], c["end"], c["idx"], c["has_clean"], clean_t=0.9999)
        assert got.tolist() == c["ret"], c


def test_status_and_total_steps(golden_dir):
    g = load(golden_dir)
    for c in g["status"]:
        a, b = ms.denoise_status_and_sequences(c["step"], c["num_steps"], c["chunk_num"], c["window"])
        assert [list(a), list(b)] == c["ret"], c
    for c in g["total_forward_step"]:
        assert ms.total_forward_step(c["num_steps"], c["chunk_num"], c["window"]) == c["ret"]


def test_kv_ranges_bit_exact(golden_dir):
    n = 0
    for c in load(golden_dir)["kvrange"]:
        if c["kind"] == "prefix":
            got = ms.kvrange_for_prefix_video(c["range_num"], c["ctn"], c["clean_kv"], c["n2c"])
        else:
            steps_each = ms.get_denoise_step_of_each_chunk(c["num_steps"], c["dps"], c["t_start"], c["t_end"], c["idx"],
                                                           c["has_clean"])
            assert steps_each == c["steps_each"]
            got = ms.kvrange_for_denoising_video(c["slice_point"], len(steps_each), c["ctn"], steps_each,
                                                 c["num_steps"], c["n2c"], c["clean_kv"])
        assert got.dtype == torch.int32 and got.tolist() == c["ret"], c
        n += 1
    assert n > 200


def test_integrate_euler_step():
    t_total = ms.init_t(dict(tSchedulerFunc="sd3", shift=3.0), 16)
    x = torch.randn(2, 4, 12, 3, 3)
    v = torch.randn_like(x)
    out = ms.integrate(x, v, t_total, 4, 0, 2, 1, chunk_width=6)
    dt = ms.get_timestep(t_total, 4, 0, 2, 2) - ms.get_timestep(t_total, 4, 0, 2, 1)
    assert torch.allclose(out[:, :, :6], x[:, :, :6] + v[:, :, :6] * dt[0])
    assert torch.allclose(out[:, :, 6:], x[:, :, 6:] + v[:, :, 6:] * dt[1])
