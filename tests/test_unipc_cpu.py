"""FlowUniPCMultistepScheduler restatement vs trajectories of the reference's own class (tests/golden/unipc.pt, generated
by oracle/make_golden_unipc.py from /root/reference/.../fm_solvers_unipc.py): timestep / sigma tables and every
intermediate sample, fp32 and bf16 latents, 50 / 8 / 20 steps — same torch arithmetic in the same order, so bit-exact."""
import pytest
import torch

from inferix_b200.unipc import FlowUniPCMultistepScheduler


def fake_flow(x, sigma):       # the synthetic "model" of the generator script
    return torch.tanh(x * 0.7) * (0.5 + sigma) + 0.1 * torch.sin(3.0 * x) - 0.3 * sigma


def test_unipc_matches_reference_trajectories(golden_dir):
    g = torch.load(golden_dir / "unipc.pt", weights_only=False)
    for case in g["cases"]:
        s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        s.set_timesteps(case["steps"], device="cpu", shift=case["shift"])
        assert torch.equal(s.timesteps, case["timesteps"]) and torch.equal(s.sigmas, case["sigmas"])
        x = case["x0"].clone()
        for i, t in enumerate(s.timesteps):
            x = s.step(fake_flow(x, float(t) / 1000.0), t, x, return_dict=False)[0]
            assert x.dtype == case["x0"].dtype
            assert torch.equal(x, case["xs"][i]), (case["steps"], i, (x.float() - case["xs"][i].float()).abs().max())


def test_unipc_refuses_other_configurations():
    with pytest.raises(NotImplementedError):
        FlowUniPCMultistepScheduler(solver_type="bh1")
    with pytest.raises(NotImplementedError):
        FlowUniPCMultistepScheduler(predict_x0=False)
    with pytest.raises(ValueError):
        FlowUniPCMultistepScheduler().step(torch.zeros(1), 0, torch.zeros(1))
