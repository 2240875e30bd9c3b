"""Generate tests/golden/unipc.pt from the reference's own FlowUniPCMultistepScheduler
(/root/reference/inferix/models/wan_base/utils/fm_solvers_unipc.py) — ORACLE tooling, build container only.

The class derives from diffusers' SchedulerMixin / ConfigMixin (not installed): both are stubbed with the minimum
they contribute here — `register_to_config` stores the constructor arguments in `self.config`.  The scheduler's
arithmetic is the reference's file, executed as is, on a synthetic "model": flow = a fixed smooth function of
(sample, sigma), so the trajectory exercises warm-up, the order-2 predictor / corrector and the lower-order final step.
"""
import functools
import importlib.util
import inspect
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/inferix/models/wan_base/utils/fm_solvers_unipc.py")


def register_to_config(init):
    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        self.config = types.SimpleNamespace(**cfg)
        self.register_to_config = lambda **kw: self.config.__dict__.update(kw)
        init(self, *args, **kwargs)
    return wrapper


def load_reference_class():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=type("ConfigMixin", (), {}), register_to_config=register_to_config)
    mod("diffusers.schedulers")
    mod("diffusers.schedulers.scheduling_utils", KarrasDiffusionSchedulers=[], SchedulerMixin=type("SchedulerMixin", (), {}),
        SchedulerOutput=lambda prev_sample: types.SimpleNamespace(prev_sample=prev_sample))
    mod("diffusers.utils", deprecate=lambda *a, **k: None, is_scipy_available=lambda: False)
    spec = importlib.util.spec_from_file_location("ref_unipc", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.FlowUniPCMultistepScheduler


def fake_flow(x, sigma):
    return torch.tanh(x * 0.7) * (0.5 + sigma) + 0.1 * torch.sin(3.0 * x) - 0.3 * sigma


def trajectory(cls, steps, shift, x0):
    s = cls(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    s.set_timesteps(steps, device="cpu", shift=shift)
    x, xs = x0.clone(), []
    for t in s.timesteps:
        sigma = float(t) / 1000.0
        x = s.step(fake_flow(x, sigma), t, x, return_dict=False)[0]
        xs.append(x.clone())
    return s.timesteps.clone(), s.sigmas.clone(), xs


def main():
    cls = load_reference_class()
    g = torch.Generator().manual_seed(0)
    cases = []
    for steps, shift, dtype in ((50, 5.0, torch.float32), (8, 3.0, torch.float32), (20, 8.0, torch.bfloat16)):
        x0 = torch.randn(1, 3, 4, 4, 4, generator=g).to(dtype)
        ts, sig, xs = trajectory(cls, steps, shift, x0)
        cases.append(dict(steps=steps, shift=shift, dtype=str(dtype), x0=x0, timesteps=ts, sigmas=sig, xs=xs))
    torch.save(dict(cases=cases, torch_version=torch.__version__), ROOT / "tests" / "golden" / "unipc.pt")
    print("wrote tests/golden/unipc.pt:", [(c["steps"], c["shift"], float(c["xs"][-1].float().norm())) for c in cases])


if __name__ == "__main__":
    main()
