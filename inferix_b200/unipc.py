"""FlowUniPCMultistepScheduler — the many-step sampler of the reference's CFG pipeline
(inferix/models/wan_base/utils/fm_solvers_unipc.py:20-800, used by
pipeline/self_forcing/CausalDiffusionInferencePipeline.py:364-372 with solver_order 2, "bh2", flow prediction, x0
parameterisation, lower_order_final, final sigma 0).

Restated for exactly that configuration (anything else raises); same class / method names, same fp32 tensor arithmetic
in the same order, so the trajectory matches the reference scheduler to rounding (tests/golden/unipc.pt, generated from
the reference class by oracle/make_golden_unipc.py).  Host-side torch math on a handful of scalars plus a few
elementwise latent updates per step: negligible next to the two DiT forwards a step costs.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

__all__ = ["FlowUniPCMultistepScheduler"]


class FlowUniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, prediction_type: str = "flow_prediction",
                 shift: Optional[float] = 1.0, use_dynamic_shifting: bool = False, thresholding: bool = False,
                 predict_x0: bool = True, solver_type: str = "bh2", lower_order_final: bool = True,
                 disable_corrector: Optional[List[int]] = None, final_sigmas_type: str = "zero"):
        if (prediction_type != "flow_prediction" or use_dynamic_shifting or thresholding or not predict_x0
                or solver_type != "bh2" or final_sigmas_type != "zero" or solver_order not in (1, 2)):
            raise NotImplementedError("FlowUniPCMultistepScheduler: only the configuration the reference pipeline uses is "
                                      "built (flow prediction, x0, bh2, order <= 2, static shift, final sigma 0)")
        self.num_train_timesteps, self.solver_order, self.shift = num_train_timesteps, solver_order, shift
        self.lower_order_final = lower_order_final
        self.disable_corrector = list(disable_corrector or [])
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()        # :105-112
        sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.sigmas = sigmas.to("cpu")
        self.timesteps = sigmas * num_train_timesteps
        self.sigma_min, self.sigma_max = self.sigmas[-1].item(), self.sigmas[0].item()
        self.num_inference_steps = None
        self._reset()

    def _reset(self):
        self.model_outputs = [None] * self.solver_order
        self.timestep_list = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None
        self.this_order = 1

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps: int, device=None, sigmas=None, mu=None, shift: Optional[float] = None):
        """:160-228."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas).to("cpu")
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self._reset()

    @staticmethod
    def _alpha_sigma(sigma):
        return 1 - sigma, sigma                                                                   # :272-274

    def convert_model_output(self, model_output, sample):
        """flow -> x0 (:316-320)."""
        return sample - self.sigmas[self.step_index] * model_output

    def _coefficients(self, h, rks, order):
        """R, b, h_phi_1, B_h of :424-447 / :565-588 (x0 parameterisation: hh = -h; bh2: B_h = expm1(hh))."""
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        factorial_i = 1
        B_h = torch.expm1(hh)
        R, b = [], []
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        return torch.stack(R), torch.tensor(b, device=rks.device), h_phi_1, B_h

    def multistep_uni_p_bh_update(self, sample, order):
        """:350-484."""
        m0, x = self.model_outputs[-1], sample
        sigma_t, sigma_s0 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        alpha_t, sigma_t = self._alpha_sigma(sigma_t)
        alpha_s0, sigma_s0 = self._alpha_sigma(sigma_s0)
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        device = sample.device
        rks, D1s = [], []
        for i in range(1, order):
            si = self.step_index - i
            mi = self.model_outputs[-(i + 1)]
            alpha_si, sigma_si = self._alpha_sigma(self.sigmas[si])
            rk = (torch.log(alpha_si) - torch.log(sigma_si) - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks, device=device)
        _R, _b, h_phi_1, B_h = self._coefficients(h, rks, order)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        if D1s:
            rhos_p = torch.tensor([0.5], dtype=x.dtype, device=device)                            # order 2: simplified
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, torch.stack(D1s, dim=1))
        else:
            pred_res = 0
        return (x_t_ - alpha_t * B_h * pred_res).to(x.dtype)

    def multistep_uni_c_bh_update(self, this_model_output, last_sample, this_sample, order):
        """:486-626."""
        m0, x, model_t = self.model_outputs[-1], last_sample, this_model_output
        sigma_t, sigma_s0 = self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        alpha_t, sigma_t = self._alpha_sigma(sigma_t)
        alpha_s0, sigma_s0 = self._alpha_sigma(sigma_s0)
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        device = this_sample.device
        rks, D1s = [], []
        for i in range(1, order):
            si = self.step_index - (i + 1)
            mi = self.model_outputs[-(i + 1)]
            alpha_si, sigma_si = self._alpha_sigma(self.sigmas[si])
            rk = (torch.log(alpha_si) - torch.log(sigma_si) - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks, device=device)
        R, b, h_phi_1, B_h = self._coefficients(h, rks, order)
        if order == 1:
            rhos_c = torch.tensor([0.5], dtype=x.dtype, device=device)
        else:
            rhos_c = torch.linalg.solve(R, b).to(device).to(x.dtype)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        corr_res = torch.einsum("k,bkc...->bc...", rhos_c[:-1], torch.stack(D1s, dim=1)) if D1s else 0
        D1_t = model_t - m0
        return (x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * D1_t)).to(x.dtype)

    def index_for_timestep(self, timestep):
        indices = (self.timesteps == timestep).nonzero()
        return indices[1 if len(indices) > 1 else 0].item()

    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor,
             return_dict: bool = True, generator=None):
        """:655-739."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.step_index is None:
            if isinstance(timestep, torch.Tensor):
                timestep = timestep.to(self.timesteps.device)
            self._step_index = self.index_for_timestep(timestep)
        use_corrector = (self.step_index > 0 and self.step_index - 1 not in self.disable_corrector
                         and self.last_sample is not None)
        model_output_convert = self.convert_model_output(model_output, sample)
        if use_corrector:
            sample = self.multistep_uni_c_bh_update(model_output_convert, self.last_sample, sample, self.this_order)
        for i in range(self.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
            self.timestep_list[i] = self.timestep_list[i + 1]
        self.model_outputs[-1] = model_output_convert
        self.timestep_list[-1] = timestep
        this_order = (min(self.solver_order, len(self.timesteps) - self.step_index) if self.lower_order_final
                      else self.solver_order)
        self.this_order = min(this_order, self.lower_order_nums + 1)
        assert self.this_order > 0
        self.last_sample = sample
        prev_sample = self.multistep_uni_p_bh_update(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev_sample,)
        import types
        return types.SimpleNamespace(prev_sample=prev_sample)

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def __len__(self):
        return self.num_train_timesteps
