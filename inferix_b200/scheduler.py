"""FlowMatchScheduler (reference inferix/models/schedulers/flow_match.py:100-193), host-side sigma table + add_noise.

Negligible cost on the path; kept in PyTorch with the reference's arithmetic and rounding.
"""
from __future__ import annotations

import torch


class FlowMatchScheduler:
    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, shift=3.0, sigma_max=1.0,
                 sigma_min=0.003 / 1.002, inverse_timesteps=False, extra_one_step=False, reverse_sigmas=False):
        self.num_train_timesteps = num_train_timesteps
        self.shift, self.sigma_max, self.sigma_min = shift, sigma_max, sigma_min
        self.inverse_timesteps, self.extra_one_step, self.reverse_sigmas = inverse_timesteps, extra_one_step, reverse_sigmas
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False):
        sigma_start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            self.sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps + 1)[:-1]
        else:
            self.sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps)
        if self.inverse_timesteps:
            self.sigmas = torch.flip(self.sigmas, dims=[0])
        self.sigmas = self.shift * self.sigmas / (1 + (self.shift - 1) * self.sigmas)
        if self.reverse_sigmas:
            self.sigmas = 1 - self.sigmas
        self.timesteps = self.sigmas * self.num_train_timesteps
        if training:
            x = self.timesteps
            y = torch.exp(-2 * ((x - num_inference_steps / 2) / num_inference_steps) ** 2)
            y_shifted = y - y.min()
            self.linear_timesteps_weights = y_shifted * (num_inference_steps / y_shifted.sum())

    def _sigma_of(self, timestep, device):
        if timestep.ndim == 2:
            timestep = timestep.flatten(0, 1)
        self.sigmas = self.sigmas.to(device)
        self.timesteps = self.timesteps.to(device)
        idx = torch.argmin((self.timesteps.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)
        return idx, self.sigmas[idx].reshape(-1, 1, 1, 1)

    def step(self, model_output, timestep, sample, to_final=False):
        idx, sigma = self._sigma_of(timestep, model_output.device)
        if to_final or (idx + 1 >= len(self.timesteps)).any():
            sigma_ = 1 if (self.inverse_timesteps or self.reverse_sigmas) else 0
        else:
            sigma_ = self.sigmas[idx + 1].reshape(-1, 1, 1, 1)
        return sample + model_output * (sigma_ - sigma)

    def add_noise(self, original_samples, noise, timestep):
        """(1 - sigma) x0 + sigma noise in fp32-promoted arithmetic, cast to noise's dtype (reference :159-176)."""
        if (noise.is_cuda and noise.dtype == torch.bfloat16 and original_samples.dtype == torch.bfloat16
                and original_samples.shape == noise.shape and self.sigmas.dtype == torch.float32):
            from . import ops      # native kernel: sigma lookup + (1 - sigma) x0 + sigma noise in one launch
            if timestep.ndim == 2:
                timestep = timestep.flatten(0, 1)
            n = original_samples.shape[0]
            if timestep.numel() in (1, n):      # one sigma per leading-dim slice, or one for all (broadcast, :170-172)
                self.sigmas = self.sigmas.to(noise.device)
                self.timesteps = self.timesteps.to(noise.device)
                t64 = timestep.reshape(-1).to(torch.float64).expand(n).contiguous()
                return ops.add_noise(original_samples, noise, t64, self.timesteps.contiguous(), self.sigmas.contiguous())
        _, sigma = self._sigma_of(timestep, noise.device)
        return ((1 - sigma) * original_samples + sigma * noise).type_as(noise)

    def training_target(self, sample, noise, timestep):
        return noise - sample
