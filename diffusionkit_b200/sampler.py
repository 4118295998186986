"""Sampler schedules — same classes and methods as the reference (python/src/diffusionkit/mlx/sampler.py:10-77).

Host-side scalar arithmetic (numpy fp32, like the reference's fp32 mx arrays); the per-step tensor math runs in the
fused CUDA sampler kernels (csrc/elementwise.cu: dk_sampler_prepare / dk_sampler_step).
"""
import numpy as np


class ModelSamplingDiscreteFlow:
    """Helper for sampler scheduling (timestep/sigma calculations) for Discrete Flow models (sampler.py:10-42)."""

    _t_first = 1

    def __init__(self, shift=1.0):
        self.shift = shift
        timesteps = 1000
        self.sigmas = self.sigma(np.arange(self._t_first, timesteps + 1, 1, dtype=np.float32))

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]

    def timestep(self, sigma):
        return sigma * np.float32(1000)

    def sigma(self, timestep):
        timestep = np.asarray(timestep, dtype=np.float32) / np.float32(1000.0)
        if self.shift == 1.0:
            return timestep
        s = np.float32(self.shift)
        return s * timestep / (np.float32(1) + (s - np.float32(1)) * timestep)

    def calculate_denoised(self, sigma, model_output, model_input):
        # tensor form lives in dk_sampler_step; this scalar/array form is kept for API parity (sampler.py:37-39)
        return model_input - model_output * sigma

    def noise_scaling(self, sigma, noise, latent_image, max_denoise=False):
        return sigma * noise + (1.0 - sigma) * latent_image


class FluxSampler(ModelSamplingDiscreteFlow):
    """Helper for sampler scheduling for Flux models (sampler.py:45-77): table over t = 0..1000."""

    _t_first = 0
