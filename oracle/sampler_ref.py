"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's sampler, schedules, noise, CFG + Euler loop and latent formats:
python/src/diffusionkit/mlx/sampler.py (all) and mlx/__init__.py:253-292, 553-584, 674-788.

PARITY STATUS: the schedule / noise functions are closed-form and pinned against known-answer values derived from the
reference formulas (tests/golden/schedule_kats.json); the sampler classes and the whole
denoise_latents -> sample_euler -> CFGDenoiser loop are pinned against the reference's own source (mlx/sampler.py,
mlx/__init__.py) executed from /root/reference on the torch-backed MLX stand-in: tests/golden/
reference_mlxsrc_sampler.json and reference_mlxsrc_{flux,sd3}_pipeline.npz, tests/test_reference_mlxsrc_pin_cpu.py.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch


class ModelSamplingDiscreteFlowRef:
    """sampler.py:10-42 (SD3).  sigma table over t = 1..1000."""

    is_flux = False

    def __init__(self, shift: float = 1.0):
        self.shift = shift
        self.sigmas = self.sigma(torch.arange(1, 1001, dtype=torch.float32))

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]

    def timestep(self, sigma):
        return sigma * 1000

    def sigma(self, timestep):
        t = timestep / 1000.0
        if self.shift == 1.0:
            return t
        return self.shift * t / (1 + (self.shift - 1) * t)

    def calculate_denoised(self, sigma, model_output, model_input):
        return model_input - model_output * sigma

    def noise_scaling(self, sigma, noise, latent_image):
        return sigma * noise + (1.0 - sigma) * latent_image


class FluxSamplerRef(ModelSamplingDiscreteFlowRef):
    """sampler.py:45-77 (FLUX).  sigma table over t = 0..1000."""

    is_flux = True

    def __init__(self, shift: float = 1.0):
        self.shift = shift
        self.sigmas = self.sigma(torch.arange(0, 1001, dtype=torch.float32))


def get_sigmas(sampler, num_steps: int) -> torch.Tensor:
    """__init__.py:559-571"""
    start = float(sampler.timestep(sampler.sigma_max))
    end = float(sampler.timestep(sampler.sigma_min))
    if sampler.is_flux:
        num_steps += 1
    ts = torch.linspace(start, end, num_steps, dtype=torch.float32)
    sigs = [float(sampler.sigma(t)) for t in ts]
    if not sampler.is_flux:
        sigs += [0.0]
    return torch.tensor(sigs, dtype=torch.float32)


def get_noise(seed: int, h: int, w: int, c: int = 16) -> torch.Tensor:
    """__init__.py:553-557 — numpy global RNG, drawn in NCHW order, returned NHWC fp32"""
    np.random.seed(seed)
    noise = np.random.randn(1, c, h, w)
    return torch.from_numpy(noise).permute(0, 2, 3, 1).to(torch.float32).contiguous()


def get_empty_latent(h: int, w: int) -> torch.Tensor:
    """__init__.py:573-574"""
    return torch.ones((1, h, w, 16), dtype=torch.float32) * 0.0609


LATENT_FORMATS = {"sd3": (1.5305, 0.0609), "flux": (0.3611, 0.1159)}  # __init__.py:736-747


def process_out(latent, fmt: str):
    scale, shift = LATENT_FORMATS[fmt]
    return latent / scale + shift


def sample_euler(
    mmdit_call: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor],
    cache_modulation: Callable[[torch.Tensor, torch.Tensor], None],
    x: torch.Tensor,
    sigmas: torch.Tensor,
    conditioning: torch.Tensor,
    pooled: torch.Tensor,
    cfg_weight: float,
    act_dtype: Optional[torch.dtype],
) -> torch.Tensor:
    """__init__.py:761-788 with CFGDenoiser.__call__ (:691-719) inlined.

    x: (B, H, W, 16) fp32 (the reference has B = 1; a batch is B independent images sharing the schedule)
    conditioning: (Bc, T, 4096), pooled: (Bc, P) with Bc = B (cfg <= 0) or 2B ordered [positive(B) | negative(B)].
    """

    def cast(t):
        return t if act_dtype is None else t.to(act_dtype).to(torch.float32)

    timesteps = cast(sigmas * 1000.0)                                    # :769-771 (quirk Q5)
    cache_modulation(pooled, timesteps)                                  # :772
    for i in range(len(sigmas) - 1):
        sigma = float(sigmas[i])
        if cfg_weight <= 0:
            xin = cast(x)                                                # :700-702
        else:
            xin = cast(torch.cat([x, x], dim=0))                         # :704-706
        t = timesteps[i].reshape(1).repeat(xin.shape[0])                 # :710
        out = mmdit_call(xin, conditioning, t)
        den = xin - out * sigma                                          # calculate_denoised, sampler.py:37-39
        if cfg_weight > 0:
            den_text, den_neg = den.chunk(2, dim=0)                      # :718
            den = den_neg + cfg_weight * (den_text - den_neg)
        d = (x - den) / sigma                                            # to_d :756-758
        x = x + d * (float(sigmas[i + 1]) - sigma)                       # :779-781
    return x


def compute_psnr(reference: np.ndarray, proxy: np.ndarray) -> float:
    """diffusionkit/utils.py:70-82 (note: its "mse" is an RMSE)"""
    reference = np.asarray(reference, dtype=np.float64).flatten()
    proxy = np.asarray(proxy, dtype=np.float64).flatten()
    peak = np.abs(reference).max()
    rmse = np.sqrt(np.mean((reference - proxy) ** 2))
    return float(20 * np.log10((peak + 1e-5) / (rmse + 1e-10)))
