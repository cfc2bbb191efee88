"""Noise schedules for the MDLM sampler.

Interface parity with the reference's ``slm/utils/noise_utils.py``: a schedule is an ``nn.Module``
whose call returns ``(total_noise(t), rate_noise(t))`` (``Noise`` :99-119).  ``LogLinearNoise``
(:188-213) is the one ``configs/experiment/mdlm.yaml:35-36`` selects and the only one the ddpm
benchmark path uses; ``CosineNoise`` (:122-135) is what ``MaskedDiffusionLanguageModeling``
falls back to when no schedule is given (model.py:346-348); ``CosineSqrNoise``, ``Linear`` and
``GeometricNoise`` (:138-185) are carried for configs that name them (same two-method interface).

Schedules are host-side scalar maths (a few flops per diffusion step).  They are evaluated with
the same torch ops as the reference so per-step sigma and move chances agree bit for bit; none of
this is on the device hot path.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class Noise(nn.Module):
    """t in [0, 1]  ->  (sigma(t), d sigma / dt)."""

    def total_noise(self, t):
        raise NotImplementedError

    def rate_noise(self, t):
        raise NotImplementedError

    def forward(self, t):
        return self.total_noise(t), self.rate_noise(t)


class LogLinearNoise(Noise):
    """sigma(t) = -log(1 - (1 - eps) t)  =>  move chance 1 - exp(-sigma) = (1 - eps) t."""

    def __init__(self, eps: float = 1e-3):
        super().__init__()
        self.eps = eps
        self.sigma_max = self.total_noise(torch.tensor(1.0))
        self.sigma_min = self.eps + self.total_noise(torch.tensor(0.0))

    def total_noise(self, t):
        return -torch.log1p(-(1 - self.eps) * t)

    def rate_noise(self, t):
        keep = 1 - self.eps
        return keep / (1 - keep * t)

    def importance_sampling_transformation(self, t):
        """t -> the time whose move chance is log-linear between sigma_min and sigma_max (noise_utils.py:207-212;
        ``model_step`` with ``importance_sampling=True``, model.py:524-525)."""
        f_T = torch.log1p(-torch.exp(-self.sigma_max))
        f_0 = torch.log1p(-torch.exp(-self.sigma_min))
        sigma_t = -torch.log1p(-torch.exp(t * f_T + (1 - t) * f_0))
        return -torch.expm1(-sigma_t) / (1 - self.eps)


class CosineNoise(Noise):
    """sigma(t) = -log(eps + (1 - eps) cos(pi t / 2))."""

    def __init__(self, eps: float = 1e-3):
        super().__init__()
        self.eps = eps

    def total_noise(self, t):
        return -torch.log(self.eps + (1 - self.eps) * torch.cos(t * math.pi / 2))

    def rate_noise(self, t):
        half_pi = math.pi / 2
        keep = 1 - self.eps
        return half_pi * keep * torch.sin(t * half_pi) / (keep * torch.cos(t * half_pi) + self.eps)


class CosineSqrNoise(Noise):
    """sigma(t) = -log(eps + (1 - eps) cos^2(pi t / 2))  (noise_utils.py:138-152)."""

    def __init__(self, eps: float = 1e-3):
        super().__init__()
        self.eps = eps

    def total_noise(self, t):
        return -torch.log(self.eps + (1 - self.eps) * torch.cos(t * torch.pi / 2) ** 2)

    def rate_noise(self, t):
        keep = 1 - self.eps
        num = keep * torch.sin(t * torch.pi)
        den = keep * (torch.cos(t * torch.pi / 2) ** 2) + self.eps
        return (torch.pi / 2) * num / den


class Linear(Noise):
    """sigma(t) = sigma_min + t (sigma_max - sigma_min)  (noise_utils.py:155-172)."""

    def __init__(self, sigma_min=0, sigma_max=10, dtype=torch.float32):
        super().__init__()
        self.sigma_min = torch.tensor(sigma_min, dtype=dtype)
        self.sigma_max = torch.tensor(sigma_max, dtype=dtype)

    def total_noise(self, t):
        return self.sigma_min + t * (self.sigma_max - self.sigma_min)

    def rate_noise(self, t):
        return self.sigma_max - self.sigma_min

    def importance_sampling_transformation(self, t):
        f_T = torch.log1p(-torch.exp(-self.sigma_max))
        f_0 = torch.log1p(-torch.exp(-self.sigma_min))
        sigma_t = -torch.log1p(-torch.exp(t * f_T + (1 - t) * f_0))
        return (sigma_t - self.sigma_min) / (self.sigma_max - self.sigma_min)


class GeometricNoise(Noise):
    """sigma(t) = sigma_min^(1-t) sigma_max^t  (noise_utils.py:175-185)."""

    def __init__(self, sigma_min=1e-3, sigma_max=1):
        super().__init__()
        self.sigmas = 1.0 * torch.tensor([sigma_min, sigma_max])

    def total_noise(self, t):
        return self.sigmas[0] ** (1 - t) * self.sigmas[1] ** t

    def rate_noise(self, t):
        return self.total_noise(t) * (self.sigmas[1].log() - self.sigmas[0].log())


SCHEDULES = {"LogLinearNoise": LogLinearNoise, "CosineNoise": CosineNoise, "CosineSqrNoise": CosineSqrNoise,
             "Linear": Linear, "GeometricNoise": GeometricNoise}
