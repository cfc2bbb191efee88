"""Noise schedules for the MDLM sampler.

Interface parity with the reference's ``slm/utils/noise_utils.py``: a schedule is an ``nn.Module``
whose call returns ``(total_noise(t), rate_noise(t))`` (``Noise`` :99-119).  ``LogLinearNoise``
(:188-213) is the one ``configs/experiment/mdlm.yaml:35-36`` selects and the only one the ddpm
benchmark path uses; ``CosineNoise`` (:122-135) is what ``MaskedDiffusionLanguageModeling``
falls back to when no schedule is given (model.py:346-348).

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


SCHEDULES = {"LogLinearNoise": LogLinearNoise, "CosineNoise": CosineNoise}
