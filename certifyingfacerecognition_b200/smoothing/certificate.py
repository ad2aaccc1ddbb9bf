"""Drop-in for the reference's smoothing/certificate.py (Certificate ABC :6-47, L2Certificate :50-75)."""
from __future__ import annotations

import torch
from scipy.stats import norm
from torch.distributions.normal import Normal


class Certificate:
    """smoothing/certificate.py:6-47 -- abstract interface taken by Smooth.__init__."""

    def compute_proxy_gap(self, logits: torch.Tensor):
        raise NotImplementedError("base class does not implement this method")

    def sample_noise(self, batch: torch.Tensor, repeated_theta: torch.Tensor):
        raise NotImplementedError("base class does not implement this method")

    def compute_gap(self, pABar: float):
        raise NotImplementedError("base class does not implement this method")

    def compute_radius_estimate(self, logits: torch.Tensor, theta: torch.Tensor):
        raise NotImplementedError("base class does not implement this method")


class L2Certificate(Certificate):
    """smoothing/certificate.py:50-75.  ``sample_noise`` keeps the reference's torch semantics for callers that
    use it directly; the fused device path (Smooth with a WrappedModel) draws the same N(0, theta^2) noise inside
    the noise/projection kernel from a Philox stream instead (one counter per global sample index)."""
    norm = "l2"

    def __init__(self, batch_size: int, device: str = "cuda:0"):
        self.m = Normal(torch.zeros(batch_size).to(device), torch.ones(batch_size).to(device))
        self.device = device

    def compute_proxy_gap(self, logits: torch.Tensor) -> torch.Tensor:
        return self.m.icdf(logits[:, 0].clamp_(0.001, 0.999)) - self.m.icdf(logits[:, 1].clamp_(0.001, 0.999))

    def sample_noise(self, batch: torch.Tensor, repeated_theta: torch.Tensor) -> torch.Tensor:
        return torch.randn_like(batch, device=self.device) * repeated_theta

    def compute_gap(self, pABar: float) -> float:
        return norm.ppf(pABar)

    def compute_radius_estimate(self, logits: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
        return theta / 2 * self.compute_proxy_gap(logits)
