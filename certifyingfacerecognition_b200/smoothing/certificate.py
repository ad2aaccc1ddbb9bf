"""Certificates accepted by `Smooth` -- same class names, methods and return types as the reference's
smoothing/certificate.py (Certificate :6-47, L2Certificate :50-75)."""
from __future__ import annotations

import math

import torch
from scipy.stats import norm as _gauss

_UNIMPLEMENTED = "base class does not implement this method"


class Certificate:
    """Interface `Smooth.__init__` expects; every method raises until a subclass provides it."""

    def _missing(self, *_a, **_k):
        raise NotImplementedError(_UNIMPLEMENTED)

    compute_proxy_gap = _missing            # (logits) -> differentiable gap proxy
    sample_noise = _missing                 # (batch, repeated_theta) -> noise like `batch`
    compute_gap = _missing                  # (pABar) -> certified gap
    compute_radius_estimate = _missing      # (logits, theta) -> differentiable radius estimate


class L2Certificate(Certificate):
    """Gaussian smoothing, l2 certificate (Cohen et al.): gap = Phi^-1(pA_lower).

    `sample_noise` keeps the reference's torch semantics for callers that use it directly; the fused device path
    (`Smooth` on a `WrappedModel`) draws the same N(0, theta^2) noise inside the noise / projection kernel from a
    Philox stream instead, one counter per global sample index."""

    norm = "l2"
    _CLAMP = (0.001, 0.999)                 # certificate.py:62 clamps the probabilities before the inverse CDF

    def __init__(self, batch_size: int, device="cuda:0"):
        self.batch_size = batch_size
        self.device = device

    @staticmethod
    def _probit(p: torch.Tensor) -> torch.Tensor:
        return math.sqrt(2.0) * torch.erfinv(2.0 * p - 1.0)              # == Normal(0, 1).icdf(p)

    def compute_proxy_gap(self, logits: torch.Tensor) -> torch.Tensor:
        top, runner_up = (logits[:, k].clamp_(*self._CLAMP) for k in (0, 1))   # in place, like the reference
        return self._probit(top) - self._probit(runner_up)

    def sample_noise(self, batch: torch.Tensor, repeated_theta: torch.Tensor) -> torch.Tensor:
        return repeated_theta * torch.randn_like(batch, device=self.device)

    def compute_gap(self, pABar: float) -> float:
        return _gauss.ppf(pABar)

    def compute_radius_estimate(self, logits: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
        return 0.5 * theta * self.compute_proxy_gap(logits)
