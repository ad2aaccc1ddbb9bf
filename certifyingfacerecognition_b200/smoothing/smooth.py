"""Drop-in for the reference's smoothing/smooth.py (Smooth :11-160): same constructor, same
certify / predict / _sample_noise / _count_arr / _lower_confidence_bound signatures and return types.

When ``base_classifier`` is our WrappedModel and the certificate is an L2Certificate, ``_sample_noise`` runs
the whole loop body (noise -> latent -> StyleGAN -> resize -> ArcFace -> gallery argmin -> vote) as ONE C-ABI
call (``cfr_sample_votes``) and only the int64 vote counts come back.  Any other duck-typed base classifier
goes through the reference's generic loop (smooth.py:126-137) unchanged.
"""
from __future__ import annotations

from math import ceil
from typing import Optional

import numpy as np
import torch
from scipy.stats import beta as _beta
from scipy.stats import binomtest as _binomtest

from .certificate import Certificate, L2Certificate


def lower_confidence_bound(NA: int, N: int, alpha: float) -> float:
    """Clopper-Pearson lower bound == statsmodels ``proportion_confint(NA, N, alpha=2*alpha, method="beta")[0]``
    (smooth.py:148-160): Beta.ppf(alpha; NA, N-NA+1), and 0 when NA == 0."""
    if NA <= 0:
        return 0.0
    return float(_beta.ppf(alpha, NA, N - NA + 1))


class Smooth:
    """A smoothed classifier g (smooth.py:11-37)."""

    ABSTAIN = -1

    def __init__(self, base_classifier, num_classes: int, sigma: torch.Tensor, certificate: Certificate,
                 seed: int = 1234, process_group=None):
        self.base_classifier = base_classifier
        self.num_classes = num_classes
        self.sigma = sigma
        self.certificate = certificate
        # --- additions (keyword-only in practice; the reference's 4 positional args are unchanged) ---
        self.seed = seed                    # Philox key of the device noise stream
        self.process_group = process_group  # torch.distributed group: MC samples are sharded across its ranks
        self._draws = 0                     # global sample counter -> Philox subsequence offset
        self._replay = None                 # injected noise rows still to be consumed (parity runs), see inject_noise
        self.samples_classified = 0         # bookkeeping for throughput reports

    # ------------------------------------------------------------------------------------------ certify
    def certify(self, z: torch.Tensor, x: torch.Tensor, label: torch.Tensor, n0: int, n: int, alpha: float,
                batch_size: int, device: torch.device = torch.device("cuda:0")):
        """smooth.py:39-77 -> (predicted class | ABSTAIN, gap)."""
        self.base_classifier.eval()
        counts_selection = self._sample_noise(z, x, n0, batch_size, device=device)
        cAHat = counts_selection.argmax().item()
        if cAHat != label.item():
            return cAHat, 0.0
        counts_estimation = self._sample_noise(z, x, n, batch_size, device=device)
        nA = counts_estimation[cAHat].item()
        pABar = self._lower_confidence_bound(nA, n, alpha)
        if pABar < 0.5:
            return Smooth.ABSTAIN, 0.0
        return cAHat, self.certificate.compute_gap(pABar)

    def predict(self, z: torch.Tensor, x: torch.Tensor, n: int, alpha: float, batch_size: int,
                device: torch.device = torch.device("cuda:0")) -> int:
        """smooth.py:79-107.  ``scipy.stats.binom_test`` no longer exists (SciPy >= 1.12); ``binomtest(...).pvalue``
        is its two-sided replacement."""
        self.base_classifier.eval()
        counts = self._sample_noise(z, x, n, batch_size, device=device)
        top2 = counts.argsort()[::-1][:2]
        count1 = counts[top2[0]]
        count2 = counts[top2[1]]
        if _binomtest(int(count1), int(count1 + count2), p=0.5).pvalue > alpha:
            return Smooth.ABSTAIN
        return top2[0]

    def certify_many(self, z: torch.Tensor, x: torch.Tensor, labels: torch.Tensor, n0: int, n: int, alpha: float,
                     batch_size: int = 0, device: torch.device = torch.device("cuda:0")):
        """``certify`` (smooth.py:39-77) for G identities at once: z [G,512], x [1,5] or [G,5], labels [G] -> list of G
        (prediction | ABSTAIN, gap) tuples, each what ``certify`` returns for that identity.

        Not in the reference (its loop, certify.py:120-141, is strictly one identity at a time).  With the MC samples of
        an identity split over R ranks a selection pass is only N0 / R samples per rank (12-13 for N0 = 100, R = 8) while a
        program run costs a whole chunk; here the selection passes of all G identities share program runs and ONE
        all-reduce of the [G, num_classes] counts, then the estimation passes of the identities whose selection matched
        their label do the same.  Identity g owns the Philox block [D + g (n0 + n), D + (g + 1)(n0 + n)): its draws do not
        depend on which other identities take the early exit."""
        if not self._fused() or self._replay is not None:
            z = z.reshape(-1, 1, z.shape[-1])
            xs = x.reshape(-1, 1, x.shape[-1])
            return [self.certify(z[g], xs[g if xs.shape[0] > 1 else 0], labels.reshape(-1)[g:g + 1], n0, n, alpha,
                                 batch_size, device=device) for g in range(z.shape[0])]
        self.base_classifier.eval()
        z = z.reshape(-1, z.shape[-1])
        G = z.shape[0]
        labels = [int(v) for v in labels.reshape(-1).tolist()]
        rank, world = 0, 1
        if self.process_group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(self.process_group), dist.get_world_size(self.process_group)
        base = [self._draws + g * (n0 + n) for g in range(G)]
        self._draws += G * (n0 + n)
        xs = x.reshape(-1, x.shape[-1])

        def run(ids, num, extra):
            lo, hi = (num * rank) // world, (num * (rank + 1)) // world
            counts = self.base_classifier.sample_votes_multi(z[ids], xs if xs.shape[0] == 1 else xs[ids], self.sigma,
                                                             [hi - lo] * len(ids), seed=self.seed,
                                                             sample_offsets=[base[g] + extra + lo for g in ids])
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.process_group)
            self.samples_classified += (hi - lo) * len(ids)
            return counts.cpu().numpy()

        with torch.no_grad():
            c0 = run(list(range(G)), n0, 0)
            chat = c0.argmax(axis=1)
            out = [(int(chat[g]), 0.0) for g in range(G)]
            alive = [g for g in range(G) if int(chat[g]) == labels[g]]
            if alive:
                c1 = run(alive, n, n0)
                for row, g in zip(c1, alive):
                    pABar = self._lower_confidence_bound(int(row[chat[g]]), n, alpha)
                    out[g] = (Smooth.ABSTAIN, 0.0) if pABar < 0.5 else (int(chat[g]), self.certificate.compute_gap(pABar))
        return out

    def inject_noise(self, noise: Optional[torch.Tensor]) -> None:
        """Parity runs on identical noise tensors: the next ``_sample_noise`` calls consume the rows of ``noise``
        ([total, 5], already scaled by sigma -- what ``certificate.sample_noise`` returned in the run being replayed)
        in order instead of drawing from the device Philox stream.  ``None`` switches back."""
        self._replay = None if noise is None else noise.reshape(-1, noise.shape[-1]).clone()

    # ------------------------------------------------------------------------------------------ MC loop
    def _fused(self) -> bool:
        return getattr(self.base_classifier, "supports_fused_votes", False) and \
            isinstance(self.certificate, L2Certificate)

    def _sample_noise(self, z: torch.Tensor, x: torch.Tensor, num: int, batch_size,
                      device: torch.device = torch.device("cuda:0")) -> np.ndarray:
        """smooth.py:109-138 -> np.ndarray[num_classes] float64 (integer valued)."""
        with torch.no_grad():
            if self._fused():
                return self._sample_noise_fused(z, x, num)
            counts = torch.zeros(self.num_classes, dtype=float, device=device)
            for _ in range(ceil(num / batch_size)):
                this_batch_size = min(batch_size, num)
                num -= this_batch_size
                batch = x.repeat((this_batch_size, 1, 1, 1))
                noise = self.certificate.sample_noise(batch, self.sigma)
                predictions = self.base_classifier(z, batch + noise).argmax(1)
                counts += self._count_arr(predictions, device, self.num_classes)
                self.samples_classified += this_batch_size
            return counts.cpu().numpy()

    def _sample_noise_fused(self, z: torch.Tensor, x: torch.Tensor, num: int) -> np.ndarray:
        """One cfr_sample_votes call per rank; MC samples [0,num) are split contiguously over the ranks of
        ``process_group`` (Philox counter = global sample index, so the union does not depend on the rank
        count) and the int64 per-identity counts are summed with one all-reduce."""
        rank, world = 0, 1
        if self.process_group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(self.process_group), dist.get_world_size(self.process_group)
        lo = (num * rank) // world
        hi = (num * (rank + 1)) // world
        if self._replay is not None:
            if self._replay.shape[0] < num:
                raise ValueError(f"inject_noise: {self._replay.shape[0]} rows left, {num} needed")
            rows, self._replay = self._replay[:num], self._replay[num:]
            counts = self.base_classifier.sample_votes(z, x, self.sigma, hi - lo, noise=rows[lo:hi])
        else:
            counts = self.base_classifier.sample_votes(z, x, self.sigma, hi - lo, seed=self.seed,
                                                       sample_offset=self._draws + lo)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.process_group)
        self._draws += num
        self.samples_classified += hi - lo
        return counts.cpu().numpy().astype(np.float64)

    def _count_arr(self, arr: torch.Tensor, device: torch.device, length: int) -> torch.Tensor:
        """smooth.py:140-146."""
        counts = torch.zeros(length, dtype=torch.long, device=device)
        unique, c = arr.unique(sorted=False, return_counts=True)
        counts[unique] = c
        return counts

    def _lower_confidence_bound(self, NA: int, N: int, alpha: float) -> float:
        """smooth.py:148-160."""
        return lower_confidence_bound(int(NA), int(N), alpha)
