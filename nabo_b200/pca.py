"""PCA fit on the device: the exact counterpart of the reference's
``IncrementalPCA.partial_fit`` loop (nabo/_dataset.py:917-983) for references that are fitted where they live.

The scaled values of the kept cells (``get_scaled_values``, a chunk of cells at a time on the GPU) are reduced to
their first two moments in FP64 - column sums and the G x G Gram matrix - and the principal axes are the leading
eigenvectors of the covariance (cuSOLVER ``syevd`` through ``torch.linalg.eigh``: a plain library call for a one-off
G x G problem, G ~ 2 000).  Conventions are scikit-learn's, so the model drops into ``transform_pca`` unchanged:
``components_`` rows sorted by decreasing variance, signs by ``svd_flip(u_based_decision=False)`` (the entry of
largest magnitude of every row is positive), ``explained_variance_`` with the n - 1 divisor, no whitening.

Parity: the reference's incremental fit is itself an approximation of this decomposition (it truncates to
``n_comps`` after every batch), so the two agree at the level of the subspace, not bit for bit - with a single batch
(``batch_size >= n_cells``) IncrementalPCA is exact and the components agree to ~1e-10
(tests/test_gpu_facade.py::test_device_pca_fit_matches_sklearn).  ``Dataset.fit_ipca(..., method='sklearn')`` stays
the default for bit-compatibility with upstream.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

__all__ = ["DevicePCA", "MomentAccumulator"]


class MomentAccumulator:
    """Column sums and Gram matrix of row blocks, FP64 on the device."""

    def __init__(self, n_features: int, device):
        self.n = 0
        self.s = torch.zeros(n_features, dtype=torch.float64, device=device)
        self.g = torch.zeros((n_features, n_features), dtype=torch.float64, device=device)

    def add(self, z: torch.Tensor) -> None:
        self.n += z.shape[0]
        self.s += z.sum(0)
        self.g.addmm_(z.T, z)


class DevicePCA:
    """The slice of ``sklearn.decomposition.IncrementalPCA`` the path reads: ``components_``, ``mean_``,
    ``explained_variance_``, ``explained_variance_ratio_``, ``singular_values_``, ``var_``, ``n_samples_seen_``,
    ``n_components``, ``whiten`` (False) and ``transform``; ``genes`` is set by ``Dataset.fit_ipca``."""

    whiten = False

    def __init__(self, n_components: int):
        self.n_components = int(n_components)
        self.genes: Optional[List[str]] = None

    def fit_moments(self, acc: MomentAccumulator) -> "DevicePCA":
        n = acc.n
        if n < 2:
            raise ValueError("ERROR: at least two cells are needed to fit a PCA")
        mean = acc.s / n
        cov = (acc.g - torch.outer(acc.s, mean)) / (n - 1)
        cov = 0.5 * (cov + cov.T)
        ev, vec = torch.linalg.eigh(cov)                                  # ascending
        order = torch.arange(cov.shape[0] - 1, cov.shape[0] - 1 - self.n_components, -1, device=cov.device)
        comps = vec[:, order].T.contiguous()
        lam = ev[order].clamp_min(0.0)
        # svd_flip(u_based_decision=False): make the largest-magnitude entry of every component positive
        piv = comps.abs().argmax(1)
        sign = torch.sign(comps[torch.arange(comps.shape[0], device=comps.device), piv])
        sign[sign == 0] = 1.0
        comps *= sign[:, None]
        total = torch.diagonal(cov).sum()
        self.components_ = comps.cpu().numpy()
        self.mean_ = mean.cpu().numpy()
        self.var_ = (torch.diagonal(cov) * (n - 1) / n).cpu().numpy()      # population variance, as sklearn's var_
        self.explained_variance_ = lam.cpu().numpy()
        self.explained_variance_ratio_ = (lam / total).cpu().numpy()
        self.singular_values_ = torch.sqrt(lam * (n - 1)).cpu().numpy()
        self.n_samples_seen_ = int(n)
        self.n_components_ = self.n_components
        return self

    def transform(self, x) -> np.ndarray:
        """``X @ components_.T - mean_ @ components_.T`` (sklearn >= 1.x evaluation order)."""
        x = np.asarray(x, dtype=np.float64)
        return x @ self.components_.T - self.mean_ @ self.components_.T
