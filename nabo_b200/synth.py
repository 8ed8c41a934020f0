"""Seeded synthetic inputs for tests and benchmarks.

No datasets ship with the reference (the tutorial data is downloaded from OSF),
so every parity case and every bench workload is drawn from these generators.
Everything is a pure function of its integer seed via ``numpy.random.Generator``
(PCG64), so the GPU box regenerates exactly what the golden fixtures were made
from; fixtures carry a SHA-256 of the inputs to prove it.
"""
from __future__ import annotations

import hashlib
from typing import List

import numpy as np

__all__ = ["nb_counts", "pc_mixture", "cell_names", "sha256_of", "size_factors"]


def nb_counts(n_cells: int, n_genes: int, seed: int, n_programs: int = 8,
              gene_seed: int = 0) -> np.ndarray:
    """Negative-binomial count matrix with latent-program structure.

    gene means mu_g ~ Gamma(0.3, 2); cell depth ~ LogNormal(0, 0.3);
    programs Z ~ Gamma(2, 0.5) (cells x P), W ~ Gamma(0.5, 1) (P x genes);
    X ~ Poisson(Gamma(shape=2, scale=lambda/2))  (dispersion 2).
    Gene-level parameters depend on ``gene_seed`` only, so a reference and its
    targets share one gene universe.
    """
    g = np.random.default_rng(gene_seed)
    mu = g.gamma(0.3, 2.0, size=n_genes)
    w = g.gamma(0.5, 1.0, size=(n_programs, n_genes))
    r = np.random.default_rng(seed)
    depth = r.lognormal(0.0, 0.3, size=n_cells)
    z = r.gamma(2.0, 0.5, size=(n_cells, n_programs))
    lam = depth[:, None] * mu[None, :] * (z @ w) / n_programs
    x = r.poisson(r.gamma(2.0, lam / 2.0 + 1e-12))
    return x.astype(np.int32)


def size_factors(counts: np.ndarray, size_scale: float = 1000.0) -> np.ndarray:
    """sf = size_scale / rowsum as float32 (nabo/_dataset.py:571-584)."""
    s = counts.sum(axis=1).astype(np.float64)
    s[s == 0] = 1
    return (size_scale / s).astype(np.float32)


def pc_mixture(n_cells: int, n_comps: int, seed: int, n_clusters: int = 32,
               centre_seed: int = 7, sigma_hi: float = 8.0, sigma_lo: float = 1.0,
               spread: float = 0.35) -> np.ndarray:
    """PCA-space coordinates: mixture of ``n_clusters`` centres drawn from
    N(0, diag(sigma_j^2)) with sigma_j geometric from sigma_hi to sigma_lo, plus
    within-cluster noise of ``spread`` * sigma_j.  float64, C-contiguous."""
    sig = sigma_hi * (sigma_lo / sigma_hi) ** (np.arange(n_comps) / max(n_comps - 1, 1))
    c = np.random.default_rng(centre_seed)
    centres = c.normal(size=(n_clusters, n_comps)) * sig[None, :]
    r = np.random.default_rng(seed)
    lab = r.integers(0, n_clusters, size=n_cells)
    x = centres[lab] + r.normal(size=(n_cells, n_comps)) * (spread * sig)[None, :]
    return np.ascontiguousarray(x, dtype=np.float64)


def cell_names(n: int, prefix: str) -> List[str]:
    """Zero-padded names, so that bytewise name order == row order."""
    w = max(4, len(str(n - 1)))
    return ["%s%0*d" % (prefix, w, i) for i in range(n)]


def sha256_of(*arrays: np.ndarray) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()
