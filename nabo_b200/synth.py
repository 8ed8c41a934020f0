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

__all__ = ["nb_counts", "pc_mixture", "pc_mixture_device", "nb_counts_device", "cell_names", "sha256_of", "size_factors"]


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


def pc_mixture_device(n_cells: int, n_comps: int, seed: int, device, n_clusters: int = 32, centre_seed: int = 7,
                      sigma_hi: float = 8.0, sigma_lo: float = 1.0, spread: float = 0.35):
    """``pc_mixture`` drawn on the GPU (same centres and scales; labels and noise from torch's Philox generator,
    a pure function of ``seed`` on a given torch / GPU generation) for the shapes that are too big to draw on the
    host and copy: 10 M x 50 float64 is 4 GB.  Returns a CUDA float64 (n_cells, n_comps) tensor."""
    import torch
    sig = sigma_hi * (sigma_lo / sigma_hi) ** (np.arange(n_comps) / max(n_comps - 1, 1))
    centres = np.random.default_rng(centre_seed).normal(size=(n_clusters, n_comps)) * sig[None, :]
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    lab = torch.randint(0, n_clusters, (n_cells,), generator=gen, device=device)
    x = torch.randn((n_cells, n_comps), generator=gen, device=device, dtype=torch.float64)
    x *= torch.from_numpy(spread * sig).to(device)
    x += torch.from_numpy(centres).to(device)[lab]
    return x.contiguous()


def nb_counts_device(n_cells: int, n_genes: int, seed: int, device, n_programs: int = 8, gene_seed: int = 0,
                     dtype=None):
    """``nb_counts`` drawn on the GPU (same gene-level parameters; cell-level draws from torch's generator):
    Gamma-Poisson counts with latent-program structure, returned as a CUDA float32 (cells, genes) tensor -
    the dense count block ``nabo_project_dense`` takes.  BASELINE config 5 draws 500 000 x 2 000 per sample."""
    import torch
    g = np.random.default_rng(gene_seed)
    mu = torch.from_numpy(g.gamma(0.3, 2.0, size=n_genes)).to(device)
    w = torch.from_numpy(g.gamma(0.5, 1.0, size=(n_programs, n_genes))).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    depth = torch.exp(0.3 * torch.randn(n_cells, generator=gen, device=device, dtype=torch.float64))
    # Gamma(2, 0.5) = sum of two exponentials of scale 0.5
    u = torch.rand((2, n_cells, n_programs), generator=gen, device=device, dtype=torch.float64).clamp_min(1e-300)
    z = -0.5 * (torch.log(u[0]) + torch.log(u[1]))
    lam = (depth[:, None] * mu[None, :] * (z @ w) / n_programs).to(torch.float32)
    u = torch.rand((2,) + tuple(lam.shape), generator=gen, device=device, dtype=torch.float32).clamp_min(1e-30)
    rate = -(lam / 2.0 + 1e-12) * (torch.log(u[0]) + torch.log(u[1]))        # Gamma(2, lam / 2)
    del u, lam
    x = torch.poisson(rate, generator=gen)
    return x if dtype is None else x.to(dtype)


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
