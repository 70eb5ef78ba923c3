"""Ritz values, SLQ weights and spectral densities from the Lanczos tridiagonal.

The tridiagonal eigensolve stays on the host (LAPACK ?stev through SciPy): T is
at most a few hundred rows, so this is microseconds and not worth a kernel.
Reference: ``eigvals, eigvects = torch.linalg.eigh(T); gammas = eigvects[0,:]**2``
(gpt2_hessian_cpu.py:215-216, lanczostrain_hand.py:208-209).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from scipy.linalg import lapack


def tridiag_eigh(alphas, betas) -> Tuple[np.ndarray, np.ndarray]:
    """Eigen-decomposition of the symmetric tridiagonal (diag=alphas, off-diag=betas[:m-1]).

    Returns (eigvals ascending [m], Y [m, m] with eigenvectors in columns), float64.
    """
    d = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
    m = d.shape[0]
    e = np.ascontiguousarray(np.asarray(betas, dtype=np.float64)[: max(m - 1, 0)])
    if m == 0:
        return np.zeros(0), np.zeros((0, 0))
    if m == 1:
        return d.copy(), np.ones((1, 1))
    vals, vecs, info = lapack.dstev(d, e, compute_v=1)
    if info != 0:
        raise RuntimeError(f"LAPACK dstev failed with info={info}")
    return vals, vecs


def ritz_values(alphas, betas, dtype=torch.float32):
    """(eigvals, gammas) as torch CPU tensors in the reference's layout: eigvals ascending,
    gammas[i] = (first component of eigenvector i)^2, sum(gammas) = 1."""
    vals, vecs = tridiag_eigh(alphas, betas)
    gammas = vecs[0, :] ** 2 if vecs.size else np.zeros(0)
    return torch.from_numpy(vals).to(dtype), torch.from_numpy(gammas).to(dtype), vecs


def dense_T(alphas, betas, dtype=torch.float32) -> torch.Tensor:
    """Dense m x m T exactly as the reference fills it (lanczostrain_hand.py:171,184,191-192,201)."""
    a = torch.as_tensor(np.asarray(alphas, dtype=np.float64))
    m = a.numel()
    T = torch.zeros(m, m, dtype=torch.float64)
    idx = torch.arange(m)
    T[idx, idx] = a
    if m > 1:
        b = torch.as_tensor(np.asarray(betas, dtype=np.float64))[: m - 1]
        T[idx[:-1], idx[1:]] = b
        T[idx[1:], idx[:-1]] = b
    return T.to(dtype)


def slq_density(eigvals_list: Sequence, gammas_list: Sequence, grid: Optional[np.ndarray] = None,
                sigma: Optional[float] = None, num_points: int = 1024, margin: float = 0.05):
    """Stochastic-Lanczos-quadrature spectral density: average over probes of
    sum_i gamma_i * N(x; lambda_i, sigma^2).  (The reference only stem-plots single probes,
    `GPT2 spectrum.ipynb` cell 3; multi-probe protocol: d.sh:4-11.)

    Returns (grid, density) as numpy arrays; density integrates to ~1.
    """
    ev = [np.asarray(torch.as_tensor(e).cpu(), dtype=np.float64) for e in eigvals_list]
    gm = [np.asarray(torch.as_tensor(g).cpu(), dtype=np.float64) for g in gammas_list]
    lo = min(e.min() for e in ev)
    hi = max(e.max() for e in ev)
    span = max(hi - lo, 1e-12)
    if grid is None:
        grid = np.linspace(lo - margin * span, hi + margin * span, num_points)
    if sigma is None:
        sigma = 0.01 * span
    dens = np.zeros_like(grid)
    for e, g in zip(ev, gm):
        z = (grid[:, None] - e[None, :]) / sigma
        dens += (np.exp(-0.5 * z * z) * g[None, :]).sum(axis=1) / (sigma * np.sqrt(2 * np.pi))
    dens /= len(ev)
    return grid, dens
