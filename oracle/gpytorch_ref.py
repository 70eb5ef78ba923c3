"""Oracle: the Lanczos tridiagonalisation the reference gets from its third-party dependency.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED.  ``gpytorch.utils.lanczos.lanczos_tridiag`` (in current releases a re-export of
``linear_operator.utils.lanczos.lanczos_tridiag``) is called at /root/reference/gpt2_hessian_cpu.py:207-213
and 17 other sites, but it is not vendored under /root/reference, no requirements / lock
file pins its version, it is not installed in this image and cannot be fetched (no network), and nothing in
the reference's own tests or notebooks pins one of its outputs reproducibly (its only recorded T,
Discrepancy.ipynb cell 2, depends on the library's internal random start vector).  This file therefore
restates the library's PUBLISHED algorithm as documented in SURVEY Appendix B -- from documentation, not from
the source -- for a single probe vector, so that the product's ``lanczos_tridiag(..., reorth_tol=...)`` shim
is at least pinned to the documented behaviour:

  * q_0 = init / |init|;  r = A q_0;  alpha_0 = q_0 . r;  r -= alpha_0 q_0;  beta_0 = |r|;  q_1 = r / beta_0
  * for k = 1 .. m-1:   r = A q_k - beta_{k-1} q_{k-1};   alpha_k = q_k . r      (beta removed BEFORE alpha)
        if k + 1 < m:   r -= alpha_k q_k
                        one classical Gram-Schmidt pass against q_0..q_k:  r -= Q^T (Q r)
                        beta_k = |r|  (recorded HERE, after the first pass);  r /= beta_k
                        up to 10 more [CGS pass + renormalise] while ANY (q_i . r) > tol   -- a SIGNED compare
                        q_{k+1} = r;   stop if |beta_k| < 1e-6 or the 10 passes did not suffice
  * T is m' x m' with m' = number of HVPs performed (<= max_iter); Q is [P, m'].
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

Matvec = Callable[[torch.Tensor], torch.Tensor]


def gpytorch_like_tridiag(matvec: Matvec, init_vec: torch.Tensor, max_iter: int, tol: float = 1e-5,
                          dtype: torch.dtype = torch.float32, breakdown: float = 1e-6, max_extra_passes: int = 10):
    """SURVEY Appendix B, single probe vector.  ``matvec`` maps [P] -> [P].
    Returns dict(Q[P, m'], T[m', m'], alphas, betas, extra_passes[list per iteration], m_eff)."""
    n = init_vec.numel()
    m = min(int(max_iter), n)
    Q = torch.zeros(m, n, dtype=dtype)
    T = torch.zeros(m, m, dtype=dtype)
    extra = [0] * m
    q0 = init_vec.reshape(-1).to(dtype)
    q0 = q0 / torch.norm(q0, 2)
    Q[0] = q0
    r = matvec(q0).reshape(-1).to(dtype).clone()
    a0 = torch.sum(q0 * r)
    r -= a0 * q0
    b0 = torch.norm(r, 2)
    T[0, 0] = a0
    k = 0
    if m > 1:
        T[0, 1] = b0
        T[1, 0] = b0
        Q[1] = r / b0
        for k in range(1, m):
            q_prev, q_cur = Q[k - 1], Q[k]
            beta_prev = T[k, k - 1]
            r = matvec(q_cur).reshape(-1).to(dtype) - q_prev * beta_prev
            alpha = torch.sum(q_cur * r)
            T[k, k] = alpha
            if k + 1 < m:
                r = r - alpha * q_cur
                basis = Q[: k + 1]
                r = r - basis.t() @ (basis @ r)                 # the unconditional classical Gram-Schmidt pass
                nrm = torch.norm(r, 2)
                r = r / nrm
                T[k, k + 1] = nrm                                # beta is the norm after the FIRST pass
                T[k + 1, k] = nrm
                inner = basis @ r
                ok = False
                for _ in range(max_extra_passes):
                    if not bool((inner > tol).any()):            # signed compare, as documented
                        ok = True
                        break
                    r = r - basis.t() @ (basis @ r)
                    r = r / torch.norm(r, 2)
                    inner = basis @ r
                    extra[k] += 1
                Q[k + 1] = r
                if float(nrm.abs()) < breakdown or not ok:
                    break
    m_eff = k + 1
    T = T[:m_eff, :m_eff].clone()
    alphas = torch.diagonal(T).clone()
    betas = torch.diagonal(T, 1).clone()
    return {"Q": Q[:m_eff].t().contiguous(), "T": T, "alphas": alphas, "betas": betas,
            "extra_passes": extra[:m_eff], "m_eff": m_eff}
