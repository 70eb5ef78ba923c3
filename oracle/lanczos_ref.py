"""Oracle: Lanczos recurrences, reorthogonalisation, Ritz post-processing.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain torch CPU ops in
the reference's own arithmetic (fp32 unless a dtype is passed), one function
per reference algorithm, each citing the file:line it follows.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

Matvec = Callable[[torch.Tensor], torch.Tensor]


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16 and back: models a bf16-stored basis row."""
    return x.to(torch.bfloat16).to(x.dtype)


# --------------------------------------------------------------------------
# A.2  hand-written Lanczos, no reorthogonalisation
# --------------------------------------------------------------------------
def hand_lanczos(matvec: Matvec, v0: torch.Tensor, k: int,
                 dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Follows /root/reference/lanczostrain_hand.py:171-203 (identical loops at
    accum.py:167-199 and diego_pythia.py:157-188).

    ``k`` is the reference's ``lanczos_iters``: the loop does k+1 operator
    applications and fills a dense (k+1)x(k+1) ``T`` and a (k+1, n) ``Q``.
    alpha is taken BEFORE beta*v_old is removed (line 200 vs 202); there is no
    breakdown guard (line 190-193), exactly like the reference.
    """
    n = v0.numel()
    T = torch.zeros(k + 1, k + 1, dtype=dtype)
    Q = torch.zeros(k + 1, n, dtype=dtype)
    v = v0.to(dtype).clone()
    Q[0] = v                                      # :177
    w = matvec(v).to(dtype).clone()               # :180
    alpha = torch.dot(w, v)                       # :183
    T[0, 0] = alpha
    w -= alpha * v                                # :185
    v_old = v
    for i in range(k):                            # :188
        b = torch.norm(w, 2)                      # :190
        T[i + 1, i] = b
        T[i, i + 1] = b
        v = w / b                                 # :193
        Q[i + 1] = v
        w = matvec(v).to(dtype).clone()           # :197
        alpha = torch.dot(w, v)                   # :200
        T[i + 1, i + 1] = alpha
        w -= (alpha * v + b * v_old)              # :202
        v_old = v
    return T, Q


# --------------------------------------------------------------------------
# A.3  the reference's own full-reorth variant (one-pass modified Gram-Schmidt)
# --------------------------------------------------------------------------
def reorth_lanczos_mgs(matvec: Matvec, v: torch.Tensor, m: int,
                       dtype: torch.dtype = torch.float32,
                       tol: float = 1e-6) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Follows /root/reference/Lanczos_Scratch/Discrepancy.ipynb cell 1:30-54.

    ``v`` is NOT normalised by the caller (the notebook divides by its norm on
    the first pass).  Returns the m x m ``T`` and the list of basis vectors.
    """
    T = torch.zeros(m, m, dtype=dtype)
    r = v.to(dtype).clone()
    q_old = torch.zeros_like(r)
    b = torch.norm(r, p=2)
    basis: List[torch.Tensor] = []
    for i in range(m):
        q = r / b
        basis.append(q)
        u = matvec(q).to(dtype) - b * q_old       # 1:41
        alpha = torch.dot(u, q)
        T[i, i] = alpha
        r = u - alpha * q
        for j in range(len(basis)):               # 1:44-45  (MGS: uses the running r)
            r -= torch.dot(r, basis[j]) * basis[j]
        b = torch.norm(r, p=2)
        if i < m - 1:
            T[i, i + 1] = b
            T[i + 1, i] = b
        q_old = q
        if b < tol:                               # 1:53
            break
    return T, basis


# --------------------------------------------------------------------------
# Product algorithm: hand-loop order (A.2) + two-pass classical Gram-Schmidt
# --------------------------------------------------------------------------
def lanczos_cgs2(matvec: Matvec, v0: torch.Tensor, m: int,
                 reorth: Optional[str] = "full",
                 dtype: torch.dtype = torch.float32,
                 basis_dtype: Optional[torch.dtype] = None,
                 passes: int = 2,
                 breakdown_tol: float = 0.0):
    """What the CUDA engine computes, restated on the CPU.

    Recurrence order is the reference hand loop's
    (/root/reference/lanczostrain_hand.py:188-203): beta=||w||, v=w/beta,
    w=Hv, alpha=w.v (before removing beta*v_old), w -= alpha*v + beta*v_old.
    With ``reorth='full'`` the three-term update is followed by ``passes``
    rounds of classical Gram-Schmidt against every stored row,
    c = Q[:j+1] w ; w -= Q[:j+1]^T c  -- the bandwidth-friendly form of the
    reference's reorth loop (Discrepancy.ipynb cell 1:44-45; gpytorch does the
    same CGS form).  ``basis_dtype=torch.bfloat16`` models
    bf16 storage of the rows (the fp32 v_j still drives the HVP and the
    three-term update).

    ``m`` = number of operator applications = size of T.
    Returns dict(alphas[m], betas[m], T[m,m], Q[m,n], m_eff); betas[j] couples
    j and j+1, betas[m-1] is the final residual norm.  With ``reorth=None`` and
    m=k+1 the T/Q equal ``hand_lanczos(…, k)`` bit for bit.
    """
    n = v0.numel()
    where = v0.device                       # CPU in the tests; a CUDA device only for the full-size evidence runs
    alphas = torch.zeros(m, dtype=dtype, device=where)
    betas = torch.zeros(m, dtype=dtype, device=where)
    Q = torch.zeros(m, n, dtype=dtype, device=where)
    store = (lambda x: bf16_round(x)) if basis_dtype == torch.bfloat16 else (lambda x: x)
    v = v0.to(dtype).clone()
    Q[0] = store(v)
    v_old = torch.zeros_like(v)
    beta = torch.zeros((), dtype=dtype, device=where)
    m_eff = m
    for j in range(m):
        w = matvec(v).to(dtype).clone()
        alpha = torch.dot(w, v)
        alphas[j] = alpha
        if j == 0:
            w -= alpha * v
        else:
            w -= (alpha * v + beta * v_old)
        if reorth == "full":
            for _ in range(passes):
                c = Q[: j + 1] @ w
                w -= Q[: j + 1].t() @ c
        beta = torch.norm(w, 2)
        betas[j] = beta
        if breakdown_tol > 0.0 and float(beta) < breakdown_tol:
            m_eff = j + 1
            break
        if j < m - 1:
            v_old = v
            v = w / beta
            Q[j + 1] = store(v)
    T = torch.zeros(m_eff, m_eff, dtype=dtype, device=where)
    for j in range(m_eff):
        T[j, j] = alphas[j]
        if j + 1 < m_eff:
            T[j, j + 1] = betas[j]
            T[j + 1, j] = betas[j]
    return {"alphas": alphas[:m_eff], "betas": betas[:m_eff], "T": T,
            "Q": Q[:m_eff], "m_eff": m_eff}


# --------------------------------------------------------------------------
# A.4  Ritz values, SLQ weights, Ritz vectors
# --------------------------------------------------------------------------
def ritz(T: torch.Tensor, Q: Optional[torch.Tensor] = None):
    """Follows /root/reference/gpt2_hessian_cpu.py:215-217 (hand-loop flavour
    /root/reference/lanczostrain_hand.py:208-210, Q stored as rows):
    eigvals ascending, gammas = first-row squares, V = Y^T Q (rows = Ritz vectors).
    """
    eigvals, eigvects = torch.linalg.eigh(T)
    gammas = eigvects[0, :] ** 2
    V = eigvects.t() @ Q if Q is not None else None
    return eigvals, gammas, V


# --------------------------------------------------------------------------
# Low-rank gradient adjustment (the reference's only native kernel)
# --------------------------------------------------------------------------
def lowrank_adjust(grad: torch.Tensor, V: torch.Tensor, eigvals: torch.Tensor,
                   delta: float) -> torch.Tensor:
    """Follows /root/reference/gpt2_hessian_cpu.py:224-229 and the CUDA
    statement of the same sum, /root/reference/vector_adjust.cu:2-15:
    out = g + sum_i (1/lam_i - 1/(lam_i+delta)) (g . V_i) V_i ; dots use the
    ORIGINAL g for every i.
    """
    out = grad.clone()
    for i in range(eigvals.numel()):
        lam = eigvals[i]
        coeff = (1 / lam - 1 / (lam + delta)) * torch.dot(grad, V[i])
        out += coeff * V[i]
    return out


# --------------------------------------------------------------------------
# flatten / split (gather / scatter)
# --------------------------------------------------------------------------
def flatten_tensors(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """/root/reference/gpt2_hessian_cpu.py:109 and :200 — torch.cat of views."""
    return torch.cat([t.reshape(-1) for t in tensors]).view(-1)


def split_like(vec: torch.Tensor, like: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """/root/reference/gpt2_hessian_cpu.py:79-82 and :231-233 — slice + view."""
    out, off = [], 0
    for t in like:
        out.append(vec[off: off + t.numel()].view(t.shape))
        off += t.numel()
    return out
