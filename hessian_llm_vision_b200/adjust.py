"""Low-rank "Lanczos-preconditioned" gradient adjustment -- the consumer of the Ritz pairs in
the reference's optimiser scripts, and the one place the reference has native code.

    g_adj = g + sum_i (1/lam_i - 1/(lam_i + delta)) (g . V_i) V_i
        torch loop ........ gpt2_hessian_cpu.py:224-229 (k H2D copies of V_i per batch)
        CUDA attempt ...... vector_adjust.cu:2-15 via gpt_hessian_cuda.py:27-54 (O(k n^2) loads)

Here it is two streaming passes over V (the same kernels as the reorthogonalisation) --
``hlv_vector_adjust_f32`` keeps the reference kernel's argument order.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels


def cuda_vector_adjust(grad_vector: torch.Tensor, V: torch.Tensor, eigvals: torch.Tensor,
                       adjusted_grad_vector: torch.Tensor, delta: float,
                       ws: Optional[kernels.Workspace] = None) -> torch.Tensor:
    """Same name, arguments and in-place ``+=`` semantics as gpt_hessian_cuda.py:27-54."""
    if ws is None:
        ws = kernels.Workspace(grad_vector.device, max_rows=max(int(eigvals.numel()), 1))
    return kernels.vector_adjust(grad_vector.contiguous(), V, eigvals.contiguous().to(torch.float32),
                                 adjusted_grad_vector, delta, ws)


def adjust_gradient(grad_vector: torch.Tensor, V: torch.Tensor, eigvals: torch.Tensor, delta: float,
                    ws: Optional[kernels.Workspace] = None) -> torch.Tensor:
    """Out-of-place form used by the training loops: clone, then adjust (gpt2_hessian_cpu.py:202,224-229)."""
    out = grad_vector.clone()
    return cuda_vector_adjust(grad_vector, V, eigvals.to(grad_vector.device), out, delta, ws)


def ema_ritz(V: torch.Tensor, eigvals: torch.Tensor, V_old: Optional[torch.Tensor],
             eigvals_old: Optional[torch.Tensor], momentum: float):
    """Optional EMA of the Ritz pairs across refreshes (gpt2_hessian_cpu.py:218-222)."""
    if momentum > 0 and V_old is not None:
        V = torch.lerp(V, V_old, momentum)            # momentum*V_old + (1-momentum)*V
        eigvals = momentum * eigvals_old + (1 - momentum) * eigvals
    return V, eigvals


def adjust_gradient_implicit(grad_vector: torch.Tensor, Q: torch.Tensor, m: int, Y, eigvals, delta: float,
                             select=None, ws: Optional[kernels.Workspace] = None,
                             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The same adjustment WITHOUT forming the Ritz vectors.  With V = Y^T Q (gpt2_hessian_cpu.py:217),
        sum_i s_i (g . V_i) V_i = Q^T ( Y diag(s) Y^T ) (Q g),      s_i = 1/lam_i - 1/(lam_i + delta)
    so it is one projection pass and one update pass over the Lanczos basis Q (the reorthogonalisation
    kernels), plus an m x m product on the device -- 2*m*s*n bytes however many Ritz pairs are used,
    instead of materialising k vectors of length n.

    Q: [>=m, ld] basis rows (fp32 or bf16); Y: [m, m] eigenvectors of T in columns (ritz.tridiag_eigh);
    eigvals: [m]; select: indices of the Ritz pairs to use (default: all)."""
    dev = grad_vector.device
    n = grad_vector.numel()
    if ws is None:
        ws = kernels.Workspace(dev, max_rows=max(int(m), 1))
    Yd = torch.as_tensor(Y, dtype=torch.float64).to(dev)[:m, :m]
    lam = torch.as_tensor(eigvals, dtype=torch.float64).to(dev)[:m]
    s = 1.0 / lam - 1.0 / (lam + float(delta))
    if select is not None:
        keep = torch.zeros(m, dtype=torch.bool, device=dev)
        keep[torch.as_tensor(select, device=dev, dtype=torch.long)] = True
        s = torch.where(keep, s, torch.zeros_like(s))
    c = torch.empty(m, dtype=torch.float64, device=dev)
    kernels.cgs_project(Q, m, grad_vector.contiguous(), c, ws)          # c = Q g
    c2 = (Yd * s) @ (Yd.t() @ c)                                        # Y diag(s) Y^T c  (m x m, fp64)
    if out is None:
        out = grad_vector.clone()
    elif out.data_ptr() != grad_vector.data_ptr():
        out.copy_(grad_vector)
    kernels.cgs_update(Q, m, c2.contiguous(), out, None, ws, sign=1.0)  # g += Q^T c2
    return out
