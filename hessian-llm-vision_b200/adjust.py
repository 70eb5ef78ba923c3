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
