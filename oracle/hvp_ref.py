"""Oracle: Hessian-vector products by torch double-backward, on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  These restate the
reference's ``hess_vec`` family op for op (slice -> loss -> grad with graph ->
sum(v*g) -> backward -> cat of .grad), device-agnostic so they run on host
cores.  ``loss_of(model, batch)`` abstracts the two loss shapes the reference
uses: ``model(input_ids=ids, labels=ids).loss`` for the language models and
``criterion(model(x), y)`` for the CIFAR nets.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

import torch


def lm_loss(model, batch):
    """/root/reference/gpt2_hessian_cpu.py:94-97 — labels = input ids, no mask."""
    ids = batch["input_ids"] if isinstance(batch, dict) else batch
    loss = model(input_ids=ids, labels=ids).loss
    if loss.dim() > 0:
        loss = loss.mean()
    return loss


def _slice_views(vector: torch.Tensor, params: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    views, off = [], 0
    for p in params:                                           # :79-82
        views.append(vector[off: off + p.numel()].detach().view_as(p).to(p.device))
        off += p.numel()
    return views


def hess_vec(vector: torch.Tensor, batch, model,
             loss_of: Callable = lm_loss) -> torch.Tensor:
    """Single-batch HVP.  Follows /root/reference/gpt2_hessian_cpu.py:75-109."""
    params = list(model.parameters())
    views = _slice_views(vector, params)
    model.eval()                                               # :84
    model.zero_grad()                                          # :88
    loss = loss_of(model, batch)                               # :94-97
    grads = torch.autograd.grad(loss, params, create_graph=True)   # :102
    s = torch.zeros(1, device=vector.device if vector.device == params[0].device else params[0].device)
    for v, g in zip(views, grads):                             # :104-105
        s = s + torch.sum(v * g)
    s.backward()                                               # :106
    return torch.cat([p.grad.view(-1) for p in params]).view(-1)   # :109


def hess_vec_dataset(vector: torch.Tensor, batches: Iterable, model,
                     weights: Optional[Sequence[float]] = None,
                     loss_of: Callable = lm_loss) -> torch.Tensor:
    """Dataset-averaged HVP.  Follows /root/reference/gpt2_savehessian.py:130-163
    with the B_i/N weighting of /root/reference/diego_pythia.py:114 and
    /root/reference/train_savespec.py:82 (the ``len(batch)/N`` of
    gpt2_savehessian.py:154 counts dict keys -- quirk Q6, not replicated).
    ``weights[i]`` multiplies batch i's loss; default B_i / sum(B).
    """
    params = list(model.parameters())
    views = _slice_views(vector, params)
    batches = list(batches)
    if weights is None:
        sizes = [float((b["input_ids"] if isinstance(b, dict) else b[0] if isinstance(b, (tuple, list)) else b).shape[0])
                 for b in batches]
        tot = sum(sizes)
        weights = [s / tot for s in sizes]
    model.eval()
    model.zero_grad()
    for wgt, batch in zip(weights, batches):                   # :145
        loss = loss_of(model, batch) * wgt                     # :149-154
        grads = torch.autograd.grad(loss, params, create_graph=True)
        s = torch.zeros(1, device=params[0].device)
        for v, g in zip(views, grads):
            s = s + torch.sum(v * g)
        s.backward()                                           # :160 accumulates into .grad
    return torch.cat([p.grad.view(-1) for p in params]).view(-1)


def hess_vec_subset(vector: torch.Tensor, batches: Iterable, model,
                    params: Sequence[torch.nn.Parameter],
                    weights: Optional[Sequence[float]] = None,
                    loss_of: Callable = lm_loss) -> torch.Tensor:
    """HVP restricted to a parameter subset (one transformer block).
    Follows /root/reference/ipynbs/visual-eigen.ipynb cell 10:5-42:
    full forward, grads only wrt ``params``, vector has sum(numel(params)) entries.
    """
    params = list(params)
    views = _slice_views(vector, params)
    batches = list(batches)
    if weights is None:
        sizes = [float((b["input_ids"] if isinstance(b, dict) else b).shape[0]) for b in batches]
        tot = sum(sizes)
        weights = [s / tot for s in sizes]
    model.eval()
    model.zero_grad()
    for wgt, batch in zip(weights, batches):
        loss = loss_of(model, batch) * wgt
        grads = torch.autograd.grad(loss, params, create_graph=True, retain_graph=True)
        s = torch.zeros(1, device=params[0].device)
        for v, g in zip(views, grads):
            s = s + torch.sum(v * g)
        s.backward()
    return torch.cat([p.grad.view(-1) for p in params]).view(-1)


def hess_vec_per_tensor(vector: torch.Tensor, batch, model,
                        loss_of: Callable = lm_loss) -> torch.Tensor:
    """Block-diagonal-by-parameter-tensor HVP for ONE batch.
    Follows /root/reference/gpt2_savehessian_layer.py:155-173 (and the
    single-tensor form /root/reference/lanczostrain_layer_hand.py:74-91):
    each tensor i is differentiated alone, g_i = dL/dtheta_i with graph, then
    d(g_i . v_i)/dtheta_i.
    """
    params = list(model.parameters())
    views = _slice_views(vector, params)
    model.eval()
    model.zero_grad()
    loss = loss_of(model, batch)
    out = torch.zeros_like(vector)
    off = 0
    for p, v in zip(params, views):
        g = torch.autograd.grad(loss, p, create_graph=True, retain_graph=True)[0]   # :157
        gv = torch.sum(g * v)                                                       # :163
        g2 = torch.autograd.grad(gv, p, retain_graph=True)[0]                       # :166
        out[off: off + p.numel()] = g2.reshape(-1)                                  # :169
        off += p.numel()
    return out
