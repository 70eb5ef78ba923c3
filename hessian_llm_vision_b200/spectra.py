"""Spectrum drivers on top of ``lanczos``: multi-probe stochastic Lanczos quadrature and
per-transformer-block spectra.

  multi-probe SLQ ....... the reference runs probes as separate processes, sequentially
                          (d.sh:4-11: 3 data seeds x 3 vector seeds of diego_pythia_tiny.py);
                          probe vector = randn(P)/norm under torch.manual_seed(vector_seed)
                          (diego_pythia.py:147-149)
  per-block spectra ..... ipynbs/visual-eigen.ipynb cell 12: one Lanczos run per transformer block,
                          operator restricted to the block's parameters (cell 10)

Both are embarrassingly parallel over probes / blocks: with a process group they are dealt
round-robin to the ranks with NO data-path collective ("replicas only"); the (eigvals, gammas)
pairs are exchanged once at the end.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import ritz as _ritz
from .hvp import HessianVectorProduct, lm_loss
from .lanczos import Comm, LanczosEngine, LanczosResult


def probe_vector(n: int, seed: int, device) -> torch.Tensor:
    """diego_pythia.py:147-149: ``torch.manual_seed(vector_seed); randn(P, device); / norm``."""
    g = torch.Generator(device=device).manual_seed(int(seed))
    v = torch.randn(n, device=device, generator=g)
    return v / torch.linalg.vector_norm(v)


@dataclass
class SLQResult:
    seeds: List[int]
    eigvals: List[torch.Tensor]
    gammas: List[torch.Tensor]
    tridiagonals: List[torch.Tensor] = field(default_factory=list)

    def density(self, **kw):
        """Gaussian-broadened spectral density averaged over probes -> (grid, density)."""
        return _ritz.slq_density(self.eigvals, self.gammas, **kw)

    def eigeninfo(self):
        """Probe-averaged result dict in the reference's layout: all Ritz values (ascending) with
        weights gamma / n_probes, so that sum(gammas) = 1 still holds."""
        ev = torch.cat(self.eigvals)
        gm = torch.cat(self.gammas) / len(self.eigvals)
        order = torch.argsort(ev)
        return {"eigvals": ev[order], "gammas": gm[order]}


def _exchange(comm: Optional[Comm], items: list) -> list:
    if comm is None or comm.world == 1:
        return items
    import torch.distributed as dist
    out = [None] * comm.world
    dist.all_gather_object(out, items, group=comm.group)
    merged = [x for part in out for x in part]
    return sorted(merged, key=lambda t: t[0])


def slq(hvp: Callable, n: int, n_iter: int, seeds: Sequence[int], device, reorth: Optional[str] = "full",
        basis_dtype: torch.dtype = torch.float32, replicas: Optional[Comm] = None,
        on_probe: Optional[Callable[[int, LanczosResult], None]] = None) -> SLQResult:
    """One Lanczos run of ``n_iter`` iterations per probe seed; the engine (and its basis
    allocation) is reused across probes.  ``replicas``: deal probes round-robin over the ranks of
    a process group (each rank must hold the full operator).
    The LanczosResult handed to ``on_probe`` shares the engine's basis buffer: its ``Q`` / ``ritz_vectors()`` are
    valid only inside the callback (the next probe overwrites the rows); clone what must outlive it."""
    rank, world = (replicas.rank, replicas.world) if replicas is not None else (0, 1)
    eng = LanczosEngine(hvp, n, n_iter, device, reorth=reorth, basis_dtype=basis_dtype)
    mine = []
    for k, seed in enumerate(seeds):
        if k % world != rank:
            continue
        eng.start(probe_vector(n, seed, device))
        for j in range(n_iter):
            eng.step(j)
        res = eng.result()
        if on_probe is not None:
            on_probe(int(seed), res)
        mine.append((k, int(seed), res.eigvals, res.gammas, res.T))
    allr = _exchange(replicas, mine)
    return SLQResult(seeds=[s for _, s, _, _, _ in allr], eigvals=[e for _, _, e, _, _ in allr],
                     gammas=[g for _, _, _, g, _ in allr], tridiagonals=[t for _, _, _, _, t in allr])


def per_block_spectra(model: torch.nn.Module, batches, n_iter: int, blocks: Optional[Sequence[torch.nn.Module]] = None,
                      seed: int = 0, loss_fn: Callable = lm_loss, reorth: Optional[str] = "full",
                      replicas: Optional[Comm] = None, **op_kwargs):
    """visual-eigen.ipynb cell 12: for every transformer block, Lanczos on the Hessian restricted to
    that block's parameters.  Returns (all_eigvals, all_gammas) lists ordered by block index."""
    if blocks is None:
        blocks = list(model.transformer.h)
    rank, world = (replicas.rank, replicas.world) if replicas is not None else (0, 1)
    mine = []
    for i, blk in enumerate(blocks):
        if i % world != rank:
            continue
        params = list(blk.parameters())
        op = HessianVectorProduct(model, batches, loss_fn=loss_fn, params=params, **op_kwargs)
        dev = params[0].device
        eng = LanczosEngine(op, op.n, n_iter, dev, reorth=reorth)
        eng.start(probe_vector(op.n, seed + i, dev))
        for j in range(n_iter):
            eng.step(j)
        res = eng.result()
        mine.append((i, seed + i, res.eigvals, res.gammas, res.T))
        del eng
    allr = _exchange(replicas, mine)
    return [e for _, _, e, _, _ in allr], [g for _, _, _, g, _ in allr]
