"""Lanczos tridiagonalisation around a Hessian-vector-product callable, with the
recurrence, reorthogonalisation and gather running as sm_100a kernels (libhlv).

Public surface (mirrors what the reference scripts hand-roll or get from gpytorch):

    lanczos(hvp, n_iter, v0, reorth=None|'full', ...) -> LanczosResult
        follows the reference's hand loop (lanczostrain_hand.py:171-203): n_iter = number of
        HVPs = size of T (the hand loop with lanczos_iters=k corresponds to n_iter=k+1).
    lanczos_tridiag(matmul_closure, max_iter, dtype, device, matrix_shape, ...) -> (Q[P,m], T[m,m])
        signature-compatible shim for gpytorch.utils.lanczos.lanczos_tridiag as called at
        gpt2_hessian_cpu.py:207-213 (full reorthogonalisation, closure sees [P,1]).

Nothing in the loop synchronises with the host: alpha/beta/coefficients live in device
doubles, kernels read them from there, and T is copied back once at the end (or every
`check_every` iterations to detect breakdown).  With a process group (`comm`), the basis
and all vectors are sharded along the parameter dimension: each rank runs the same
kernels on its shard and the partial scalars are combined with k-float all-reduces.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import kernels as _kernels      # the CUDA library front end (raises at first use if libhlv.so is not built)
from . import ritz as _ritz
from ._lib import CH_HV as _CH_HV, CH_V as _CH_V

_ALIGN = 8          # elements: keeps fp32 and bf16 rows 16-byte aligned


# ----------------------------------------------------------------------------
# process-group wrapper
# ----------------------------------------------------------------------------
class Comm:
    """Thin wrapper over torch.distributed for the three collectives the path needs.
    ``Comm(None)`` with no initialised process group is the single-process identity."""

    def __init__(self, group=None, enabled: Optional[bool] = None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        on = dist.is_available() and dist.is_initialized() if enabled is None else enabled
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.backend = dist.get_backend(group) if on else "none"

    def all_reduce_sum(self, t: torch.Tensor) -> None:
        if self.world > 1:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)

    def reduce_scatter_sum(self, out: torch.Tensor, full: torch.Tensor) -> None:
        """out = (sum over ranks of full)[rank*len(out) : (rank+1)*len(out)]"""
        if self.world == 1:
            if out.data_ptr() != full.data_ptr():
                out.copy_(full[: out.numel()])
            return
        if self.backend == "gloo":            # gloo has no reduce_scatter: all-reduce and slice
            self._dist.all_reduce(full, op=self._dist.ReduceOp.SUM, group=self.group)
            out.copy_(full[self.rank * out.numel(): (self.rank + 1) * out.numel()])
        else:
            self._dist.reduce_scatter_tensor(out, full, op=self._dist.ReduceOp.SUM, group=self.group)

    def all_gather(self, full: torch.Tensor, shard: torch.Tensor) -> None:
        if self.world == 1:
            if full.data_ptr() != shard.data_ptr():
                full[: shard.numel()].copy_(shard)
            return
        self._dist.all_gather_into_tensor(full, shard, group=self.group)

    def barrier(self) -> None:
        if self.world > 1:
            self._dist.barrier(group=self.group)


# ----------------------------------------------------------------------------
# CUDA-event phase timer (optional)
# ----------------------------------------------------------------------------
class _Phases:
    """Per-phase CUDA-event timing (profile=True) and NVTX ranges (HLV_NVTX=1) around the kernels of an
    iteration.  The reference only has wall-clock prints around each HVP (gpt2_savehessian.py:178-188)."""

    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.nvtx = os.environ.get("HLV_NVTX", "0") == "1"
        self.pairs: List[Tuple[str, int, Any, Any]] = []
        self._open: Dict[str, Any] = {}

    def start(self, name: str) -> None:
        if self.nvtx:
            torch.cuda.nvtx.range_push(f"hlv:{name}")
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._open[name] = e

    def stop(self, name: str, rows: int = 0) -> None:
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.pairs.append((name, rows, self._open.pop(name), e))
        if self.nvtx:
            torch.cuda.nvtx.range_pop()

    def summary(self) -> Dict[str, Dict[str, float]]:
        out: Dict[str, Dict[str, float]] = {}
        if not self.enabled:
            return out
        torch.cuda.synchronize()
        for name, rows, a, b in self.pairs:
            d = out.setdefault(name, {"ms": 0.0, "calls": 0, "rows": 0})
            d["ms"] += a.elapsed_time(b)
            d["calls"] += 1
            d["rows"] += rows
        return out


# ----------------------------------------------------------------------------
# result
# ----------------------------------------------------------------------------
@dataclass
class LanczosResult:
    """Outcome of a Lanczos run.

    alphas[j] = T[j,j]; betas[j] = T[j,j+1] for j < m-1 and betas[m-1] is the final residual
    norm.  eigvals (ascending) / gammas follow gpt2_hessian_cpu.py:215-216.  ``basis`` holds the
    Lanczos vectors as ROWS (the hand loop's Q, lanczostrain_hand.py:176) on the device, local
    shard only when the run was sharded."""
    alphas: torch.Tensor
    betas: torch.Tensor
    eigvals: torch.Tensor
    gammas: torch.Tensor
    Y: np.ndarray
    m: int
    n: int
    breakdown: bool = False
    basis: Optional[torch.Tensor] = None
    n_local: int = 0
    timings: Dict[str, Dict[str, float]] = field(default_factory=dict)
    conditional_passes: Optional[int] = None      # reorth_tol runs: iterations whose last Gram-Schmidt pass was applied

    @property
    def T(self) -> torch.Tensor:
        return _ritz.dense_T(self.alphas.numpy(), self.betas.numpy())

    @property
    def Q(self) -> torch.Tensor:
        """[m, n_local] rows = Lanczos vectors (view into the device basis)."""
        if self.basis is None:
            raise RuntimeError("this run did not keep its basis (keep_basis=False)")
        return self.basis[: self.m, : self.n_local]

    def ritz_vectors(self, which: Optional[Sequence[int]] = None) -> torch.Tensor:
        """Rows = Ritz vectors  V = Y^T Q  (gpt2_hessian_cpu.py:217) for the eigenvalue indices in
        ``which`` (indices into the ascending ``eigvals``; default all), as a device fp32
        [len(which), n_local] tensor.  Streams the basis once per 8 requested vectors."""
        if self.basis is None:
            raise RuntimeError("this run did not keep its basis (keep_basis=False)")
        idx = list(range(self.m)) if which is None else [int(i) % self.m for i in which]
        dev = self.basis.device
        Ysel = torch.from_numpy(np.ascontiguousarray(self.Y[:, idx])).to(torch.float32).to(dev).contiguous()
        ld = (self.n_local + _ALIGN - 1) // _ALIGN * _ALIGN
        out = torch.empty(len(idx), ld, dtype=torch.float32, device=dev)
        _kernels.ritz_vectors(self.basis, self.m, Ysel, out, self.n_local)
        return out[:, : self.n_local]

    def eigeninfo(self, basis: bool = False) -> Dict[str, torch.Tensor]:
        """The reference's result dict (gpt2_savehessian.py:216-223; 'V' as train_savespec.py:328-333)."""
        d = {"eigvals": self.eigvals.clone(), "gammas": self.gammas.clone()}
        if basis:
            d["V"] = self.ritz_vectors()
        return d


# ----------------------------------------------------------------------------
# engine
# ----------------------------------------------------------------------------
class LanczosEngine:
    """Device-resident state of one Lanczos run; ``step(j)`` performs iteration j."""

    def __init__(self, hvp: Callable, n: int, n_iter: int, device, reorth: Optional[str] = None,
                 basis_dtype: torch.dtype = torch.float32, keep_basis: Optional[bool] = None,
                 breakdown_tol: Optional[float] = None, comm: Optional[Comm] = None,
                 profile: bool = False, column_vectors: bool = False, cgs_passes: int = 2, fused_cgs: bool = True,
                 reorth_tol: Optional[float] = None, exchange: str = "auto", peer=None, peer_allgather: Optional[bool] = None):
        if reorth not in (None, "full"):
            raise ValueError("reorth must be None or 'full'")
        if basis_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("basis_dtype must be torch.float32 or torch.bfloat16")
        if n_iter < 1:
            raise ValueError("n_iter must be >= 1")
        ops = self.ops = _kernels
        ops.require_device(device)            # no CPU path
        self.hvp = hvp
        self.n, self.m = int(n), int(n_iter)
        self.device = torch.device(device)
        self.reorth = reorth
        self.cgs_passes = int(cgs_passes)
        self.basis_dtype = basis_dtype
        self.keep_basis = (reorth == "full") if keep_basis is None else bool(keep_basis or reorth == "full")
        self.breakdown_tol = (1e-6 if reorth == "full" else 0.0) if breakdown_tol is None else float(breakdown_tol)
        self.comm = comm if comm is not None else Comm(enabled=False)
        self.column_vectors = column_vectors
        self.phases = _Phases(profile)
        G = self.comm.world
        self.shard_n = (-(-self.n // G) + _ALIGN - 1) // _ALIGN * _ALIGN      # per-rank length (zero padded)
        self.n_pad = self.shard_n * G
        self.lo = self.comm.rank * self.shard_n
        f32, f64 = dict(dtype=torch.float32, device=self.device), dict(dtype=torch.float64, device=self.device)
        sn, m = self.shard_n, self.m
        # --- basis and current/previous vectors (local shard) ---
        self.basis = None
        self.fp32_rows = self.keep_basis and basis_dtype == torch.float32
        if self.keep_basis:
            self.basis = torch.zeros(m, sn, dtype=basis_dtype, device=self.device)
        if not self.fp32_rows:
            self.ring = [torch.zeros(sn, **f32) for _ in range(2)]
        # --- full-length buffers (only distinct from the shard when sharded) ---
        self.v_full = None                                # allocated below, once the exchange mode is known
        self.hv_full = None
        self.w = torch.zeros(sn, **f32)
        # --- device scalars ---
        self.alphas = torch.zeros(m, **f64)
        self.betas = torch.zeros(m + 1, **f64)          # betas[j+1] = ||w|| after iteration j; betas[0] unused
        self.norm2 = torch.zeros(1, **f64)
        self.coef = torch.zeros(max(m, 1), **f64)
        self.coef2 = torch.zeros(max(m, 1), **f64)
        self.breakdown_iter = torch.full((1,), -1, dtype=torch.int32, device=self.device)
        self.ws = ops.Workspace(self.device, max_rows=max(m, 1) + 1)
        # fused middle pass of CGS2 (TMA-staged slab, basis read 3x instead of 4x per iteration)
        # Measured policy (profiles/r01_cgs_kernels.json): 1.5-1.9x over the update+project pair for fp32 rows
        # and 1.2-1.5x for bf16 rows; break-even near 4 rows.  fused_cgs="force" uses it from 1 row.
        want = fused_cgs == "force" or bool(fused_cgs)
        self.fused = (want and self.keep_basis and hasattr(ops, "cgs_update_project")
                      and sn < 2 ** 31 - 4096)            # the fused pass addresses tiles with 32-bit TMA coordinates
        self.fused_max_rows = ops.fused_max_rows(basis_dtype) if self.fused else 0
        self.fused_min_rows = 1 if fused_cgs == "force" else 4
        # reorth_tol (None = unconditional two-pass CGS, the default): the last pass of CGS2 is applied only when the
        # projection of the once-orthogonalised w still has a component > reorth_tol * |w| -- gpytorch's
        # "re-orthogonalise while any q_i . r > tol" (tol = 1e-5 there).  The test and the predication live on the
        # device (hlv_cgs_needs_pass, hlv_cgs_update_if_*): no host round trip, and a skipped pass costs no traffic.
        self.reorth_tol = None if reorth_tol is None else float(reorth_tol)
        if self.reorth_tol is not None:
            if not (self.fused and self.cgs_passes == 2):
                raise ValueError("reorth_tol needs the fused two-pass Gram-Schmidt path (reorth='full', cgs_passes=2, fused_cgs on)")
            self.pass_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.pass_count = torch.zeros(1, dtype=torch.int32, device=self.device)    # iterations whose last pass was applied
            self.norm2_b = torch.zeros(1, **f64)
        # the three-term update folded into the first projection (hlv_x_update_project): one launch and one pass less
        self.fold_update = self.reorth == "full" and hasattr(ops, "x_update_project")
        # how shards / coefficients move between ranks: "nccl" = torch.distributed collectives; "peer" = inside the libhlv
        # kernels over NVLink peer memory (csrc/hlv_peer.cuh)
        self.peer = None
        self.multicast = False
        self.multicast_store = False
        self.peer_allgather = peer_allgather
        self._v_pending = False
        self.exchange_mode = "none" if G == 1 else "nccl"
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        if G > 1 and exchange != "nccl":
            why = self._peer_unsupported()
            if why is None:
                try:
                    if peer is None:
                        from . import peer as _peer
                        peer = _peer.connect(self.comm, self.device, self.n_pad)
                    self._adopt_peer(peer)
                except Exception as e:  # noqa: BLE001 -- no symmetric memory on this system: keep the collectives
                    why = f"{type(e).__name__}: {e}"
            if self.peer is None:
                if exchange == "peer":
                    raise RuntimeError(f"exchange='peer' is not available: {why}")
                self.exchange_mode = f"nccl (peer exchange unavailable: {str(why)[:200]})"
        if G > 1 and self.peer is None:
            self.v_full = torch.zeros(self.n_pad, **f32)
            self.hv_full = torch.zeros(self.n_pad, **f32)
        self.j = 0
        self.launches = 0

    def _peer_unsupported(self) -> Optional[str]:
        """The fused exchange covers the default recurrences: no reorthogonalisation, or two-pass CGS with the fused
        middle pass at every depth.  Anything else keeps the torch.distributed collectives."""
        if getattr(self.comm, "backend", "nccl") == "gloo" or self.device.type != "cuda":
            return "needs CUDA devices"
        if not hasattr(self.ops, "x_reduce_scatter_dot"):
            return "kernel front end has no exchange-aware entry points"
        if self.reorth == "full" and not (self.fused and self.cgs_passes == 2 and self.m <= self.fused_max_rows):
            return "needs the fused two-pass Gram-Schmidt path at every depth"
        if self.reorth_tol is not None:
            return "the conditional last pass is not exchange-aware"
        if self.comm.world > 16:
            return "more than 16 ranks"
        return None

    def _adopt_peer(self, peer) -> None:
        if peer.world != self.comm.world or peer.rank != self.comm.rank or peer.hv_full.numel() < self.n_pad:
            raise ValueError("peer context does not match this engine (world, rank, vector length)")
        self.peer = peer
        self.hv_full = peer.hv_full[: self.n_pad]
        self.v_full = peer.v_full[: self.n_pad]
        # Measured defaults (profiles/r02_bench_n8_k20_exchange_*.json, r02_bench_n8_k20_ab2_*.json; knobs for A/B runs):
        #  * the reduce-scatter + alpha kernel and the in-kernel coefficient exchange are always used;
        #  * v_{j+1}: with ONE peer the stores from the normalise kernel (one CTA per SM) move 248 MB in 0.39 ms, NCCL's
        #    all-gather takes 0.58 ms -> stores at 2 ranks; with 7 peers every thread fans out to 7 destinations and the
        #    stores drop to 275 GB/s (1.58 ms; multimem.st 1.14 ms) against 0.71 ms for NCCL's NVLS all-gather -> NCCL
        #    from 3 ranks.  HLV_PEER_ALLGATHER=peer|nccl overrides;
        #  * HLV_MULTICAST=0/1 forces the in-switch add of the reduce-scatter off / on (default: on from 4 ranks -- with one
        #    peer there is nothing to reduce in the switch), HLV_MULTICAST_STORE the multimem.st variant of the stores.
        mc = os.environ.get("HLV_MULTICAST", "auto")
        self.multicast = bool(peer.hv_multicast and peer.v_multicast) and (mc == "1" or (mc == "auto" and peer.world >= 4))
        ms = os.environ.get("HLV_MULTICAST_STORE", "auto")     # multimem.st for the v stores, separately from the in-switch add
        self.multicast_store = bool(peer.v_multicast) and (ms == "1" or (ms == "auto" and self.multicast))
        if self.peer_allgather is None:
            self.peer_allgather = os.environ.get("HLV_PEER_ALLGATHER", "peer" if peer.world == 2 else "nccl") == "peer"
        self.exchange_mode = (("peer+multicast" if self.multicast else "peer")
                              + (("+multicast_stores" if self.multicast_store else "+peer_stores") if self.peer_allgather else "+nccl_allgather"))

    # -- vectors ---------------------------------------------------------------
    def _v_shard(self, j: int) -> torch.Tensor:
        return self.basis[j] if self.fp32_rows else self.ring[j & 1]

    def _v_for_hvp(self, j: int) -> torch.Tensor:
        v = self.v_full if self.comm.world > 1 else self._v_shard(j)
        v = v[: self.n]
        return v.unsqueeze(1) if self.column_vectors else v

    def start(self, v0: torch.Tensor, normalize: bool = False) -> None:
        """Install the start vector (global, length n).  The hand loop assumes ||v0|| = 1
        (lanczostrain_hand.py:162-163 normalises before the loop)."""
        v0 = v0.reshape(-1)
        if v0.numel() != self.n:
            raise ValueError(f"v0 has {v0.numel()} elements, operator dimension is {self.n}")
        v0 = v0.to(device=self.device, dtype=torch.float32)
        if normalize:
            v0 = v0 / torch.linalg.vector_norm(v0)
        self.j = 0
        if hasattr(self.hvp, "invalidate"):
            self.hvp.invalidate()               # a captured operator redoes its v-independent half for this run
        self.alphas.zero_(); self.betas.zero_(); self.breakdown_iter.fill_(-1)
        if self.reorth_tol is not None:
            self.pass_count.zero_()
        local = torch.zeros(self.shard_n, dtype=torch.float32, device=self.device)
        hi = min(self.lo + self.shard_n, self.n)
        if hi > self.lo:
            local[: hi - self.lo] = v0[self.lo: hi]
        if self.fp32_rows:
            self.basis[0].copy_(local)
        else:
            self.ring[0].copy_(local)
            if self.keep_basis:
                self.basis[0].copy_(local.to(self.basis_dtype))
        if self.comm.world > 1:
            if self.peer is not None:
                # run boundary: no rank may still be reading this rank's Hv (or writing its v) from the previous run
                torch.cuda.synchronize(self.device)
                self.comm.barrier()
                self._v_pending = False
            self.v_full.zero_()
            self.v_full[: self.n].copy_(v0)

    # -- one operator application: fills self.w (local shard of H v_j) and alphas[j] ------
    def _hvp_into(self, j: int, target: torch.Tensor, fused_dot: bool) -> bool:
        """target[:n] = (this rank's share of) H v_j.  Returns True when alphas[j] still has to be computed."""
        ops, n, ph = self.ops, self.n, self.phases
        v_in = self._v_for_hvp(j)
        v_sh = self._v_shard(j)
        a_out = self.alphas[j: j + 1]
        tgt = target[:n] if target.numel() != n else target
        ph.start("hvp")
        if hasattr(self.hvp, "accumulate_into"):
            # library operator: runs the double-backward per micro-batch and gathers the per-tensor
            # pieces straight into the flat vector (fused alpha on the last micro-batch)
            self.hvp.accumulate_into(v_in, tgt, dot_with=v_sh[:n] if fused_dot else None,
                                     dot_out=a_out if fused_dot else None, ws=self.ws, ops=ops, phases=ph)
            ph.stop("hvp")
            return not fused_dot
        r = self.hvp(v_in)
        ph.stop("hvp")
        if isinstance(r, (list, tuple)):
            ph.start("gather")
            ops.gather(list(r), tgt, dot_with=v_sh[:n] if fused_dot else None,
                       dot_out=a_out if fused_dot else None, ws=self.ws)
            ph.stop("gather")
            return not fused_dot
        r = r.detach().reshape(-1)
        if r.numel() != n:
            raise ValueError(f"hvp returned {r.numel()} elements, expected {n}")
        # always a copy into the engine's own buffer (4n bytes: 0.15 ms at GPT-2 size): an operator may return its
        # input, a view of it, or a buffer it keeps -- w is updated in place by every kernel that follows
        tgt.copy_(r)                                                 # also converts dtype / device (a closure that ends in .cpu())
        return True

    def _apply(self, j: int):
        """w = local shard of H v_j, alphas[j] = <w, v_j>: a generator (one yield per exchange point)."""
        ops, G, ph, comm = self.ops, self.comm.world, self.phases, self.comm
        v_sh = self._v_shard(j)
        a_out = self.alphas[j: j + 1]
        if self.peer is not None:
            # peer exchange: wait until every rank's shard of v_j has landed in v_full, apply, tell the ranks, then read
            # this rank's shard of every rank's Hv over NVLink (sum in rank order) with the alpha partial in the same pass
            if self._v_pending:
                ops.peer_wait(self.peer, _CH_V)
                self._v_pending = False
            self._hvp_into(j, self.hv_full, fused_dot=False)
            ops.peer_signal(self.peer, _CH_HV)
            yield
            ph.start("reduce_scatter_alpha")
            ops.x_reduce_scatter_dot(self.peer, self.peer.hv_ptrs, self.lo, self.w, v_sh, a_out, self.ws,
                                     hv_multicast=self.peer.hv_multicast if self.multicast else 0)
            ph.stop("reduce_scatter_alpha")
            yield
            return
        need_dot = self._hvp_into(j, self.w if G == 1 else self.hv_full, fused_dot=G == 1)
        if G > 1:
            ph.start("reduce_scatter")
            comm.reduce_scatter_sum(self.w, self.hv_full)
            ph.stop("reduce_scatter")
        if need_dot:
            ph.start("dot")
            ops.dot(self.w, v_sh, a_out, self.ws)
            ph.stop("dot")
            comm.all_reduce_sum(a_out)
        yield

    def step(self, j: Optional[int] = None, store_next: bool = True) -> None:
        """Iteration j of the hand loop (lanczostrain_hand.py:188-203 order):
        w = H v_j; alpha_j = w.v_j; w -= alpha_j v_j + beta_j v_{j-1}; [CGS2 vs rows 0..j];
        beta_{j+1} = ||w||; v_{j+1} = w / beta_{j+1}."""
        for _ in self.step_phases(j, store_next):
            pass

    def step_phases(self, j: Optional[int] = None, store_next: bool = True):
        """``step`` as a generator that yields after every launch whose result other ranks consume.  One rank per GPU
        simply exhausts it; a single-GPU emulation of several ranks (tests) advances all ranks' generators in
        lockstep, so that every push has been issued before the kernel that waits for it."""
        j = self.j if j is None else j
        ops, ph, comm, peer = self.ops, self.phases, self.comm, self.peer
        yield from self._apply(j)
        v_j = self._v_shard(j)
        v_jm1 = self._v_shard(j - 1) if j > 0 else None
        a_j = self.alphas[j: j + 1]
        b_j = self.betas[j: j + 1] if j > 0 else None
        x_path = peer is not None or self.fold_update          # hlv_x_* entry points (peer may be None: single GPU)
        rows = j + 1
        fused = self.reorth == "full" and self.fused and (peer is not None or self.fused_min_rows <= rows) and rows <= self.fused_max_rows
        norm_reduced = peer is not None
        if self.reorth != "full":
            ph.start("update")
            if x_path:
                ops.x_lanczos_update(peer, self.w, v_j, v_jm1, a_j, b_j, self.norm2, self.ws)
            else:
                ops.lanczos_update(self.w, v_j, v_jm1, a_j, b_j, self.norm2, self.ws)
            ph.stop("update")
            yield
        else:
            cur, nxt_c = self.coef, self.coef2
            if x_path:          # three-term update folded into the first projection: one launch, one pass
                ph.start("update_project")
                ops.x_update_project(peer, self.basis, rows, self.w, v_j, v_jm1, a_j, b_j, cur, self.ws)
                ph.stop("update_project", rows)
            else:
                ph.start("update")
                ops.lanczos_update(self.w, v_j, v_jm1, a_j, b_j, self.norm2, self.ws)
                ph.stop("update")
                ph.start("cgs_project")
                ops.cgs_project(self.basis, rows, self.w, cur, self.ws)
                ph.stop("cgs_project", rows)
            if peer is None:
                comm.all_reduce_sum(cur[:rows])
            yield
            for p in range(self.cgs_passes - 1):
                if fused:           # update with c_p and project for c_{p+1} in ONE pass over the basis
                    ph.start("cgs_update_project")
                    if peer is not None:
                        ops.x_cgs_update_project(peer, self.basis, rows, cur, self.w, nxt_c, self.norm2, self.ws)
                    else:
                        ops.cgs_update_project(self.basis, rows, cur, self.w, nxt_c, self.norm2, self.ws)
                    ph.stop("cgs_update_project", rows)
                else:
                    ph.start("cgs_update")
                    ops.cgs_update(self.basis, rows, cur, self.w, None, self.ws)
                    ph.stop("cgs_update", rows)
                    ph.start("cgs_project")
                    ops.cgs_project(self.basis, rows, self.w, nxt_c, self.ws)
                    ph.stop("cgs_project", rows)
                if peer is None:
                    comm.all_reduce_sum(nxt_c[:rows])
                cur, nxt_c = nxt_c, cur
                yield
            if self.reorth_tol is not None and fused:
                # cur = V w' and norm2 = |w'|^2 were measured by the fused pass: is w' orthogonal enough already?
                comm.all_reduce_sum(self.norm2)
                ops.cgs_needs_pass(cur, rows, self.norm2, self.reorth_tol, self.pass_flag)
                self.pass_count.add_(self.pass_flag)
                self.norm2_b.zero_()
                ph.start("cgs_update")
                ops.cgs_update(self.basis, rows, cur, self.w, self.norm2_b, self.ws, run_flag=self.pass_flag)
                ph.stop("cgs_update", rows)
                comm.all_reduce_sum(self.norm2_b)
                self.norm2.copy_(torch.where(self.pass_flag != 0, self.norm2_b, self.norm2))
                norm_reduced = True
            else:
                ph.start("cgs_update")
                if peer is not None:
                    ops.x_cgs_update(peer, self.basis, rows, cur, self.w, self.norm2, self.ws)
                else:
                    ops.cgs_update(self.basis, rows, cur, self.w, self.norm2, self.ws)
                ph.stop("cgs_update", rows)
            yield
        if not norm_reduced:
            comm.all_reduce_sum(self.norm2)
        if store_next:
            ph.start("normalize")
            nxt = j + 1
            if nxt < self.m:
                v_out = self.basis[nxt] if self.fp32_rows else self.ring[nxt & 1]
                row16 = self.basis[nxt] if (self.keep_basis and not self.fp32_rows) else None
            else:                       # last iteration: only beta_m (residual norm) is needed
                v_out, row16 = None, None
            if peer is not None:
                ops.x_normalize_store(peer, self.w, self.norm2, self.betas[nxt: nxt + 1], v_out, row16,
                                      peer.v_ptrs if self.peer_allgather else None, self.lo,
                                      self.breakdown_tol, self.breakdown_iter, j, self.ws,
                                      v_multicast=peer.v_multicast if (self.multicast_store and self.peer_allgather) else 0)
                self._v_pending = nxt < self.m and self.peer_allgather
            else:
                ops.normalize_store(self.w, self.norm2, self.betas[nxt: nxt + 1], v_out, row16,
                                    self.breakdown_tol, self.breakdown_iter, j)
            ph.stop("normalize")
            if (peer is None or not self.peer_allgather) and comm.world > 1 and nxt < self.m:
                ph.start("all_gather")
                comm.all_gather(self.v_full, v_out)
                ph.stop("all_gather")
        elif peer is not None:
            raise RuntimeError("store_next=False is not available with the peer exchange (the norm total is consumed by the normalise kernel)")
        self.j = j + 1
        yield

    # -- checkpoint / resume ---------------------------------------------------------
    def state_dict(self, include_basis: bool = True) -> Dict[str, Any]:
        """Everything needed to continue at iteration ``self.j``: (T so far, j, v_j, v_{j-1}[, rows
        0..j of the basis]) as CPU tensors (local shard when sharded).  The reference only ever
        WRITES T per iteration (diego_pythia.py:192) and cannot resume."""
        j = self.j
        sd: Dict[str, Any] = {"j": j, "n": self.n, "m": self.m, "reorth": self.reorth,
                              "basis_dtype": str(self.basis_dtype), "world": self.comm.world, "rank": self.comm.rank,
                              "alphas": self.alphas.cpu(), "betas": self.betas.cpu(),
                              "breakdown_iter": self.breakdown_iter.cpu()}
        if j < self.m:
            sd["v_cur"] = self._v_shard(j).float().cpu()
            if j > 0:
                sd["v_prev"] = self._v_shard(j - 1).float().cpu()
        if include_basis and self.keep_basis:
            sd["basis_rows"] = self.basis[: min(j + 1, self.m)].cpu()
        return sd

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        if (sd["n"], sd["m"], sd["world"], sd["rank"]) != (self.n, self.m, self.comm.world, self.comm.rank):
            raise ValueError("checkpoint does not match this engine (n, m, world, rank)")
        if (sd.get("reorth", self.reorth), sd.get("basis_dtype", str(self.basis_dtype))) != (self.reorth, str(self.basis_dtype)):
            raise ValueError("checkpoint does not match this engine (reorth, basis_dtype)")
        if self.keep_basis and "basis_rows" not in sd:
            raise ValueError("engine keeps a basis but the checkpoint has none")
        j = int(sd["j"])
        self.alphas.copy_(sd["alphas"]); self.betas.copy_(sd["betas"]); self.breakdown_iter.copy_(sd["breakdown_iter"])
        if self.keep_basis:
            rows = sd["basis_rows"]
            self.basis[: rows.shape[0]].copy_(rows)
        if not self.fp32_rows and j < self.m:
            self.ring[j & 1].copy_(sd["v_cur"])
            if j > 0:
                self.ring[(j - 1) & 1].copy_(sd["v_prev"])
        if self.comm.world > 1 and j < self.m:
            self.comm.all_gather(self.v_full, self._v_shard(j))
        self.j = j

    # -- host read-back -----------------------------------------------------------
    def broke_down(self) -> int:
        """Iteration index at which beta fell below breakdown_tol, or -1 (synchronises)."""
        return int(self.breakdown_iter.item())

    def tridiagonal(self, upto: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        """(alphas[:k], betas[:k]) computed so far as float64 numpy (synchronises)."""
        k = self.j if upto is None else upto
        a = self.alphas[:k].cpu().numpy().copy()
        b = self.betas[1: k + 1].cpu().numpy().copy()
        return a, b

    def result(self) -> LanczosResult:
        if self.peer is not None and self.peer.error():
            raise RuntimeError(f"peer exchange: a wait on channel {self.peer.error() - 1} ran into its time limit; results are invalid")
        bd = self.broke_down()
        m_eff = self.j if bd < 0 else min(self.j, bd + 1)
        a, b = self.tridiagonal(m_eff)
        # fp32 round trip: T is an fp32 matrix in the reference
        a32, b32 = a.astype(np.float32), b.astype(np.float32)
        eigvals, gammas, Y = _ritz.ritz_values(a32, b32)
        return LanczosResult(alphas=torch.from_numpy(a32.astype(np.float64)), betas=torch.from_numpy(b32.astype(np.float64)),
                             eigvals=eigvals, gammas=gammas, Y=Y, m=m_eff, n=self.n, breakdown=bd >= 0,
                             basis=self.basis if self.keep_basis else None, n_local=self.shard_n if self.comm.world > 1 else self.n,
                             timings=self.phases.summary(),
                             conditional_passes=int(self.pass_count.item()) if self.reorth_tol is not None else None)


def lanczos(hvp: Callable, n_iter: int, v0: torch.Tensor, reorth: Optional[str] = None, *,
            basis_dtype: torch.dtype = torch.float32, keep_basis: Optional[bool] = None,
            breakdown_tol: Optional[float] = None, check_every: int = 16, normalize_v0: bool = False,
            comm: Optional[Comm] = None, profile: bool = False, column_vectors: bool = False,
            on_iteration: Optional[Callable[[int, LanczosEngine], None]] = None, fused_cgs: bool = True,
            reorth_tol: Optional[float] = None) -> LanczosResult:
    """Run ``n_iter`` Lanczos iterations of the symmetric operator ``hvp`` from ``v0``.

    hvp: callable v[P] -> Hv.  It may return a flat [P] (or [P,1]) tensor, or -- to use the fused
         gather + alpha kernel -- the list of per-parameter pieces in ``model.parameters()``
         order (what ``torch.autograd.grad`` returns), or be a library operator
         (``hvp.HessianVectorProduct``).
    reorth: None = the reference hand loop (lanczostrain_hand.py:171-203);
            'full' = hand-loop order + two-pass classical Gram-Schmidt against all stored rows.
    reorth_tol: None = both passes always; a float = apply the second pass's update only when some projection
            coefficient of the once-orthogonalised w exceeds reorth_tol * |w| (gpytorch's rule with tol=1e-5),
            decided and predicated on the device.  Measured at GPT-2 size, m=100: tol=1e-5 never applies it and the
            Ritz values stay within 1.4e-6*|T| of a float64 recurrence; orthogonality of the stored rows under the
            option is only verified at test size, so keep the default when the Ritz vectors matter.
    v0 must have unit norm (pass normalize_v0=True otherwise), as in the reference.  Normalise it on the device
    (probe_vector / normalize_v0): torch's CPU float32 norm is 1.4% off at 1.2e8 elements.
    """
    dev = v0.device
    eng = LanczosEngine(hvp, v0.numel(), n_iter, dev, reorth=reorth, basis_dtype=basis_dtype,
                        keep_basis=keep_basis, breakdown_tol=breakdown_tol, comm=comm,
                        profile=profile, column_vectors=column_vectors, fused_cgs=fused_cgs, reorth_tol=reorth_tol)
    eng.start(v0, normalize=normalize_v0)
    for j in range(n_iter):
        eng.step(j)
        if on_iteration is not None:
            on_iteration(j, eng)
        if eng.breakdown_tol > 0 and check_every > 0 and (j + 1) % check_every == 0 and j + 1 < n_iter:
            if eng.broke_down() >= 0:
                break
    return eng.result()


def lanczos_tridiag(matmul_closure: Callable, max_iter: int, dtype=torch.float32, device="cuda",
                    matrix_shape=None, batch_shape=None, init_vecs: Optional[torch.Tensor] = None,
                    num_init_vecs: int = 1, tol: float = 1e-5, **engine_kwargs):
    """Drop-in for ``gpytorch.utils.lanczos.lanczos_tridiag`` as the reference calls it
    (gpt2_hessian_cpu.py:207-213, gpt2_savehessian.py:202-208): returns ``(Q[P,m], T[m,m])`` with
    full reorthogonalisation, calling ``matmul_closure`` with [P,1] column vectors.

    Differences, all documented in INTEGRATION.md: the recurrence always runs on the current
    CUDA device (``device='cpu'`` only moves the RESULTS to the host, as gpt2_hessian_cpu.py
    expects); a single probe vector (num_init_vecs=1, no batch_shape) is supported; ``init_vecs``
    defaults to randn(P,1) and is normalised, exactly like gpytorch -- the reference's
    ``init_vec`` swap inside CurvVecProduct (quirk F3) is therefore not needed: pass init_vecs.
    ``tol`` is accepted for signature compatibility: by default BOTH Gram-Schmidt passes always run, which is
    at least as strong as gpytorch's "re-orthogonalise while any q_i . r > tol"; pass ``reorth_tol=tol`` to get
    the conditional rule itself (checked in tests/ against a restatement of SURVEY Appendix B -- parity
    unpinned at that boundary).  Like gpytorch, at most ``matrix_shape[-1]`` iterations run."""
    if dtype != torch.float32:
        raise NotImplementedError("lanczos_tridiag: only float32 (the reference's dtype) is supported")
    if num_init_vecs != 1 or (batch_shape is not None and len(tuple(batch_shape)) > 0):
        raise NotImplementedError("lanczos_tridiag: a single probe vector is supported")
    if matrix_shape is None:
        raise ValueError("matrix_shape is required")
    P = int(tuple(matrix_shape)[-1])
    if init_vecs is None and getattr(matmul_closure, "init_vec", None) is not None:
        init_vecs = matmul_closure.init_vec          # CurvVecProduct(loader, model, init_vec=...) -- see hvp.py
    cuda_dev = _kernels.compute_device(device)       # device='cpu' (gpt2_hessian_cpu.py:209) only moves the RESULTS
    if init_vecs is None:
        init_vecs = torch.randn(P, 1, dtype=torch.float32, device=cuda_dev)
    res = lanczos(matmul_closure, min(int(max_iter), P), init_vecs.reshape(-1).to(cuda_dev), reorth="full",
                  normalize_v0=True, column_vectors=True, breakdown_tol=1e-6, **engine_kwargs)
    Q = res.Q.t()                       # [P, m] view of the row-major device basis
    T = res.T
    out_dev = torch.device(device)
    return Q.to(out_dev), T.to(out_dev)
