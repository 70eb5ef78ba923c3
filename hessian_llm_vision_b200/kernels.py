"""Torch-tensor front end of the libhlv C ABI (include/hlv.h).

Every function takes CUDA tensors, passes raw device pointers plus torch's
current stream to the C entry point, and returns without synchronising.
Scalars (dot products, squared norms, Gram-Schmidt coefficients) live in
caller-provided float64 CUDA tensors.  There is no CPU path: tensors that are
not on a CUDA device are rejected.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import HLVError  # noqa: F401

#: how many times any libhlv kernel entry point has been called in this process
launch_count = 0


def require_device(device) -> None:
    """The recurrence runs on a CUDA device or not at all."""
    if torch.device(device).type != "cuda":
        raise RuntimeError(f"lanczos: vectors must live on a CUDA device (got {device}); this engine has no CPU path")


def compute_device(requested) -> torch.device:
    """Where the recurrence runs when a caller names ``requested`` as the place it wants the results."""
    d = torch.device(requested)
    return d if d.type == "cuda" else torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cuda(t: torch.Tensor, dtype, name: str) -> int:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor (libhlv has no CPU path), got "
                        f"{type(t).__name__} on {getattr(t, 'device', None)}")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t.data_ptr()


def _opt(t: Optional[torch.Tensor], dtype, name: str) -> Optional[int]:
    return None if t is None else _cuda(t, dtype, name)


def _basis(V: torch.Tensor, rows: int, n: int, name: str):
    """Row-major 2-D basis (or 1-D when rows == 1); returns (ptr, ldv, suffix)."""
    if not V.is_cuda:
        raise TypeError(f"{name}: basis must be a CUDA tensor")
    if V.dtype == torch.float32:
        sfx = "f32"
    elif V.dtype == torch.bfloat16:
        sfx = "bf16"
    else:
        raise TypeError(f"{name}: basis dtype must be float32 or bfloat16, got {V.dtype}")
    if V.dim() == 1:
        V = V.unsqueeze(0)
    if V.dim() != 2 or V.stride(1) != 1:
        raise ValueError(f"{name}: basis must be 2-D with unit stride along the parameter dimension")
    if V.shape[0] < rows or V.shape[1] < n:
        raise ValueError(f"{name}: basis shape {tuple(V.shape)} smaller than rows={rows}, n={n}")
    # a single row has no meaningful pitch: hand the library any aligned value >= n
    ldv = V.stride(0) if V.shape[0] > 1 else (max(V.shape[1], n) + 7) // 8 * 8
    return V.data_ptr(), ldv, sfx


class Workspace:
    """Scratch for the cross-CTA reductions of one stream (hlv.h: `ws`)."""

    def __init__(self, device, max_rows: int = 128):
        lib = _lib.load()
        self.max_rows = int(max_rows)
        self.nbytes = int(lib.hlv_workspace_bytes(self.max_rows))
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(self.buf.device):
            _lib.call("hlv_workspace_init", self.buf.data_ptr(), self.nbytes, _stream())

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()


def device_info():
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    _lib.call("hlv_device_info", C.byref(sm), C.byref(maj), C.byref(mnr))
    return sm.value, maj.value, mnr.value


class TensorList:
    """Host-side pointer/size table of a per-tensor list (reused across calls when the
    tensors are static, e.g. CUDA-graph outputs)."""

    def __init__(self, tensors: Sequence[torch.Tensor]):
        self.n_tensors = len(tensors)
        ptrs, numels = [], []
        for i, t in enumerate(tensors):
            _cuda(t, torch.float32, f"tensor[{i}]")
            ptrs.append(t.data_ptr())
            numels.append(t.numel())
        self.total = sum(numels)
        self.ptrs = (C.c_void_p * max(1, self.n_tensors))(*ptrs)
        self.numels = (C.c_int64 * max(1, self.n_tensors))(*numels)
        self._keep = list(tensors)


def gather(tensors, dst: torch.Tensor, *, scale: float = 1.0, accumulate: bool = False,
           dot_with: Optional[torch.Tensor] = None, dot_out: Optional[torch.Tensor] = None,
           ws: Optional[Workspace] = None) -> None:
    """dst = cat(tensors) (bit-exact), or dst (+)= scale*cat(tensors); optional fused
    dot_out[0] = <dst_new, dot_with>.  Replaces torch.cat (gpt2_hessian_cpu.py:109,200)."""
    global launch_count
    tl = tensors if isinstance(tensors, TensorList) else TensorList(tensors)
    d = _cuda(dst, torch.float32, "dst")
    v = _opt(dot_with, torch.float32, "dot_with")
    o = _opt(dot_out, torch.float64, "dot_out")
    if (v is None) != (o is None):
        raise ValueError("gather: dot_with and dot_out go together")
    if v is not None and ws is None:
        raise ValueError("gather: fused dot needs a Workspace")
    with torch.cuda.device(dst.device):
        _lib.call("hlv_gather_f32", tl.ptrs, tl.numels, tl.n_tensors, d, dst.numel(), float(scale),
                  int(bool(accumulate)), v, o, ws.ptr if ws else None, ws.nbytes if ws else 0, _stream())
    launch_count += 1


def scatter(src: torch.Tensor, tensors) -> None:
    """tensors[t][...] = src[off_t : off_t + numel_t] (bit-exact; gpt2_hessian_cpu.py:79-82, :231-233)."""
    global launch_count
    tl = tensors if isinstance(tensors, TensorList) else TensorList(tensors)
    s = _cuda(src, torch.float32, "src")
    with torch.cuda.device(src.device):
        _lib.call("hlv_scatter_f32", s, src.numel(), tl.ptrs, tl.numels, tl.n_tensors, _stream())
    launch_count += 1


def dot(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, ws: Workspace) -> None:
    global launch_count
    n = a.numel()
    if b.numel() != n:
        raise ValueError("dot: length mismatch")
    with torch.cuda.device(a.device):
        _lib.call("hlv_dot_f32", _cuda(a, torch.float32, "a"), _cuda(b, torch.float32, "b"), n,
                  _cuda(out, torch.float64, "out"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def lanczos_update(w: torch.Tensor, vj: torch.Tensor, vjm1: Optional[torch.Tensor], alpha: torch.Tensor,
                   beta: Optional[torch.Tensor], norm2_out: torch.Tensor, ws: Workspace) -> None:
    """w -= alpha*vj + beta*vjm1 ; norm2_out = sum w^2  (lanczostrain_hand.py:202 + :190)."""
    global launch_count
    n = w.numel()
    if vj.numel() != n or (vjm1 is not None and vjm1.numel() != n):
        raise ValueError("lanczos_update: length mismatch")
    with torch.cuda.device(w.device):
        _lib.call("hlv_lanczos_update_f32", _cuda(w, torch.float32, "w"), _cuda(vj, torch.float32, "vj"),
                  _opt(vjm1, torch.float32, "vjm1"), _cuda(alpha, torch.float64, "alpha"),
                  _opt(beta, torch.float64, "beta"), n, _cuda(norm2_out, torch.float64, "norm2_out"),
                  ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def normalize_store(w: torch.Tensor, norm2: torch.Tensor, beta_out: torch.Tensor,
                    v_out: Optional[torch.Tensor], row_bf16: Optional[torch.Tensor] = None,
                    breakdown_tol: float = 0.0, breakdown_iter: Optional[torch.Tensor] = None,
                    it: int = 0) -> None:
    """beta = sqrt(norm2); v_out = w / beta [; row_bf16 = bf16(v_out)]  (lanczostrain_hand.py:190-194)."""
    global launch_count
    n = w.numel()
    with torch.cuda.device(w.device):
        _lib.call("hlv_normalize_store_f32", _cuda(w, torch.float32, "w"), _cuda(norm2, torch.float64, "norm2"), n,
                  _cuda(beta_out, torch.float64, "beta_out"), _opt(v_out, torch.float32, "v_out"),
                  _opt(row_bf16, torch.bfloat16, "row_bf16"), float(breakdown_tol),
                  _opt(breakdown_iter, torch.int32, "breakdown_iter"), int(it), _stream())
    launch_count += 1


def cgs_project(V: torch.Tensor, rows: int, w: torch.Tensor, c_out: torch.Tensor, ws: Workspace) -> None:
    """c_out[:rows] = V[:rows, :n] @ w  -- one streaming pass over the basis."""
    global launch_count
    n = w.numel()
    p, ldv, sfx = _basis(V, rows, n, "cgs_project")
    if c_out.numel() < rows:
        raise ValueError("cgs_project: c_out too small")
    with torch.cuda.device(w.device):
        _lib.call(f"hlv_cgs_project_{sfx}", p, ldv, int(rows), _cuda(w, torch.float32, "w"), n,
                  _cuda(c_out, torch.float64, "c_out"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def cgs_update(V: torch.Tensor, rows: int, c: torch.Tensor, w: torch.Tensor,
               norm2_out: Optional[torch.Tensor], ws: Workspace, sign: float = -1.0,
               run_flag: Optional[torch.Tensor] = None) -> None:
    """w += sign * V[:rows, :n]^T c ; norm2_out = sum w^2.  With ``run_flag`` (device int32[1]) the pass is
    predicated on the device: flag 0 = leave w and norm2_out untouched (no host round trip)."""
    global launch_count
    n = w.numel()
    p, ldv, sfx = _basis(V, rows, n, "cgs_update")
    with torch.cuda.device(w.device):
        if run_flag is None:
            _lib.call(f"hlv_cgs_update_{sfx}", p, ldv, int(rows), _cuda(c, torch.float64, "c"), float(sign),
                      _cuda(w, torch.float32, "w"), n, _opt(norm2_out, torch.float64, "norm2_out"),
                      ws.ptr, ws.nbytes, _stream())
        else:
            _lib.call(f"hlv_cgs_update_if_{sfx}", p, ldv, int(rows), _cuda(c, torch.float64, "c"), float(sign),
                      _cuda(w, torch.float32, "w"), n, _opt(norm2_out, torch.float64, "norm2_out"),
                      _cuda(run_flag, torch.int32, "run_flag"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def cgs_needs_pass(c: torch.Tensor, rows: int, norm2: torch.Tensor, tol: float, flag_out: torch.Tensor) -> None:
    """flag_out[0] = 1 if some |c[i]| > tol * sqrt(norm2[0]) (another Gram-Schmidt pass is needed), else 0 --
    gpytorch's "while any q_i . r > tol" test, evaluated on the device."""
    global launch_count
    with torch.cuda.device(c.device):
        _lib.call("hlv_cgs_needs_pass", _cuda(c, torch.float64, "c"), int(rows), _cuda(norm2, torch.float64, "norm2"),
                  float(tol), _cuda(flag_out, torch.int32, "flag_out"), _stream())
    launch_count += 1


def fused_max_rows(dtype: torch.dtype) -> int:
    """Largest row count the fused update+project kernel can stage in shared memory."""
    return int(_lib.load().hlv_cgs_fused_max_rows(4 if dtype == torch.float32 else 2))


def cgs_update_project(V: torch.Tensor, rows: int, c_in: torch.Tensor, w: torch.Tensor, c_out: torch.Tensor,
                       norm2_out: torch.Tensor, ws: Workspace) -> None:
    """Middle pass of CGS2 in ONE read of the basis: w -= V^T c_in ; c_out = V w ; norm2_out = |w|^2
    (TMA-staged shared-memory slab).  ws must have been created with max_rows >= rows + 1."""
    global launch_count
    n = w.numel()
    p, ldv, sfx = _basis(V, rows, n, "cgs_update_project")
    if c_out.numel() < rows:
        raise ValueError("cgs_update_project: c_out too small")
    with torch.cuda.device(w.device):
        _lib.call(f"hlv_cgs_update_project_{sfx}", p, ldv, int(rows), _cuda(c_in, torch.float64, "c_in"),
                  _cuda(w, torch.float32, "w"), n, _cuda(c_out, torch.float64, "c_out"),
                  _cuda(norm2_out, torch.float64, "norm2_out"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def vector_adjust(grad_vector: torch.Tensor, V: torch.Tensor, eigvals: torch.Tensor,
                  adjusted_grad_vector: torch.Tensor, delta: float, ws: Workspace,
                  coef_scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Drop-in for the reference's cuda_vector_adjust (gpt_hessian_cuda.py:27-54 /
    vector_adjust.cu:2-15): adjusted += sum_i (1/lam_i - 1/(lam_i+delta)) (g.V_i) V_i, in place."""
    global launch_count
    k = eigvals.numel()
    n = grad_vector.numel()
    p, ldv, sfx = _basis(V, k, n, "vector_adjust")
    if sfx != "f32":
        raise TypeError("vector_adjust: V must be float32")
    if coef_scratch is None:
        coef_scratch = torch.empty(k, dtype=torch.float64, device=grad_vector.device)
    with torch.cuda.device(grad_vector.device):
        _lib.call("hlv_vector_adjust_f32", _cuda(grad_vector, torch.float32, "grad_vector"), p,
                  _cuda(eigvals, torch.float32, "eigvals"),
                  _cuda(adjusted_grad_vector, torch.float32, "adjusted_grad_vector"), int(k), n, float(delta),
                  ldv, _cuda(coef_scratch, torch.float64, "coef_scratch"), ws.ptr, ws.nbytes, _stream())
    launch_count += 3
    return adjusted_grad_vector


def ritz_vectors(Q: torch.Tensor, m: int, Y: torch.Tensor, out: torch.Tensor, n: int) -> None:
    """out[r, :n] = sum_i Y[i, r] * Q[i, :n]   (V = eigvects.t() @ Q, lanczostrain_hand.py:210)."""
    global launch_count
    p, ldq, sfx = _basis(Q, m, n, "ritz_vectors")
    if Y.dim() != 2 or Y.shape[0] < m or Y.stride(1) != 1:
        raise ValueError("ritz_vectors: Y must be [m, nvec] row-major")
    nvec = Y.shape[1]
    if out.dim() != 2 or out.shape[0] < nvec or out.shape[1] < n or out.stride(1) != 1:
        raise ValueError("ritz_vectors: out must be [nvec, >=n]")
    with torch.cuda.device(out.device):
        _lib.call(f"hlv_ritz_vectors_{sfx}", p, ldq, int(m), _cuda(Y, torch.float32, "Y"), Y.stride(0), int(nvec),
                  out.data_ptr(), out.stride(0) if out.shape[0] > 1 else (max(out.shape[1], n) + 7) // 8 * 8, int(n), _stream())
    launch_count += (nvec + 7) // 8


# ---------------------------------------------------------------------------------------------------------------
# exchange-aware entry points (hlv_x_*): the same kernels with the inter-rank exchange fused in.  ``peer`` is a
# peer.PeerContext or None (single GPU: alpha / coefficients are read from the given device buffers as usual).
# ---------------------------------------------------------------------------------------------------------------
def _ctx(peer):
    return None if peer is None else C.byref(peer.ctx)


def peer_xchg_bytes() -> int:
    return int(_lib.load().hlv_peer_xchg_bytes())


def peer_xchg_init(xchg: torch.Tensor) -> None:
    with torch.cuda.device(xchg.device):
        _lib.call("hlv_peer_xchg_init", _cuda(xchg, torch.uint8, "xchg"), _stream())


def peer_error(peer) -> int:
    """0, or 1 + channel of a wait that ran into its time limit (synchronises)."""
    err = C.c_int(0)
    with torch.cuda.device(peer.xchg.device):
        _lib.call("hlv_peer_xchg_error", peer.xchg.data_ptr(), C.byref(err), _stream())
    return err.value


def peer_signal(peer, channel: int) -> None:
    global launch_count
    with torch.cuda.device(peer.xchg.device):
        _lib.call("hlv_peer_signal", _ctx(peer), int(channel), _stream())
    launch_count += 1


def peer_wait(peer, channel: int) -> None:
    global launch_count
    with torch.cuda.device(peer.xchg.device):
        _lib.call("hlv_peer_wait", _ctx(peer), int(channel), _stream())
    launch_count += 1


def x_reduce_scatter_dot(peer, hv_ptrs, shard_lo: int, w: torch.Tensor, v: torch.Tensor, alpha_out: torch.Tensor,
                         ws: Workspace, hv_multicast: int = 0) -> None:
    """w = sum over ranks (rank order) of Hv_p[shard_lo : shard_lo + len(w)] read through peer memory;
    alpha_out = <w, v> partial, pushed to every rank.  hv_ptrs: ctypes array of the ranks' full-length Hv buffers
    as mapped in this process (or a 1-element array with the local pointer when peer is None)."""
    global launch_count
    n = w.numel()
    with torch.cuda.device(w.device):
        _lib.call("hlv_x_reduce_scatter_dot_f32", _ctx(peer), hv_ptrs, hv_multicast or None, int(shard_lo), n, _cuda(w, torch.float32, "w"),
                  _cuda(v, torch.float32, "v"), _cuda(alpha_out, torch.float64, "alpha_out"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def x_update_project(peer, V: torch.Tensor, rows: int, w: torch.Tensor, vj: torch.Tensor, vjm1: Optional[torch.Tensor],
                     alpha: torch.Tensor, beta: Optional[torch.Tensor], c_out: torch.Tensor, ws: Workspace) -> None:
    """w -= alpha*vj + beta*vjm1 (lanczostrain_hand.py:202, torch's rounding) folded into the first projection
    c_out = V[:rows] w_new; with peers alpha is the rank-ordered total and c_out is pushed."""
    global launch_count
    n = w.numel()
    p, ldv, sfx = _basis(V, rows, n, "x_update_project")
    if vj.numel() != n or (vjm1 is not None and vjm1.numel() != n):
        raise ValueError("x_update_project: length mismatch")
    with torch.cuda.device(w.device):
        _lib.call(f"hlv_x_update_project_{sfx}", _ctx(peer), p, ldv, int(rows), _cuda(w, torch.float32, "w"), n,
                  _cuda(vj, torch.float32, "vj"), _opt(vjm1, torch.float32, "vjm1"), _cuda(alpha, torch.float64, "alpha"),
                  _opt(beta, torch.float64, "beta"), _cuda(c_out, torch.float64, "c_out"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def x_lanczos_update(peer, w, vj, vjm1, alpha, beta, norm2_out, ws: Workspace) -> None:
    global launch_count
    n = w.numel()
    with torch.cuda.device(w.device):
        _lib.call("hlv_x_lanczos_update_f32", _ctx(peer), _cuda(w, torch.float32, "w"), _cuda(vj, torch.float32, "vj"),
                  _opt(vjm1, torch.float32, "vjm1"), _cuda(alpha, torch.float64, "alpha"), _opt(beta, torch.float64, "beta"), n,
                  _cuda(norm2_out, torch.float64, "norm2_out"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def x_cgs_update_project(peer, V, rows: int, c_in, w, c_out, norm2_out, ws: Workspace) -> None:
    global launch_count
    n = w.numel()
    p, ldv, sfx = _basis(V, rows, n, "x_cgs_update_project")
    with torch.cuda.device(w.device):
        _lib.call(f"hlv_x_cgs_update_project_{sfx}", _ctx(peer), p, ldv, int(rows), _cuda(c_in, torch.float64, "c_in"),
                  _cuda(w, torch.float32, "w"), n, _cuda(c_out, torch.float64, "c_out"),
                  _cuda(norm2_out, torch.float64, "norm2_out"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def x_cgs_update(peer, V, rows: int, c, w, norm2_out, ws: Workspace) -> None:
    global launch_count
    n = w.numel()
    p, ldv, sfx = _basis(V, rows, n, "x_cgs_update")
    with torch.cuda.device(w.device):
        _lib.call(f"hlv_x_cgs_update_{sfx}", _ctx(peer), p, ldv, int(rows), _cuda(c, torch.float64, "c"),
                  _cuda(w, torch.float32, "w"), n, _cuda(norm2_out, torch.float64, "norm2_out"), ws.ptr, ws.nbytes, _stream())
    launch_count += 1


def x_normalize_store(peer, w, norm2, beta_out, v_out, row_bf16, v_full_ptrs, shard_lo: int, breakdown_tol: float,
                      breakdown_iter, it: int, ws: Workspace, v_multicast: int = 0) -> None:
    """normalize_store with |w|^2 = the ranks' total; the normalised shard also goes straight into every rank's
    full-length vector (v_full_ptrs) and HLV_CH_V is raised when it is on its way."""
    global launch_count
    n = w.numel()
    with torch.cuda.device(w.device):
        _lib.call("hlv_x_normalize_store_f32", _ctx(peer), _cuda(w, torch.float32, "w"), _cuda(norm2, torch.float64, "norm2"), n,
                  _cuda(beta_out, torch.float64, "beta_out"), _opt(v_out, torch.float32, "v_out"),
                  _opt(row_bf16, torch.bfloat16, "row_bf16"), v_full_ptrs, v_multicast or None, int(shard_lo), float(breakdown_tol),
                  _opt(breakdown_iter, torch.int32, "breakdown_iter"), int(it), ws.ptr, ws.nbytes, _stream())
    launch_count += 1
