"""fp32-accurate GEMMs on the tensor cores for the torch HVP ("3xTF32" error-compensated split).

The HVP stays a torch double-backward, but on B200 its fp32 GEMMs run on the SIMT pipe
(`cutlass3x_sm100_simt_sgemm`: 64% of the HVP's GPU time, profiles/r01_hvp_torch_profile_B8.txt)
because plain TF32 is too coarse for Lanczos (9e-4 deviation in Hv; the parity bar is 1e-5).
The classic remedy keeps fp32 accuracy on the TF32 tensor-core path:

    a = a_hi + a_lo,  a_hi = a with the low 13 mantissa bits cleared (exactly TF32-representable)
    a @ b  ~=  a_lo @ b_hi  +  a_hi @ b_lo  +  a_hi @ b_hi          (a_lo @ b_lo ~ 2^-22 dropped)

with fp32 accumulation inside each GEMM: per-product relative error ~2^-21, i.e. below the
sqrt(K)*2^-24 accumulation noise any fp32 GEMM already has.

``enable()`` overrides the CUDA kernels of aten::mm / addmm / bmm / baddbmm (the functional
overloads only -- the ``.out`` overloads keep the stock cuBLAS kernels and are what the override
itself calls, with TF32 switched on just around those calls).  Because the override sits at the
backend dispatch key, every GEMM autograd issues in the first AND second backward goes through it.
Opt-in; the reference arm never enables it.
"""
from __future__ import annotations

import contextlib

import torch

_LIB = None
_MASK = -8192            # 0xFFFFE000: keep sign, exponent and the top 10 mantissa bits
calls = 0


def _split(x: torch.Tensor):
    hi = (x.view(torch.int32) & _MASK).view(torch.float32)
    return hi, x - hi


def _eligible(*ts) -> bool:
    return all(t.dtype == torch.float32 for t in ts)


def _tf32(on: bool):
    torch.backends.cuda.matmul.allow_tf32 = on


def _mm(a, b):
    global calls
    if not _eligible(a, b):
        return torch.ops.aten.mm.out(a, b, out=a.new_empty((a.shape[0], b.shape[1])))
    calls += 1
    a_hi, a_lo = _split(a)
    b_hi, b_lo = _split(b)
    out = a.new_empty((a.shape[0], b.shape[1]))
    _tf32(True)
    try:
        torch.ops.aten.mm.out(a_lo, b_hi, out=out)
        torch.ops.aten.addmm.out(out, a_hi, b_lo, out=out)
        torch.ops.aten.addmm.out(out, a_hi, b_hi, out=out)
    finally:
        _tf32(False)
    return out


def _addmm(bias, a, b, *, beta=1, alpha=1):
    global calls
    if not _eligible(bias, a, b) or alpha != 1:
        out = a.new_empty((a.shape[0], b.shape[1]))
        return torch.ops.aten.addmm.out(bias, a, b, beta=beta, alpha=alpha, out=out)
    calls += 1
    a_hi, a_lo = _split(a)
    b_hi, b_lo = _split(b)
    out = a.new_empty((a.shape[0], b.shape[1]))
    _tf32(True)
    try:
        torch.ops.aten.addmm.out(bias, a_lo, b_hi, beta=beta, out=out)
        torch.ops.aten.addmm.out(out, a_hi, b_lo, out=out)
        torch.ops.aten.addmm.out(out, a_hi, b_hi, out=out)
    finally:
        _tf32(False)
    return out


def _bmm(a, b):
    global calls
    if not _eligible(a, b):
        return torch.ops.aten.bmm.out(a, b, out=a.new_empty((a.shape[0], a.shape[1], b.shape[2])))
    calls += 1
    a_hi, a_lo = _split(a)
    b_hi, b_lo = _split(b)
    out = a.new_empty((a.shape[0], a.shape[1], b.shape[2]))
    _tf32(True)
    try:
        torch.ops.aten.bmm.out(a_lo, b_hi, out=out)
        torch.ops.aten.baddbmm.out(out, a_hi, b_lo, out=out)
        torch.ops.aten.baddbmm.out(out, a_hi, b_hi, out=out)
    finally:
        _tf32(False)
    return out


def _baddbmm(bias, a, b, *, beta=1, alpha=1):
    global calls
    out = a.new_empty((a.shape[0], a.shape[1], b.shape[2]))
    if not _eligible(bias, a, b) or alpha != 1:
        return torch.ops.aten.baddbmm.out(bias, a, b, beta=beta, alpha=alpha, out=out)
    calls += 1
    a_hi, a_lo = _split(a)
    b_hi, b_lo = _split(b)
    _tf32(True)
    try:
        torch.ops.aten.baddbmm.out(bias, a_lo, b_hi, beta=beta, out=out)
        torch.ops.aten.baddbmm.out(out, a_hi, b_lo, out=out)
        torch.ops.aten.baddbmm.out(out, a_hi, b_hi, out=out)
    finally:
        _tf32(False)
    return out


def enable() -> None:
    """Route fp32 CUDA GEMMs of this process through the 3xTF32 split (idempotent)."""
    global _LIB
    if _LIB is not None:
        return
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lib = torch.library.Library("aten", "IMPL")
        lib.impl("mm", _mm, "CUDA")
        lib.impl("addmm", _addmm, "CUDA")
        lib.impl("bmm", _bmm, "CUDA")
        lib.impl("baddbmm", _baddbmm, "CUDA")
    _LIB = lib


def disable() -> None:
    global _LIB
    if _LIB is not None:
        _LIB._destroy()
        _LIB = None


@contextlib.contextmanager
def fp32_tensor_core_gemms():
    enable()
    try:
        yield
    finally:
        disable()
