"""TEST INFRASTRUCTURE ONLY.  Runs the reference's own native source
(/root/reference/vector_adjust.cu), built by oracle/Makefile into oracle/_ref/:
  * on host cores, through the g++/shim build (``vector_adjust_cpu``);
  * on the GPU, the nvcc-built cubin launched with the reference's own geometry
    (block 256, grid ceil(n/256): gpt_hessian_cuda.py:38-52) through the CUDA driver API
    (``vector_adjust_gpu``; cuda-python).
The kernel is O(k*n^2): keep n at a few thousand.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
CPU_LIB = os.path.join(_REF_DIR, "libvector_adjust_ref.so")
CUBIN = os.path.join(_REF_DIR, "vector_adjust.sm_100a.cubin")


def have_cpu_ref() -> bool:
    return os.path.exists(CPU_LIB)


def have_cubin() -> bool:
    return os.path.exists(CUBIN)


def vector_adjust_cpu(grad: np.ndarray, V: np.ndarray, eigvals: np.ndarray, adjusted: np.ndarray,
                      delta: float, block: int = 256) -> np.ndarray:
    """adjusted += reference kernel result (in place, like the kernel); float32 C-contiguous arrays."""
    lib = C.CDLL(CPU_LIB)
    fp = C.POINTER(C.c_float)
    lib.ref_vector_adjust_cpu.argtypes = [fp, fp, fp, fp, C.c_int, C.c_int, C.c_float, C.c_int]
    lib.ref_vector_adjust_cpu.restype = None
    for a in (grad, V, eigvals, adjusted):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    k, n = V.shape
    lib.ref_vector_adjust_cpu(grad.ctypes.data_as(fp), V.ctypes.data_as(fp), eigvals.ctypes.data_as(fp),
                              adjusted.ctypes.data_as(fp), int(k), int(n), float(delta), int(block))
    return adjusted


def vector_adjust_gpu(grad, V, eigvals, adjusted, delta: float, block: int = 256):
    """Launch the reference cubin on torch CUDA tensors (in place on ``adjusted``)."""
    import torch
    from cuda.bindings import driver as cu

    def ok(res):
        err = res[0]
        if int(err) != 0:
            raise RuntimeError(f"CUDA driver error {err}")
        return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)

    torch.cuda.synchronize()
    data = open(CUBIN, "rb").read()
    mod = ok(cu.cuModuleLoadData(data))
    fn = ok(cu.cuModuleGetFunction(mod, b"vector_adjust"))
    k, n = V.shape
    args = [np.array([grad.data_ptr()], dtype=np.uint64), np.array([V.data_ptr()], dtype=np.uint64),
            np.array([eigvals.data_ptr()], dtype=np.uint64), np.array([adjusted.data_ptr()], dtype=np.uint64),
            np.array([k], dtype=np.int32), np.array([n], dtype=np.int32), np.array([delta], dtype=np.float32)]
    argv = np.array([a.ctypes.data for a in args], dtype=np.uint64)
    grid = (n + block - 1) // block
    stream = torch.cuda.current_stream().cuda_stream
    ok(cu.cuLaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, stream, argv.ctypes.data, 0))
    torch.cuda.synchronize()
    ok(cu.cuModuleUnload(mod))
    return adjusted
