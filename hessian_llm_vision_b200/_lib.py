"""ctypes binding of libhlv.so (include/hlv.h).  No CPU fallback: a missing or
unloadable library raises at first use, loudly."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhlv.so")

HLV_OK = 0
HLV_VERSION = 200          # must equal include/hlv.h's HLV_VERSION: a stale libhlv.so is refused at load
HLV_MAX_ROWS = 1024


class HLVError(RuntimeError):
    """Raised when a libhlv entry point returns a non-zero status."""

    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with status {code}: {msg}")
        self.fn, self.code, self.msg = fn, code, msg


HLV_MAX_PEERS = 16
CH_HV, CH_ALPHA, CH_C1, CH_C2, CH_NORM, CH_V = 0, 1, 2, 3, 4, 5


class PeerCtx(C.Structure):
    """hlv_peer_ctx (include/hlv.h)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("spin_timeout_ms", C.c_uint32), ("reserved", C.c_uint32),
                ("xchg", C.c_void_p * HLV_MAX_PEERS)]


_lib: Optional[C.CDLL] = None

_vp, _i64, _i32, _f32, _f64, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/hlv.h one to one
SIGNATURES = {
    "hlv_version": (C.c_int, []),
    "hlv_last_error_string": (C.c_char_p, []),
    "hlv_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "hlv_workspace_bytes": (_sz, [_i32]),
    "hlv_workspace_init": (C.c_int, [_vp, _sz, _vp]),
    "hlv_gather_f32": (C.c_int, [_vp, _vp, _i32, _vp, _i64, _f32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "hlv_scatter_f32": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _vp]),
    "hlv_dot_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_lanczos_update_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_normalize_store_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _f64, _vp, _i32, _vp]),
    "hlv_cgs_project_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_cgs_project_bf16": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_cgs_update_f32": (C.c_int, [_vp, _i64, _i32, _vp, _f32, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_cgs_update_bf16": (C.c_int, [_vp, _i64, _i32, _vp, _f32, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_cgs_needs_pass": (C.c_int, [_vp, _i32, _vp, _f64, _vp, _vp]),
    "hlv_cgs_update_if_f32": (C.c_int, [_vp, _i64, _i32, _vp, _f32, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "hlv_cgs_update_if_bf16": (C.c_int, [_vp, _i64, _i32, _vp, _f32, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "hlv_cgs_fused_max_rows": (C.c_int, [_i32]),
    "hlv_cgs_update_project_f32": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "hlv_cgs_update_project_bf16": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "hlv_peer_xchg_bytes": (_sz, []),
    "hlv_peer_xchg_init": (C.c_int, [_vp, _vp]),
    "hlv_peer_xchg_error": (C.c_int, [_vp, C.POINTER(C.c_int), _vp]),
    "hlv_peer_signal": (C.c_int, [_vp, _i32, _vp]),
    "hlv_peer_wait": (C.c_int, [_vp, _i32, _vp]),
    "hlv_x_reduce_scatter_dot_f32": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hlv_x_update_project_f32": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hlv_x_update_project_bf16": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hlv_x_lanczos_update_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_x_cgs_update_project_f32": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "hlv_x_cgs_update_project_bf16": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "hlv_x_cgs_update_f32": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_x_cgs_update_bf16": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "hlv_x_normalize_store_f32": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _f64, _vp, _i32, _vp, _sz, _vp]),
    "hlv_vector_adjust_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _f32, _i64, _vp, _vp, _sz, _vp]),
    "hlv_ritz_vectors_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _i64, _vp]),
    "hlv_ritz_vectors_bf16": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _i64, _vp]),
}


def load() -> C.CDLL:
    """Load libhlv.so (once).  Raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "This package has no CPU or pure-PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        if lib.hlv_version() != HLV_VERSION:
            raise ImportError(f"{LIB_PATH} is version {lib.hlv_version()}, this package needs {HLV_VERSION}: rebuild it "
                              "(python __graft_entry__.py --force)")
        _lib = lib
    return _lib


def check(fn: str, status: int) -> None:
    if status != HLV_OK:
        msg = load().hlv_last_error_string()
        raise HLVError(fn, status, msg.decode() if msg else "")


def call(fn: str, *args) -> None:
    check(fn, getattr(load(), fn)(*args))
