"""ctypes binding of the C-ABI in ``include/updes_b200.h`` (``libupdes_b200.so``, built in-tree).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present, every
compute entry point raises.  Device memory and streams come from PyTorch (plumbing only).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libupdes_b200.so")

c_double_p = ctypes.c_void_p
c_int32_p = ctypes.c_void_p


class UpdesRows(ctypes.Structure):
    """struct UpdesRows (include/updes_b200.h): device pointers of the row descriptors."""
    _fields_ = [("pts", ctypes.c_void_p), ("p1", ctypes.c_void_p), ("p2", ctypes.c_void_p),
                ("cphi1", ctypes.c_void_p), ("cphi2", ctypes.c_void_p), ("cpol1", ctypes.c_void_p),
                ("cpol2", ctypes.c_void_p), ("skip", ctypes.c_void_p)]


# name -> (restype, argtypes): every symbol the header declares
_I64, _I32, _VP, _DBL = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_double
SIGNATURES = {
    "updes_assemble_rows": (_I32, [_I32, _DBL, _I32, _I32, _VP, ctypes.POINTER(UpdesRows), _I64, _I64, _I32, _VP, _I64, _VP]),
    "updes_assemble_block": (_I32, [_I32, _DBL, _I32, _I32, _VP, ctypes.POINTER(UpdesRows), _I64, _I64, _I64, _I64, _I32, _VP, _I64, _VP]),
    "updes_eval_jets_workspace_bytes": (ctypes.c_size_t, [_I32, _I32, _I32]),
    "updes_eval_jets": (_I32, [_I32, _DBL, _I32, _I32, _VP, _VP, _I64, _I32, _VP, _I32, _VP, _VP, _VP, _VP, _VP]),
    "updes_lu_create": (_I32, [ctypes.POINTER(_VP), _I64, _I64]),
    "updes_lu_destroy": (_I32, [_VP]),
    "updes_lu_factor": (_I32, [_VP, _VP, _VP, _VP, _VP]),
    "updes_lu_factor_scaled": (_I32, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "updes_row_absmax": (_I32, [_VP, _I64, _I64, _I64, _VP, _VP]),
    "updes_scale_from_absmax": (_I32, [_VP, _I64, _VP, _VP]),
    "updes_row_scale": (_I32, [_VP, _I64, _I64, _I64, _VP, _VP]),
    "updes_lu_set_row_scale": (_I32, [_VP, _VP]),
    "updes_lu_status": (_I32, [_VP, ctypes.POINTER(ctypes.c_int32), _VP]),
    "updes_lu_solve": (_I32, [_VP, _VP, _VP, _VP, _I64, _I32, _I32, _VP]),
    "updes_dgemm_sub": (_I32, [_VP, _VP, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _VP]),
    "updes_lu_panel": (_I32, [_VP, _VP, _I64, _I64, _VP, _VP, _VP]),
    "updes_lu_bind": (_I32, [_VP, _I32, _VP, _I64, _I64]),
    "updes_lu_set_gemm_ctas": (_I32, [_VP, _I32]),
    "updes_lu_set_gemm_variant": (_I32, [_VP, _I32]),
    "updes_lu_set_panel_capacity": (_I32, [_VP, _I64]),
    "updes_lu_set_solve_variant": (_I32, [_VP, _I32]),
    "updes_lu_set_panel_variant": (_I32, [_VP, _I32]),
    "updes_lu_set_trsm_base": (_I32, [_VP, _I32]),
    "updes_lu_panel_factor": (_I32, [_VP, _I32, _I64, _I64, _I64, _VP, _VP, _VP]),
    "updes_lu_apply_swaps": (_I32, [_VP, _I32, _I64, _I64, _I64, _I64, _VP, _VP]),
    "updes_lu_trsm": (_I32, [_VP, _I32, _I64, _I64, _I64, _I32, _I64, _I64, _I64, _VP]),
    "updes_lu_gemm": (_I32, [_VP, _I32, _I64, _I64, _I32, _I64, _I64, _I32, _I64, _I64, _I64, _I64, _I64, _VP]),
    "updes_lu_set_pivots": (_I32, [_VP, _VP, _VP]),
    "updes_lu_permute_rhs": (_I32, [_VP, _VP, _I64, _I32, _VP, _VP]),
    "updes_tri_block_sweep": (_I32, [_VP, _I32, _I32, _I64, _I64, _I64, _VP, _I32, _VP]),
    "updes_block_gemv": (_I32, [_VP, _I32, _I64, _I64, _I64, _I64, _VP, _VP, _VP]),
    "updes_tri_diag_solve": (_I32, [_VP, _I32, _I32, _I64, _I64, _I64, _VP, _VP]),
    "updes_b200_version": (ctypes.c_char_p, []),
    "updes_launch_count": (_I64, []),
    "updes_profile_enable": (_I32, [_I32]),
    "updes_assemble_set_variant": (_I32, [_I32]),
    "updes_profile_read": (_I32, [_I32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "updes_profile_records": (ctypes.c_int64, [_I32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_int64]),
}

_lib = None


def load():
    """Load the shared library (no GPU needed for loading; needed for any call that computes)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "updes_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C updes_b200/csrc`).  There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise ValueError("updes_b200.%s: bad argument #%d" % (what, -rc))
    raise RuntimeError("updes_b200.%s: CUDA error %d while enqueuing" % (what, rc))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("updes_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().updes_launch_count())


PROF_CLASSES = {"gemm": 0, "panel": 1, "swap": 2, "trsm": 3, "assemble": 4, "solve": 5}


def profile_enable(on: bool):
    load().updes_profile_enable(1 if on else 0)


def profile_read(name: str):
    """(milliseconds, work, launches) of one kernel class since profile_enable(True)."""
    ms, work, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    check(load().updes_profile_read(PROF_CLASSES[name], ctypes.byref(ms), ctypes.byref(work), ctypes.byref(cnt)),
          "updes_profile_read")
    return ms.value, work.value, cnt.value


def profile_records(name: str, limit: int = 1 << 20):
    """Per-launch (ms, work) arrays of one kernel class since profile_enable(True), in launch order."""
    import numpy as np
    ms = np.zeros(limit)
    work = np.zeros(limit)
    dp = ctypes.POINTER(ctypes.c_double)
    cnt = int(load().updes_profile_records(PROF_CLASSES[name], ms.ctypes.data_as(dp), work.ctypes.data_as(dp), limit))
    if cnt < 0:
        raise RuntimeError("updes_profile_records failed: CUDA error %d" % -cnt)
    cnt = min(cnt, limit)
    return ms[:cnt], work[:cnt]
