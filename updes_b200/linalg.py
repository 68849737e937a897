"""Dense FP64 LU with partial pivoting on the GPU: factor once, solve many.

Replaces the reference's ``jnp.linalg.inv`` (assembly.py:87-90) + ``diffMat @ inv_A`` (:399) +
``lineax.QR`` solve (operators.py:612-613) + second ``inv(A)`` matvec (:616) by one in-place
factorisation ``P K = L U`` of the collocation system and two triangular sweeps per right-hand side.
"""
from __future__ import annotations

import ctypes

from . import _lib


class LUFactorization:
    """Owns the factors (in place in ``K``), the pivots and the native handle."""

    def __init__(self, K, n=None):
        torch = _lib.require_cuda()
        self._lib = _lib.load()
        assert K.dtype == torch.float64 and K.is_cuda and K.dim() == 2 and K.is_contiguous()
        self.n = K.shape[0] if n is None else n
        self.ld = K.shape[1]
        self.K = K
        self.ipiv = torch.empty(self.n, dtype=torch.int32, device=K.device)
        self.info = torch.zeros(1, dtype=torch.int32, device=K.device)
        handle = ctypes.c_void_p()
        _lib.check(self._lib.updes_lu_create(ctypes.byref(handle), self.n, self.ld), "updes_lu_create")
        self._handle = handle
        self.factored = False

    def factor(self, equilibrate=False):
        """Enqueue the factorisation on the current stream (asynchronous).  ``equilibrate``: scale every
        row by the power of two that brings its largest magnitude into [1, 2) before pivoting; the factors
        are kept in ``self.scale`` and applied to right-hand sides by ``solve``."""
        if equilibrate:
            torch = _lib.require_cuda()
            self.scale = torch.empty(self.n, dtype=torch.float64, device=self.K.device)
            rc = self._lib.updes_lu_factor_scaled(self._handle, self.K.data_ptr(), self.ipiv.data_ptr(),
                                                  self.scale.data_ptr(), self.info.data_ptr(), _lib.stream_ptr())
            _lib.check(rc, "updes_lu_factor_scaled")
        else:
            self.scale = None
            rc = self._lib.updes_lu_factor(self._handle, self.K.data_ptr(), self.ipiv.data_ptr(), self.info.data_ptr(),
                                           _lib.stream_ptr())
            _lib.check(rc, "updes_lu_factor")
        self.factored = True
        return self

    def zero_pivot(self) -> int:
        """0, or the 1-based index of the first exactly-zero pivot (synchronises).  Negative values are
        internal failures (see ``check``)."""
        return int(self.info.item())

    def check(self) -> int:
        """Raise ``RuntimeError`` if a device-side wait timed out (the panel kernel's grid barrier: info < 0;
        a triangular sweep: status bit 0) -- results are garbage then and must not be returned.  Otherwise
        the LAPACK-style status: 0, or the 1-based column of the first exactly-zero pivot."""
        info = int(self.info.item())
        if info < 0:
            raise RuntimeError("updes_b200: the panel kernel's grid barrier timed out (info = %d); the factors are invalid" % info)
        flags = ctypes.c_int32(0)
        _lib.check(self._lib.updes_lu_status(self._handle, ctypes.byref(flags), _lib.stream_ptr()), "updes_lu_status")
        if flags.value:
            raise RuntimeError("updes_b200: a triangular sweep timed out waiting for a solved block (status %d)" % flags.value)
        return info

    def solve(self, B, transpose=False):
        """Solve K X = B (or K^T X = B) in place.  B: (nrhs, ldb>=n) or (n,) CUDA float64; returns B."""
        assert self.factored
        single = B.dim() == 1
        Bm = B.view(1, -1) if single else B
        assert Bm.is_contiguous() and Bm.shape[1] >= self.n
        rc = self._lib.updes_lu_solve(self._handle, self.K.data_ptr(), self.ipiv.data_ptr(), Bm.data_ptr(),
                                      Bm.shape[1], Bm.shape[0], 1 if transpose else 0, _lib.stream_ptr())
        _lib.check(rc, "updes_lu_solve")
        return B

    def set_panel_variant(self, variant: int):
        _lib.check(self._lib.updes_lu_set_panel_variant(self._handle, variant), "updes_lu_set_panel_variant")

    def set_trsm_base(self, rows: int):
        _lib.check(self._lib.updes_lu_set_trsm_base(self._handle, rows), "updes_lu_set_trsm_base")

    def set_solve_variant(self, variant: int):
        _lib.check(self._lib.updes_lu_set_solve_variant(self._handle, variant), "updes_lu_set_solve_variant")

    def set_panel_capacity(self, rows: int):
        _lib.check(self._lib.updes_lu_set_panel_capacity(self._handle, rows), "updes_lu_set_panel_capacity")

    def set_gemm_variant(self, variant: int):
        _lib.check(self._lib.updes_lu_set_gemm_variant(self._handle, variant), "updes_lu_set_gemm_variant")

    def gemm_sub(self, rc_, cc, ra, ca, rb, cb, m, n, k):
        """C -= A @ B on sub-blocks of the bound matrix (exposed for kernel-level tests)."""
        rc = self._lib.updes_dgemm_sub(self._handle, self.K.data_ptr(), rc_, cc, ra, ca, rb, cb, m, n, k,
                                       _lib.stream_ptr())
        _lib.check(rc, "updes_dgemm_sub")

    def panel(self, r0, nc):
        rc = self._lib.updes_lu_panel(self._handle, self.K.data_ptr(), r0, nc, self.ipiv.data_ptr(),
                                      self.info.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_lu_panel")

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.updes_lu_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
