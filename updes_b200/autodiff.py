"""Differentiable solves with a factored collocation system.

The reference's differentiable-physics demos (demos/Laplace/10_laplace_with_diff_phys.py:90-104)
differentiate a loss through ``pde_solver`` w.r.t. a boundary-condition array; with JAX that replays
autodiff through inv(A), the GEMM and the QR.  Here the factorisation is reused: for c = K^-1 b the
vector-Jacobian product is  b_bar = K^-T c_bar  -- one transposed solve with the same LU factors
(``updes_lu_solve(..., transpose=1)``).  PyTorch autograd is only the tape; the arithmetic is the CUDA
triangular solves.
"""
from __future__ import annotations

from . import _lib


def _function_class():
    torch = _lib.require_cuda()

    class LinearSolve(torch.autograd.Function):
        @staticmethod
        def forward(ctx, b, lu):
            ctx.lu = lu
            return lu.solve(b.detach().clone().contiguous())

        @staticmethod
        def backward(ctx, grad_out):
            return ctx.lu.solve(grad_out.detach().clone().contiguous(), transpose=True), None

    return LinearSolve


_CLS = None


def linear_solve(lu, b):
    """x = K^-1 b, differentiable w.r.t. ``b`` (CUDA float64 tensor, shape (n,) or (nrhs, n));
    ``lu`` is a factored ``updes_b200.linalg.LUFactorization``."""
    global _CLS
    if _CLS is None:
        _CLS = _function_class()
    return _CLS.apply(b, lu)
