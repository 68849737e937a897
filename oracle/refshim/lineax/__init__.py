"""lineax stand-in: MatrixLinearOperator + linear_solve(..., solver=QR()) as Householder QR + back-substitution
(what lineax's QR solver does for a square full-rank matrix), on torch float64."""
import torch


class MatrixLinearOperator:
    def __init__(self, matrix):
        self.matrix = matrix


class QR:
    pass


class GMRES:
    def __init__(self, *a, **k):
        raise NotImplementedError("the stand-in only provides the QR solver the reference's pde_solver uses")


class _Solution:
    def __init__(self, value):
        self.value = value


def linear_solve(operator, vector, solver=None):
    assert isinstance(solver, QR), "the reference calls lx.linear_solve(..., solver=lx.QR())"
    Q, R = torch.linalg.qr(operator.matrix)
    return _Solution(torch.linalg.solve_triangular(R, (Q.T @ vector).unsqueeze(-1), upper=True).squeeze(-1))
