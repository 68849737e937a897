"""Host side of the dense assembly (reference ``updes/assembly.py``).

The reference fills Phi, P, op(Phi), op(P), bd(Phi), bd(P) with serial ``fori_loop``s over rows and
autodiff per entry, then forms ``A``, ``inv(A)`` and ``B = diffMat @ inv(A)``.  Here the host only
builds *row descriptors* -- for every collocation row the evaluation point(s), five coefficients on
the jet (phi, phi_x, phi_y, phi_xx, phi_yy) and the skipped diagonal column -- and the CUDA kernel
(``csrc/assemble.cu``) writes the matrix in one pass.  The assembled system is

    K = [[op(Phi) op(P)], [bd(Phi) bd(P)], [P^T 0]]      ((N+M) x (N+M))

whose solution c satisfies ``K c = [q; 0]`` and gives ``vals = [Phi P] c`` -- the same ``vals`` and
``coeffs`` the reference obtains through ``inv(A)``, a GEMM and a QR solve.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from .rbf import RBF_CODES

JET_VAL, JET_GRAD, JET_HESS, JET_ISO = 1, 2, 4, 8


def padded_ld(n: int) -> int:
    """Leading dimension: rows are whole 128-byte lines (TMA tiles, 16-byte vector stores)."""
    return (n + 15) // 16 * 16


def _mask_of(coef: np.ndarray) -> int:
    """Which parts of the jet a block of coefficient rows uses (kernel specialisation)."""
    if coef.size == 0:
        return JET_VAL
    has_val = bool(np.any(coef[:, 0] != 0))
    has_grad = bool(np.any(coef[:, 1:3] != 0))
    has_hess = bool(np.any(coef[:, 3:5] != 0))
    if has_hess and not has_grad and np.array_equal(coef[:, 3], coef[:, 4]):
        return JET_ISO | (JET_VAL if has_val else 0)      # c0 phi + c3 lap(phi): closed-form radial Laplacian
    m = (JET_VAL if has_val else 0) | (JET_GRAD if has_grad else 0) | ((JET_HESS | JET_GRAD) if has_hess else 0)
    return m or JET_VAL


@dataclass
class RowTable:
    """Row descriptors of the N collocation rows (host copy; see ``struct UpdesRows``)."""
    p1: np.ndarray
    p2: np.ndarray
    cphi1: np.ndarray
    cphi2: np.ndarray
    cpol1: np.ndarray
    cpol2: np.ndarray
    skip: np.ndarray
    Ni: int

    def masks(self):
        """jet masks of the internal and the boundary row ranges."""
        Ni = self.Ni
        return (_mask_of(self.cphi1[:Ni]),
                _mask_of(np.concatenate([self.cphi1[Ni:], self.cphi2[Ni:]], axis=0)))


def build_operator_rows(cloud, coef_phi, coef_pol=None, betas=None) -> RowTable:
    """Row descriptors of ``diffMat`` (reference assembly.py:93-137 for rows < Ni, :141-362 after).

    coef_phi / coef_pol: (Ni, 5) lowered operator coefficients on RBF / monomial columns.
    betas: (Nr,) Robin coefficients in sorted node order (operators.py:512-541).
    """
    N, Ni, Nd, Nn, Nr = cloud.N, cloud.Ni, cloud.Nd, cloud.Nn, cloud.Nr
    Np = list(cloud.Np)
    son = np.asarray(cloud.sorted_outward_normals, dtype=np.float64).reshape(-1, 2)
    coef_phi = np.asarray(coef_phi, dtype=np.float64).reshape(Ni, 5)
    coef_pol = coef_phi if coef_pol is None else np.asarray(coef_pol, dtype=np.float64).reshape(Ni, 5)
    p1 = np.arange(N, dtype=np.int32)
    p2 = np.full(N, -1, dtype=np.int32)
    skip = np.arange(N, dtype=np.int32)
    cphi1 = np.zeros((N, 5)); cphi2 = np.zeros((N, 5)); cpol1 = np.zeros((N, 5)); cpol2 = np.zeros((N, 5))
    cphi1[:Ni] = coef_phi
    cpol1[:Ni] = coef_pol
    # Dirichlet rows: phi (assembly.py:169-176, :283-287)
    d = slice(Ni, Ni + Nd)
    cphi1[d, 0] = 1.0; cpol1[d, 0] = 1.0
    # Neumann rows: grad phi . n_i (assembly.py:178-190, :290-299)
    if Nn:
        nn = slice(Ni + Nd, Ni + Nd + Nn)
        cphi1[nn, 1:3] = son[0:Nn]; cpol1[nn, 1:3] = son[0:Nn]
    # Robin rows: beta_i phi + grad phi . n.  bd(Phi) reads normals[i-Ni-Nd-Nn] (assembly.py:206),
    # i.e. the slot of a Neumann node when Nn > 0; bd(P) uses the node's own normal (:303).  Kept (Q3).
    if Nr:
        rr = slice(Ni + Nd + Nn, Ni + Nd + Nn + Nr)
        b = np.zeros(Nr) if betas is None else np.asarray(betas, dtype=np.float64).reshape(Nr)
        cphi1[rr, 0] = b; cpol1[rr, 0] = b
        cphi1[rr, 1:3] = son[0:Nr]
        cpol1[rr, 1:3] = son[Nn:Nn + Nr]
    # Periodic rows (assembly.py:215-267, :319-359): value rows of every group, then flux rows
    start = Ni + Nd + Nn + Nr
    half = sum(Np) // 2
    node0, row0 = start, start
    for nb in Np:
        nc = nb // 2
        i1 = np.arange(node0, node0 + nc)
        rows_v = np.arange(row0, row0 + nc)
        rows_f = rows_v + half
        for rows, kind in ((rows_v, "v"), (rows_f, "f")):
            p1[rows] = i1; p2[rows] = i1 + nc; skip[rows] = -1        # all columns written (Q2)
        cphi1[rows_v] = 0; cphi2[rows_v] = 0
        cphi1[rows_v, 0] = 1.0; cphi2[rows_v, 0] = -1.0
        n1 = son[i1 - Ni - Nd]; n2 = son[i1 - Ni - Nd + nc]
        cphi1[rows_f] = 0; cphi2[rows_f] = 0
        cphi1[rows_f, 1:3] = n1
        cphi2[rows_f, 1:3] = n2                                       # grad2 . n2: minus of -(n2), :256
        cpol1[rows_v] = cphi1[rows_v]; cpol2[rows_v] = cphi2[rows_v]
        cpol1[rows_f] = cphi1[rows_f]; cpol2[rows_f] = cphi2[rows_f]
        node0 += nb
        row0 += nc
    return RowTable(p1, p2, cphi1, cphi2, cpol1, cpol2, skip, Ni)


def build_interpolation_rows(cloud) -> RowTable:
    """Rows of ``A = [[Phi P], [P^T 0]]`` (assembly.py:10-85): phi at every node, own column skipped."""
    N = cloud.N
    c = np.zeros((N, 5)); c[:, 0] = 1.0
    z = np.zeros((N, 5))
    return RowTable(np.arange(N, dtype=np.int32), np.full(N, -1, dtype=np.int32), c, z, c.copy(), z.copy(),
                    np.arange(N, dtype=np.int32), N)


def require_global_support(cloud):
    """The product assembles the GLOBAL collocation system: every node in every support (``support_size="max"``, i.e. N).
    A cloud built with a smaller support (RBF-FD, reference cloud.py:97-112) is refused here, before any device work."""
    ss = getattr(cloud, "support_size", "max")
    if ss not in ("max", None) and int(ss) != cloud.N:
        raise NotImplementedError(
            "updes_b200 implements the global collocation path only (support_size='max' or N = %d, got %d); "
            "local RBF-FD supports are out of scope (reference README lists them as ill-conditioned)" % (cloud.N, int(ss)))


class DeviceRows:
    """Row descriptors + centres resident in HBM, ready to hand to the C-ABI."""

    def __init__(self, cloud, table: RowTable, device="cuda"):
        require_global_support(cloud)
        torch = _lib.require_cuda()
        self.torch = torch
        self.N = cloud.N
        self.table = table
        f = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(device, non_blocking=False)
        self.centres = f(cloud.sorted_nodes, torch.float64)
        self.p1 = f(table.p1, torch.int32); self.p2 = f(table.p2, torch.int32); self.skip = f(table.skip, torch.int32)
        self.cphi1 = f(table.cphi1, torch.float64); self.cphi2 = f(table.cphi2, torch.float64)
        self.cpol1 = f(table.cpol1, torch.float64); self.cpol2 = f(table.cpol2, torch.float64)
        self.struct = _lib.UpdesRows(self.centres.data_ptr(), self.p1.data_ptr(), self.p2.data_ptr(),
                                     self.cphi1.data_ptr(), self.cphi2.data_ptr(), self.cpol1.data_ptr(),
                                     self.cpol2.data_ptr(), self.skip.data_ptr())
        self.mask_internal, self.mask_boundary = table.masks()


def assemble_system(rows: DeviceRows, rbf_kind: str, rbf_param: float, M: int, out=None):
    """Assemble the (N+M) x (N+M) system described by ``rows`` into HBM (row-major, padded ld).

    Three launches over row ranges so each gets the narrowest jet specialisation: internal rows,
    boundary rows, P^T rows.  Returns the torch tensor of shape (N+M, ld)."""
    torch = rows.torch
    lib = _lib.load()
    N, n = rows.N, rows.N + M
    ld = padded_ld(n)
    if out is None:
        out = torch.empty((n, ld), dtype=torch.float64, device=rows.centres.device)
    assert out.shape == (n, ld) and out.is_contiguous()
    code = RBF_CODES[rbf_kind]
    st = _lib.stream_ptr()
    Ni = rows.table.Ni
    ranges = [(0, Ni, rows.mask_internal), (Ni, N - Ni, rows.mask_boundary), (N, M, 7)]
    for r0, nr, mask in ranges:
        if nr <= 0:
            continue
        rc = lib.updes_assemble_rows(code, float(rbf_param), N, M, rows.centres.data_ptr(), ctypes.byref(rows.struct),
                                     r0, nr, mask, out.data_ptr() + 8 * r0 * ld, ld, st)
        _lib.check(rc, "updes_assemble_rows")
    return out


def eval_jets(rbf_kind, rbf_param, centres, coeffs, pts, skip=None):
    """Matrix-free jet sums (replaces the reference's field evaluators, operators.py:118-351).

    centres (N,2), coeffs (nf, N+M) and pts (R,2) are CUDA float64 tensors; skip an int32 tensor or
    None.  Returns (jphi, jpol), each (nf, R, 5): RBF and polynomial parts of
    (value, d/dx, d/dy, d2/dx2, d2/dy2) at every point."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    N = centres.shape[0]
    nf, ncoef = coeffs.shape
    M = ncoef - N
    R = pts.shape[0]
    coeffs = coeffs.contiguous(); pts = pts.contiguous()
    jphi = torch.empty((nf, R, 5), dtype=torch.float64, device=centres.device)
    jpol = torch.empty((nf, R, 5), dtype=torch.float64, device=centres.device)
    ws_bytes = lib.updes_eval_jets_workspace_bytes(N, R, nf)
    ws = torch.empty((max(ws_bytes, 8) + 7) // 8, dtype=torch.float64, device=centres.device)
    rc = lib.updes_eval_jets(RBF_CODES[rbf_kind], float(rbf_param), N, M, centres.data_ptr(), coeffs.data_ptr(), ncoef,
                             nf, pts.data_ptr(), R, skip.data_ptr() if skip is not None else None, jphi.data_ptr(),
                             jpol.data_ptr(), ws.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "updes_eval_jets")
    return jphi, jpol


def apply_rows(rows: DeviceRows, rbf_kind, rbf_param, M, coeffs):
    """Matrix-free product of the assembled system with coefficient vectors: returns K @ c for
    c given as (nf, N+M) -> (nf, N+M).  Used for residuals / backward error without a second matrix."""
    torch = rows.torch
    N = rows.N
    t = rows.table
    # a point skips its own column exactly when some row that evaluates there does (non-periodic nodes)
    if not hasattr(rows, "_pt_skip"):
        pt_skip = np.full(N, -1, dtype=np.int32)
        sel = t.skip >= 0
        pt_skip[t.p1[sel]] = t.skip[sel]
        rows._pt_skip = torch.as_tensor(pt_skip).to(rows.centres.device)
    jphi, jpol = eval_jets(rbf_kind, rbf_param, rows.centres, coeffs, rows.centres, rows._pt_skip)
    p1 = rows.p1.long()
    out = (jphi[:, p1, :] * rows.cphi1[None]).sum(-1) + (jpol[:, p1, :] * rows.cpol1[None]).sum(-1)
    has2 = rows.p2 >= 0
    if bool(has2.any()):
        p2 = rows.p2.clamp(min=0).long()
        second = (jphi[:, p2, :] * rows.cphi2[None]).sum(-1) + (jpol[:, p2, :] * rows.cpol2[None]).sum(-1)
        out = out + second * has2[None].to(out.dtype)
    # P^T rows: sum_j monomial_m(x_j) c_j  (tiny: M dot products of length N)
    from .rbf import MONOMIAL_EXPONENTS
    x, y = rows.centres[:, 0], rows.centres[:, 1]
    pt = torch.stack([(x ** a) * (y ** b) for a, b in MONOMIAL_EXPONENTS[:M]], dim=0) if M else x.new_zeros((0, N))
    tail = coeffs[:, :N] @ pt.T
    return torch.cat([out, tail], dim=1)
