"""CPU emulation of the single-GPU part of the C-ABI (include/updes_b200.h) -- TEST INFRASTRUCTURE, never shipped.

Why: the build container has no GPU, and part of the host layer (updes_b200/operators.py, explicit.py, cloud.py) and
of the `-m gpu` tests was written after the round's GPU minutes were spent.  `install()` lets that host code and those
tests run HERE, unchanged, on CPU tensors: every entry point of libupdes_b200.so the operator path calls is replaced by
a numpy / LAPACK function with the semantics the header documents, and "cuda" devices are mapped to CPU memory.  What
this checks is the host orchestration and the TEST LOGIC (shapes, golden keys, tolerances, launch-count assertions);
the CUDA kernels themselves are only ever checked on a B200 (`pytest -m gpu`).  The jets are written from the closed
forms in the header / DESIGN.md, independently of oracle/updes_oracle.c, and `self_check()` compares the two.

Nothing under updes_b200/ imports this file; it is activated only by tests/run_gpu_tests_on_cpu.py and by
tests/test_host_on_emulated_abi.py (through a subprocess, so the monkey-patching never leaks into other tests).
"""
import ctypes
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PHI0 = {0: 0.0, 1: 0.0, 2: 1.0, 3: 1.0, 4: 1.0}
EXPONENTS = [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2), (3, 0), (2, 1), (1, 2), (0, 3),
             (4, 0), (3, 1), (2, 2), (1, 3), (0, 4)]


def _ptr(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return int(p)


def _f64(p, count):
    return np.ctypeslib.as_array((ctypes.c_double * int(count)).from_address(_ptr(p)))


def _i32(p, count):
    return np.ctypeslib.as_array((ctypes.c_int32 * int(count)).from_address(_ptr(p)))


def rbf_jets(kind, param, pts, centres):
    """(R, N, 5): jet (phi, phi_x, phi_y, phi_xx, phi_yy) w.r.t. the evaluation point; derivatives 0 at r == 0."""
    dx = pts[:, None, 0] - centres[None, :, 0]
    dy = pts[:, None, 1] - centres[None, :, 1]
    s = dx * dx + dy * dy
    zero = s == 0.0
    ss = np.where(zero, 1.0, s)
    if kind == 0:
        p = 2 * int(param) + 1
        r = np.sqrt(ss)
        phi, g, h = r ** p, p * r ** (p - 2), p * (p - 2) * r ** (p - 4)
    elif kind == 1:
        a = int(param)
        q = 2 * a
        L = 0.5 * np.log(ss)
        t = q * L + 1.0
        phi, g, h = ss ** a * L, ss ** (a - 1) * t, ss ** (a - 2) * ((q - 2) * t + q)
    elif kind == 2:
        e2 = param * param
        phi = np.exp(-e2 * ss)
        g, h = -2.0 * e2 * phi, 4.0 * e2 * e2 * phi
    elif kind == 3:
        e2 = param * param
        w = 1.0 + e2 * ss
        phi, g, h = np.sqrt(w), e2 / np.sqrt(w), -(e2 * e2) * w ** -1.5
    else:
        e2 = param * param
        w = 1.0 + e2 * ss
        phi, g, h = w ** -0.5, -e2 * w ** -1.5, 3.0 * e2 * e2 * w ** -2.5
    J = np.stack([phi, g * dx, g * dy, g + h * dx * dx, g + h * dy * dy], axis=-1)
    J[zero] = (PHI0[kind], 0.0, 0.0, 0.0, 0.0)
    return J


def monomial_jets(M, pts):
    """(R, M, 5)"""
    x, y = pts[:, 0], pts[:, 1]
    out = np.zeros((pts.shape[0], M, 5))

    def pw(v, e):
        return np.ones_like(v) if e == 0 else v ** e

    for m, (a, b) in enumerate(EXPONENTS[:M]):
        out[:, m, 0] = pw(x, a) * pw(y, b)
        if a >= 1:
            out[:, m, 1] = a * pw(x, a - 1) * pw(y, b)
        if b >= 1:
            out[:, m, 2] = b * pw(x, a) * pw(y, b - 1)
        if a >= 2:
            out[:, m, 3] = a * (a - 1) * pw(x, a - 2) * pw(y, b)
        if b >= 2:
            out[:, m, 4] = b * (b - 1) * pw(x, a) * pw(y, b - 2)
    return out


class _Handle:
    def __init__(self, n, ld):
        self.n, self.ld, self.scale_ptr = n, ld, 0
        self.slots = {}                  # slot -> (ptr, rows, ld)
        self.perm = None


class EmulatedLib:
    """Same entry-point names and argument meaning as libupdes_b200.so (subset: assembly, evaluators, whole-matrix LU)."""

    def __init__(self):
        self.launches = 0
        self.handles = {}
        self.next_handle = 1
        self.prof = [0] * 6                              # launches per class since profile_enable(1)

    # ---- assembly -----------------------------------------------------------------------------------------------
    def _rows(self, rows_ref, N):
        st = rows_ref._obj if hasattr(rows_ref, "_obj") else rows_ref
        g = lambda name, n, f: f(getattr(st, name), n).copy()
        return dict(p1=g("p1", N, _i32), p2=g("p2", N, _i32), skip=g("skip", N, _i32),
                    cphi1=g("cphi1", 5 * N, _f64).reshape(N, 5), cphi2=g("cphi2", 5 * N, _f64).reshape(N, 5),
                    cpol1=g("cpol1", 5 * N, _f64).reshape(N, 5), cpol2=g("cpol2", 5 * N, _f64).reshape(N, 5))

    def _block(self, kind, param, N, M, centres, rows_ref, row0, nrows, col0, ncols):
        ctr = _f64(centres, 2 * N).reshape(N, 2)
        out = np.zeros((nrows, ncols))
        n_coll = max(0, min(row0 + nrows, N) - row0)
        cols = np.arange(col0, col0 + ncols)
        rbf_cols, pol_cols = cols < N, (cols >= N) & (cols < N + M)
        if n_coll > 0:
            R = self._rows(rows_ref, N)
            sl = slice(row0, row0 + n_coll)
            pts_all = _f64((rows_ref._obj if hasattr(rows_ref, "_obj") else rows_ref).pts, 2 * N).reshape(N, 2)
            p1, p2 = R["p1"][sl], R["p2"][sl]
            x1 = pts_all[p1]
            has2 = p2 >= 0
            x2 = pts_all[np.where(has2, p2, 0)]
            cj = cols[rbf_cols]
            if cj.size:
                J1 = rbf_jets(kind, param, x1, ctr[cj])
                blk = np.einsum("rjk,rk->rj", J1, R["cphi1"][sl])
                if has2.any():
                    J2 = rbf_jets(kind, param, x2, ctr[cj])
                    blk += np.einsum("rjk,rk->rj", J2, R["cphi2"][sl]) * has2[:, None]
                blk[R["skip"][sl][:, None] == cj[None, :]] = 0.0
                out[:n_coll, rbf_cols] = blk
            if pol_cols.any():
                mi = cols[pol_cols] - N
                P1 = monomial_jets(M, x1)[:, mi]
                blk = np.einsum("rmk,rk->rm", P1, R["cpol1"][sl])
                if has2.any():
                    P2 = monomial_jets(M, x2)[:, mi]
                    blk += np.einsum("rmk,rk->rm", P2, R["cpol2"][sl]) * has2[:, None]
                out[:n_coll, pol_cols] = blk
        # P^T rows
        for r in range(max(row0, N), min(row0 + nrows, N + M)):
            cj = cols[rbf_cols]
            out[r - row0, rbf_cols] = monomial_jets(M, ctr[cj])[:, r - N, 0]
        return out

    def updes_assemble_rows(self, kind, param, N, M, centres, rows_ref, row0, nrows, mask, out, ld, st):
        if kind < 0 or kind > 4: return -1
        if nrows < 0 or row0 < 0 or row0 + nrows > N + M: return -7
        if ld % 16 or ld < N + M: return -11
        self.launches += 1
        self.prof[4] += 1
        o = _f64(out, nrows * ld).reshape(nrows, ld)
        o[:] = self._block(kind, param, N, M, centres, rows_ref, row0, nrows, 0, ld)
        return 0

    def updes_assemble_block(self, kind, param, N, M, centres, rows_ref, row0, nrows, col0, ncols, mask, out, ld, st):
        self.launches += 1
        self.prof[4] += 1
        o = _f64(out, (nrows - 1) * ld + ncols).reshape(-1) if nrows else None
        blk = self._block(kind, param, N, M, centres, rows_ref, row0, nrows, col0, ncols)
        for r in range(nrows):
            o[r * ld:r * ld + ncols] = blk[r]
        return 0

    def updes_assemble_set_variant(self, v):
        return 0

    # ---- evaluators ---------------------------------------------------------------------------------------------
    def updes_eval_jets_workspace_bytes(self, N, npts, nf):
        return 64

    def updes_eval_jets(self, kind, param, N, M, centres, coeffs, ldc, nf, pts, npts, skip, jphi, jpol, ws, st):
        self.launches += 2
        ctr = _f64(centres, 2 * N).reshape(N, 2)
        C = _f64(coeffs, (nf - 1) * ldc + N + M) if nf else np.zeros(0)
        P = _f64(pts, 2 * npts).reshape(npts, 2)
        jp = _f64(jphi, nf * npts * 5).reshape(nf, npts, 5)
        jq = _f64(jpol, nf * npts * 5).reshape(nf, npts, 5)
        sk = _i32(skip, npts) if _ptr(skip) else None
        step = max(1, 4_000_000 // max(N, 1))
        for lo in range(0, npts, step):
            hi = min(npts, lo + step)
            J = rbf_jets(kind, param, P[lo:hi], ctr)
            if sk is not None:
                J[sk[lo:hi, None] == np.arange(N)[None, :]] = 0.0
            Q = monomial_jets(M, P[lo:hi])
            for f in range(nf):
                c = C[f * ldc:f * ldc + N + M]
                jp[f, lo:hi] = np.einsum("rjk,j->rk", J, c[:N])
                jq[f, lo:hi] = np.einsum("rmk,m->rk", Q, c[N:]) if M else 0.0
        return 0

    # ---- LU -----------------------------------------------------------------------------------------------------
    def updes_lu_create(self, href, n, ld):
        if n <= 0: return -2
        if ld < 16 or ld % 16: return -3
        h = self.next_handle
        self.next_handle += 1
        self.handles[h] = _Handle(n, ld)
        href._obj.value = h
        return 0

    def updes_lu_destroy(self, h):
        self.handles.pop(_ptr(h), None)
        return 0

    def _factor(self, h, K, ipiv, info):
        H = self.handles[_ptr(h)]
        A = _f64(K, H.n * H.ld).reshape(H.n, H.ld)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lu, piv = sla.lu_factor(A[:, :H.n].copy(), check_finite=False)
        A[:, :H.n] = lu
        _i32(ipiv, H.n)[:] = piv
        d = np.flatnonzero(np.diag(lu) == 0.0)
        _i32(info, 1)[0] = int(d[0]) + 1 if d.size else 0
        k = 1 + H.n // 128                           # a factorisation is many launches; a solve is a handful
        self.launches += 20 * k
        for c, m in ((0, 6), (1, 1), (2, 2), (3, 6)):
            self.prof[c] += m * k
        return 0

    def updes_lu_factor(self, h, K, ipiv, info, st):
        self.handles[_ptr(h)].scale_ptr = 0
        return self._factor(h, K, ipiv, info)

    def updes_lu_factor_scaled(self, h, K, ipiv, scale, info, st):
        if not _ptr(scale): return -4
        H = self.handles[_ptr(h)]
        A = _f64(K, H.n * H.ld).reshape(H.n, H.ld)
        m = np.abs(A[:, :H.n]).max(axis=1)
        sc = np.where(m > 0, np.ldexp(1.0, -np.floor(np.log2(np.where(m > 0, m, 1.0))).astype(int)), 1.0)
        _f64(scale, H.n)[:] = sc
        A[:, :H.n] *= sc[:, None]
        H.scale_ptr = _ptr(scale)
        self.launches += 3
        return self._factor(h, K, ipiv, info)

    def updes_lu_set_row_scale(self, h, scale):
        self.handles[_ptr(h)].scale_ptr = _ptr(scale)
        return 0

    def updes_lu_status(self, h, flags_ref, st):
        flags_ref._obj.value = 0
        return 0

    def updes_lu_solve(self, h, LU, ipiv, B, ldb, nrhs, transpose, st):
        H = self.handles[_ptr(h)]
        if ldb < H.n: return -5
        A = _f64(LU, H.n * H.ld).reshape(H.n, H.ld)[:, :H.n]
        piv = _i32(ipiv, H.n).copy()
        Bm = _f64(B, (nrhs - 1) * ldb + H.n) if nrhs else None
        sc = _f64(H.scale_ptr, H.n) if H.scale_ptr else None
        for f in range(nrhs):
            b = Bm[f * ldb:f * ldb + H.n]
            if transpose:
                x = sla.lu_solve((A, piv), b, trans=1, check_finite=False)
                b[:] = x * sc if sc is not None else x
            else:
                b[:] = sla.lu_solve((A, piv), b * sc if sc is not None else b, check_finite=False)
        self.launches += 3
        self.prof[5] += 3
        return 0

    def updes_lu_set_panel_variant(self, h, v): return 0
    def updes_lu_set_trsm_base(self, h, v): return 0
    def updes_lu_set_solve_variant(self, h, v): return 0
    def updes_lu_set_panel_capacity(self, h, v): return 0
    def updes_lu_set_gemm_variant(self, h, v): return 0
    def updes_lu_set_gemm_ctas(self, h, v): return 0

    # ---- building blocks on bound buffers (multi-GPU drivers: distributed.py, grid2d.py) -------------------------
    def _slot(self, h, slot):
        ptr, rows, ld = self.handles[_ptr(h)].slots[slot]
        return _f64(ptr, rows * ld).reshape(rows, ld), rows

    def updes_lu_bind(self, h, slot, ptr, rows, ld):
        if slot < 0 or slot > 3: return -2
        if ld % 16: return -5
        self.handles[_ptr(h)].slots[slot] = (_ptr(ptr), int(rows), int(ld))
        return 0

    def updes_lu_panel_factor(self, h, slot, r0, c0, nc, ipiv, info, st):
        A, rows = self._slot(h, slot)
        piv = _i32(ipiv, r0 + nc)
        inf = _i32(info, 1)
        for j in range(nc):
            p = r0 + j + int(np.argmax(np.abs(A[r0 + j:rows, c0 + j])))
            piv[r0 + j] = p
            if A[p, c0 + j] == 0.0 and inf[0] == 0:
                inf[0] = r0 + j + 1
            if p != r0 + j:
                A[[r0 + j, p], c0:c0 + nc] = A[[p, r0 + j], c0:c0 + nc]
            if A[r0 + j, c0 + j] != 0.0:
                A[r0 + j + 1:rows, c0 + j] /= A[r0 + j, c0 + j]
                A[r0 + j + 1:rows, c0 + j + 1:c0 + nc] -= np.outer(A[r0 + j + 1:rows, c0 + j], A[r0 + j, c0 + j + 1:c0 + nc])
        self.launches += 1; self.prof[1] += 1
        return 0

    def updes_lu_apply_swaps(self, h, slot, c_lo, c_hi, k0, npiv, ipiv, st):
        A, rows = self._slot(h, slot)
        piv = _i32(ipiv, k0 + npiv)
        for t in range(npiv):
            p = int(piv[k0 + t])
            if p != k0 + t:
                A[[k0 + t, p], c_lo:c_hi] = A[[p, k0 + t], c_lo:c_hi]
        self.launches += 1; self.prof[2] += 1
        return 0

    def updes_lu_trsm(self, h, slot_l, rl, cl, n1, slot_b, rb, cb, ncols, st):
        L, _ = self._slot(h, slot_l)
        B, _ = self._slot(h, slot_b)
        L11 = np.tril(L[rl:rl + n1, cl:cl + n1], -1) + np.eye(n1)
        B[rb:rb + n1, cb:cb + ncols] = sla.solve_triangular(L11, B[rb:rb + n1, cb:cb + ncols], lower=True, unit_diagonal=True)
        self.launches += 1; self.prof[3] += 1
        return 0

    def updes_lu_gemm(self, h, slot_a, ra, ca, slot_b, rb, cb, slot_c, rc, cc, m, n, k, st):
        A, _ = self._slot(h, slot_a)
        B, _ = self._slot(h, slot_b)
        C, crows = self._slot(h, slot_c)
        if rc + m > crows: return -9
        C[rc:rc + m, cc:cc + n] -= A[ra:ra + m, ca:ca + k] @ B[rb:rb + k, cb:cb + n]
        self.launches += 1; self.prof[0] += 1
        return 0

    def updes_lu_set_pivots(self, h, ipiv, st):
        H = self.handles[_ptr(h)]
        piv = _i32(ipiv, H.n)
        perm = np.arange(H.n)
        for k in range(H.n):
            q = int(piv[k])
            perm[[k, q]] = perm[[q, k]]
        H.perm = perm
        self.launches += 1
        return 0

    def updes_lu_permute_rhs(self, h, B, ldb, nrhs, X, st):
        H = self.handles[_ptr(h)]
        sc = _f64(H.scale_ptr, H.n) if H.scale_ptr else np.ones(H.n)
        b = _f64(B, (nrhs - 1) * ldb + H.n)
        x = _f64(X, nrhs * H.n)
        for f in range(nrhs):
            x[f * H.n:(f + 1) * H.n] = (b[f * ldb:f * ldb + H.n] * sc)[H.perm]
        self.launches += 1; self.prof[5] += 1
        return 0

    def updes_tri_block_sweep(self, h, slot, upper, r0, c0, width, X, nrhs, st):
        A, rows = self._slot(h, slot)
        n = self.handles[_ptr(h)].n
        x = _f64(X, nrhs * n).reshape(nrhs, n)
        T = A[r0:r0 + width, c0:c0 + width]
        for f in range(nrhs):
            if not upper:
                x[f, r0:r0 + width] = np.linalg.solve(np.tril(T, -1) + np.eye(width), x[f, r0:r0 + width])
                x[f, r0 + width:n] -= A[r0 + width:n, c0:c0 + width] @ x[f, r0:r0 + width]
            else:
                x[f, r0:r0 + width] = np.linalg.solve(np.triu(T), x[f, r0:r0 + width])
                x[f, :r0] -= A[:r0, c0:c0 + width] @ x[f, r0:r0 + width]
        self.launches += 1; self.prof[5] += 1
        return 0

    def updes_block_gemv(self, h, slot, r0, nrows, c_lo, c_hi, x, out, st):
        A, _ = self._slot(h, slot)
        o = _f64(out, nrows)
        if c_hi <= c_lo:
            o[:] = 0.0
        else:
            o[:] = A[r0:r0 + nrows, c_lo:c_hi] @ _f64(x, c_hi)[c_lo:c_hi]
        self.launches += 1; self.prof[5] += 1
        return 0

    def updes_tri_diag_solve(self, h, slot, upper, r0, c0, width, X, st):
        A, _ = self._slot(h, slot)
        T = A[r0:r0 + width, c0:c0 + width]
        x = _f64(X, r0 + width)
        x[r0:r0 + width] = np.linalg.solve(np.triu(T) if upper else np.tril(T, -1) + np.eye(width), x[r0:r0 + width])
        self.launches += 1; self.prof[5] += 1
        return 0

    def updes_row_absmax(self, A, rows, cols, ld, out, st):
        a = _f64(A, rows * ld).reshape(rows, ld)
        _f64(out, rows)[:] = np.abs(a[:, :cols]).max(axis=1)
        self.launches += 1
        return 0

    def updes_scale_from_absmax(self, absmax, n, scale, st):
        m = _f64(absmax, n).copy()
        sc = np.ones(n)
        ok = (m > 0) & np.isfinite(m)
        sc[ok] = np.ldexp(1.0, 1 - np.frexp(m[ok])[1])
        _f64(scale, n)[:] = sc
        self.launches += 1
        return 0

    def updes_row_scale(self, A, rows, cols, ld, scale, st):
        a = _f64(A, rows * ld).reshape(rows, ld)
        a[:, :cols] *= _f64(scale, rows)[:, None]
        self.launches += 1
        return 0

    # ---- misc ---------------------------------------------------------------------------------------------------
    def updes_launch_count(self):
        return self.launches

    def updes_profile_enable(self, on):
        if on:
            self.prof = [0] * 6
        return 0

    def updes_profile_read(self, cat, ms, work, cnt):
        ms._obj.value, work._obj.value, cnt._obj.value = 1e-3 * self.prof[cat], 1.0 * self.prof[cat], self.prof[cat]
        return 0

    def updes_profile_records(self, cat, ms, work, limit):
        return 0

    def updes_b200_version(self):
        return b"cpu-emulation (tests only)"


_installed = None


def install():
    """Route updes_b200._lib to the emulation and map "cuda" onto CPU memory.  Process-wide and irreversible:
    call it only in a process dedicated to this purpose."""
    global _installed
    if _installed is not None:
        return _installed
    import torch
    if torch.cuda.is_available():
        raise RuntimeError("cpu_abi_emulation: a CUDA device is present -- the emulation exists for GPU-less containers only "
                           "and must never stand in for libupdes_b200.so where the real path can run")
    from updes_b200 import _lib
    emu = EmulatedLib()
    _lib.load = lambda: emu
    _lib._lib = emu
    _lib.require_cuda = lambda: torch
    _lib.stream_ptr = lambda: None

    def is_cuda_dev(d):
        return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")

    def strip(args, kwargs):
        if is_cuda_dev(kwargs.get("device")):
            kwargs["device"] = "cpu"
        args = tuple("cpu" if is_cuda_dev(a) else a for a in args)
        return args, kwargs

    def wrap(fn):
        def f(*a, **k):
            a, k = strip(a, k)
            return fn(*a, **k)
        return f

    for name in ("zeros", "empty", "ones", "full", "arange", "tensor", "as_tensor", "zeros_like", "empty_like", "eye",
                 "rand", "randn", "linspace"):
        setattr(torch, name, wrap(getattr(torch, name)))
    torch.Tensor.to = wrap(torch.Tensor.to)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.device_count = lambda: int(os.environ.get("UPDES_EMULATED_GPUS", "1"))
    torch.cuda.current_device = lambda: 0

    class _Props:
        multi_processor_count = 148
        total_memory = 180 << 30
        name = "emulated"

    torch.cuda.get_device_properties = lambda *a, **k: _Props()
    import torch.distributed as dist
    real_init = dist.init_process_group

    def init_process_group(backend=None, *a, **k):
        k.pop("device_id", None)
        return real_init("gloo", *a, **k)            # NCCL collectives -> gloo on CPU tensors

    dist.init_process_group = init_process_group
    torch.cuda.mem_get_info = lambda *a, **k: (160 << 30, 180 << 30)
    def live_tensor_bytes(*a, **k):
        # stands in for the caching allocator's counter: bytes of all distinct live tensor storages
        import gc
        gc.collect()
        seen, total = set(), 0
        for o in gc.get_objects():
            try:
                if isinstance(o, torch.Tensor):
                    st = o.untyped_storage()
                    if st.data_ptr() not in seen:
                        seen.add(st.data_ptr())
                        total += st.nbytes()
            except Exception:
                pass
        return total

    torch.cuda.memory_allocated = live_tensor_bytes
    torch.cuda.max_memory_allocated = lambda *a, **k: 0
    torch.cuda.reset_peak_memory_stats = lambda *a, **k: None

    class _Event:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self, *a):
            import time
            self.t = time.perf_counter()

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    torch.cuda.Event = _Event
    _installed = emu
    return emu


def self_check():
    """The emulation's closed forms against the oracle's C restatement (independent code), all kernels."""
    from oracle import oracle as O
    O.build()
    rng = np.random.default_rng(0)
    pts, ctr = rng.random((7, 2)), rng.random((9, 2))
    ctr[3] = pts[2]
    worst = 0.0
    for name, code, params in (("polyharmonic", 0, (0, 1, 2, 3)), ("thin_plate", 1, (1, 2, 3)), ("gaussian", 2, (0.5, 3.0)),
                               ("multiquadric", 3, (1.0, 2.5)), ("inverse_multiquadric", 4, (1.0, 2.5))):
        for p in params:
            J = rbf_jets(code, float(p), pts, ctr)
            for i in range(7):
                for j in range(9):
                    ref = O.rbf_jet(name, float(p), pts[i], ctr[j])
                    if i == 2 and j == 3:
                        ref = np.array([PHI0[code], 0, 0, 0, 0.0])      # nan_to_num at r = 0 (Q4)
                    worst = max(worst, float(np.max(np.abs(J[i, j] - ref) / (1.0 + np.abs(ref)))))
    Q = monomial_jets(15, pts)
    for i in range(7):
        for m in range(15):
            worst = max(worst, float(np.max(np.abs(Q[i, m] - O.monomial_jet(m, pts[i])))))
    return worst


if __name__ == "__main__":
    print("closed forms vs oracle C: max scaled difference %.2e" % self_check())
