#!/usr/bin/env python
"""Build and run the reference-side jax.ffi adapter (integration/updes_jax_ffi.cc + integration/updes_jax.py) WITHOUT
JAX: the handlers are compiled against a mock of XLA's FFI binding API (tests/mock_xla/xla/ffi/api/ffi.h) and driven by
a jax.ffi stand-in (oracle/refshim/jax/ffi.py) that does what XLA does around a custom call.  Test infrastructure.

    python tests/run_jax_adapter.py --emulated    # build container: the handlers' updes_* calls are forwarded to the
                                                  # CPU emulation of the C-ABI (tests/cpu_abi_emulation.py)
    python tests/run_jax_adapter.py               # B200 box: handlers linked against libupdes_b200.so, CUDA buffers

Checks the adapter's pde_solver against updes_b200.pde_solver_jit and against what the reference's own pde_solver_jit
returned (tests/golden/ref_config1_30x20.npz, ref_robin_11x8.npz)."""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
MOCK = os.path.join(ROOT, "tests", "mock_xla")
SRC = os.path.join(ROOT, "integration", "updes_jax_ffi.cc")


def build(outdir, emulated):
    so = os.path.join(outdir, "libupdes_jax_ffi.so")
    inc = ["-I", MOCK, "-I", os.path.join(ROOT, "include")]
    if emulated:
        tramp = os.path.join(outdir, "tramp.o")
        subprocess.check_call(["gcc", "-O1", "-fPIC", "-c", "-I", os.path.join(ROOT, "include"),
                               os.path.join(MOCK, "cpu", "abi_trampolines.c"), "-o", tramp])
        cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", os.path.join(MOCK, "cpu")] + inc + [SRC, tramp, "-o", so]
    else:
        cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        libdir = os.path.join(ROOT, "updes_b200")
        cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", os.path.join(cuda, "include")] + inc + \
              [SRC, "-L", libdir, "-lupdes_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
               "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", so]
    subprocess.check_call(cmd)
    return so


def forward_to_emulation(so_path, emu):
    """Point the trampolines at the emulated entry points (ctypes callbacks with the header's signatures)."""
    from updes_b200._lib import UpdesRows
    so = ctypes.CDLL(so_path)
    C, VP, I64, I32, DBL = ctypes.CFUNCTYPE, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    RowsP = ctypes.POINTER(UpdesRows)

    def lu_create(href, n, ld):
        class _Ref:                      # what ctypes.byref(handle) looks like to the emulation
            _obj = ctypes.c_void_p()
        rc = emu.updes_lu_create(_Ref, n, ld)
        ctypes.cast(href, ctypes.POINTER(ctypes.c_void_p))[0] = _Ref._obj.value
        return rc

    cbs = [
        C(I32, I32, DBL, I32, I32, VP, RowsP, I64, I64, I32, VP, I64, VP)(
            lambda k, p, N, M, ctr, rows, r0, nr, mask, out, ld, st: emu.updes_assemble_rows(k, p, N, M, ctr, rows.contents, r0, nr, mask, out, ld, st)),
        C(I32, VP, I64, I64)(lu_create),
        C(I32, VP)(lambda h: emu.updes_lu_destroy(h)),
        C(I32, VP, VP, VP, VP, VP)(lambda h, K, ipiv, info, st: emu.updes_lu_factor(h, K, ipiv, info, st)),
        C(I32, VP, VP, VP, VP, I64, I32, I32, VP)(lambda h, LU, ipiv, B, ldb, nrhs, tr, st: emu.updes_lu_solve(h, LU, ipiv, B, ldb, nrhs, tr, st)),
        C(ctypes.c_size_t, I32, I32, I32)(lambda N, R, nf: emu.updes_eval_jets_workspace_bytes(N, R, nf)),
        C(I32, I32, DBL, I32, I32, VP, VP, I64, I32, VP, I32, VP, VP, VP, VP, VP)(
            lambda k, p, N, M, ctr, cf, ldc, nf, pts, npts, skip, jphi, jpol, ws, st: emu.updes_eval_jets(k, p, N, M, ctr, cf, ldc, nf, pts, npts, skip, jphi, jpol, ws, st)),
    ]
    so.updes_mock_register.argtypes = [VP] * 7
    so.updes_mock_register(*[ctypes.cast(c, VP) for c in cbs])
    return so, cbs          # keep the callbacks alive


def main():
    emulated = "--emulated" in sys.argv
    keep = None
    with tempfile.TemporaryDirectory() as td:
        so_path = build(td, emulated)
        if emulated:
            import cpu_abi_emulation
            emu = cpu_abi_emulation.install()
            keep = forward_to_emulation(so_path, emu)
        os.environ["UPDES_JAX_FFI_LIB"] = so_path
        sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))        # `import jax` -> the stand-in
        sys.path.insert(0, os.path.join(ROOT, "integration"))
        import updes_jax as J
        import updes_b200 as u
        from functools import partial
        to_np = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)

        # ---- config 1 (README Laplace, 30x20) -------------------------------------------------------------------
        cloud = u.SquareCloud(Nx=30, Ny=20, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
        op = lambda x, center, rbf, monomial, fields: u.nodal_laplacian(x, center, rbf, monomial)
        rhs = lambda x, centers, rbf, fields: 0.0
        bcs = {"South": lambda c: 0.0, "West": lambda c: 0.0, "North": lambda c: np.sin(np.pi * c[0]), "East": lambda c: 0.0}
        a = J.pde_solver(op, rhs, cloud, bcs, u.polyharmonic, 1)
        b = u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1)
        g = np.load(os.path.join(ROOT, "tests", "golden", "ref_config1_30x20.npz"))
        scale = np.max(np.abs(g["vals"]))
        d_prod = float(np.max(np.abs(to_np(a.vals) - b.vals)) / scale)
        d_ref = float(np.max(np.abs(to_np(a.vals) - g["vals"])) / scale)
        print("config 1: adapter vs product %.2e, adapter vs the reference's own result %.2e" % (d_prod, d_ref))
        assert d_prod <= 1e-8 and d_ref <= 1e-8
        assert to_np(a.coeffs).shape == (cloud.N + 3,) and a.mat is None

        # ---- phi(0) != 0 kernel (gaussian), Robin + Neumann, degree 2: exercises the self-term correction of vals --------
        cloud = u.SquareCloud(Nx=11, Ny=8, facet_types={"South": "n", "West": "d", "North": "r", "East": "d"})
        rbf = partial(u.gaussian, eps=3.0)
        op2 = lambda x, center, rbf, monomial, fields: u.nodal_laplacian(x, center, rbf, monomial) - 2.0 * u.nodal_value(x, center, rbf, monomial)
        rhs2 = lambda x, centers, rbf, fields: 1.0
        bcs2 = {"South": lambda c: 0.0, "West": lambda c: c[1], "North": (lambda c: 1.0, lambda c: 2.0 + c[0]), "East": lambda c: 0.5}
        a = J.pde_solver(op2, rhs2, cloud, bcs2, rbf, 2)
        b = u.pde_solver_jit(op2, rhs2, cloud, bcs2, rbf, 2)
        d = float(np.max(np.abs(to_np(a.vals) - b.vals)) / np.max(np.abs(b.vals)))
        dc = float(np.max(np.abs(to_np(a.coeffs) - b.coeffs)) / np.max(np.abs(b.coeffs)))
        print("gaussian Robin 11x8: adapter vs product vals %.2e, coeffs %.2e" % (d, dc))
        # cond(K) ~ 1e10 here and the adapter solves without equilibration / refinement: agreement to ~cond * eps.  What this
        # case guards is the phi(0) self-term correction of vals -- dropping it changes vals by O(max |c_i|), i.e. O(1) relative
        assert np.max(np.abs(b.coeffs[: cloud.N])) > 1e-2 * np.max(np.abs(b.vals))      # so a wrong self term shows at O(1e-2) >> 1e-4
        assert d <= 1e-4 and dc <= 1e-4

        # ---- a mis-bound call must be refused by the frame check, not crash ---------------------------------------------
        import jax
        try:
            jax.ffi.ffi_call("UpdesEvalJets", jax.ShapeDtypeStruct((1,), jax.numpy.float64))(jax.numpy.zeros((2, 2)), kind=np.int32(0))
            raise AssertionError("a call with the wrong arity was accepted")
        except RuntimeError as e:
            assert "binding" in str(e)
        print("jax.ffi adapter ok (%s)" % ("emulated C-ABI, CPU" if emulated else "libupdes_b200.so, CUDA"))
    return keep


if __name__ == "__main__":
    main()
