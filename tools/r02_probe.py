"""Round-2 kernel probe (GPU box): head-to-head timings of the kernel generations, CUDA events, warm.
  asm     assembly at nx (Laplace rows and a general 5-term jet), columns-per-thread variants
  solve   triangular sweeps, solve_variant 2 / 1 at n
  lu      whole LU at several n, panel_variant 2 / 1, with the per-class breakdown, beside cuSOLVER Dgetrf
Writes one JSON object to stdout.  Exploration / evidence for profiles/, not a bench line."""
import ctypes
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import updes_b200 as u
from updes_b200 import _lib, assembly as asm
from updes_b200.assembly import padded_ld
from updes_b200.linalg import LUFactorization

FACETS = {"South": "n", "West": "d", "North": "d", "East": "d"}


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def probe_asm(nx):
    lib = _lib.load()
    cloud = u.SquareCloud(Nx=nx, Ny=nx, facet_types=FACETS)
    n = cloud.N + 3
    out = {"nx": nx, "n": n}
    K = torch.empty((n, padded_ld(n)), dtype=torch.float64, device="cuda")
    for name, coefrow, kind, param in (("laplace_r3", [0, 0, 0, 1.0, 1.0], "polyharmonic", 1.0),
                                        ("helmholtz_r3", [2.0, 0, 0, 1.0, 1.0], "polyharmonic", 1.0),
                                        ("general_r3", [1e4, 100.0, 3.0, -0.08, -0.05], "polyharmonic", 1.0),
                                        ("navier_stokes_r3", [0.0, 1.0, 0.5, -0.01, -0.01], "polyharmonic", 1.0),
                                        ("general_r5", [1e4, 100.0, 3.0, -0.08, -0.05], "polyharmonic", 2.0),
                                        ("general_gaussian", [1e4, 100.0, 3.0, -0.08, -0.05], "gaussian", 3.0),
                                        ("laplace_gaussian", [0, 0, 0, 1.0, 1.0], "gaussian", 3.0),
                                        ("general_mq", [1e4, 100.0, 3.0, -0.08, -0.05], "multiquadric", 2.0)):
        rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, np.tile(coefrow, (cloud.Ni, 1))))
        for variant in (0, 3):
            lib.updes_assemble_set_variant(variant)
            ms = timeit(lambda: asm.assemble_system(rows, kind, param, 3, out=K))
            out["%s_cpt%s" % (name, "2/4" if variant == 0 else "4/2")] = {"ms": round(ms, 3), "gbs": round(8.0 * n * n / ms * 1e-6, 1)}
        lib.updes_assemble_set_variant(0)
    return out


def probe_solve(n):
    K = torch.randn((n, padded_ld(n)), dtype=torch.float64, device="cuda")
    K[:, :n] += n ** 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    lu = LUFactorization(K, n).factor()
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    out = {"n": n}
    ref = None
    for variant in (2, 1):
        lu.set_solve_variant(variant)
        y = x.clone()
        ms = timeit(lambda: lu.solve(y.copy_(x)))
        out["variant%d" % variant] = {"ms": round(ms, 3), "gbs": round(8.0 * n * n / ms * 1e-6, 1)}
        if ref is None:
            ref = y.clone()
        else:
            out["variants_agree_rel"] = float((y - ref).abs().max() / ref.abs().max())
    x4 = torch.randn((4, n), dtype=torch.float64, device="cuda")
    lu.set_solve_variant(2)
    y4 = x4.clone()
    ms = timeit(lambda: lu.solve(y4.copy_(x4)))
    out["variant2_4rhs"] = {"ms": round(ms, 3)}
    out["status"] = lu.check()
    return out


def cusolver_getrf_ms(n):
    A = torch.randn((n, n), dtype=torch.float64, device="cuda")
    lib = ctypes.CDLL("libcusolver.so.11")
    h = ctypes.c_void_p()
    lib.cusolverDnCreate(ctypes.byref(h))
    lib.cusolverDnSetStream(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    lwork = ctypes.c_int(0)
    vp = ctypes.c_void_p
    lib.cusolverDnDgetrf_bufferSize(h, n, n, vp(A.data_ptr()), n, ctypes.byref(lwork))
    work = torch.empty(max(lwork.value, 1), dtype=torch.float64, device="cuda")
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    A0 = A.clone()

    def run():
        A.copy_(A0)
        lib.cusolverDnDgetrf(h, n, n, vp(A.data_ptr()), n, vp(work.data_ptr()), vp(ipiv.data_ptr()), vp(info.data_ptr()))
    t_copy = timeit(lambda: A.copy_(A0))
    ms = timeit(run) - t_copy
    lib.cusolverDnDestroy(h)
    return ms


def probe_lu(n, with_cusolver=True):
    g = torch.Generator(device="cuda").manual_seed(n)
    A = torch.randn((n, padded_ld(n)), dtype=torch.float64, device="cuda", generator=g)
    K = torch.empty_like(A)
    out = {"n": n}
    for variant in (2, 1):
        lu = LUFactorization(K, n)
        lu.set_panel_variant(variant)

        def run():
            K.copy_(A)
            lu.factor()
        t_copy = timeit(lambda: K.copy_(A))
        ms = timeit(run) - t_copy
        _lib.profile_enable(True)
        K.copy_(A); lu.factor(); torch.cuda.synchronize()
        prof = {k: (round(_lib.profile_read(k)[0], 2), _lib.profile_read(k)[2]) for k in ("gemm", "panel", "swap", "trsm")}
        _lib.profile_enable(False)
        out["panel_variant%d" % variant] = {"lu_ms": round(ms, 2), "tflops": round(2 / 3 * n ** 3 / ms * 1e-9, 2), "status": lu.check(),
                                            "class_ms_launches": prof}
        del lu
    if with_cusolver:
        ms = cusolver_getrf_ms(n)
        out["cusolver_getrf"] = {"ms": round(ms, 2), "tflops": round(2 / 3 * n ** 3 / ms * 1e-9, 2)}
    return out


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = {}
    if what in ("asm", "all"):
        res["asm"] = [probe_asm(nx) for nx in (200,)]
    if what in ("solve", "all"):
        res["solve"] = [probe_solve(n) for n in (20000, 60000)]
    if what in ("lu", "all"):
        res["lu"] = [probe_lu(n) for n in (1388, 4096, 8192, 16384, 32768)]
    print(json.dumps(res))
