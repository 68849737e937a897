"""Runs ONE hot kernel a few times so that `ncu -k regex:<name> -s 1 -c 1` captures a warm, representative launch.
  gemm   dgemm_sub_kernel, default variant, LU-representative trailing update: m = n = 32768 - 2048, k = 2048
  gemm1k the same with k = 1024 (the inner dimension of the 8-GPU column blocks)
  sweep  tri_sweep2_kernel (row-block streaming triangular sweep), n = 60000
  asm    assemble_phi_kernel, 200x200 cloud: Laplace rows (closed-form radial Laplacian) then a general 5-term jet
  panel  lu_panel2_kernel, one 32-column base panel over 90 003 rows
  trsm   trsm_base_big_kernel, 128-row block against 8192 columns
Profiling helper for profiles/, not a bench."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import updes_b200 as u
from updes_b200 import _lib, assembly as asm
from updes_b200.assembly import padded_ld
from updes_b200.linalg import LUFactorization

what = sys.argv[1]
reps = 3
if what in ("gemm", "gemm1k"):
    n, k = 32768, (2048 if what == "gemm" else 1024)
    K = torch.randn((n, padded_ld(n)), dtype=torch.float64, device="cuda")
    lu = LUFactorization(K, n)
    for _ in range(reps):
        lu.gemm_sub(k, k, k, 0, 0, k, n - k, n - k, k)
elif what == "sweep":
    n = 60000
    K = torch.randn((n, padded_ld(n)), dtype=torch.float64, device="cuda")
    K[:, :n] += n ** 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    lu = LUFactorization(K, n).factor()
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        lu.solve(x)
elif what == "asm":
    cloud = u.SquareCloud(Nx=200, Ny=200, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
    n = cloud.N + 3
    K = torch.empty((n, padded_ld(n)), dtype=torch.float64, device="cuda")
    for coefrow in ([0, 0, 0, 1.0, 1.0], [1e4, 100.0, 3.0, -0.08, -0.05]):
        rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, np.tile(coefrow, (cloud.Ni, 1))))
        for _ in range(reps):
            asm.assemble_system(rows, "polyharmonic", 1.0, 3, out=K)
elif what == "panel":
    n = 90003
    K = torch.randn((n, 64), dtype=torch.float64, device="cuda")
    K0 = K.clone()
    import ctypes
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.updes_lu_create(ctypes.byref(h), n, 64), "create")
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        K.copy_(K0)
        _lib.check(lib.updes_lu_panel(h, K.data_ptr(), 0, 32, ipiv.data_ptr(), info.data_ptr(), _lib.stream_ptr()), "panel")
    torch.cuda.synchronize()
    lib.updes_lu_destroy(h)
elif what == "trsm":
    n = 8192 + 128
    K = torch.randn((n, padded_ld(n)), dtype=torch.float64, device="cuda")
    lu = LUFactorization(K, n)
    import ctypes
    lib = _lib.load()
    for _ in range(reps):
        _lib.check(lib.updes_lu_bind(lu._handle, 0, K.data_ptr(), n, K.shape[1]), "bind")
        _lib.check(lib.updes_lu_trsm(lu._handle, 0, 0, 0, 128, 0, 0, 128, 8192, _lib.stream_ptr()), "trsm")
torch.cuda.synchronize()
print("done", what)
