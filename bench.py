#!/usr/bin/env python
"""Benchmark of the Updes global RBF-collocation hot path on B200 (contract: task brief, section 4).

A *step* is one pass of the hot path over the synthetic Laplace SquareCloud of BASELINE.json
(configs[3]: 300x300 = 90 000 nodes, polyharmonic a=1, max_degree=1, n = 90 003): dense assembly of
the collocation system in HBM, in-place LU with partial pivoting, one right-hand-side solve.

  metric   assemble_lu_solve_fp64_tflops = (2/3 n^3 flop) / (assemble + LU + solve seconds), whole job;
           ms_per_step carries BASELINE.json's "seconds at N=90k" directly.
  value    inputs (node coordinates, row descriptors, right-hand side) already resident in HBM.
  e2e      the same pass through the public API pde_solver_jit with HOST (numpy) inputs and outputs:
           operator lowering, row-descriptor upload, assembly, row equilibration, LU, solve + one refinement
           step, solution download.
  roofline the dominant kernel (the DMMA trailing-update GEMM): algorithmic flops / its summed launch
           durations, measured with CUDA events on the launching stream inside the timed region.
  library_baseline  cuSOLVER Dgetrf + Dgetrs at the same n on the same GPU in the same run (N = 1).
  small_configs     BASELINE.json configs 1-3 (the sizes Updes users actually run) end to end, in ms.
  cpu_baseline / --impl reference: the reference *formulation* (inv(A), B = D inv(A), QR, inv(A)[u;0];
           updes/assembly.py:366-410, operators.py:602-616) restated on the CPU oracle with LAPACK on all
           host cores, on bounded samples, in the same unit (2/3 n_s^3)/seconds; N in {600, 2 500, 10 000}
           are timed and the 90k figure is EXTRAPOLATED from the largest (n^2 and n^3 parts scaled separately).
           JAX is not installed in this image, so this is the oracle port, not the reference package.

Multi-GPU (--gpus N, one process per GPU under torchrun): STRONG scaling of the same 90k-node problem
("90 k fits one GPU => use it for the 1/2/4/8 strong-scaling series"), sharded
column-block-cyclically (updes_b200/distributed.py): assembly of the owned column blocks (no communication),
LU with one NCCL broadcast per panel and look-ahead, distributed solve.  At N = 8 the line also carries
`config5`: one timed pass of BASELINE.json's 500x500 / 250k-node problem (500 GB, 62.5 GB per GPU).
"""
import os
import sys


def _restore_host_threads_for_the_cpu_arm():
    """torchrun exports OMP_NUM_THREADS=1 before Python starts, and an already-initialised BLAS ignores later
    changes (round 1: the reference arm ran 2x slower at N >= 2).  The reference arm only needs rank 0 and the
    host cores: non-zero ranks exit at once, rank 0 re-executes itself with the thread limits lifted."""
    if "--impl" not in sys.argv or "reference" not in sys.argv:
        return
    if int(os.environ.get("RANK", "0")) != 0:
        sys.exit(0)
    if os.environ.get("UPDES_BENCH_REEXEC") == "1":
        return
    ncpu = str(os.cpu_count() or 1)
    if os.environ.get("OMP_NUM_THREADS", ncpu) != ncpu:
        env = dict(os.environ)
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            env[k] = ncpu
        env["UPDES_BENCH_REEXEC"] = "1"
        os.execve(sys.executable, [sys.executable] + sys.argv, env)


_restore_host_threads_for_the_cpu_arm()

import argparse  # noqa: E402
import ctypes  # noqa: E402
import json  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import tempfile  # noqa: E402
import time  # noqa: E402
import traceback  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

FACETS = {"South": "n", "West": "d", "North": "d", "East": "d"}
METRIC, UNIT = "assemble_lu_solve_fp64_tflops", "TFLOP/s"
HEADLINE_NX = 300


def lu_flops(n):
    return 2.0 / 3.0 * float(n) ** 3


def workload_name(nx):
    return "Synthetic Laplace SquareCloud %dx%d (%d nodes), polyharmonic a=1, max_degree=1, FP64" % (nx, nx, nx * nx)


# ------------------------------------------------------------------------------------------------
# CPU leg: the reference formulation on the oracle (bounded samples)
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(nx, ny=None, parts=None):
    from oracle import oracle as O
    ny = ny or nx
    cloud = O.RefSquareCloud(nx, ny, FACETS) if nx * ny <= 1600 else _fast_ref_cloud(nx, ny)
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    xy = cloud.sorted_nodes
    t0 = time.perf_counter()
    bc = {f: (np.sin(np.pi * xy[ids, 0]) if f == "North" else np.zeros(len(ids))) for f, ids in cloud.facet_nodes.items()}
    q = O.assemble_q(cloud, np.zeros(cloud.Ni), bc)
    vals, coeffs, _ = O.reference_solve(cloud, "polyharmonic", 1.0, 1, coef, q, timings=parts)
    dt = time.perf_counter() - t0
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    return dt, cloud.N + 3, float(np.max(np.abs(vals - exact)))


def _fast_ref_cloud(nx, ny):
    """Cloud arrays for the CPU leg at sizes where the oracle's literal dict loops are slow: the
    product's vectorised SquareCloud yields identical arrays (tests/test_host.py checks that)."""
    import updes_b200 as u
    return u.SquareCloud(Nx=nx, Ny=ny, facet_types=FACETS)


def use_all_host_threads():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(os.cpu_count() or 1)
    except Exception:
        pass


def blas_threads():
    try:
        import numpy  # noqa: F401  (make sure the BLAS is loaded before asking)
        import scipy.linalg  # noqa: F401
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_size_sweep(sizes, n_target):
    """time the reference formulation at a few sizes and extrapolate to the headline n.  The pass
    has an O(n^2) part (assembling diffMat and A, single-threaded closed forms) and an O(n^3) part (inv, GEMM, QR
    in LAPACK on all host threads), timed separately; each is scaled from the LARGEST measured size with its own
    exponent (t = t_asm (n/n_s)^2 + t_linalg (n/n_s)^3).  A free least-squares fit of a n^3 + b n^2 through three
    points is not used: LAPACK's efficiency is still rising below n ~ 10^4, which such a fit books as an n^2 term
    and then under-predicts the 90k time several-fold.  Returns (records, extrapolated seconds, model)."""
    recs = []
    for nx, ny in sizes:
        parts = {}
        dt, n_s, err = cpu_reference_pass(nx, ny, parts)
        lin = parts.get("linalg_s", dt)
        recs.append({"cloud": "%dx%d" % (nx, ny), "n": n_s, "seconds": dt, "assemble_seconds": parts.get("assemble_s", 0.0),
                     "linalg_seconds": lin, "linalg_gflops": 20.0 / 3.0 * float(n_s) ** 3 / lin * 1e-9,
                     "max_err_vs_analytic": err})
    big = max(recs, key=lambda r: r["n"])
    s = float(n_target) / big["n"]
    other = big["seconds"] - big["linalg_seconds"]
    t = other * s ** 2 + big["linalg_seconds"] * s ** 3
    model = {"from": big["cloud"], "n2_seconds_at_sample": other, "n3_seconds_at_sample": big["linalg_seconds"],
             "formula": "t(n) = n2_seconds (n/n_s)^2 + n3_seconds (n/n_s)^3",
             "linalg_flops_counted": "20/3 n^3 (inv 2, D inv(A) 2, QR with explicit Q 8/3)"}
    return recs, t, model


def parse_sizes(txt):
    out = []
    for tok in txt.split(","):
        tok = tok.strip()
        if tok:
            a, b = tok.split("x")
            out.append((int(a), int(b)))
    return out


def run_reference_arm(args):
    from oracle import oracle as O
    O.build()
    use_all_host_threads()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nx = args.cpu_nx
    for _ in range(args.warmup):
        cpu_reference_pass(min(nx, 30))
    times = []
    for _ in range(args.steps):
        dt, n_s, err = cpu_reference_pass(nx)
        times.append(dt)
    t = sum(times) / len(times)
    val = lu_flops(n_s) / t * 1e-12
    n_head = HEADLINE_NX * HEADLINE_NX + 3
    sweep = None
    if not args.no_cpu_sweep:
        recs, t90, fit = cpu_size_sweep(parse_sizes(args.cpu_sizes), n_head)
        sweep = {"measured": recs, "extrapolation_model": fit, "extrapolated_seconds_at_n_%d" % n_head: t90,
                 "extrapolated_value_at_headline_n": lu_flops(n_head) / t90 * 1e-12,
                 "note": "EXTRAPOLATED: the reference formulation needs ~4 x 65 GB of host RAM at n = 90 003 and cannot run there"}
    sample = "reference formulation (inv+GEMM+QR, oracle port with LAPACK) on SquareCloud %dx%d, n=%d; %.2f s per pass; " \
             "rate = (2/3 n^3)/t, the same normalisation as the GPU arm" % (nx, nx, n_s, t)
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.nx or HEADLINE_NX), "sample": "SquareCloud %dx%d" % (nx, nx),
                       "host_threads_env": os.environ.get("OMP_NUM_THREADS"), "launched_under_torchrun_world": world},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": blas_threads(), "kind": "port", "sample": sample,
                             "size_sweep": sweep},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.split(",") for l in open(self.f.name).read().splitlines() if l.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any("Active" == r[5 + k].strip() for r in rows)]
        out.update(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=float(rows[0][2]), reasons=reasons,
                   power_w_max=max(float(r[3]) for r in rows), samples=len(rows))
        return out


class CusolverLU:
    """Library comparator (NOT the product): cuSOLVER Dgetrf / Dgetrs called in place on a caller-owned
    column-major matrix, on torch's current stream."""

    def __init__(self):
        import torch
        self.torch = torch
        torch.linalg.lu_factor(torch.eye(4, dtype=torch.float64, device="cuda"))     # makes torch load its libcusolver
        self.lib = lib = ctypes.CDLL("libcusolver.so.11")
        self.h = ctypes.c_void_p()
        assert lib.cusolverDnCreate(ctypes.byref(self.h)) == 0
        assert lib.cusolverDnSetStream(self.h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0

    def factor_solve(self, A, n, lda, b):
        """A: CUDA float64 buffer holding an n x n column-major matrix with leading dimension lda; b: (n,) rhs.
        Returns (getrf ms, getrs ms)."""
        torch, lib = self.torch, self.lib
        lwork = ctypes.c_int(0)
        vp = ctypes.c_void_p
        assert lib.cusolverDnDgetrf_bufferSize(self.h, n, n, vp(A.data_ptr()), lda, ctypes.byref(lwork)) == 0
        work = torch.empty(max(lwork.value, 1), dtype=torch.float64, device="cuda")
        ipiv = torch.empty(n, dtype=torch.int32, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        rc = lib.cusolverDnDgetrf(self.h, n, n, vp(A.data_ptr()), lda, vp(work.data_ptr()), vp(ipiv.data_ptr()), vp(info.data_ptr()))
        e[1].record()
        rc2 = lib.cusolverDnDgetrs(self.h, 0, n, 1, vp(A.data_ptr()), lda, vp(ipiv.data_ptr()), vp(b.data_ptr()), n, vp(info.data_ptr()))
        e[2].record()
        torch.cuda.synchronize()
        assert rc == 0 and rc2 == 0 and int(info.item()) == 0, (rc, rc2, int(info.item()))
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), int(lwork.value)

    def close(self):
        self.lib.cusolverDnDestroy(self.h)


def fp64_peak():
    peak, src = 37.0, "B200 FP64 tensor nominal 37 TF (fallback)"
    try:
        c = json.load(open(os.path.join(ROOT, "profiles", "r01_ceilings.json")))
        peak = float(c["dmma_tflops_8warps"])
        src = "measured on this pool: raw DMMA issue rate %.2f TF (profiles/r01_ceilings.json; cuBLAS DGEMM %.2f TF); " \
              "MEASURED_PEAKS.json has no FP64 figure" % (peak, c.get("cublas_dgemm_tflops_n16384", 0))
    except Exception:
        pass
    return peak, src


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json (driver-written)"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md"


def gemm_traffic():
    """dram bytes per launch of the CURRENT default GEMM at an LU-representative launch, from the committed ncu summary."""
    for name in ("r02_ncu_summary.json", "r01_ncu_summary.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            return d.get("dgemm_traffic_bytes_per_launch"), "profiles/" + name + ": " + str(d.get("dgemm_traffic_note", ""))
        except Exception:
            continue
    return None, None


def headline_problem(u, asm, nx):
    cloud = u.SquareCloud(Nx=nx, Ny=nx, facet_types=FACETS)
    M = 3
    n = cloud.N + M
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    table = asm.build_operator_rows(cloud, coef)
    rows = asm.DeviceRows(cloud, table)
    xy = cloud.sorted_nodes
    q = np.zeros(n)
    north = np.asarray(cloud.facet_nodes["North"])
    q[north] = np.sin(np.pi * xy[north, 0])
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    return cloud, M, n, table, rows, q, exact


def api_problem(u, cloud):
    xy = cloud.sorted_nodes
    north = np.asarray(cloud.facet_nodes["North"])
    op = lambda xx, center, rbf, monomial, fields: u.nodal_laplacian(xx, center, rbf, monomial)
    rhs = lambda xx, centers, rbf, fields: 0.0
    bcs = {"South": np.zeros(len(cloud.facet_nodes["South"])), "West": np.zeros(len(cloud.facet_nodes["West"])),
           "North": np.sin(np.pi * xy[north, 0]), "East": np.zeros(len(cloud.facet_nodes["East"]))}
    return op, rhs, bcs


def small_configs(u, torch):
    """BASELINE.json configs 1-3 end to end through the public API (numpy in, numpy out), median wall-clock ms."""
    import configs
    from updes_b200 import _lib

    def timed(fn, reps, before=None):
        """median wall-clock ms of fn(); `before` (dropping the factor cache: cudaFree of cached blocks, whose cost
        depends on what else the process has allocated) runs outside the timed region"""
        ts = []
        for _ in range(reps):
            if before:
                before()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        return statistics.median(ts)

    out = {}
    cloud, solve = configs.config1(u)

    u.clear_cache()
    solve()
    l0 = _lib.launch_count()
    out["config1_laplace_30x20"] = {"N": cloud.N, "e2e_ms": timed(solve, 5, before=u.clear_cache), "launches": (_lib.launch_count() - l0) // 5,
                                    "what": "lowering + assembly + LU + solve + refinement; the factor cache is emptied before every call"}
    out["config1_laplace_30x20"]["e2e_ms_factor_cached"] = timed(solve, 5)
    cloud, u0, step, _ = configs.config2(u)
    state = {"u": u0}

    def c2_first():
        state["u"] = step(u0).vals

    def c2_step():
        state["u"] = step(state["u"]).vals
    u.clear_cache()
    c2_first()
    out["config2_advdiff_periodic_35x35"] = {"N": cloud.N, "first_step_ms": timed(c2_first, 3, before=u.clear_cache), "per_step_ms_factor_cached": timed(c2_step, 10),
                                             "what": "first step factors K and A; later steps: coefficients of u (A solve), rhs, two sweeps"}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import cloud_from_golden
        cv, _ = cloud_from_golden("mesh_msh_cloud_vel.npz")
        cp, _ = cloud_from_golden("mesh_msh_cloud_phi.npz")

        def c3():
            configs.config3_projection_loop(u, cv, cp, nb_iter=2)
        u.clear_cache()
        c3()
        out["config3_navier_stokes_projection_1385"] = {"N": cv.N, "ms_per_iteration": timed(c3, 3, before=u.clear_cache) / 2.0,
                                                        "what": "u, v (re-assembled + re-factored: the matrix depends on the previous velocity) and phi "
                                                                "(factor cached) solves per iteration, mesh.msh clouds, 2 iterations averaged"}
    except Exception as e:  # the mesh fixture lives under tests/golden
        out["config3_navier_stokes_projection_1385"] = {"unavailable": repr(e)}
    u.clear_cache()
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm, one GPU
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import updes_b200 as u
    from updes_b200 import _lib, assembly as asm
    from updes_b200.linalg import LUFactorization

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL only moves one panel per block step; keep its kernels small so they do not take SMs
        # from the update GEMM they overlap with (the GEMM's dynamic tile scheduler absorbs the rest)
        os.environ.setdefault("NCCL_MAX_CTAS", "8")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if args.grid:
            P, Q = (int(v) for v in args.grid.lower().split("x"))
            if P > 1:
                return run_gpu_arm_grid2d(args, world, rank, local, P, Q)
        return run_gpu_arm_distributed(args, world, rank, local)

    nx = args.nx if args.nx else HEADLINE_NX
    cloud, M, n, table, rows, q, exact = headline_problem(u, asm, nx)
    xy = cloud.sorted_nodes
    b = torch.as_tensor(q).cuda()
    K = torch.empty((n, asm.padded_ld(n)), dtype=torch.float64, device="cuda")
    lu = LUFactorization(K, n)
    if args.trsm_base:
        lu.set_trsm_base(args.trsm_base)
    if args.gemm_variant >= 0:
        lu.set_gemm_variant(args.gemm_variant)
    x = torch.empty_like(b)

    def step():
        asm.assemble_system(rows, "polyharmonic", 1.0, M, out=K)
        lu.factor()
        x.copy_(b)
        lu.solve(x)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    _lib.profile_enable(True)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    prof = {k: _lib.profile_read(k) for k in ("gemm", "panel", "swap", "trsm", "assemble", "solve")}
    _lib.profile_enable(False)
    ms_step = ms_total / args.steps
    value = lu_flops(n) / (ms_step * 1e-3) * 1e-12
    status = lu.check()

    # correctness beside the timing: matrix-free backward error and the analytic solution
    r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
    own = torch.arange(cloud.N, dtype=torch.int32, device="cuda")
    jphi, jpol = asm.eval_jets("polyharmonic", 1.0, rows.centres, x.view(1, -1), rows.centres, own)
    vals = (jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy()
    max_err = float(np.max(np.abs(vals - exact)))
    asm.assemble_system(rows, "polyharmonic", 1.0, M, out=K)
    knorm = float(K[:, :n].abs().sum(dim=1).max().item())
    berr = float(r.abs().max().item() / (knorm * x.abs().max().item() + b.abs().max().item()))

    # ---- library comparator: cuSOLVER on the same buffer (as a column-major matrix: K^T, same size and conditioning) ----
    library = None
    if not args.no_library:
        try:
            cs = CusolverLU()
            bl = b.clone()
            f_ms, s_ms, lwork = cs.factor_solve(K, n, K.shape[1], bl)
            cs.close()
            library = {"what": "cusolverDnDgetrf + cusolverDnDgetrs, same n, same GPU, same run (matrix = K^T in place)",
                       "n": n, "getrf_seconds": f_ms * 1e-3, "getrs_seconds": s_ms * 1e-3, "getrf_tflops": lu_flops(n) / (f_ms * 1e-3) * 1e-12,
                       "ours_lu_seconds": sum(prof[k][0] for k in ("gemm", "panel", "swap", "trsm")) / args.steps * 1e-3,
                       "ours_solve_seconds": prof["solve"][0] / args.steps * 1e-3, "workspace_doubles": lwork}
            del bl
        except Exception as e:
            library = {"unavailable": repr(e)}

    # ---- e2e through the public API with host buffers -----------------------------------------------
    del K, lu
    torch.cuda.empty_cache()
    op, rhs, bcs = api_problem(u, cloud)
    e2e_times = []
    sol = None
    from updes_b200 import operators as ops
    e2e_phases = None
    for i in range(args.e2e_steps + 1):
        sol = None
        u.clear_cache()
        last = i == args.e2e_steps
        ops.TRACE = [] if last else None          # phase marks (a device synchronize each) on the last call only
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops._trace_t[0] = t0
        sol = u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if last:
            e2e_phases = [[k.strip(), round(v * 1e3, 3)] for k, v in ops.TRACE]
            ops.TRACE = None
        if i > 0 or args.e2e_steps == 0:
            e2e_times.append(dt)
    u.clear_cache()
    e2e_s = max(e2e_times)
    e2e_val = lu_flops(n) / e2e_s * 1e-12
    h2d = int(xy.nbytes + sum(getattr(table, k).nbytes for k in ("p1", "p2", "cphi1", "cphi2", "cpol1", "cpol2", "skip")) + 8 * n * 2)
    d2h = int(8 * n + 8 * cloud.N + 4)
    e2e_err = float(np.max(np.abs(sol.vals - exact)))

    small = None if args.no_small else small_configs(u, torch)

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peak, peak_src = fp64_peak()
    hbm, hbm_src = hbm_peak()
    g_ms, g_flops, g_cnt = prof["gemm"]
    gemm_tf = g_flops / (g_ms * 1e-3) * 1e-12 if g_ms > 0 else 0.0
    a_ms, a_bytes, a_cnt = prof["assemble"]
    s_ms, s_bytes, s_cnt = prof["solve"]
    traffic, traffic_src = gemm_traffic()
    roofline = {"kernel": "dgemm_sub_kernel (DMMA m8n8k4 + TMA; <64,4> ping-pong for wide updates), %d launches/step" % (g_cnt // max(args.steps, 1)),
                "bound": "tensor", "achieved": gemm_tf, "peak": peak, "unit": "TFLOP/s", "frac": gemm_tf / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "share_of_step": g_ms / ms_total}
    breakdown = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] // max(args.steps, 1)} for k, v in prof.items()}
    breakdown["assemble"].update(gbs=a_bytes / (a_ms * 1e-3) * 1e-9 if a_ms else None,
                                 frac_of_hbm=(a_bytes / (a_ms * 1e-3) * 1e-9 / hbm) if a_ms else None, hbm_peak=hbm,
                                 hbm_peak_source=hbm_src)
    breakdown["solve"].update(gbs=s_bytes / (s_ms * 1e-3) * 1e-9 if s_ms else None,
                              frac_of_hbm=(s_bytes / (s_ms * 1e-3) * 1e-9 / hbm) if s_ms else None)
    breakdown["lu_tflops"] = lu_flops(n) / (sum(prof[k][0] for k in ("gemm", "panel", "swap", "trsm")) / args.steps * 1e-3) * 1e-12

    cpu = None
    if not args.no_cpu:
        from oracle import oracle as O
        O.build()
        use_all_host_threads()
        cpu_reference_pass(30)
        recs, t90, fit = cpu_size_sweep(parse_sizes(args.cpu_sizes), n)
        big = recs[-1]
        cpu = {"value": lu_flops(big["n"]) / big["seconds"] * 1e-12, "unit": UNIT, "cores": blas_threads(), "kind": "port",
               "sample": "reference formulation (inv+GEMM+QR; oracle port, LAPACK) timed at " +
                         ", ".join("%s (n=%d): %.2f s" % (r["cloud"], r["n"], r["seconds"]) for r in recs) +
                         "; value = (2/3 n^3)/t at the largest; EXTRAPOLATED to n = %d (n^2 part and n^3 part scaled separately from the largest size): %.0f s (%.1f h)" % (n, t90, t90 / 3600),
               "size_sweep": recs, "extrapolation_model": fit, "extrapolated_seconds_at_headline_n": t90,
               "extrapolated_value_at_headline_n": lu_flops(n) / t90 * 1e-12}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(nx), "n": n, "matrix_bytes": 8 * n * n, "lu_flops": lu_flops(n),
                       "parallelism": "1 GPU",
                       "l2": "matrix (%.1f GB) is far larger than L2; no flush needed" % (8 * n * n / 1e9)},
            "seconds_per_step": ms_step * 1e-3,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "seconds_per_step": e2e_s, "api": "updes_b200.pde_solver_jit (numpy in, numpy out; row equilibration + 1 refinement step)",
                    "max_err_vs_analytic": e2e_err, "phases_ms_last_call": e2e_phases},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "breakdown": breakdown,
            "cpu_baseline": cpu, "library_baseline": library, "small_configs": small,
            "correctness": {"backward_error": berr, "max_err_vs_analytic": max_err, "zero_pivot": status}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm, N > 1: one sharded problem
# ------------------------------------------------------------------------------------------------
def _timeline_summary(dist, dlu, world):
    """Per-rank sums of the per-panel phases (ms) of ONE instrumented factorisation, plus the tail where the
    serial panel chain is no longer hidden behind the trailing update."""
    tl = dlu.timeline_ms()
    per_rank = [None] * world
    dist.all_gather_object(per_rank, tl)
    nblk = len(per_rank[0])
    out = {"per_rank_ms": []}
    for r, recs in enumerate(per_rank):
        out["per_rank_ms"].append({"rank": r, "wait_for_panel": sum(x["wait"] for x in recs),
                                   "lookahead_update": sum(x.get("lookahead", 0.0) for x in recs),
                                   "panel_factor_and_pack": sum(x.get("panel", 0.0) for x in recs),
                                   "trailing_update": sum(x["update"] for x in recs)})
    # per panel: the owner's serial chain (look-ahead update + panel factor + pack) vs the longest trailing update of any rank
    chain, upd, exposed = [], [], 0
    for k in range(nblk):
        c = max((recs[k].get("lookahead", 0.0) + recs[k].get("panel", 0.0)) for recs in per_rank)
        m = max(recs[k]["update"] for recs in per_rank)
        chain.append(c); upd.append(m)
        if c > m:
            exposed += 1
    out.update(panels=nblk, owner_chain_ms_total=sum(chain), max_trailing_update_ms_total=sum(upd),
               panels_where_owner_chain_exceeds_update=exposed,
               first_panel={"owner_chain_ms": chain[0], "update_ms": upd[0]},
               last_panels=[{"k": k, "owner_chain_ms": chain[k], "update_ms": upd[k]} for k in range(max(0, nblk - 4), nblk)])
    return out


def _dist_pass(torch, dist, asm, be, dlu, rows, M, b, phases=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if phases is not None else None
    if ev: ev[0].record()
    be.info.zero_()
    be.assemble(rows, "polyharmonic", 1.0, M)
    if ev: ev[1].record()
    dlu.factor()
    if ev: ev[2].record()
    x = dlu.solve(b)
    if ev:
        ev[3].record()
        phases.append(ev)
    return x


def _dist_correctness(torch, dist, asm, be, rows, cloud, M, b, x, exact):
    r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
    own = torch.arange(cloud.N, dtype=torch.int32, device="cuda")
    jphi, jpol = asm.eval_jets("polyharmonic", 1.0, rows.centres, x.view(1, -1), rows.centres, own)
    vals = (jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy()
    max_err = float(np.max(np.abs(vals - exact)))
    be.assemble(rows, "polyharmonic", 1.0, M)              # ||K||_inf from the owned column blocks (row sums add across ranks)
    rs = be.local[:, :be.cols].abs().sum(dim=1)
    dist.all_reduce(rs)
    berr = float(r.abs().max().item() / (rs.max().item() * x.abs().max().item() + b.abs().max().item()))
    lo = be.info.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    hi = be.info.clone(); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    be.check_sweeps()
    return berr, max_err, (int(lo.item()) if int(lo.item()) < 0 else int(hi.item()))


def run_gpu_arm_distributed(args, world, rank, local):
    import torch
    import torch.distributed as dist
    import updes_b200 as u
    from updes_b200 import _lib, assembly as asm
    from updes_b200.distributed import ColumnBlockCyclic, CudaBackend, DistributedLU
    from updes_b200.operators import default_block_width

    nx = args.nx if args.nx else HEADLINE_NX
    cloud, M, n, table, rows, q, exact = headline_problem(u, asm, nx)
    xy = cloud.sorted_nodes
    nb = args.nb or default_block_width(n, world)
    layout = ColumnBlockCyclic(n, nb, world)
    be = CudaBackend(layout, rank, gemm_sms_reserved=args.reserve_sms)
    dlu = DistributedLU(layout, rank, be)
    dlu.solve_variant = args.dist_solve
    b = be.vector(q)
    state = {}

    phases = []

    def step():
        state["x"] = _dist_pass(torch, dist, asm, be, dlu, rows, M, b, phases)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    phases.clear()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.profile_enable(True)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    prof = {k: _lib.profile_read(k) for k in ("gemm", "panel", "swap", "trsm", "assemble")}
    _lib.profile_enable(False)
    value = lu_flops(n) / (ms_step * 1e-3) * 1e-12
    # phase times on this rank's stream (CUDA events), mean over the timed steps, max over ranks
    ph = torch.tensor([sum(e[i].elapsed_time(e[i + 1]) for e in phases) / max(len(phases), 1) for i in range(3)],
                      device="cuda", dtype=torch.float64)
    dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    phase_ms = {"assemble": float(ph[0].item()), "lu": float(ph[1].item()), "solve": float(ph[2].item())}

    berr, max_err, status = _dist_correctness(torch, dist, asm, be, rows, cloud, M, b, state["x"], exact)

    # one more, instrumented, factorisation: where the time goes per panel (not part of the timed region)
    timeline = None
    if not args.no_timeline:
        be.assemble(rows, "polyharmonic", 1.0, M)
        dlu.timeline = []
        barrier()
        dlu.factor()
        barrier()
        timeline = _timeline_summary(dist, dlu, world)
        dlu.timeline = None

    # e2e: the public API with host inputs, sharded the same way (every rank calls pde_solver_jit)
    state.clear()
    be.close()
    del be, dlu
    torch.cuda.empty_cache()
    u.enable_distributed()
    op, rhs, bcs = api_problem(u, cloud)
    e2e_s, sol = 0.0, None
    for i in range(args.e2e_steps + 1):
        sol = None
        u.clear_cache(); barrier()
        t0 = time.perf_counter()
        sol = u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i > 0 or args.e2e_steps == 0:
            e2e_s = max(e2e_s, dt)
    u.clear_cache()
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_err = float(np.max(np.abs(sol.vals - exact)))
    h2d = int(world * (xy.nbytes + sum(getattr(table, k).nbytes for k in ("p1", "p2", "cphi1", "cphi2", "cpol1", "cpol2", "skip")) + 16 * n))
    d2h = int(world * (8 * n + 8 * cloud.N + 4))
    del rows
    torch.cuda.empty_cache()

    # ---- config 5: BASELINE.json's 500x500 / 250k-node problem, one timed pass (N = 8 by default) -------------------
    config5 = None
    if args.config5 == "on" or (args.config5 == "auto" and world >= 8):
        config5 = run_config5(args, world, rank, torch, dist, u, asm, _lib)

    if rank == 0:
        peak, peak_src = fp64_peak()
        g_ms, g_flops, g_cnt = prof["gemm"]
        gemm_tf = g_flops / (g_ms * 1e-3) * 1e-12 if g_ms > 0 else 0.0
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(nx), "n": n, "matrix_bytes": 8 * n * n, "lu_flops": lu_flops(n),
                           "parallelism": "1x%d column-block-cyclic, nb=%d, look-ahead 1, NCCL panel broadcast" % (world, nb),
                           "matrix_bytes_per_gpu": 8 * n * n // world,
                           "l2": "local matrix (%.1f GB) is far larger than L2; no flush needed" % (8 * n * n / world / 1e9)},
                "seconds_per_step": ms_step * 1e-3,
                "e2e": {"value": lu_flops(n) / e2e_s * 1e-12, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "seconds_per_step": e2e_s, "api": "updes_b200.pde_solver_jit on every rank after enable_distributed() (numpy in, numpy out)",
                        "max_err_vs_analytic": e2e_err},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"kernel": "dgemm_sub_kernel (DMMA m8n8k4 + TMA) on rank 0", "bound": "tensor", "achieved": gemm_tf, "peak": peak,
                             "unit": "TFLOP/s", "frac": gemm_tf / peak, "traffic": None,
                             "share_of_step": g_ms / (ms_step * args.steps), "peak_source": peak_src},
                "breakdown": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] // max(args.steps, 1)} for k, v in prof.items()},
                "phase_ms_per_step": phase_ms,
                "per_gpu_tflops": value / world, "cpu_baseline": None, "timeline": timeline, "config5": config5,
                "correctness": {"backward_error": berr, "max_err_vs_analytic": max_err, "zero_pivot": status}}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def run_gpu_arm_grid2d(args, world, rank, local, P, Q):
    """Comparison hook (`--grid PxQ`, P > 1): the same strong-scaling step on the 2-D block-cyclic layout of
    updes_b200/grid2d.py.  NOT the default and not part of the driver's series: the measured layout is 1 x Q."""
    import torch
    import torch.distributed as dist
    import updes_b200 as u
    from updes_b200 import _lib, assembly as asm
    from updes_b200.grid2d import BlockCyclic2D, CudaKernels2D, DistributedLU2D
    from updes_b200.operators import default_block_width
    assert P * Q == world, "--grid PxQ must match the number of ranks"
    nx = args.nx if args.nx else HEADLINE_NX
    cloud, M, n, table, rows, q, exact = headline_problem(u, asm, nx)
    nb = args.nb or default_block_width(n, max(P, Q))
    layout = BlockCyclic2D(n, nb, P, Q)
    dlu = DistributedLU2D(layout, rank, CudaKernels2D())
    b = torch.as_tensor(q).cuda()
    state = {}

    def step():
        dlu.assemble(rows, "polyharmonic", 1.0, M)
        dlu.factor()
        state["x"] = dlu.solve(b)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.profile_enable(True)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    prof = {k: _lib.profile_read(k) for k in ("gemm", "panel", "swap", "trsm", "assemble", "solve")}
    _lib.profile_enable(False)
    x = state["x"]
    r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
    own = torch.arange(cloud.N, dtype=torch.int32, device="cuda")
    jphi, jpol = asm.eval_jets("polyharmonic", 1.0, rows.centres, x.view(1, -1), rows.centres, own)
    max_err = float(np.max(np.abs((jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy() - exact)))
    dlu.assemble(rows, "polyharmonic", 1.0, M)
    rs = dlu.local[:, :dlu.cvalid].abs().sum(dim=1)                     # ||K||_inf: row sums add along the process row
    dist.all_reduce(rs, group=dlu.row_groups[dlu.p])
    kn = rs.max().reshape(1)
    dist.all_reduce(kn, op=dist.ReduceOp.MAX)
    berr = float(r.abs().max().item() / (kn.item() * x.abs().max().item() + b.abs().max().item()))
    status = dlu.zero_pivot()
    if rank == 0:
        peak, peak_src = fp64_peak()
        g_ms, g_flops, g_cnt = prof["gemm"]
        gemm_tf = g_flops / (g_ms * 1e-3) * 1e-12 if g_ms > 0 else 0.0
        value = lu_flops(n) / (ms_step * 1e-3) * 1e-12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(nx), "n": n, "matrix_bytes": 8 * n * n, "lu_flops": lu_flops(n),
                           "parallelism": "%dx%d 2-D block-cyclic, nb=%d, centralised panel factorisation, no look-ahead (comparison hook)" % (P, Q, nb),
                           "l2": "local matrix (%.1f GB) is far larger than L2; no flush needed" % (8 * n * n / world / 1e9)},
                "seconds_per_step": ms_step * 1e-3, "e2e": None, "gpu_launches": launches, "clocks": clocks,
                "roofline": {"kernel": "dgemm_sub_kernel (DMMA m8n8k4 + TMA) on rank 0", "bound": "tensor", "achieved": gemm_tf, "peak": peak,
                             "unit": "TFLOP/s", "frac": gemm_tf / peak, "traffic": None,
                             "share_of_step": g_ms / (ms_step * args.steps), "peak_source": peak_src},
                "breakdown": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] // max(args.steps, 1)} for k, v in prof.items()},
                "per_gpu_tflops": value / world, "cpu_baseline": None,
                "correctness": {"backward_error": berr, "max_err_vs_analytic": max_err, "zero_pivot": status}}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def run_config5(args, world, rank, torch, dist, u, asm, _lib):
    from updes_b200.distributed import ColumnBlockCyclic, CudaBackend, DistributedLU
    from updes_b200.operators import default_block_width
    nx = args.config5_nx
    cloud, M, n, table, rows, q, exact = headline_problem(u, asm, nx)
    nb = args.nb or default_block_width(n, world)
    layout = ColumnBlockCyclic(n, nb, world)
    be = CudaBackend(layout, rank, gemm_sms_reserved=args.reserve_sms)
    dlu = DistributedLU(layout, rank, be)
    b = be.vector(q)
    dist.barrier(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    be.info.zero_()
    be.assemble(rows, "polyharmonic", 1.0, M)
    ev[1].record()
    dlu.factor()
    ev[2].record()
    x = dlu.solve(b)
    ev[3].record()
    dist.barrier(); torch.cuda.synchronize()
    ts = torch.tensor([ev[0].elapsed_time(ev[3]), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])],
                      device="cuda", dtype=torch.float64)
    dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    berr, max_err, status = _dist_correctness(torch, dist, asm, be, rows, cloud, M, b, x, exact)
    total_s = float(ts[0].item()) * 1e-3
    out = {"workload": workload_name(nx), "n": n, "matrix_bytes": 8 * n * n, "matrix_bytes_per_gpu": 8 * n * n // world,
           "passes": 1, "seconds": total_s, "assemble_seconds": float(ts[1].item()) * 1e-3, "lu_seconds": float(ts[2].item()) * 1e-3,
           "solve_seconds": float(ts[3].item()) * 1e-3, "tflops": lu_flops(n) / total_s * 1e-12,
           "per_gpu_tflops": lu_flops(n) / total_s * 1e-12 / world, "nb": nb,
           "backward_error": berr, "max_err_vs_analytic": max_err, "zero_pivot": status,
           "note": "one timed pass (max over ranks, CUDA events), NCCL and kernels already warm from the 90k steps"}
    be.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=0, help="SquareCloud side (default 300: the 90k-node headline config, at every GPU count)")
    ap.add_argument("--nb", type=int, default=0, help="column-block width of the multi-GPU layout (default: auto)")
    ap.add_argument("--grid", default="", help="comparison hook (multi-GPU): PxQ with P > 1 runs the 2-D block-cyclic layout of updes_b200/grid2d.py instead of 1xQ")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs left free by the update GEMM for NCCL (multi-GPU)")
    ap.add_argument("--cpu-nx", type=int, default=70, help="side of the bounded CPU sample of the reference arm's steps")
    ap.add_argument("--cpu-sizes", default="30x20,50x50,100x100", help="clouds of the CPU size sweep (BASELINE.md 4.3: N = 600, 2 500, 10 000)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--config5", default="auto", choices=["auto", "on", "off"], help="250k-node pass inside the line (auto: at 8 GPUs)")
    ap.add_argument("--config5-nx", type=int, default=500)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cpu-sweep", action="store_true")
    ap.add_argument("--no-library", action="store_true")
    ap.add_argument("--no-small", action="store_true")
    ap.add_argument("--no-timeline", action="store_true")
    ap.add_argument("--trsm-base", type=int, default=0, help="tuning hook (1 GPU): rows of the unit-lower solve's base kernel (32 / 128)")
    ap.add_argument("--gemm-variant", type=int, default=-1, help="tuning hook (1 GPU): updes_lu_set_gemm_variant bits")
    ap.add_argument("--dist-solve", default="left", choices=["left", "right"], help="multi-GPU substitution: left-looking (default) or the first-generation column sweep")
    args = ap.parse_args()
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_gpu_arm(args)
    except Exception:
        # make the failing rank's traceback visible in the driver's tail (round 1: N=4 died without one)
        msg = "[bench.py rank %s] FAILED\n%s" % (os.environ.get("RANK", "0"), traceback.format_exc())
        print(msg, file=sys.stderr, flush=True)
        print(msg, flush=True)
        raise


if __name__ == "__main__":
    main()
