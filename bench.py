#!/usr/bin/env python
"""Benchmark of the Updes global RBF-collocation hot path on B200 (contract: task brief, section 4).

A *step* is one pass of the hot path over the synthetic Laplace SquareCloud of BASELINE.json
(configs[3]: 300x300 = 90 000 nodes, polyharmonic a=1, max_degree=1, n = 90 003): dense assembly of
the collocation system in HBM, in-place LU with partial pivoting, one right-hand-side solve.

  metric   assemble_lu_solve_fp64_tflops = (2/3 n^3 flop) / (assemble + LU + solve seconds), summed
           over ranks; ms_per_step carries BASELINE.json's "seconds at N=90k" directly.
  value    inputs (node coordinates, row descriptors, right-hand side) already resident in HBM.
  e2e      the same pass through the public API pde_solver_jit with HOST (numpy) inputs and outputs:
           operator lowering, row-descriptor upload, assembly, LU, solve, solution download.
  roofline the dominant kernel (the DMMA trailing-update GEMM), algorithmic flops / its summed launch
           durations measured with CUDA events on the launching stream inside the timed region.
  cpu_baseline / --impl reference: the reference *formulation* (inv(A), B = D inv(A), QR, inv(A)[u;0];
           updes/assembly.py:366-410, operators.py:602-616) restated on the CPU oracle with LAPACK on
           all host cores, on a bounded sample (70x70 = 4 900 nodes, the size the reference's own demo
           quotes as "19 minutes"), expressed in the same unit as (2/3 n_s^3) / seconds.  JAX is not
           installed in this image, so this is the oracle port, not the reference package itself.

Multi-GPU (--gpus N, one process per GPU under torchrun): ONE problem sharded column-block-cyclically
over the N GPUs (updes_b200/distributed.py): assembly of the owned column blocks (no communication),
LU with one NCCL broadcast per panel, distributed solve.  The problem grows with N so that every GPU
keeps the 1-GPU HBM footprint (64.8 GB of matrix): side = 300 N^(1/4) -> 300, 357, 424, 500 nodes per
side, i.e. BASELINE.json's 500x500 / 250k-node configuration at N = 8 ("scaling": "weak" in memory per
GPU; flops per GPU grow as sqrt(N), so the metric is TFLOP/s, not seconds).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FACETS = {"South": "n", "West": "d", "North": "d", "East": "d"}
METRIC, UNIT = "assemble_lu_solve_fp64_tflops", "TFLOP/s"


def lu_flops(n):
    return 2.0 / 3.0 * float(n) ** 3


def default_side(gpus):
    """SquareCloud side per GPU count: constant HBM footprint per GPU (side = 300 N^(1/4)), with the
    BASELINE.json configurations at the ends: 300x300 on 1 GPU, 500x500 on 8."""
    return {1: 300, 2: 357, 4: 424, 8: 500}.get(gpus, int(round(300 * gpus ** 0.25)))


def workload_name(nx):
    return "Synthetic Laplace SquareCloud %dx%d (%d nodes), polyharmonic a=1, max_degree=1, FP64" % (nx, nx, nx * nx)


# ------------------------------------------------------------------------------------------------
# CPU leg: the reference formulation on the oracle (bounded sample)
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(nx):
    from oracle import oracle as O
    cloud = O.RefSquareCloud(nx, nx, FACETS) if nx <= 40 else _fast_ref_cloud(nx)
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    xy = cloud.sorted_nodes
    t0 = time.perf_counter()
    bc = {f: (np.sin(np.pi * xy[ids, 0]) if f == "North" else np.zeros(len(ids))) for f, ids in cloud.facet_nodes.items()}
    q = O.assemble_q(cloud, np.zeros(cloud.Ni), bc)
    vals, coeffs, _ = O.reference_solve(cloud, "polyharmonic", 1.0, 1, coef, q)
    dt = time.perf_counter() - t0
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    return dt, cloud.N + 3, float(np.max(np.abs(vals - exact)))


def _fast_ref_cloud(nx):
    """Cloud arrays for the CPU leg at sizes where the oracle's literal dict loops are slow: the
    product's vectorised SquareCloud yields identical arrays (tests/test_host.py checks that)."""
    import updes_b200 as u
    return u.SquareCloud(Nx=nx, Ny=nx, facet_types=FACETS)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core."""
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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    use_all_host_threads()
    nx = args.cpu_nx
    for _ in range(args.warmup):
        cpu_reference_pass(min(nx, 30))
    times = []
    for _ in range(args.steps):
        dt, n_s, err = cpu_reference_pass(nx)
        times.append(dt)
    t = sum(times) / len(times)
    val = lu_flops(n_s) / t * 1e-12
    sample = "reference formulation (inv+GEMM+QR, oracle port with LAPACK) on SquareCloud %dx%d, n=%d; %.2f s per pass; " \
             "rate = (2/3 n^3)/t, the same normalisation as the GPU arm" % (nx, nx, n_s, t)
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.nx or default_side(args.gpus)), "sample": "SquareCloud %dx%d" % (nx, nx)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": blas_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
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

    nx = args.nx if args.nx else default_side(world)
    if world > 1:
        return run_gpu_arm_distributed(args, world, rank, local, nx)
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
    b = torch.as_tensor(q).cuda()
    K = torch.empty((n, asm.padded_ld(n)), dtype=torch.float64, device="cuda")
    lu = LUFactorization(K, n)
    x = torch.empty_like(b)

    def step():
        asm.assemble_system(rows, "polyharmonic", 1.0, M, out=K)
        lu.factor()
        x.copy_(b)
        lu.solve(x)

    def barrier():
        if world > 1:
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
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    prof = {k: _lib.profile_read(k) for k in ("gemm", "panel", "swap", "trsm", "assemble", "solve")}
    _lib.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * lu_flops(n) / (ms_step * 1e-3) * 1e-12

    # correctness beside the timing: matrix-free backward error and the analytic solution
    r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
    own = torch.arange(cloud.N, dtype=torch.int32, device="cuda")
    jphi, jpol = asm.eval_jets("polyharmonic", 1.0, rows.centres, x.view(1, -1), rows.centres, own)
    vals = (jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy()
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    max_err = float(np.max(np.abs(vals - exact)))
    asm.assemble_system(rows, "polyharmonic", 1.0, M, out=K)
    knorm = float(K[:, :n].abs().sum(dim=1).max().item())
    berr = float(r.abs().max().item() / (knorm * x.abs().max().item() + b.abs().max().item()))

    # ---- e2e through the public API with host buffers -----------------------------------------------
    del K, lu
    torch.cuda.empty_cache()
    op = lambda xx, center, rbf, monomial, fields: u.nodal_laplacian(xx, center, rbf, monomial)
    rhs = lambda xx, centers, rbf, fields: 0.0
    bcs = {"South": np.zeros(len(cloud.facet_nodes["South"])), "West": np.zeros(len(cloud.facet_nodes["West"])),
           "North": np.sin(np.pi * xy[north, 0]), "East": np.zeros(len(cloud.facet_nodes["East"]))}
    e2e_times = []
    for i in range(args.e2e_steps + 1):
        u.clear_cache()
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        sol = u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i > 0 or args.e2e_steps == 0:
            e2e_times.append(dt)
    u.clear_cache()
    e2e_s = max(e2e_times)
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * lu_flops(n) / e2e_s * 1e-12
    h2d = int(xy.nbytes + sum(getattr(table, k).nbytes for k in ("p1", "p2", "cphi1", "cphi2", "cpol1", "cpol2", "skip")) + 8 * n * 2)
    d2h = int(8 * n + 8 * cloud.N + 4)
    e2e_err = float(np.max(np.abs(sol.vals - exact)))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    ceil_path = os.path.join(ROOT, "profiles", "r01_ceilings.json")
    fp64_peak, peak_src = 37.0, "B200 FP64 tensor nominal 37 TF (fallback)"
    try:
        c = json.load(open(ceil_path))
        fp64_peak = float(c["dmma_tflops_8warps"])
        peak_src = "measured on this pool: raw DMMA issue rate %.2f TF (profiles/r01_ceilings.json; cuBLAS DGEMM %.2f TF); " \
                   "MEASURED_PEAKS.json has no FP64 figure" % (fp64_peak, c.get("cublas_dgemm_tflops_n16384", 0))
    except Exception:
        pass
    hbm_peak, hbm_src = 6650.0, "fallback"
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); hbm_src = "measured"
    except Exception:
        pass
    g_ms, g_flops, g_cnt = prof["gemm"]
    gemm_tf = g_flops / (g_ms * 1e-3) * 1e-12 if g_ms > 0 else 0.0
    a_ms, a_bytes, a_cnt = prof["assemble"]
    s_ms, s_bytes, s_cnt = prof["solve"]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_summary.json"))).get("dgemm_traffic_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "dgemm_sub_kernel (DMMA m8n8k4 + TMA; <64,4> ping-pong for wide updates), %d launches/step" % (g_cnt // max(args.steps, 1)),
                "bound": "tensor", "achieved": gemm_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": gemm_tf / fp64_peak,
                "traffic": traffic, "peak_source": peak_src, "share_of_step": g_ms / ms_total}
    breakdown = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] // max(args.steps, 1)} for k, v in prof.items()}
    breakdown["assemble"].update(gbs=a_bytes / (a_ms * 1e-3) * 1e-9 if a_ms else None,
                                 frac_of_hbm=(a_bytes / (a_ms * 1e-3) * 1e-9 / hbm_peak) if a_ms else None, hbm_peak=hbm_peak,
                                 hbm_peak_source=hbm_src)
    breakdown["solve"].update(gbs=s_bytes / (s_ms * 1e-3) * 1e-9 if s_ms else None)
    breakdown["lu_tflops"] = lu_flops(n) / (sum(prof[k][0] for k in ("gemm", "panel", "swap", "trsm")) / args.steps * 1e-3) * 1e-12

    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle as O
        O.build()
        use_all_host_threads()
        cpu_reference_pass(30)
        dt, n_s, cerr = cpu_reference_pass(args.cpu_nx)
        cpu = {"value": lu_flops(n_s) / dt * 1e-12, "unit": UNIT, "cores": blas_threads(), "kind": "port",
               "sample": "reference formulation (inv+GEMM+QR; oracle port, LAPACK) on SquareCloud %dx%d, n=%d: %.2f s; "
                         "rate = (2/3 n^3)/t; max err vs analytic %.1e" % (args.cpu_nx, args.cpu_nx, n_s, dt, cerr)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(nx), "n": n, "matrix_bytes": 8 * n * n, "lu_flops": lu_flops(n),
                       "parallelism": "1 GPU",
                       "l2": "matrix (%.1f GB) is far larger than L2; no flush needed" % (8 * n * n / 1e9)},
            "seconds_per_step": ms_step * 1e-3,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "seconds_per_step": e2e_s, "api": "updes_b200.pde_solver_jit (numpy in, numpy out)",
                    "max_err_vs_analytic": e2e_err},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "breakdown": breakdown,
            "cpu_baseline": cpu,
            "correctness": {"backward_error": berr, "max_err_vs_analytic": max_err, "zero_pivot": 0}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_gpu_arm_distributed(args, world, rank, local, nx):
    """N > 1: one sharded problem (column-block-cyclic), timed as the max over ranks."""
    import torch
    import torch.distributed as dist
    import updes_b200 as u
    from updes_b200 import _lib, assembly as asm
    from updes_b200.distributed import ColumnBlockCyclic, CudaBackend, DistributedLU
    from updes_b200.operators import default_block_width

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
    nb = args.nb or default_block_width(n, world)
    layout = ColumnBlockCyclic(n, nb, world)
    be = CudaBackend(layout, rank, gemm_sms_reserved=args.reserve_sms)
    dlu = DistributedLU(layout, rank, be)
    b = be.vector(q)
    state = {}

    def step():
        be.info.zero_()
        be.assemble(rows, "polyharmonic", 1.0, M)
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
    prof = {k: _lib.profile_read(k) for k in ("gemm", "panel", "swap", "trsm", "assemble")}
    _lib.profile_enable(False)
    value = lu_flops(n) / (ms_step * 1e-3) * 1e-12

    x = state["x"]
    r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
    own = torch.arange(cloud.N, dtype=torch.int32, device="cuda")
    jphi, jpol = asm.eval_jets("polyharmonic", 1.0, rows.centres, x.view(1, -1), rows.centres, own)
    vals = (jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy()
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    max_err = float(np.max(np.abs(vals - exact)))
    # ||K||_inf from the owned column blocks (row sums add across ranks)
    be.assemble(rows, "polyharmonic", 1.0, M)
    rs = be.local[:, :be.cols].abs().sum(dim=1)
    dist.all_reduce(rs)
    berr = float(r.abs().max().item() / (rs.max().item() * x.abs().max().item() + b.abs().max().item()))
    zero_piv = be.info.clone(); dist.all_reduce(zero_piv, op=dist.ReduceOp.MAX)

    # e2e: the public API with host inputs, sharded the same way (every rank calls pde_solver_jit)
    del be, dlu
    torch.cuda.empty_cache()
    op = lambda xx, center, rbf, monomial, fields: u.nodal_laplacian(xx, center, rbf, monomial)
    rhs = lambda xx, centers, rbf, fields: 0.0
    bcs = {"South": np.zeros(len(cloud.facet_nodes["South"])), "West": np.zeros(len(cloud.facet_nodes["West"])),
           "North": np.sin(np.pi * xy[north, 0]), "East": np.zeros(len(cloud.facet_nodes["East"]))}
    e2e_s = 0.0
    for i in range(args.e2e_steps + 1):
        u.clear_cache(); torch.cuda.empty_cache(); barrier()
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
    h2d = int(world * (xy.nbytes + sum(getattr(table, k).nbytes for k in ("p1", "p2", "cphi1", "cphi2", "cpol1", "cpol2", "skip")) + 16 * n))
    d2h = int(world * (8 * n + 8 * cloud.N + 4))
    if rank == 0:
        fp64_peak = 37.0
        try:
            fp64_peak = float(json.load(open(os.path.join(ROOT, "profiles", "r01_ceilings.json")))["dmma_tflops_8warps"])
        except Exception:
            pass
        g_ms, g_flops, g_cnt = prof["gemm"]
        gemm_tf = g_flops / (g_ms * 1e-3) * 1e-12 if g_ms > 0 else 0.0
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(nx), "n": n, "matrix_bytes": 8 * n * n, "lu_flops": lu_flops(n),
                           "parallelism": "1x%d column-block-cyclic, nb=%d, look-ahead 1, NCCL panel broadcast" % (world, nb),
                           "matrix_bytes_per_gpu": 8 * n * n // world,
                           "l2": "local matrix (%.1f GB) is far larger than L2; no flush needed" % (8 * n * n / world / 1e9)},
                "seconds_per_step": ms_step * 1e-3,
                "e2e": {"value": lu_flops(n) / e2e_s * 1e-12, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "seconds_per_step": e2e_s, "api": "updes_b200.pde_solver_jit on every rank (numpy in, numpy out)",
                        "max_err_vs_analytic": float(np.max(np.abs(sol.vals - exact)))},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"kernel": "dgemm_sub_kernel (DMMA m8n8k4 + TMA) on rank 0", "bound": "tensor", "achieved": gemm_tf, "peak": fp64_peak,
                             "unit": "TFLOP/s", "frac": gemm_tf / fp64_peak, "traffic": None,
                             "share_of_step": g_ms / (ms_step * args.steps),
                             "peak_source": "raw DMMA issue rate per GPU, profiles/r01_ceilings.json"},
                "breakdown": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] // max(args.steps, 1)} for k, v in prof.items()},
                "per_gpu_tflops": value / world, "cpu_baseline": None,
                "correctness": {"backward_error": berr, "max_err_vs_analytic": max_err, "zero_pivot": int(zero_piv.item())}}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=0, help="SquareCloud side (default 300 * gpus^(1/4): 300 -> the 90k-node headline config)")
    ap.add_argument("--nb", type=int, default=0, help="column-block width of the multi-GPU layout (default: auto)")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs left free by the update GEMM for NCCL (multi-GPU)")
    ap.add_argument("--cpu-nx", type=int, default=70, help="side of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
