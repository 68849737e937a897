"""Wall-clock of the public API on the reference's own small configurations (BASELINE.json configs 0-2)
next to the CPU oracle's reference formulation on the same box.  Exploration, not a bench line."""
import json, os, sys, time
from functools import partial
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import updes_b200 as u
from helpers import cloud_from_golden
from oracle import oracle as O
O.build()
try:
    from threadpoolctl import threadpool_limits; threadpool_limits(limits=os.cpu_count())
except Exception: pass

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r

out = {}
# config 0: README Laplace 30x20
cloud = u.SquareCloud(Nx=30, Ny=20, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
lap = lambda x, c, r, m, f: u.nodal_laplacian(x, c, r, m)
bcs = {"South": lambda c: 0.0, "West": lambda c: 0.0, "North": lambda c: np.sin(np.pi * c[0]), "East": lambda c: 0.0}
def c0():
    u.clear_cache(); return u.pde_solver_jit(lap, lambda x, cs, r, f: 0.0, cloud, bcs, u.polyharmonic, 1)
t, sol = timed(c0)
coef = np.tile([0., 0, 0, 1, 1], (cloud.Ni, 1)); q = O.assemble_q(cloud, np.zeros(cloud.Ni), u.boundary_conditions_func_to_arr(bcs, cloud))
t0 = time.perf_counter(); vals, _, _ = O.reference_solve(cloud, "polyharmonic", 1, 1, coef, q); tc = time.perf_counter() - t0
out["config0_laplace_30x20"] = {"N": cloud.N, "gpu_e2e_ms": round(t * 1e3, 2), "cpu_reference_formulation_ms": round(tc * 1e3, 1), "rel_diff": float(np.max(np.abs(sol.vals - vals)) / np.max(np.abs(vals)))}
# config 1: periodic advection-diffusion 35x35, 10 time steps (factor once, solve per step)
DT = 1e-4
cloud2 = u.SquareCloud(Nx=35, Ny=35, facet_types={"South": "p1", "North": "p1", "West": "p2", "East": "p2"}, noise_key=11)
rbf = partial(u.polyharmonic, a=1)
def adv(x, c, r, m, f):
    return u.nodal_value(x, c, r, m) / DT + np.dot(np.array([100.0, 0.0]), u.nodal_gradient(x, c, r, m)) - 0.08 * u.nodal_laplacian(x, c, r, m)
rhs = lambda x, cs, r, f: u.value(x, f[:, 0], cs, r) / DT
xy = cloud2.sorted_nodes
u0 = np.exp(-((xy[:, 0] - 0.35) ** 2 + (xy[:, 1] - 0.5) ** 2) / 0.02)
bz = {k: (lambda c: 0.0) for k in cloud2.facet_types}
u.clear_cache()
t0 = time.perf_counter(); s = u.pde_solver_jit(adv, rhs, cloud2, bz, rbf, 0, rhs_args=[u0]); torch.cuda.synchronize(); first = time.perf_counter() - t0
uu = s.vals; t0 = time.perf_counter()
for _ in range(10):
    uu = u.pde_solver_jit(adv, rhs, cloud2, bz, rbf, 0, rhs_args=[uu]).vals
torch.cuda.synchronize(); per = (time.perf_counter() - t0) / 10
coef2 = np.tile([1 / DT, 100.0, 0.0, -0.08, -0.08], (cloud2.Ni, 1))
t0 = time.perf_counter()
A = O.assemble_A(cloud2, "polyharmonic", 1, 1); cprev = np.linalg.solve(A, np.concatenate([u0, np.zeros(1)]))
qi = O.eval_field(xy[:cloud2.Ni], xy, cprev, "polyharmonic", 1, "value") / DT
q2 = O.assemble_q(cloud2, qi, {k: np.zeros(len(v)) for k, v in cloud2.facet_nodes.items()})
O.reference_solve(cloud2, "polyharmonic", 1, 0, coef2, q2); tc = time.perf_counter() - t0
out["config1_advdiff_periodic_35x35"] = {"N": cloud2.N, "gpu_first_step_ms": round(first * 1e3, 2), "gpu_per_step_ms_factor_cached": round(per * 1e3, 2), "cpu_reference_formulation_per_step_ms": round(tc * 1e3, 1)}
# config 2: mesh.msh cloud, pressure Poisson solve
cloud3, _ = cloud_from_golden("mesh_msh_cloud_phi.npz")
src = np.random.default_rng(2).normal(size=cloud3.Ni)
bc3 = {k: np.zeros(len(v)) for k, v in cloud3.facet_nodes.items()}
def c3():
    u.clear_cache(); return u.pde_solver_jit(lap, lambda x, cs, r, f: src, cloud3, bc3, rbf, 1)
t, sol3 = timed(c3, reps=3)
coef3 = np.tile([0., 0, 0, 1, 1], (cloud3.Ni, 1)); q3 = O.assemble_q(cloud3, src, bc3)
t0 = time.perf_counter(); v3, _, _ = O.reference_solve(cloud3, "polyharmonic", 1, 1, coef3, q3); tc = time.perf_counter() - t0
out["config2_gmsh_poisson_1385"] = {"N": cloud3.N, "gpu_e2e_ms": round(t * 1e3, 2), "cpu_reference_formulation_ms": round(tc * 1e3, 1), "rel_diff": float(np.max(np.abs(sol3.vals - v3)) / np.max(np.abs(v3)))}
out["host_cores"] = os.cpu_count()
print(json.dumps(out, indent=1))
