"""The host layer and the `-m gpu` tests themselves, dry-run on a CPU emulation of the C-ABI (no GPU in the build
container).  tests/cpu_abi_emulation.py replaces every libupdes_b200.so entry point by numpy / LAPACK with the header's
semantics; the runners execute the GPU test files unchanged on it, in subprocesses (the emulation monkey-patches torch
and must not leak into this process).  This proves the orchestration around the kernels and the tests' own logic --
golden keys, tolerances, launch-count assertions, world-size > 1 drivers under gloo -- not the kernels: those are
checked on a B200 by `pytest -m gpu`."""
import atexit
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the emulation refuses to install where a GPU exists (there the real `-m gpu` suite is the check)
needs_no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="a CUDA device is present: run pytest -m gpu instead")


JOBS = {                    # the four dry runs are independent processes: started together, collected one per test
    "single_a": ("run_gpu_tests_on_cpu.py", "tests/test_gpu_solver.py", "tests/test_gpu_assembly.py"),
    "single_b": ("run_gpu_tests_on_cpu.py", "--smoke", "tests/test_gpu_zy_reference_golden.py", "tests/test_gpu_zz_cache_lru.py",
                 "tests/test_gpu_zz_advection_demos.py", "tests/test_gpu_zz_fullsize_entries.py"),
    "multi": ("run_multi_gpu_tests_on_cpu.py", "--quick"),
    "jax": ("run_jax_adapter.py", "--emulated"),
    "fuzz": ("run_solver_fuzz_on_cpu.py", "7", "12"),
    "demos": ("run_reference_demo_on_product.py",),
}
REFERENCE = "/root/reference"
_procs = {}


def _reap():
    for p in _procs.values():
        if p.poll() is None:
            p.kill()


atexit.register(_reap)


def _run(job, timeout=900):
    if not _procs:
        env = dict(os.environ, OMP_NUM_THREADS="2")
        for name, cmd in JOBS.items():
            if name == "demos" and not os.path.isdir(REFERENCE):
                continue
            _procs[name] = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", cmd[0]), *cmd[1:]], cwd=ROOT, env=env,
                                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    p = _procs[job]
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        p.kill()
        raise
    assert p.returncode == 0, out[-3000:] + err[-3000:]
    return out


def test_emulation_closed_forms_agree_with_the_oracle(oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_abi_emulation as emu                # importing does not install anything
    assert emu.self_check() <= 1e-13
    # and the emulated assembly reproduces the oracle's K on a cloud with every row type but periodic
    import updes_b200 as u
    from updes_b200 import assembly as asm
    cloud = u.SquareCloud(Nx=9, Ny=8, facet_types={"South": "n", "West": "d", "North": "r", "East": "d"})
    rng = np.random.default_rng(1)
    coef, betas = rng.normal(size=(cloud.Ni, 5)), rng.normal(size=cloud.Nr)
    t = asm.build_operator_rows(cloud, coef, betas=betas)
    lib = emu.EmulatedLib()
    N, M = cloud.N, 6
    ld = asm.padded_ld(N + M)
    out = np.full((N + M, ld), np.nan)
    ctr = np.ascontiguousarray(cloud.sorted_nodes)
    keep = [np.ascontiguousarray(getattr(t, k)) for k in ("p1", "p2", "cphi1", "cphi2", "cpol1", "cpol2", "skip")]
    from updes_b200._lib import UpdesRows
    st = UpdesRows(ctr.ctypes.data, *[a.ctypes.data for a in keep])
    assert lib.updes_assemble_rows(3, 1.5, N, M, ctr.ctypes.data, st, 0, N + M, 7, out.ctypes.data, ld, None) == 0
    Kref = oracle.assemble_K(cloud, "multiquadric", 1.5, M, coef, betas)
    assert np.max(np.abs(out[:, :N + M] - Kref)) <= 1e-12 * np.max(np.abs(Kref)) and not out[:, N + M:].any()


@needs_no_gpu
def test_single_gpu_test_files_and_smoke_on_the_emulated_abi():
    a, b = _run("single_a"), _run("single_b")
    assert " passed" in a and "failed" not in a
    assert "smoke ok" in b and " passed" in b and "failed" not in b


@needs_no_gpu
def test_multi_gpu_test_files_on_the_emulated_abi_under_gloo():
    out = _run("multi")
    assert "hold on the emulated ABI + gloo" in out


@needs_no_gpu
def test_jax_ffi_adapter_compiles_against_the_mock_xla_api_and_runs_on_the_emulated_abi():
    """integration/updes_jax_ffi.cc compiled against tests/mock_xla (the Bind() chains must match the handlers'
    signatures) with its updes_* calls forwarded to the emulation; integration/updes_jax.py over the jax.ffi stand-in:
    its pde_solver equals the product's and the reference's own result."""
    out = _run("jax")
    assert "jax.ffi adapter ok (emulated C-ABI, CPU)" in out


@needs_no_gpu
def test_randomised_solves_of_the_host_layer_against_oracle_assembled_systems():
    out = _run("fuzz")
    assert "0 outside the bounds" in out


@needs_no_gpu
def test_reference_demo_scripts_run_unmodified_on_the_product():
    """The reference's own three tests (updes/tests/test_*.py), then demos/Laplace/00_laplace_with_rbf.py,
    demos/Darcy/00_darcy_flow.py and the projection loop of demos/NavierStokes/30_... of the reference executed as they are, with
    `updes` aliased to `updes_b200` (tests/run_reference_demo_on_product.py): same solution as the reference computed for
    the same script (1.4e-10 / 1.4e-8), same error figure printed (4.559171e-07 vs 4.559172e-07)."""
    if not os.path.isdir(REFERENCE):
        pytest.skip("needs /root/reference (build container)")
    out = _run("demos")
    assert "reference demos ran unmodified on the product" in out
    assert "the reference's own 3 tests pass on the product" in out          # updes/tests/test_{interpolation,integrals,operators}.py
