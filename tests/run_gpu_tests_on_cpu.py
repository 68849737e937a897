#!/usr/bin/env python
"""Run `-m gpu` test files (and smoke()) on the CPU emulation of the C-ABI -- a dry run of the HOST logic and of the
tests themselves in a container without a GPU (tests/cpu_abi_emulation.py says what that does and does not prove).

    python tests/run_gpu_tests_on_cpu.py                      # the default file list below + smoke()
    python tests/run_gpu_tests_on_cpu.py tests/test_gpu_solver.py -k coupled
    python tests/run_gpu_tests_on_cpu.py --smoke tests/test_gpu_zy_reference_golden.py      # listed files, then smoke()

Not collected by pytest (no test_ prefix); tests/test_host_on_emulated_abi.py runs a bounded subset of it in a subprocess.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# files whose tests go through the emulated entry points only (whole-matrix LU, assembly, evaluators);
# test_gpu_lu.py / test_gpu_distributed.py / test_gpu_zx_grid2d.py drive kernel building blocks that are not emulated
DEFAULT = ["tests/test_gpu_solver.py", "tests/test_gpu_assembly.py", "tests/test_gpu_zy_reference_golden.py", "tests/test_gpu_zz_cache_lru.py", "tests/test_gpu_zz_advection_demos.py",
           "tests/test_gpu_zz_fullsize_entries.py"]          # (the last one at 40 x 40 instead of 300 x 300)


def main(argv):
    os.environ.setdefault("UPDES_FULLSIZE_NX", "40")
    import cpu_abi_emulation as emu
    emu.install()
    import pytest
    smoke = not argv or "--smoke" in argv
    argv = [a for a in argv if a != "--smoke"]
    args = argv or DEFAULT
    rc = pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + [os.path.join(ROOT, a) if a.startswith("tests/") else a for a in args])
    if smoke and rc == 0:
        import __graft_entry__ as g
        g.smoke()
    return int(rc)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
