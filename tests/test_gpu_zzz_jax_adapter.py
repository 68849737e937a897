"""The reference-side jax.ffi adapter (integration/updes_jax_ffi.cc + integration/updes_jax.py) on a B200, without JAX:
handlers compiled against the mock of XLA's FFI binding API (tests/mock_xla) and linked against libupdes_b200.so, driven
by the jax.ffi stand-in of oracle/refshim on CUDA buffers (tests/run_jax_adapter.py says what is checked).  Runs in a
subprocess: the stand-in `jax` package must not leak into this process.  Written after the round's GPU minutes were
spent: dry-run on the emulated C-ABI only (tests/test_host_on_emulated_abi.py); last in file order for that reason."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jax_ffi_adapter_against_product_and_reference_golden():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_jax_adapter.py")], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    print(p.stdout)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "jax.ffi adapter ok (libupdes_b200.so, CUDA)" in p.stdout
