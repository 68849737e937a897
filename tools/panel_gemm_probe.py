"""Does the row pitch of the panel buffer (A operand of the distributed update GEMM) matter?
C[m x n] -= A[m x k] B[k x n] with A in an auxiliary buffer of leading dimension ldp."""
import ctypes, json, sys
import torch
sys.path.insert(0, ".")
from updes_b200 import _lib

lib = _lib.load()
m, n, k = 120000, 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rows = m + k
ld = 8192 + 16
local = torch.randn((rows, ld), dtype=torch.float64, device="cuda")
out = {}
for ldp in (k, k + 16, k + 48):
    buf = torch.randn((rows, ldp), dtype=torch.float64, device="cuda")
    h = ctypes.c_void_p()
    _lib.check(lib.updes_lu_create(ctypes.byref(h), rows, ld), "create")
    _lib.check(lib.updes_lu_bind(h, 0, local.data_ptr(), rows, ld), "bind0")
    _lib.check(lib.updes_lu_bind(h, 1, buf.data_ptr(), rows, ldp), "bind1")
    st = _lib.stream_ptr()
    for rep in range(3):
        if rep == 1:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        _lib.check(lib.updes_lu_gemm(h, 1, k, 0, 0, 0, 0, 0, k, 0, m, n, k, st), "gemm")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    out["ldp%d" % ldp] = {"ms": round(ms, 2), "tflops": round(2.0 * m * n * k / ms * 1e-9, 2)}
    lib.updes_lu_destroy(h)
    del buf
print(json.dumps({"m": m, "n": n, "k": k, **out}))
