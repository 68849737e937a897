"""Times (and, under ncu, profiles) the DMMA trailing-update GEMM alone on LU-shaped sub-blocks."""
import json
import sys

import torch

sys.path.insert(0, ".")
from updes_b200.assembly import padded_ld
from updes_b200.linalg import LUFactorization

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ld = padded_ld(n)
K = torch.randn((n, ld), dtype=torch.float64, device="cuda")
lu = LUFactorization(K, n)
out = {}
for variant in (3, 11):   # 3 = default (RED epilogue), 11 = read-modify-write epilogue with L2 prefetch
  lu.set_gemm_variant(variant)
  for k in (128, 512, 1024, 2048, 8192):
    m = nn = n - k
    lu.gemm_sub(k, k, k, 0, 0, k, m, nn, k)          # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lu.gemm_sub(k, k, k, 0, 0, k, m, nn, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out["v%d_k%d" % (variant, k)] = {"ms": round(ms, 3), "tflops": round(2.0 * m * nn * k / ms * 1e-9, 2)}
lu.set_gemm_variant(0)
# tall-skinny shapes of the panel recursion
for (m, nn, k) in ((n - 64, 32, 32), (n - 128, 64, 64), (n - 256, 128, 128), (n - 512, 256, 256)):
    lu.gemm_sub(k, k, k, 0, 0, k, m, nn, k); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lu.gemm_sub(k, k, k, 0, 0, k, m, nn, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out["panel_%dx%dx%d" % (m, nn, k)] = {"ms": round(ms, 4), "tflops": round(2.0 * m * nn * k / ms * 1e-9, 2)}
print(json.dumps({"n": n, **out}))
