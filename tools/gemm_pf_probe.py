"""A/B of the GEMM producer's L2 prefetch (distance x operand mask) on LU-shaped trailing updates."""
import json
import sys

import torch

sys.path.insert(0, ".")
from updes_b200.assembly import padded_ld
from updes_b200.linalg import LUFactorization

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K = torch.randn((n, padded_ld(n)), dtype=torch.float64, device="cuda")
lu = LUFactorization(K, n)
out = {}
for k in (2048, 512, 8192):
    m = nn = n - k
    for dist, mask in ((0, 0), (2, 3), (4, 3), (8, 3), (4, 1), (4, 2), (12, 3)):
        lu.set_gemm_variant(3 | (dist << 2) | (mask << 6))
        lu.gemm_sub(k, k, k, 0, 0, k, m, nn, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lu.gemm_sub(k, k, k, 0, 0, k, m, nn, k)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out["k%d_pf%d_mask%d" % (k, dist, mask)] = round(2.0 * m * nn * k / ms * 1e-9, 2)
print(json.dumps({"n": n, "tflops": out}))
