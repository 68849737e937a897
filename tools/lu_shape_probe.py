"""Where the time of a mid-size LU goes, per GEMM shape class (GPU box).  The recursion of lu_driver.cu /
lu_aux.cu is replayed on the host to label every GEMM launch (Schur update or TRSM update, inner dimension k);
the per-launch CUDA-event times come from the library's profiler in launch order.  Exploration tool."""
import json
import sys
from collections import defaultdict

import torch

sys.path.insert(0, ".")
from updes_b200 import _lib
from updes_b200.assembly import padded_ld
from updes_b200.linalg import LUFactorization


def model(n, base=128):
    g = []

    def up32(x):
        return (x // 2 + 31) // 32 * 32

    def trsm_rec(n1, ncols):
        if n1 <= base:
            return
        h = up32(n1)
        trsm_rec(h, ncols)
        g.append((n1 - h, ncols, h, "trsm"))
        trsm_rec(n1 - h, ncols)

    def lu(r0, nc):
        if nc <= 32:
            return
        n1 = up32(nc)
        n2 = nc - n1
        lu(r0, n1)
        trsm_rec(n1, n2)
        g.append((n - (r0 + n1), n2, n1, "schur"))
        lu(r0 + n1, n2)
    lu(0, n)
    return g


def probe(n, base=128):
    gen = torch.Generator(device="cuda").manual_seed(n)
    A = torch.randn((n, padded_ld(n)), dtype=torch.float64, device="cuda", generator=gen)
    K = A.clone()
    lu = LUFactorization(K, n)
    lu.set_trsm_base(base)
    lu.factor(); torch.cuda.synchronize()
    K.copy_(A)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lu.factor(); e1.record(); torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1)
    K.copy_(A)
    _lib.profile_enable(True)
    lu.factor(); torch.cuda.synchronize()
    ms, work = _lib.profile_records("gemm")
    other = {k: _lib.profile_read(k) for k in ("panel", "swap", "trsm")}
    _lib.profile_enable(False)
    shapes = model(n, base)
    assert len(shapes) == len(ms), (len(shapes), len(ms))
    cls = defaultdict(lambda: [0, 0.0, 0.0])
    for (m, nn, k, w), t, fl in zip(shapes, ms, work):
        assert abs(2.0 * m * nn * k - fl) < 1, (m, nn, k, fl)
        c = cls[(w, k)]
        c[0] += 1; c[1] += t; c[2] += fl
    out = {"n": n, "trsm_base": base, "lu_ms_unprofiled": round(plain_ms, 2), "gemm_ms_profiled": round(float(ms.sum()), 2),
           "other": {k: [round(v[0], 2), v[2]] for k, v in other.items()}, "classes": []}
    for key in sorted(cls):
        c = cls[key]
        out["classes"].append({"kind": key[0], "k": key[1], "launches": c[0], "ms": round(c[1], 3),
                               "us_per_launch": round(1e3 * c[1] / c[0], 1), "tflops": round(c[2] / c[1] * 1e-9, 2)})
    return out


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [8192]
    print(json.dumps([probe(n, b) for n in sizes for b in (128, 64, 32)]))
