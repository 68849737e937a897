#!/usr/bin/env python
"""Where the FP64 tensor time of one LU goes, from the launch structure alone (CPU, no GPU needed).

Replays the host recursion of updes_b200/csrc/lu_driver.cu (lu_recursive, trsm_unit_lower: same split rules) for a
given n, lists every trailing-update GEMM (m, n, k) it launches, and prices each launch with the GEMM rates measured on
B200 (profiles/r02_gemm_probe_epilogues.json: square-ish updates at k = 128 ... 8192; tall-skinny in-panel updates at
k = n = 32 ... 256).  Output: launch count (to compare with `breakdown.gemm.launches_per_step` of the bench line), flops and
predicted time per k class, and the predicted GEMM time next to the measured one.

    python tools/lu_time_model.py [--n 90003] [--bench profiles/r02_bench_1gpu.json]
"""
import argparse
import json
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = 32            # base panel width (panels up to 94 720 rows)
TRSM_BIG = 128    # rows the unit-lower base kernel solves in one launch


def split(nc):
    return 16 if nc <= 32 else (nc // 2 + 31) // 32 * 32


def trsm(n1, ncols, gemms, bases):
    if n1 <= 32 or (n1 <= TRSM_BIG and n1 % 32 == 0):
        bases.append((n1, ncols))
        return
    hlf = (n1 // 2 + 31) // 32 * 32
    trsm(hlf, ncols, gemms, bases)
    gemms.append((n1 - hlf, ncols, hlf, "trsm"))
    trsm(n1 - hlf, ncols, gemms, bases)


def lu(r0, nc, rows, gemms, bases, panels):
    if nc <= W:
        panels.append((rows - r0, nc))
        return
    n1 = split(nc)
    lu(r0, n1, rows, gemms, bases, panels)
    n2 = nc - n1
    trsm(n1, n2, gemms, bases)
    gemms.append((rows - (r0 + n1), n2, n1, "schur"))
    lu(r0 + n1, n2, rows, gemms, bases, panels)


def rate_tf(m, n, k, probe):
    """TFLOP/s of one launch: wide updates from the k-sweep at m = n = 32 768 (log-interpolated in k), narrow in-panel
    updates (n <= 256) from the tall-skinny probes."""
    wide = sorted((int(key.split("_k")[1]), v["tflops"]) for key, v in probe.items() if key.startswith("v11_k"))
    skinny = sorted((int(key.split("x")[-1]), v["tflops"]) for key, v in probe.items() if key.startswith("panel_"))

    def interp(tab, x):
        if x <= tab[0][0]:
            return tab[0][1] * (x / tab[0][0]) ** 0.5 if x < tab[0][0] else tab[0][1]
        for (x0, y0), (x1, y1) in zip(tab[:-1], tab[1:]):
            if x <= x1:
                t = (math.log(x) - math.log(x0)) / (math.log(x1) - math.log(x0))
                return y0 + t * (y1 - y0)
        return tab[-1][1]
    if n <= 256:
        return min(interp(skinny, min(n, k)), interp(wide, k))
    r = interp(wide, k)
    tiles = math.ceil(m / 128) * math.ceil(n / 64)
    if tiles < 2 * 148:                          # fewer tiles than CTA slots: the tail of the matrix
        r *= tiles / (2 * 148)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=90003)
    ap.add_argument("--bench", default=os.path.join(ROOT, "profiles", "r02_bench_1gpu.json"))
    args = ap.parse_args()
    probe = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_probe_epilogues.json")))
    gemms, bases, panels = [], [], []
    lu(0, args.n, args.n, gemms, bases, panels)
    classes = {}
    for m, n, k, kind in gemms:
        fl = 2.0 * m * n * k
        key = ("k>=2048" if k >= 2048 else "k=1024..2047" if k >= 1024 else "k=512..1023" if k >= 512 else
               "k=128..511" if k >= 128 else "k<128")
        c = classes.setdefault(key, [0, 0.0, 0.0])
        c[0] += 1
        c[1] += fl
        c[2] += fl / (rate_tf(m, n, k, probe) * 1e12)
    tot_fl = sum(c[1] for c in classes.values())
    tot_s = sum(c[2] for c in classes.values())
    print("n = %d: %d GEMM launches (%d Schur updates, %d inside triangular solves), %d base panels, %d unit-lower base solves"
          % (args.n, len(gemms), sum(1 for g in gemms if g[3] == "schur"), sum(1 for g in gemms if g[3] == "trsm"), len(panels), len(bases)))
    print("GEMM flops %.4e of 2/3 n^3 = %.4e (%.2f %%)" % (tot_fl, 2 / 3 * args.n ** 3, 100 * tot_fl / (2 / 3 * args.n ** 3)))
    print("%-14s %9s %10s %9s %11s" % ("inner dim", "launches", "flops %", "model s", "model TF"))
    for key in ("k>=2048", "k=1024..2047", "k=512..1023", "k=128..511", "k<128"):
        if key in classes:
            c = classes[key]
            print("%-14s %9d %9.2f%% %9.3f %11.2f" % (key, c[0], 100 * c[1] / tot_fl, c[2], c[1] / c[2] * 1e-12))
    print("model: GEMM %.3f s = %.2f TF" % (tot_s, tot_fl / tot_s * 1e-12))
    try:
        line = [json.loads(l) for l in open(args.bench) if l.strip().startswith("{")][-1]
        g = line["breakdown"]["gemm"]
        if line["config"]["n"] == args.n:
            print("measured (%s): GEMM %.3f s = %.2f TF over %d launches per step"
                  % (os.path.basename(args.bench), g["ms_per_step"] * 1e-3, line["roofline"]["achieved"], g["launches_per_step"]))
    except Exception as e:
        print("(no bench line to compare with: %s)" % e)


if __name__ == "__main__":
    main()
