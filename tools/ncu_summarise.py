"""Condenses the ncu exports of tools/ncu_r02.sh (gpurun_out/r02_ncu_*_raw.csv, launch lists) into the tracked
profiles/r02_* files: one JSON of headline metrics per kernel, the per-kernel launch-time shares, and the `details`
pages as text."""
import collections
import csv
import json
import os
import shutil
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
DST = sys.argv[2] if len(sys.argv) > 2 else "profiles"
KEYS = {
    "duration_ms": "gpu__time_duration.sum",
    "dram_read_GB": "dram__bytes_read.sum",
    "dram_write_GB": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "issue_active_pct": "smsp__issue_active.avg.pct",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_math_pipe_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_membar": "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "stall_lg_throttle": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
}
UNIT_SCALE = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "second": 1e3}


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")][:110]}
        for k, name in KEYS.items():
            if name in hdr:
                i = hdr.index(name)
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    continue
                if k.endswith("_GB") or k.endswith("_ms"):
                    v *= UNIT_SCALE.get(units[i], 1.0)
                rec[k] = v
        out.append(rec)
    return out


def launch_shares(path):
    """ncu --metrics gpu__time_duration.sum --csv launch list -> {kernel: (launches, total us)}"""
    agg = collections.OrderedDict()
    rd = csv.reader(l for l in open(path) if l.startswith('"'))
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        try:
            v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "ms ": 1e3}.get(r[iu], 1.0)
        except ValueError:
            continue
        name = r[ik].split("(")[0].replace("void ", "")[:70]
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + v)
    tot = sum(t for _, t in agg.values()) or 1.0
    return [{"kernel": k, "launches": c, "total_us": round(t, 1), "share": round(t / tot, 4)}
            for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]


summary = {}
for name in ("dgemm", "dgemm_k1024", "sweep", "panel", "trsm", "assemble"):
    p = os.path.join(SRC, "r02_ncu_%s_raw.csv" % name)
    if os.path.exists(p):
        summary[name] = read_raw(p)
    d = os.path.join(SRC, "r02_ncu_%s_details.txt" % name)
    if os.path.exists(d):
        shutil.copy(d, os.path.join(DST, "r02_ncu_%s.txt" % name))
g = summary.get("dgemm", [{}])[0]
if "dram_read_GB" in g:
    summary["dgemm_traffic_bytes_per_launch"] = (g["dram_read_GB"] + g["dram_write_GB"]) * 1e9
    summary["dgemm_traffic_note"] = ("ncu --set full of ONE warm launch of the current default <64,4,2> ping-pong kernel, m = n = 30720, k = 2048 "
                                     "(tools/ncu_targets.py gemm): dram read %.2f GB + write %.2f GB; algorithmic = C read+write 15.10 GB + A, B "
                                     "once 1.01 GB" % (g["dram_read_GB"], g["dram_write_GB"]))
for tag in ("default_c400", "nx91"):
    p = os.path.join(SRC, "r02_launches_%s.csv" % tag)
    if os.path.exists(p):
        summary["launches_" + tag] = launch_shares(p)
json.dump(summary, open(os.path.join(DST, "r02_ncu_summary.json"), "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, list) else v[:6]) for k, v in summary.items()}, indent=1)[:6000])
