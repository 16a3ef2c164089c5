"""Turns the raw pages of the `ncu --set full` captures of tools/ncu_traffic.sh (gpurun_out/r02_full_c*_raw.csv) into profiles/r02_traffic.json
(DRAM bytes per replicate per driver kernel, read by bench.py's roofline block) and profiles/r02_kernel_metrics.md
(duration, issue rate, pipe utilisation, top stall reasons).  Run here after the GPU call:  python tools/ncu_traffic_summary.py"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unbiasedmcmc_b200.workloads import KERNEL_OF  # noqa: E402

NREP = {"c1": 1 << 18, "c2": 148 * 16, "c3": 148, "c4": 2368, "c5": 148 * 16}     # defaults of tools/prof_case.py
CASE = {"c1": "k=500 m=2000 lag=500", "c2": "k=400 m=2000", "c3": "k=10 m=60", "c4": "k=20 m=40 (capped at 80 iterations)",
        "c5": "k=1500 m=3000"}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum"]

traffic, lines = {}, ["# Round 2 — `ncu --set full --clock-control none` of each driver kernel on the small cases of tools/prof_case.py", "",
                      "One launch each; raw reports stay in `gpurun_out/` (scratch).  Bytes are dram__bytes_read.sum + dram__bytes_write.sum.", ""]
for c, kern in KERNEL_OF.items():
    raw = os.path.join(ROOT, "gpurun_out", f"r02_full_{c}_raw.csv")
    if not os.path.exists(raw):
        continue
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def num(key):
        v, u = m.get(key, ("nan", ""))
        x = float(v.replace(",", "")) if v not in ("", "nan") else float("nan")
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "nsecond": 1e-9}.get(u, 1.0)
        return x * scale
    rd, wr, dur = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
    traffic[kern] = {"dram_bytes_per_replicate": (rd + wr) / NREP[c],
                     "source": f"ncu --set full, one launch of tools/prof_case.py {c} ({NREP[c]} replicates, {CASE[c]}, {dur * 1e3:.1f} ms): "
                               f"{rd / 1e6:.2f} MB read + {wr / 1e6:.2f} MB written (profiles/r02_kernel_metrics.md); state stays on chip, "
                               "the traffic is the per-replicate outputs plus the model blob"}
    stalls = sorted(((float(v[0]), h) for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                     and v[0] not in ("", "nan")), reverse=True)[:5]
    lines += [f"## {c}: `{kern}` — {NREP[c]} replicates, {CASE[c]}", ""]
    lines += [f"* {k}: {m[k][0]} {m[k][1]}" for k in WANT if k in m]
    lines += ["* top stall reasons (warps per issue-active cycle): " + ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {x:.2f}" for x, h in stalls), ""]
json.dump(traffic, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", "r02_kernel_metrics.md"), "w").write("\n".join(lines) + "\n")
print(json.dumps({k: v["dram_bytes_per_replicate"] for k, v in traffic.items()}, indent=1))
