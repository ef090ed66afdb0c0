#!/usr/bin/env python
"""Turns the ncu artefacts a GPU call left in gpurun_out/ into the committed summaries under profiles/.

    python scripts/ncu_summary.py r01a [--launches gpurun_out/launches.csv] [--reps gpurun_out/prof_fwd.ncu-rep ...]

Writes profiles/<tag>_launches.md (per-kernel launch count, average device time and share of the profiled
command; ncu times are cold-cache and serialised, so SHARES are the comparable figure) and
profiles/<tag>_<rep>.md (the raw-page metrics that back the roofline numbers in bench.py / DESIGN.md)."""
from __future__ import annotations

import argparse
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tc.sum",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    # shared-memory side of the tensor-core kernels: UMMA operand reads (tc data pipe), LSU wavefronts, data-bank load
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_reads.max.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.max.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tc.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches_md(path: str, tag: str) -> str:
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ["# %s — ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)" % tag, "",
           "Source: `%s` (%d launches). ncu serialises launches and runs them cold, so the SHARE column is the "
           "figure that must agree with bench.py's CUDA-event timing, not the absolute time." % (path, len(rows)), "",
           "| kernel | launches | avg us | total us | share | grid | block |", "|---|---:|---:|---:|---:|---|---|"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f | %.1f%% | %s | %s |" % (k, a[0], a[1] / a[0] / 1e3, a[1] / 1e3,
                                                                  100 * a[1] / tot, a[2], a[3]))
    return "\n".join(out) + "\n"


def rep_md(path: str, tag: str) -> str:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    rows = [r for r in rows if r and not r[0].startswith("==")]
    hdr, units, body = rows[0], rows[1], rows[2:]
    out = ["# %s — `ncu --set full --clock-control none` capture: %s" % (tag, os.path.basename(path)), ""]
    for r in body:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        out += ["## `%s`" % name, "", "| metric | value | unit |", "|---|---:|---|"]
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append("| %s | %s | %s |" % (m, r[i], units[i]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]) * unit_scale(units[hdr.index("dram__bytes_read.sum")])
            wr = float(r[hdr.index("dram__bytes_write.sum")]) * unit_scale(units[hdr.index("dram__bytes_write.sum")])
            t = float(r[hdr.index("gpu__time_duration.sum")]) * time_scale(units[hdr.index("gpu__time_duration.sum")])
            out.append("| **DRAM traffic per launch (read + write)** | %.1f | MB |" % ((rd + wr) / 1e6))
            out.append("| **DRAM GB/s under ncu (cold, serialised)** | %.0f | GB/s |" % ((rd + wr) / t / 1e9))
        except Exception:
            pass
        out.append("")
    return "\n".join(out) + "\n"


def rep_traffic(path: str) -> dict:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw)) if r and not r[0].startswith("==")]
    hdr, units, body = rows[0], rows[1], rows[2:]
    out = {}
    for r in body:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "").split("<")[0]
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * unit_scale(units[hdr.index("dram__bytes_read.sum")])
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * unit_scale(units[hdr.index("dram__bytes_write.sum")])
        out[name] = rd + wr
    return out


def rep_l2(path: str) -> dict:
    """L2 -> SM bytes per launch (l1tex__m_xbar2l1tex_read_bytes.sum): basis rows re-read from L2 show up here"""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw)) if r and not r[0].startswith("==")]
    hdr, units, body = rows[0], rows[1], rows[2:]
    out = {}
    key = "l1tex__m_xbar2l1tex_read_bytes.sum"
    if key not in hdr:
        return out
    for r in body:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "").split("<")[0]
        out[name] = float(r[hdr.index(key)]) * unit_scale(units[hdr.index(key)])
    return out


def unit_scale(u: str) -> float:
    u = u.lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


def time_scale(u: str) -> float:
    return {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u.lower(), 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches", default=os.path.join(ROOT, "gpurun_out", "launches.csv"))
    ap.add_argument("--reps", nargs="*", default=None)
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if os.path.exists(a.launches):
        with open(os.path.join(ROOT, "profiles", "%s_launches.md" % a.tag), "w") as f:
            f.write(launches_md(a.launches, a.tag))
    reps = a.reps
    if reps is None:
        d = os.path.join(ROOT, "gpurun_out")
        reps = [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".ncu-rep")]
    traffic, l2 = {}, {}
    for rep in reps:
        name = os.path.splitext(os.path.basename(rep))[0]
        with open(os.path.join(ROOT, "profiles", "%s_%s.md" % (a.tag, name)), "w") as f:
            f.write(rep_md(rep, a.tag))
        traffic.update(rep_traffic(rep))
        l2.update(rep_l2(rep))
    if traffic:
        # dram__bytes_read.sum + dram__bytes_write.sum per launch: what bench.py reports as roofline.traffic
        import json
        with open(os.path.join(ROOT, "profiles", "%s_traffic.json" % a.tag), "w") as f:
            json.dump({"source": "ncu --set full --clock-control none, %s" % a.tag, "bytes_per_launch": traffic,
                       "l2_to_sm_bytes_per_launch": l2}, f, indent=1)
    print("wrote profiles/%s_*" % a.tag)


if __name__ == "__main__":
    sys.exit(main())
