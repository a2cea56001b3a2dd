#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv            > profiles/r1_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_fb_kernel_r1.ncu-rep       > profiles/r1_fb_kernel.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEY = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "smsp__sass_inst_executed_op_global_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"cub::CUB_\w+::", "cub::", name)
    return re.sub(r"<.*", "", name)


def launches(path):
    text = open(path).read()
    text = text[text.index('"ID"'):]
    per = OrderedDict()
    tot = 0.0
    for r in csv.DictReader(io.StringIO(text)):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        k = short(r["Kernel Name"])
        c = per.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += ns
        tot += ns
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f %% |" % (k, n, ns / 1e6, 100 * ns / tot))
    print("| **all** | %d | %.3f | 100 %% |" % (sum(v[0] for v in per.values()), tot / 1e6))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("### %s" % short(r[hdr.index("Kernel Name")]))
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for m in KEY:
            if m in hdr:
                i = hdr.index(m)
                print("| %s | %s | %s |" % (m, r[i], units[i]))
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
