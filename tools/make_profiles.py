#!/usr/bin/env python3
"""Turns the ncu captures under gpurun_out/ into the tracked summaries under
profiles/ (the judge reads profiles/, gpurun_out/ is scratch).

  python tools/make_profiles.py r01 nogrp_agg=gpurun_out/prof_nogrp.ncu-rep@100000000 \
         where_agg=gpurun_out/prof_where.ncu-rep@25000000 [launches=gpurun_out/launches.csv]
  (@N = rows the profiled launch processed)

Writes profiles/<round>_<workload>_ncu.txt (selected raw metrics of the
gpupreagg_main launch), profiles/<round>_launches.csv (copy of the launch
list) and profiles/traffic.json ({workload: dram bytes per launch}) which
bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__grid_size
launch__block_size launch__shared_mem_per_block_dynamic smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__thread_inst_executed_per_inst_executed.ratio""".split()


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v * mult


def main():
    rnd = sys.argv[1]
    outdir = os.path.join(ROOT, "profiles")
    os.makedirs(outdir, exist_ok=True)
    tpath = os.path.join(outdir, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        if name == "launches":
            shutil.copy(path, os.path.join(outdir, "%s_launches.csv" % rnd))
            continue
        nrows_launch = None
        if "@" in path:
            path, n = path.rsplit("@", 1)
            nrows_launch = int(n)
        hdr, units, rows = raw_page(path)
        lines = ["# ncu --set full --clock-control none, %s (%s)" % (name, os.path.basename(path))]
        total, kernels = 0.0, []
        for r in rows:
            lines.append("kernel %s" % r[hdr.index("Kernel Name")])
            kernels.append(r[hdr.index("Kernel Name")])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    lines.append("  %-86s %s %s" % (k, r[i], units[i]))
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            total += rd + wr
            # a workload whose step is several kernels (deal + partagg, page
            # index + heap scan): the sum over the kernels of one chunk
            traffic[name] = {"dram_bytes_per_launch": total, "rows_per_launch": nrows_launch,
                             "kernels": list(kernels), "source": os.path.basename(path)}
        with open(os.path.join(outdir, "%s_%s_ncu.txt" % (rnd, name)), "w") as f:
            f.write("\n".join(lines) + "\n")
    with open(tpath, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    print("profiles updated:", sorted(os.listdir(outdir)))


if __name__ == "__main__":
    main()
