#!/usr/bin/env python3
"""srcprof.py <report.ncu-rep> [top]: per-SASS-instruction samples and executed
counts of the profiled kernel (ncu --page source), hottest first, plus totals
per 64-instruction region.  argv[3]: index of the kernel in the output, argv[4]: kernel name regex."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kfilter = ["-k", "regex:" + sys.argv[4]] if len(sys.argv) > 4 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + kfilter,
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]; rows = [r for r in rows[1:] if len(r) == len(hdr)]
# a report with several kernels repeats the header: pick the kernel by index (argv[3])
kernels = [[]]
for r in rows:
    if r == hdr:
        kernels.append([])
    else:
        kernels[-1].append(r)
rows = kernels[int(sys.argv[3]) if len(sys.argv) > 3 else 0]
ia, isrc, isamp, iexec, ithr = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[0][ia], 16)
tot = sum(int(r[isamp]) for r in rows); texec = sum(int(r[iexec]) for r in rows)
print("total samples", tot, "total warp-inst executed", texec, "instructions", len(rows))
print("-- per region (64 inst): offset samples% exec%")
for b in range(0, len(rows), 64):
    s = sum(int(r[isamp]) for r in rows[b:b + 64]); e = sum(int(r[iexec]) for r in rows[b:b + 64])
    print("  %05x  %5.1f%%  %5.1f%%" % (int(rows[b][ia], 16) - base, 100.0 * s / tot, 100.0 * e / texec))
print("-- hottest instructions")
for r in sorted(rows, key=lambda r: -int(r[isamp]))[:top]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print("  %05x %5.2f%% exec %9s thr %5s  %-60s %s" % (int(r[ia], 16) - base, 100.0 * int(r[isamp]) / tot, r[iexec], r[ithr],
          r[isrc].strip()[:60], " ".join("%s:%d" % (n, v) for v, n in st)))
