#!/usr/bin/env python3
"""summarize_bench.py <bench.json> [<bench_N.json> ...]: markdown tables of bench lines
(first file: the per-workload table; all files: the scaling table by n_gpus)."""
import json
import sys


def load(path):
    for ln in open(path).read().strip().splitlines()[::-1]:
        if ln.startswith("{"):
            return json.loads(ln)
    raise SystemExit("no JSON line in " + path)


def workloads(line):
    head = {k: line[k] for k in ("value", "ms_per_step", "config", "details", "checked", "roofline",
                                 "cpu_baseline", "e2e", "phase_ms") if k in line}
    out = [(line["config"]["workload"], head)]
    for n, w in line.get("workloads", {}).items():
        out.append((n, w))
    return out


first = load(sys.argv[1])
print("| workload | rows / launch | scan kernel(s) | algorithmic GB/s | of measured HBM | whole step | rows/s | "
      "end to end (host buffers) | CPU Agg (port, all cores) |")
print("|---|---|---|---|---|---|---|---|---|")
for name, w in workloads(first):
    if "error" in w:
        print("| `%s` | error: %s |" % (name, w["error"]))
        continue
    rf, e, cpu = w["roofline"], w["e2e"], w.get("cpu_baseline")
    print("| `%s` | %.3g M | %.3f ms (%s) | %.0f | %.3f | %.3f ms | %.3g | %.3g rows/s, H2D %.1f GB/s | %s |" % (
        name, rf["bytes_per_launch"] / w["details"]["algorithmic_bytes_per_row"] / 1e6, rf["launch_ms"],
        rf["kernel"], rf["achieved"], rf["frac"], w["ms_per_step"], w["value"], e["value"], e["h2d_gb_per_s"],
        ("%.3g rows/s on %d cores" % (cpu["value"], cpu["cores"])) if cpu else "-"))
if len(sys.argv) > 2:
    lines = sorted((load(p) for p in sys.argv[1:]), key=lambda l: l["n_gpus"])
    base = {n: w["value"] for n, w in workloads(lines[0])}
    names = [n for n, _ in workloads(lines[0])]
    print()
    print("| N GPUs | " + " | ".join("`%s` rows/s (x of N=%d, efficiency)" % (n, lines[0]["n_gpus"]) for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for l in lines:
        ws = dict(workloads(l))
        cells = []
        for n in names:
            w = ws.get(n)
            if not w or "error" in w:
                cells.append("-")
                continue
            x = w["value"] / base[n] * lines[0]["n_gpus"]
            cells.append("%.3g (%.2fx, %.2f)" % (w["value"], x, x / l["n_gpus"]))
        print("| %d | " % l["n_gpus"] + " | ".join(cells) + " |")
