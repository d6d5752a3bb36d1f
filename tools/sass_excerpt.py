#!/usr/bin/env python3
"""sass_excerpt.py <round>: disassembles the device programs of the bench
workloads (built with NVRTC for sm_100a through the library, no GPU needed) and
writes profiles/<round>_sass_<workload>.txt: per kernel the resource usage, a
census of the mnemonics that matter (UBLKCP = cp.async.bulk / TMA bulk copy,
SYNCS = mbarrier, LDS.128, ATOMS/ATOMG/RED, STG.256 ...) and the SASS lines
around every UBLKCP / SYNCS / 256-bit store."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pg_strom_b200 import _capi, gpupreagg as gp, workloads as W  # noqa: E402

CENSUS = ["UBLKCP", "UTMALDG", "SYNCS", "LDS.128", "LDS.64", "LDS", "STS", "ATOMS", "ATOMG", "RED",
          "LDG.E.128", "LDG", "STG.E.ENL2.256", "STG.E.128", "STG", "SHFL", "VOTE", "POPC", "DFMA", "DADD",
          "IMAD.WIDE", "HMMA", "UTCHMMA", "CCTL", "BAR.SYNC", "WARPSYNC", "NANOSLEEP"]
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on", "pg_strom.perfmon": "on"}


def main():
    rnd = sys.argv[1]
    lib = _capi.load()
    for name, mk in (("nogrp_agg", W.nogrp_plan), ("where_agg", W.where_plan),
                     ("high_cardinality", W.hc_plan)):
        plan = gp.Plan(mk(), gucs=GUCS)
        prog = plan.build_program()
        n = C.c_size_t()
        p = lib.pgs_program_cubin(prog, C.byref(n))
        cubin = C.string_at(p, n.value)
        with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
            f.write(cubin)
            path = f.name
        out = ["# %s: cuobjdump of the NVRTC cubin (sm_100a) the bench runs; tools/sass_excerpt.py" % name]
        out.append("# elf: " + subprocess.run(["cuobjdump", "-lelf", path], capture_output=True,
                                               text=True).stdout.strip())
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        funcs = re.split(r"\n\s*Function : ", sass)[1:]
        for fn in funcs:
            fname = fn.split("\n", 1)[0].strip()
            lines = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip() for ln in fn.splitlines()
                     if re.search(r"/\*[0-9a-f]{4,}\*/", ln)]
            if len(lines) < 40:
                continue
            out.append("")
            out.append("== %s: %d SASS instructions" % (fname, len(lines)))
            m = re.search(r"Function %s:\n\s*(.*)" % re.escape(fname), res)
            if m:
                out.append("   " + m.group(1).strip())
            counts = []
            for mn in CENSUS:
                c = sum(1 for ln in lines if re.search(r"\b%s\b" % re.escape(mn), ln))
                if c:
                    counts.append("%s=%d" % (mn, c))
            out.append("   " + " ".join(counts))
            if fname.startswith("gpupreagg_main") or fname == "gpupreagg_partagg":
                shown = set()
                for i, ln in enumerate(lines):
                    if re.search(r"UBLKCP|SYNCS\.(ARRIVE|PHASECHK|EXCH)|\.256|ATOMS|ATOMG", ln):
                        for j in range(max(0, i - 1), min(len(lines), i + 2)):
                            if j not in shown:
                                shown.add(j)
                shown = sorted(shown)[:160]
                prev = None
                for j in shown:
                    if prev is not None and j != prev + 1:
                        out.append("        ...")
                    out.append("   " + lines[j].strip()[:120])
                    prev = j
        os.unlink(path)
        lib.pgs_program_release(prog)
        plan.free()
        dst = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (rnd, name))
        with open(dst, "w") as f:
            f.write("\n".join(out) + "\n")
        print("wrote", dst, len(out), "lines")


if __name__ == "__main__":
    main()
