#!/usr/bin/env python3
"""Extract the reference's pg_regress goldens into compact JSON fixtures.

TEST INFRASTRUCTURE.  Run once in the build container (where /root/reference
exists); the GPU box only sees the committed JSON files next to this script.

Source of the vectors (reference file -> fixture):
  expected/nogrp_agg.out    -> nogrp_agg.json      (input/sql/nogrp_agg.sql)
  expected/group_agg.out    -> group_agg.json
  expected/where_agg.out    -> where_agg.json
  expected/zero_agg.out     -> zero_agg.json
  expected/overflow_agg.out -> overflow_agg.json
  expected/recheck_agg.out  -> recheck_agg.json
  expected/explain_agg.out  -> explain_agg.json    (plans only)

Each fixture is a list of statements:
  {"sql": "...", "columns": [...], "rows": [[str|null,...],...],
   "error": str|null, "notices": [str,...]}
psql prints NULL as an empty cell; empty cells are stored as null.

Usage:  python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

ROWCOUNT = re.compile(r"^\((\d+) rows?\)$")
GUCSET = re.compile(r"^set\s+(pg_strom\.\w+)\s*(?:=|\s+to\s+)\s*([^;]+);", re.I)


def parse_out(path):
    with open(path) as f:
        lines = f.read().split("\n")
    stmts = []
    i = 0
    n = len(lines)
    while i < n:
        ln = lines[i]
        low = ln.strip().lower()
        if not (low.startswith("select") or low.startswith("explain")):
            i += 1
            continue
        # statement text may span several lines up to ';'
        sql = ln
        while not sql.rstrip().endswith(";") and i + 1 < n:
            i += 1
            sql += " " + lines[i].strip()
        i += 1
        st = {"sql": " ".join(sql.split()), "columns": None, "rows": [],
              "error": None, "notices": []}
        while i < n and (lines[i].startswith("NOTICE:") or
                         lines[i].startswith("WARNING:")):
            st["notices"].append(lines[i])
            i += 1
        if i < n and lines[i].startswith("ERROR:"):
            st["error"] = lines[i][len("ERROR:"):].strip()
            i += 1
            stmts.append(st)
            continue
        # header line, separator line, rows, "(N rows)"
        if i + 1 < n and re.match(r"^-[-+]*$", lines[i + 1] or "x"):
            st["columns"] = [c.strip() for c in lines[i].split("|")]
            i += 2
            while i < n and not ROWCOUNT.match(lines[i]):
                cells = [c.strip() for c in lines[i].split("|")]
                st["rows"].append([c if c != "" else None for c in cells])
                i += 1
            m = ROWCOUNT.match(lines[i]) if i < n else None
            if m:
                # EXPLAIN output: one text column, keep raw lines (indentation
                # is meaningful there)
                if st["columns"] == ["QUERY PLAN"]:
                    pass
                assert int(m.group(1)) == len(st["rows"]), (path, st["sql"])
                i += 1
        stmts.append(st)
    return stmts


def parse_explain(path):
    """explain_agg.out: keep each plan as raw text lines (indent matters)."""
    with open(path) as f:
        lines = f.read().split("\n")
    stmts = []
    i = 0
    n = len(lines)
    gucs = {}
    while i < n:
        ln = lines[i]
        m = GUCSET.match(ln.strip())
        if m:
            # the session's pg_strom.* settings at this point of the script
            gucs[m.group(1).lower()] = m.group(2).strip().strip("'").lower()
        if not ln.strip().lower().startswith("explain"):
            i += 1
            continue
        sql = " ".join(ln.split())
        i += 1
        if i < n and lines[i].startswith("ERROR:"):
            stmts.append({"sql": sql, "plan": None, "gucs": dict(gucs),
                          "error": lines[i][6:].strip()})
            i += 1
            continue
        assert "QUERY PLAN" in lines[i], (i, lines[i])
        i += 2
        plan = []
        while i < n and not ROWCOUNT.match(lines[i]):
            plan.append(lines[i][1:].rstrip() if lines[i].startswith(" ")
                        else lines[i].rstrip())
            i += 1
        i += 1
        stmts.append({"sql": sql, "plan": plan, "gucs": dict(gucs), "error": None})
    return stmts


def main():
    for name in ("nogrp_agg", "group_agg", "where_agg", "zero_agg",
                 "overflow_agg", "recheck_agg"):
        stmts = parse_out(os.path.join(REF, "expected", name + ".out"))
        stmts = [s for s in stmts if s["sql"].lower().startswith("select")]
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(stmts, f, separators=(",", ":"))
        print(name, len(stmts), "statements")
    ex = parse_explain(os.path.join(REF, "expected", "explain_agg.out"))
    with open(os.path.join(HERE, "explain_agg.json"), "w") as f:
        json.dump(ex, f, separators=(",", ":"))
    print("explain_agg", len(ex), "statements")


if __name__ == "__main__":
    main()
