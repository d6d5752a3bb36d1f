#!/usr/bin/env python3
"""gpu_keyheap_check.py: a quick check of long text grouping keys on a GPU
(the device key heap of kern_textlib.cuh) without pytest and without the
oracle: GROUP BY (text, character(5)) over 24 000 generated rows in column
and heap-page chunks, with and without a WHERE clause, compared with a
python dict.  Prints one JSON line per case; exit code 0 = all equal."""
import json
import os
import random
import struct
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                          # noqa: E402
from pg_strom_b200 import gpupreagg as gp, pgplan as P      # noqa: E402

GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}


def varlena(b, short):
    if short and len(b) < 127:
        return bytes([((len(b) + 1) << 1) | 1]) + b
    return struct.pack("<I", (len(b) + 4) << 2) + b


def make_rows(n, seed):
    rng = random.Random(seed)
    cats = [bytes([97 + i]) * 3 for i in range(26)] + [b"", b"a", b"abcdefg", b"caf\xc3\xa9"]
    long_cats = [b"abcdefgh", b"category-with-a-long-name", b"x" * 300] + \
        [b"long-category-%05d" % i for i in range(400)]
    codes = [c.ljust(5) for c in (b"ab", b"abcde", b"x", b"", b"a b")] + [b"caf\xc3\xa9 "]
    rows = []
    for _ in range(n):
        cat = rng.choice(long_cats) if rng.random() < 0.4 else rng.choice(cats)
        rows.append((None if rng.random() < 0.04 else cat,
                     None if rng.random() < 0.04 else rng.choice(codes),
                     rng.randrange(0, 100), rng.randrange(-10 ** 15, 10 ** 15)))
    return rows


def main():
    t = P.Table("cats", [("cat", "text"), ("code", "bpchar"), ("f", "int4"), ("v", "int8")],
                typmods={"code": 4 + 5})
    rows = make_rows(24000, 31)
    coltypes = [c for _, c in t.columns]
    gp.cuda_init()
    ok = True
    for fmt in ("column", "row"):
        for with_qual in (False, True):
            t0 = time.time()
            tree = P.make_agg_plan(
                t, [(t.col("cat"), "cat"), (t.col("code"), "code"),
                    (P.Agg("count", star=True), "count"), (P.Agg("sum", [t.col("f")]), "sum"),
                    (P.Agg("min", [t.col("v")]), "min")],
                group_by=["cat", "code"], num_groups=3000,
                where=[P.Op("<", t.col("f"), P.Const("int4", 50))] if with_qual else [])
            plan = gp.Plan(tree, gucs=GUCS)
            assert plan.num_gpupreagg == 1, plan.reject_reason
            chunks = []
            for lo in range(0, len(rows), 9000):
                part = rows[lo:lo + 9000]
                cols = []
                for c, typ in enumerate(coltypes):
                    raw = [r[c] for r in part]
                    if gp.PGTYPES[typ][0] > 0:
                        cols.append((np.array(raw, dtype=gp.PGTYPES[typ][3]), None))
                    else:
                        cols.append(([None if v is None else varlena(v, i % 2 == 1)
                                      for i, v in enumerate(raw)], None))
                chunks.append(gp.DataStore(coltypes, cols, nrows=len(part)) if fmt == "column"
                              else gp.HeapDataStore(coltypes, cols, nrows=len(part)))
            st = gp.GpuPreAggState(plan, chunks)
            try:
                device_rows = st.fetch_all()
                recheck = st.recheck_rows()
            finally:
                st.end()
            for ds in chunks:
                ds.free()
            plan.free()
            exp = {}
            for cat, code, f, v in rows:
                if with_qual and not f < 50:
                    continue
                e = exp.setdefault((cat, code), [0, 0, None])
                e[0] += 1
                e[1] += f
                e[2] = v if e[2] is None else min(e[2], v)
            got = {}
            dup = 0
            for r in device_rows:
                key = (r[0], r[1])
                dup += key in got
                e = got.setdefault(key, [0, 0, None])
                e[0] += r[4]
                e[1] += r[5]
                e[2] = r[6] if e[2] is None else min(e[2], r[6])
            same = (got == exp)
            ok = ok and same and not recheck and dup == 0
            print(json.dumps({"format": fmt, "where": with_qual, "rows": len(rows),
                              "groups_expected": len(exp), "device_rows": len(device_rows),
                              "duplicate_keys": dup, "recheck_rows": len(recheck),
                              "longest_key": max((len(k[0]) for k in got if k[0]), default=0),
                              "equal": same, "seconds": round(time.time() - t0, 2)}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
