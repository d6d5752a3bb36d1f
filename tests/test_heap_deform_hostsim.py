"""CPU: the tuple de-forming of the heap-page kernel (pgs_heap_tuple /
pgs_heap_deform in kern_gpupreagg.cuh, the counterpart of kern_get_tuple_rs /
kern_get_datum_tuple, /root/reference/opencl_common.h:817-899) cut out of the
real header and run under g++ on chunks built by the product's own builders
(KDS_FORMAT_ROW with heap pages, KDS_FORMAT_ROW_FLAT): random column type
mixes, NULL bitmaps, alignment padding, 1- and 4-byte varlena headers - every
referenced attribute of every row must come back as it went in - and pages /
tuples that do not look like heap data must be refused, not followed."""
import ctypes as C
import os
import random
import struct
import subprocess

import numpy as np
import pytest

from oracle import pg_typelib as T
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P
from test_codegen_hostsim import CSRC, GUCS, HOST_PREAMBLE, PRE, ROOT, SIM, _cut, pack

TYPES = ["bool", "int2", "int4", "int8", "float4", "float8", "date", "timestamp", "text",
         "numeric"]

WRAPPER = r'''
extern "C" int heap_row(const void *kds, unsigned int rowidx,
                        unsigned long long *vals, unsigned int *valid)
{
    pgs_heap_chunk  hc;
    kern_row_regs   rr;
    cl_uint         avail = 0;

    pgs_heap_chunk_init(hc, (const kern_data_store *)kds);
    const unsigned char *htup = pgs_heap_tuple(hc, rowidx, &avail);
    if (!htup)
        return 1;
    if (!pgs_heap_deform((const kern_data_store *)kds, htup, avail, rr))
        return 2;
    *valid = 0;
    for (int s = 0; s < GPUPREAGG_NUM_INCOLS; s++)
    {
        vals[s] = rr.v[s];
        if (rr.vbits[s] & 1U)
            *valid |= (1U << s);
    }
    return 0;
}
'''


@pytest.fixture(scope="module")
def heapsim(lib):
    os.makedirs(SIM, exist_ok=True)
    d = os.path.join(SIM, "heap")
    os.makedirs(d, exist_ok=True)
    common = os.path.join(CSRC, "kern_common.cuh")
    regs = _cut(common, "struct kern_row_regs", "\n};\n") + "\n};\n"
    with open(os.path.join(d, "kern_common.cuh"), "w") as f:
        f.write("#pragma once\n" + HOST_PREAMBLE + regs)
    heap = _cut(os.path.join(CSRC, "kern_gpupreagg.cuh"),
                "#define PGS_HEAP_HASNULL", "/*\n * gpupreagg_main_heap - heap-page chunks")
    with open(os.path.join(d, "kern_gpupreagg.cuh"), "w") as f:
        f.write("#pragma once\n" + heap)
    for name in ("kern_numeric.cuh", "kern_timelib.cuh", "kern_textlib.cuh"):
        with open(os.path.join(d, name), "w") as f:
            f.write("#pragma once\n")           # the expression runtime is not needed here
    return d


_n = [0]


def build(plan, d):
    _n[0] += 1
    src = os.path.join(d, "h%d.cpp" % _n[0])
    out = os.path.join(d, "h%d.so" % _n[0])
    text = plan.kernel_source()
    # only the compile-time description of the query is wanted, not the
    # generated functions (they need the whole expression runtime)
    text = text[:text.index('#include "kern_gpupreagg.cuh"')] + '#include "kern_gpupreagg.cuh"\n'
    with open(src, "w") as f:
        f.write(PRE + text + WRAPPER)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w",
                        "-I", d, "-I", os.path.join(ROOT, "include"), "-o", out, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:3000]
    so = C.CDLL(out)
    so.heap_row.argtypes = [C.c_char_p, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint)]
    return so


def rand_table(rng):
    ncols = rng.randrange(2, 10)
    types = [rng.choice(TYPES) for _ in range(ncols)]
    return P.Table("h", [("c%d" % i, t) for i, t in enumerate(types)])


def rand_datum(typ, rng):
    if typ == "bool":
        return rng.random() < 0.5
    if typ in ("int2", "int4", "int8"):
        b = {"int2": 15, "int4": 31, "int8": 63}[typ]
        return rng.randrange(-2 ** b, 2 ** b)
    if typ == "float4":
        return struct.unpack("f", struct.pack("f", rng.uniform(-1e6, 1e6)))[0]
    if typ == "float8":
        return rng.uniform(-1e12, 1e12)
    if typ == "date":
        return rng.randrange(-10000, 10000)
    if typ == "timestamp":
        return rng.randrange(-10 ** 15, 10 ** 15)
    if typ == "text":
        n = rng.choice([0, 1, 3, 7, 20, 126, 127, 300])
        return bytes(rng.randrange(32, 127) for _ in range(n))
    return gp.numeric_datum("%d.%02d" % (rng.randrange(-10 ** 9, 10 ** 9), rng.randrange(0, 100)))


def make_store(table, rows, flat, rng):
    coltypes = [t for _, t in table.columns]
    cols = []
    for c, typ in enumerate(coltypes):
        raw = [r[c] for r in rows]
        if gp.PGTYPES[typ][0] > 0:
            mask = np.array([v is None for v in raw], dtype=np.uint8)
            arr = np.array([0 if v is None else v for v in raw], dtype=gp.PGTYPES[typ][3])
            cols.append((arr, mask if mask.any() else None))
        elif typ == "text":
            cols.append(([None if v is None else
                          T.varlena(v, short=None if rng.random() < 0.7 else False) for v in raw],
                         None))
        else:
            cols.append((raw, None))            # numeric: already a varlena image
    return gp.HeapDataStore(coltypes, cols, nrows=len(rows), flat=flat)


def varlena_bytes(img):
    b0 = img[0]
    if b0 & 1:
        return img[:(b0 >> 1) & 0x7f]
    return img[:(struct.unpack_from("<I", img, 0)[0] >> 2) & 0x3fffffff]


def payload(img):
    return img[1:] if img[0] & 1 else img[4:]


@pytest.mark.parametrize("flat", [False, True])
def test_deform_round_trip(heapsim, flat):
    rng = random.Random(11 if flat else 12)
    checked = 0
    for _ in range(12):
        table = rand_table(rng)
        types = [t for _, t in table.columns]
        nullfrac = rng.choice([0.0, 0.1, 0.5])
        rows = [tuple(None if rng.random() < nullfrac else rand_datum(t, rng) for t in types)
                for _ in range(rng.choice([1, 7, 300]))]
        ref = sorted(rng.sample(range(len(types)), rng.randrange(1, len(types) + 1)))
        tree = P.make_agg_plan(table, [(P.Agg("count", star=True), "count")] +
                               [(P.Agg("count", [table.col("c%d" % c)]), "count") for c in ref])
        plan = gp.Plan(tree, gucs=GUCS)
        try:
            assert plan.num_gpupreagg == 1, plan.reject_reason
            incols = plan.describe()["incol_index"]
            assert incols == ref
            so = build(plan, heapsim)
        finally:
            plan.free()
        ds = make_store(table, rows, flat, rng)
        image = ds.device_image()
        assert ds.nitems == len(rows)
        vals = (C.c_uint64 * len(incols))()
        valid = C.c_uint()
        for i, row in enumerate(rows):
            assert so.heap_row(image, i, vals, C.byref(valid)) == 0, (types, i)
            for slot, c in enumerate(incols):
                v = row[c]
                assert bool((valid.value >> slot) & 1) == (v is not None), (types, i, c)
                if v is None:
                    continue
                if gp.PGTYPES[types[c]][0] > 0:
                    assert vals[slot] == pack(v, types[c]), (types, i, c, v)
                else:
                    got = varlena_bytes(image[vals[slot]:vals[slot] + 400])
                    if types[c] == "text":
                        assert payload(got) == v, (types, i, c)
                    else:
                        assert payload(got) == payload(bytes(v)), (types, i, c)
                checked += 1
        assert so.heap_row(image, len(rows) - 1, vals, C.byref(valid)) == 0
        ds.free()
    assert checked > 3000


def test_refuses_what_is_not_a_heap_page(heapsim):
    """kern_get_tuple_rs's sanity checks (opencl_common.h:866-899) and the
    bounds of the attribute walk: StromError_DataStoreCorruption instead of a
    wild address."""
    rng = random.Random(5)
    table = P.Table("h", [("a", "int4"), ("b", "text"), ("c", "int8")])
    rows = [(i, b"row %d" % i, i * 1000) for i in range(50)]
    tree = P.make_agg_plan(table, [(P.Agg("count", [table.col(n)]), "count") for n in "abc"])
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        so = build(plan, heapsim)
    finally:
        plan.free()
    ds = make_store(table, rows, False, rng)
    good = ds.device_image()
    headlen, first, total = ds.device_layout()
    ds.free()
    vals = (C.c_uint64 * 3)()
    valid = C.c_uint()
    assert so.heap_row(good, 3, vals, C.byref(valid)) == 0 and vals[0] == 3

    def patched(off, data):
        b = bytearray(good)
        b[off:off + len(data)] = data
        return bytes(b)
    kds = gp.kern_data_store.from_buffer_copy(good[:64])
    ncols = 3
    head = 48 + 8 * ncols
    head = (head + 15) & ~15
    items = head + ((16 * kds.maxblocks + 15) & ~15)
    # row item 3: block index beyond nblocks / line pointer number 0 / beyond pd_lower
    assert so.heap_row(patched(items + 4 * 3, struct.pack("<HH", 999, 4)), 3, vals, C.byref(valid)) == 1
    assert so.heap_row(patched(items + 4 * 3, struct.pack("<HH", 0, 0)), 3, vals, C.byref(valid)) == 1
    assert so.heap_row(patched(items + 4 * 3, struct.pack("<HH", 0, 2000)), 3, vals, C.byref(valid)) == 1
    # line pointer 4 of page 0: unused (flags 0) / unaligned offset / offset at the page end
    lp_off = first + 24 + 4 * 3
    lp, = struct.unpack_from("<I", good, lp_off)
    assert so.heap_row(patched(lp_off, struct.pack("<I", lp & ~(3 << 15))), 3, vals, C.byref(valid)) == 1
    assert so.heap_row(patched(lp_off, struct.pack("<I", lp | 4)), 3, vals, C.byref(valid)) == 1
    assert so.heap_row(patched(lp_off, struct.pack("<I", (lp & ~0x7fff) | 8184)), 3, vals, C.byref(valid)) == 1
    # tuple header: t_hoff too small for the header; attribute walk past the page end
    tup = first + (lp & 0x7fff)
    assert so.heap_row(patched(tup + 22, bytes([8])), 3, vals, C.byref(valid)) == 2
    assert so.heap_row(patched(tup + 22, bytes([248])), 3, vals, C.byref(valid)) in (0, 2)
    # pd_lower says the page has no line pointers at all
    assert so.heap_row(patched(first + 12, struct.pack("<H", 24)), 3, vals, C.byref(valid)) == 1
    # the untouched image still reads
    assert so.heap_row(good, 49, vals, C.byref(valid)) == 0 and vals[2] == 49000
