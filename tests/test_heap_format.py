"""CPU: the KDS_FORMAT_ROW / ROW_FLAT builders (datastore.cpp, following
datastore.c:382-470,556-710,799-823) and the synthetic heap pages they carry,
checked by an independent decoder that walks the chunk the way
kern_get_tuple_rs / kern_get_datum_tuple do (opencl_common.h:817-947)."""
import ctypes as C
import struct

import numpy as np
import pytest

from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200._capi import kern_data_store

BLCKSZ = 8192
ALIGN = lambda a, v: (v + a - 1) & ~(a - 1)     # noqa: E731


def _varsize_any(b, o):
    h = b[o]
    if h == 0x01:
        return 2 + (8 if b[o + 1] == 1 else 16)
    if h & 1:
        return (h >> 1) & 0x7f
    return (struct.unpack_from("<I", b, o)[0] >> 2) & 0x3fffffff


def deform(img, htup, coltypes):
    """kern_get_datum_tuple for every column; varlena -> payload bytes."""
    infomask2, infomask, hoff = struct.unpack_from("<HHB", img, htup + 18)
    natts = infomask2 & 0x07ff
    hasnull = infomask & 1
    off = hoff
    out = []
    for i, t in enumerate(coltypes):
        attlen, _, align, dt = gp.PGTYPES[t]
        if i >= natts or (hasnull and not (img[htup + 23 + (i >> 3)] >> (i & 7)) & 1):
            out.append(None)
            continue
        if attlen > 0:
            off = ALIGN(align, off)
            out.append(np.frombuffer(img, dtype=dt, count=1, offset=htup + off)[0])
            off += attlen
        else:
            if img[htup + off] == 0:
                off = ALIGN(align, off)
            n = _varsize_any(img, htup + off)
            hdr = 1 if img[htup + off] & 1 else 4
            out.append(bytes(img[htup + off + hdr:htup + off + n]))
            off += n
    return out


def rows_of(img, coltypes):
    kds = kern_data_store.from_buffer_copy(img[:48])
    ncols = kds.ncols
    head = ALIGN(16, 48 + 8 * ncols)
    items = head + ALIGN(16, 16 * kds.maxblocks)
    rows = []
    if kds.format == gp.KDS_FORMAT_ROW:
        first = ALIGN(BLCKSZ, items + ALIGN(16, 4 * kds.nitems))
        for r in range(kds.nitems):
            blk, item = struct.unpack_from("<HH", img, items + 4 * r)
            assert blk < kds.nblocks
            page = first + BLCKSZ * blk
            lower, = struct.unpack_from("<H", img, page + 12)
            assert 1 <= item <= (lower - 24) // 4
            lp, = struct.unpack_from("<I", img, page + 24 + 4 * (item - 1))
            assert (lp >> 15) & 3 == 1 and (lp & 0x7fff) % 8 == 0
            rows.append(deform(img, page + (lp & 0x7fff), coltypes))
    else:
        assert kds.format == gp.KDS_FORMAT_ROW_FLAT
        for r in range(kds.nitems):
            off, = struct.unpack_from("<I", img, items + 4 * r)
            assert off % 8 == 0 and off < kds.length
            rows.append(deform(img, off, coltypes))
    return kds, rows


def _table(n, seed=1):
    rng = np.random.default_rng(seed)
    coltypes = ["int4", "int2", "float8", "numeric", "int8", "float4", "bool"]
    cols = [
        (rng.integers(-2**31, 2**31 - 1, n, dtype=np.int64).astype(np.int32), rng.random(n) < 0.1),
        (rng.integers(-30000, 30000, n).astype(np.int16), None),
        (rng.standard_normal(n), rng.random(n) < 0.5),
        ([None if rng.random() < 0.2 else gp.numeric_datum("%d.%04d" % (rng.integers(0, 10**9),
                                                                        rng.integers(0, 10**4)))
          for _ in range(n)], None),
        (rng.integers(-2**62, 2**62, n, dtype=np.int64), rng.random(n) < 0.05),
        (rng.standard_normal(n).astype(np.float32), None),
        None,
    ]
    return coltypes, cols


@pytest.mark.parametrize("flat", [False, True])
def test_round_trip(lib, flat):
    n = 5000
    coltypes, cols = _table(n)
    vis = np.random.default_rng(2).random(n) < 0.9
    ds = gp.HeapDataStore(coltypes, cols, nrows=n, flat=flat, visible=vis)
    try:
        img = ds.device_image()
        kds, rows = rows_of(img, coltypes)
        assert kds.nitems == int(vis.sum()) == len(rows)
        if not flat:
            assert kds.nblocks == ds.npages and kds.nblocks <= kds.maxblocks
        src = np.flatnonzero(vis)
        for got, r in zip(rows, src):
            for c, col in enumerate(cols):
                if col is None:
                    assert got[c] is None
                    continue
                v, m = col
                isnull = (m is not None and bool(m[r])) or v[r] is None
                if isnull:
                    assert got[c] is None
                elif gp.PGTYPES[coltypes[c]][0] > 0:
                    assert got[c] == v[r] or (got[c] != got[c] and v[r] != v[r])
                else:
                    assert got[c] == v[r][4:]       # payload behind the 4-byte header
    finally:
        ds.free()


def test_row_store_full(lib):
    """pgstrom_data_store_insert_block refuses a block that does not fit
    (datastore.c:604-613)."""
    coltypes, cols = _table(400)
    ds = gp.HeapDataStore(coltypes, cols, nrows=400)
    try:
        ncols = len(coltypes)
        ln = lib.pgstrom_kds_row_length(ncols, 1, 50)
        buf = C.create_string_buffer(ln)
        assert lib.pgstrom_kds_row_init(buf, ln, ncols, ds.colmeta, 1, 50) == 0
        offs = (C.c_uint16 * 4)(1, 2, 3, 4)
        # 1 of 1 block slots may never be used ("we never use all the block slots")
        assert lib.pgstrom_kds_row_insert_block(buf, ds._pages, offs, 4) == -1
    finally:
        ds.free()


def test_row_store_block_limit(lib):
    """kern_rowitem.blk_index has 16 bits (opencl_common.h:395-401): a
    KDS_FORMAT_ROW chunk that would need more than 65536 block items is
    refused instead of wrapping the page index."""
    coltypes = ["int4"]
    colmeta = gp.make_colmeta(coltypes)
    ln = lib.pgstrom_kds_row_length(1, 65537, 10)
    buf = C.create_string_buffer(int(ln))
    assert lib.pgstrom_kds_row_init(buf, ln, 1, colmeta, 65537, 10) == 301
    ln = lib.pgstrom_kds_row_length(1, 65536, 10)
    buf = C.create_string_buffer(int(ln))
    assert lib.pgstrom_kds_row_init(buf, ln, 1, colmeta, 65536, 10) == 0
