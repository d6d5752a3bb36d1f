"""Thin Python driver over the C ABI of the GpuPreAgg path.

Everything that computes lives in libpgstrom_cuda.so (CUDA kernels built by
NVRTC for sm_100a + the C++ host layer).  This module only marshals: plan
JSON in, chunks in, partial rows out.  Names follow the reference
(gpupreagg_begin / _exec / _end, pgstrom_data_store ...).
"""
import ctypes as C
import json
import struct

import numpy as np

from . import _capi
from ._capi import kern_colmeta, kern_data_store, pgs_session_config, check

# typname -> (attlen, attbyval, attalign, numpy dtype)
PGTYPES = {
    "bool": (1, 1, 1, np.int8), "int2": (2, 1, 2, np.int16), "int4": (4, 1, 4, np.int32),
    "int8": (8, 1, 8, np.int64), "float4": (4, 1, 4, np.float32),
    "float8": (8, 1, 8, np.float64), "date": (4, 1, 4, np.int32),
    "time": (8, 1, 8, np.int64), "timestamp": (8, 1, 8, np.int64),
    "numeric": (-1, 0, 4, None), "text": (-1, 0, 4, None), "bytea": (-1, 0, 4, None),
    "bpchar": (-1, 0, 4, None),
}


def make_colmeta(coltypes):
    arr = (kern_colmeta * len(coltypes))()
    for i, t in enumerate(coltypes):
        attlen, byval, align, _ = PGTYPES[t]
        arr[i].attbyval = byval
        arr[i].attalign = align
        arr[i].attlen = attlen
        arr[i].attnum = i + 1
        arr[i].attcacheoff = -1
    return arr


def numeric_datum(text):
    """PostgreSQL numeric varlena image of a decimal literal."""
    lib = _capi.load()
    buf = C.create_string_buffer(len(text) + 64)
    n = lib.pgstrom_numeric_from_text(str(text).encode(), buf, len(buf))
    if n == 0:
        raise ValueError("bad numeric literal %r" % (text,))
    return buf.raw[:n]


class DataStore:
    """A chunk (pgstrom_data_store) in pinned host memory, KDS_FORMAT_COLUMN.

    columns: list, one entry per table column, of
       None                      -> column not loaded (not referenced)
       (values, nullmask|None)   -> numpy array (fixed-length types) or list
                                    of bytes|None (varlena), optional bool
                                    mask with True = NULL
    """

    def __init__(self, coltypes, columns, nrows=None):
        lib = _capi.load()
        self.lib = lib
        self.coltypes = list(coltypes)
        ncols = len(coltypes)
        assert len(columns) == ncols
        self.colmeta = make_colmeta(coltypes)
        vals = (C.c_void_p * ncols)()
        nuls = (C.c_void_p * ncols)()
        keep = []
        for c, col in enumerate(columns):
            if col is None:
                continue
            v, m = col
            attlen, _, _, dt = PGTYPES[coltypes[c]]
            if attlen > 0:
                a = np.ascontiguousarray(v, dtype=dt)
                n = a.shape[0]
                keep.append(a)
                vals[c] = a.ctypes.data
            else:
                n = len(v)
                ptrs = (C.c_void_p * n)()
                bufs = []
                for i, d in enumerate(v):
                    if d is None:
                        ptrs[i] = None
                    else:
                        b = C.create_string_buffer(bytes(d), len(d))
                        bufs.append(b)
                        ptrs[i] = C.addressof(b)
                keep.append((ptrs, bufs))
                vals[c] = C.addressof(ptrs)
            if nrows is None:
                nrows = n
            assert n == nrows, "column %d has %d rows, expected %d" % (c, n, nrows)
            if m is not None:
                mm = np.ascontiguousarray(m, dtype=np.uint8)
                assert mm.shape[0] == nrows
                keep.append(mm)
                nuls[c] = mm.ctypes.data
        self.nrows = int(nrows or 0)
        length = lib.pgstrom_kds_column_length(ncols, self.colmeta, self.nrows, vals, nuls)
        self.length = int(length)
        self.ptr = lib.pgs_chunk_alloc(self.length)
        if not self.ptr:
            raise MemoryError("pgs_chunk_alloc(%d)" % self.length)
        check(lib.pgstrom_kds_column_build(self.ptr, self.length, ncols, self.colmeta,
                                           self.nrows, vals, nuls))
        del keep

    def bytes(self):
        return C.string_at(self.ptr, self.length)

    def free(self):
        if self.ptr:
            self.lib.pgs_chunk_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


KDS_FORMAT_ROW, KDS_FORMAT_ROW_FLAT, KDS_FORMAT_TUPSLOT, KDS_FORMAT_COLUMN = 1, 2, 3, 4
BLCKSZ = 8192


class HeapDataStore:
    """A chunk in the reference's own input formats: KDS_FORMAT_ROW (heap
    pages referenced by kern_blkitem.page, one kern_rowitem per visible
    tuple - what pgstrom_data_store_insert_block builds, datastore.c:556-710)
    or KDS_FORMAT_ROW_FLAT (tuples packed from the tail, :799-823).

    `columns` as for DataStore; a column given as None is NULL in every row
    (PostgreSQL would store all columns, the device skips the ones the query
    does not reference either way).  `visible` optionally selects the rows a
    snapshot sees (bool array): invisible tuples stay on their page but get no
    kern_rowitem.  The synthetic pages stand in for shared buffers.
    """

    def __init__(self, coltypes, columns, nrows=None, flat=False, visible=None):
        lib = _capi.load()
        self.lib = lib
        self.coltypes = list(coltypes)
        ncols = len(coltypes)
        self.colmeta = make_colmeta(coltypes)
        lib.pgstrom_colmeta_set_cacheoff(ncols, self.colmeta)
        vals = (C.c_void_p * ncols)()
        nuls = (C.c_void_p * ncols)()
        keep = []
        blob = bytearray(b"\0" * 8)
        rowbytes = 24 + 8 + (ncols + 7) // 8
        for c, col in enumerate(columns):
            if col is None:
                continue
            v, m = col
            attlen, _, align, dt = PGTYPES[coltypes[c]]
            if attlen > 0:
                a = np.ascontiguousarray(v, dtype=dt)
                n = a.shape[0]
                keep.append(a)
                vals[c] = a.ctypes.data
                rowbytes += attlen + align
            else:
                n = len(v)
                offs = np.zeros(n, dtype=np.uint32)
                mm = np.zeros(n, dtype=np.uint8)
                maxlen = 0
                for i, d in enumerate(v):
                    if d is None:
                        mm[i] = 1
                        continue
                    while len(blob) % 4:
                        blob.append(0)
                    offs[i] = len(blob)
                    blob += bytes(d)
                    maxlen = max(maxlen, len(d))
                keep.append(offs)
                vals[c] = offs.ctypes.data
                rowbytes += maxlen + 4
                if m is None and mm.any():
                    m = mm
                elif m is not None:
                    m = np.asarray(m, dtype=np.uint8) | mm
            if nrows is None:
                nrows = n
            assert n == nrows
            if m is not None:
                mk = np.ascontiguousarray(m, dtype=np.uint8)
                keep.append(mk)
                nuls[c] = mk.ctypes.data
        self.nrows = nrows = int(nrows or 0)
        per_page = max(1, (BLCKSZ - 24) // (rowbytes + 4 + 8))
        maxpages = nrows // per_page + 2
        self._pages = lib.pgs_chunk_alloc(maxpages * BLCKSZ)
        rpp = np.zeros(maxpages, dtype=np.uint32)
        cblob = (C.c_char * len(blob)).from_buffer(blob)
        npages = lib.pgstrom_heap_form_pages(ncols, self.colmeta, nrows, vals, nuls,
                                             C.addressof(cblob), self._pages, maxpages,
                                             rpp.ctypes.data)
        if npages < 0:
            raise MemoryError("pgstrom_heap_form_pages: %d pages are not enough" % maxpages)
        del cblob
        self.npages = int(npages)
        vis = None if visible is None else np.asarray(visible, dtype=bool)
        self.format = KDS_FORMAT_ROW_FLAT if flat else KDS_FORMAT_ROW
        if not flat:
            # the reference sizes a chunk as BLCKSZ * maxblocks bytes for
            # head, items and pages together (datastore.c:604-613)
            maxblocks = self.npages + 2
            while True:
                meta = int(lib.pgstrom_kds_row_length(ncols, maxblocks, max(nrows, 1)))
                need = self.npages + (meta + BLCKSZ - 1) // BLCKSZ + 1
                if need <= maxblocks:
                    break
                maxblocks = need
            self.length = int(lib.pgstrom_kds_row_length(ncols, maxblocks, max(nrows, 1)))
            self.ptr = lib.pgs_chunk_alloc(self.length)
            check(lib.pgstrom_kds_row_init(self.ptr, self.length, ncols, self.colmeta,
                                           maxblocks, max(nrows, 1)))
            r0 = 0
            for p in range(self.npages):
                n = int(rpp[p])
                offs = np.arange(1, n + 1, dtype=np.uint16)
                if vis is not None:
                    offs = offs[vis[r0:r0 + n]]
                offs = np.ascontiguousarray(offs)
                got = lib.pgstrom_kds_row_insert_block(self.ptr, self._pages + p * BLCKSZ,
                                                       offs.ctypes.data, len(offs))
                if got < 0:
                    raise MemoryError("pgstrom_kds_row_insert_block: store is full")
                r0 += n
        else:
            self.length = int(lib.pgstrom_kds_row_length(ncols, 0, max(nrows, 1))) + \
                self.npages * BLCKSZ + 64
            self.ptr = lib.pgs_chunk_alloc(self.length)
            check(lib.pgstrom_kds_flat_init(self.ptr, self.length, ncols, self.colmeta,
                                            max(nrows, 1)))
            raw = C.string_at(self._pages, self.npages * BLCKSZ)
            r0 = 0
            for p in range(self.npages):
                n = int(rpp[p])
                for i in range(n):
                    if vis is not None and not vis[r0 + i]:
                        continue
                    lp, = struct.unpack_from("<I", raw, p * BLCKSZ + 24 + 4 * i)
                    off, ln = lp & 0x7fff, (lp >> 17) & 0x7fff
                    if not lib.pgstrom_kds_flat_insert_tuple(
                            self.ptr, self._pages + p * BLCKSZ + off, ln):
                        raise MemoryError("pgstrom_kds_flat_insert_tuple: store is full")
                r0 += n
        kds = kern_data_store.from_address(self.ptr)
        self.nitems = int(kds.nitems)
        del keep

    def device_image(self):
        """The chunk as it sits in device memory (what the CUDA layer's DMA
        assembles): bytes, for tests and device-resident runs."""
        kds = kern_data_store.from_address(self.ptr)
        if self.format == KDS_FORMAT_ROW_FLAT:
            return C.string_at(self.ptr, kds.length)
        ncols = len(self.coltypes)
        head = int(self.lib.pgstrom_kds_head_length(ncols))
        blk = (16 * kds.maxblocks + 15) & ~15
        items = (4 * kds.nitems + 15) & ~15
        first = (head + blk + items + BLCKSZ - 1) // BLCKSZ * BLCKSZ
        img = bytearray(first + self.npages * BLCKSZ)
        img[:head + blk + 4 * kds.nitems] = C.string_at(self.ptr, head + blk + 4 * kds.nitems)
        img[first:] = C.string_at(self._pages, self.npages * BLCKSZ)
        return bytes(img)

    def device_layout(self):
        """(bytes of head + items, offset of the first page, total bytes) of
        the device image."""
        kds = kern_data_store.from_address(self.ptr)
        if self.format == KDS_FORMAT_ROW_FLAT:
            return kds.length, kds.length, kds.length
        head = int(self.lib.pgstrom_kds_head_length(len(self.coltypes)))
        blk = (16 * kds.maxblocks + 15) & ~15
        items = (4 * kds.nitems + 15) & ~15
        first = (head + blk + items + BLCKSZ - 1) // BLCKSZ * BLCKSZ
        return head + blk + 4 * kds.nitems, first, first + self.npages * BLCKSZ

    def upload(self, device=0):
        """Assembles the device image in HBM; returns (pointer, bytes)."""
        headlen, first, total = self.device_layout()
        dptr = self.lib.pgs_device_alloc(device, total)
        if not dptr:
            raise MemoryError("pgs_device_alloc(%d)" % total)
        check(self.lib.pgs_device_upload(device, dptr, self.ptr, headlen))
        if self.format == KDS_FORMAT_ROW:
            check(self.lib.pgs_device_upload(device, dptr + first, self._pages,
                                             self.npages * BLCKSZ))
        return dptr, total

    def free(self):
        if getattr(self, "ptr", None):
            self.lib.pgs_chunk_free(self.ptr)
            self.ptr = None
        if getattr(self, "_pages", None):
            self.lib.pgs_chunk_free(self._pages)
            self._pages = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def decode_datum(datum, isnull, typ, typmod=-1, key_heap=(None, 0)):
    """8-byte Datum of a TUPSLOT row -> python value (by-value types; text /
    bpchar grouping keys come back as "kernel text" -> payload bytes; the
    long ones point into `key_heap` = (address, length) of the session)."""
    if isnull:
        return None
    if typ in ("text", "bpchar"):
        lib = _capi.load()
        heap, heap_len = key_heap
        size = max(64, typmod + 32)
        while True:
            buf = C.create_string_buffer(size)
            n = lib.pgstrom_fixup_kernel_text_heap(datum, typmod, heap, heap_len, buf, len(buf))
            if n == 0 and (datum >> 56) == 0x80 and heap and size < heap_len + 4 * max(typmod, 0) + 64:
                size = heap_len + 4 * max(typmod, 0) + 64   # a long key: any entry fits this
                continue
            break
        if n == 0:
            raise ValueError("bad kernel text datum %#x" % datum)
        return buf.raw[4:n]
    if typ == "bool":
        return bool(datum & 0xff)
    if typ == "int2":
        return struct.unpack("<h", struct.pack("<H", datum & 0xffff))[0]
    if typ in ("int4", "date"):
        return struct.unpack("<i", struct.pack("<I", datum & 0xffffffff))[0]
    if typ in ("int8", "time", "timestamp"):
        return struct.unpack("<q", struct.pack("<Q", datum))[0]
    if typ == "float4":
        return struct.unpack("<f", struct.pack("<I", datum & 0xffffffff))[0]
    if typ == "float8":
        return struct.unpack("<d", struct.pack("<Q", datum))[0]
    if typ == "numeric":
        # 64-bit device numeric -> Decimal via the text form numeric_in() takes
        from decimal import Decimal
        lib = _capi.load()
        buf = C.create_string_buffer(128)
        check(lib.pgstrom_fixup_kernel_numeric(datum, buf, len(buf)))
        return Decimal(buf.value.decode())
    raise NotImplementedError(typ)


class Plan:
    """Result of pgstrom_grafter_json() on a plan tree."""

    def __init__(self, plan_tree, gucs=None):
        lib = _capi.load()
        self.lib = lib
        for k, v in (gucs or {}).items():
            check(lib.pgstrom_guc_set(k.encode(), str(v).encode()))
        text = plan_tree if isinstance(plan_tree, str) else json.dumps(plan_tree)
        self.handle = lib.pgstrom_grafter_json(text.encode())
        if not self.handle:
            raise RuntimeError("pgstrom_grafter_json: " + lib.pgs_last_error().decode())

    @property
    def num_gpupreagg(self):
        return self.lib.pgs_plan_num_gpupreagg(self.handle)

    @property
    def reject_reason(self):
        return self.lib.pgs_plan_reject_reason(self.handle).decode()

    def explain(self, verbose=True):
        return self.lib.pgs_plan_explain(self.handle, 1 if verbose else 0).decode().split("\n")

    def tree(self):
        return json.loads(self.lib.pgs_plan_tree_json(self.handle).decode())

    def kernel_source(self, idx=0):
        return self.lib.pgs_plan_kernel_source(self.handle, idx).decode()

    def extra_flags(self, idx=0):
        return self.lib.pgs_plan_extra_flags(self.handle, idx)

    def describe(self, idx=0):
        return json.loads(self.lib.pgs_plan_describe_json(self.handle, idx).decode())

    def kparams(self, idx=0):
        n = C.c_size_t()
        p = self.lib.pgs_plan_kparams(self.handle, idx, C.byref(n))
        return C.string_at(p, n.value)

    def build_program(self, idx=0):
        """NVRTC build (needs no GPU); returns the opaque program handle."""
        prog = C.c_void_p()
        log = C.c_char_p()
        rc = self.lib.pgs_program_build(self.kernel_source(idx).encode(),
                                        self.extra_flags(idx), C.byref(prog), C.byref(log))
        if rc != 0:
            raise _capi.StromError(rc, (log.value or b"").decode(errors="replace")[:4000])
        return prog

    def free(self):
        if self.handle:
            self.lib.pgs_plan_free(self.handle)
            self.handle = None


_cuda_ready = False


def cuda_init(devices=None):
    """pgs_cuda_init(); raises if there is no usable CUDA device."""
    global _cuda_ready
    lib = _capi.load()
    if devices is None:
        check(lib.pgs_cuda_init(None, 0))
    else:
        arr = (C.c_int * len(devices))(*devices)
        check(lib.pgs_cuda_init(arr, len(devices)))
    _cuda_ready = True
    return lib.pgs_cuda_device_count()


class GpuPreAggState:
    """Executor half: gpupreagg_begin / gpupreagg_exec / gpupreagg_end over a
    child that hands over chunks (bulk-load protocol)."""

    def __init__(self, plan, chunks, idx=0, device=0):
        lib = _capi.load()
        self.lib = lib
        self.plan = plan
        self.desc = plan.describe(idx)
        self.coltypes = [c["type"] for c in self.desc["columns"]]
        self.typmods = [c.get("typmod", -1) for c in self.desc["columns"]]
        self._chunks = iter(chunks)
        self._held = []
        self.released = []          # addresses of the chunks handed back so far

        def release(_arg, kds):
            self.released.append(int(kds or 0))

        self._release = _capi.RELEASE_FN(release)

        def child_exec(_state, slot_p):
            slot = slot_p.contents
            try:
                item = next(self._chunks)
            except StopIteration:
                slot.kds = None
                return 0
            if isinstance(item, tuple):
                ds, rowmap = item
            else:
                ds, rowmap = item, None
            self._held.append(ds)
            slot.kds = ds.ptr
            slot.release = self._release
            slot.release_arg = None
            if rowmap is not None:
                rm = np.concatenate([np.array([len(rowmap)], dtype=np.int32),
                                     np.asarray(rowmap, dtype=np.int32)])
                self._held.append(rm)
                slot.krowmap = rm.ctypes.data
            else:
                slot.krowmap = None
            return 0

        self._cb = _capi.BULK_EXEC_FN(child_exec)
        st = C.c_void_p()
        check(lib.gpupreagg_begin(plan.handle, idx, device, self._cb, None, C.byref(st)))
        self.state = st

    def fetch_all(self):
        """Runs ExecCustomPlan until it is exhausted; returns partial rows as
        python tuples in GpuPreAgg target-list order."""
        ncols = len(self.coltypes)
        values = (C.c_uint64 * ncols)()
        isnull = C.create_string_buffer(ncols)
        rows = []
        while True:
            rc = self.lib.gpupreagg_exec(self.state, values, isnull)
            if rc == 0:
                break
            if rc < 0:
                raise _capi.StromError(-rc, self.lib.pgs_last_error().decode(errors="replace"))
            heap = self.key_heap()
            rows.append(tuple(decode_datum(values[i], isnull.raw[i] != 0, self.coltypes[i],
                                           self.typmods[i], heap)
                              for i in range(ncols)))
        return rows

    def key_heap(self):
        """(address, length) of the strings behind long text keys of the rows
        gpupreagg_exec returned."""
        heap, n = C.c_void_p(), C.c_size_t()
        check(self.lib.gpupreagg_key_heap(self.state, C.byref(heap), C.byref(n)))
        return heap.value, n.value

    def recheck_rows(self):
        n = self.lib.gpupreagg_recheck_rows(self.state, None, None, 0)
        if n <= 0:
            return []
        seq = (C.c_uint32 * n)()
        rows = (C.c_uint32 * n)()
        self.lib.gpupreagg_recheck_rows(self.state, seq, rows, n)
        return [(int(seq[i]), int(rows[i])) for i in range(n)]

    def recheck_chunk(self, seq):
        """Address of the retained chunk `seq` (0 = not held)."""
        return int(self.lib.gpupreagg_recheck_chunk(self.state, seq, None) or 0)

    def recheck_done(self, seq):
        check(self.lib.gpupreagg_recheck_done(self.state, seq))

    def explain(self, verbose=False, analyze=False):
        return self.lib.gpupreagg_explain(self.state, int(verbose), int(analyze)).decode()

    def rescan(self, chunks):
        self._chunks = iter(chunks)
        check(self.lib.gpupreagg_rescan(self.state))

    def end(self):
        notice = None
        if self.state:
            r = self.lib.gpupreagg_end(self.state)
            notice = r.decode() if r else None
            self.state = None
        self._held = []
        return notice


class Session:
    """Direct use of the session API (pgs_preagg_*) for benches: chunks may be
    device resident."""

    def __init__(self, plan, idx=0, device=0, num_groups=None, max_async_chunks=0,
                 max_chunk_rows=0, max_chunk_bytes=0):
        lib = _capi.load()
        self.lib = lib
        self.plan = plan
        self.desc = plan.describe(idx)
        self.coltypes = [c["type"] for c in self.desc["columns"]]
        self.typmods = [c.get("typmod", -1) for c in self.desc["columns"]]
        self.program = plan.build_program(idx)
        self._kparams = C.create_string_buffer(plan.kparams(idx))
        ncols = len(self.coltypes)
        self.result_colmeta = (kern_colmeta * ncols)()
        lib.pgs_plan_result_colmeta(plan.handle, idx, self.result_colmeta, ncols)
        cfg = pgs_session_config()
        cfg.device = device
        cfg.needs_grouping = 1 if self.desc["needs_grouping"] else 0
        cfg.num_groups = float(num_groups if num_groups is not None else self.desc["num_groups"])
        cfg.max_async_chunks = max_async_chunks
        cfg.max_chunk_rows = max_chunk_rows
        cfg.max_chunk_bytes = max_chunk_bytes
        cfg.result_ncols = ncols
        cfg.result_colmeta = self.result_colmeta
        self.device = device
        self._result_buf = None
        self._result_len = 0
        sess = C.c_void_p()
        check(lib.pgs_preagg_open(self.program, self._kparams, C.byref(cfg), C.byref(sess)))
        self.handle = sess

    def submit(self, ds, rowmap=None):
        t = C.c_int64()
        rm = None
        if rowmap is not None:
            rm = np.concatenate([np.array([len(rowmap)], dtype=np.int32),
                                 np.asarray(rowmap, dtype=np.int32)])
            self._rm = rm
        check(self.lib.pgs_preagg_submit(self.handle, ds.ptr,
                                         rm.ctypes.data if rm is not None else None,
                                         C.byref(t)))
        return t.value

    def submit_device_format(self, dptr, length, nitems, fmt):
        t = C.c_int64()
        check(self.lib.pgs_preagg_submit_device_format(self.handle, dptr, length, nitems, fmt,
                                                       None, C.byref(t)))
        return t.value

    def submit_device(self, dptr, length, nitems):
        t = C.c_int64()
        check(self.lib.pgs_preagg_submit_device(self.handle, dptr, length, nitems, None,
                                                C.byref(t)))
        return t.value

    def wait(self, ticket, timeout_ms=-1):
        st = C.c_int32()
        rc = self.lib.pgs_preagg_wait(self.handle, ticket, timeout_ms, C.byref(st))
        if rc == -1:
            return None
        check(rc)
        return st.value

    def recheck_rows(self, ticket):
        n = self.lib.pgs_preagg_recheck_rows(self.handle, ticket, None, 0)
        if n <= 0:
            return []
        rows = (C.c_uint32 * n)()
        self.lib.pgs_preagg_recheck_rows(self.handle, ticket, rows, n)
        return [int(r) for r in rows]

    def finish_raw(self, reset=True, nrooms=None):
        """Partial rows of everything submitted so far as a TUPSLOT
        kern_data_store in a pinned buffer the session keeps and reuses:
        returns (address, kern_data_store view)."""
        lib = self.lib
        ncols = len(self.coltypes)
        if nrooms is None:
            nrooms = int(max(16, self.desc["num_groups"] * 1.25 + 64)) \
                if self.desc["needs_grouping"] else 16
        for _ in range(4):
            length = lib.pgstrom_kds_tupslot_length(ncols, nrooms)
            if self._result_len < length:
                if self._result_buf:
                    lib.pgs_chunk_free(self._result_buf)
                self._result_buf = lib.pgs_chunk_alloc(length)
                if not self._result_buf:
                    self._result_len = 0
                    raise MemoryError("pgs_chunk_alloc(%d)" % length)
                self._result_len = length
            buf = C.c_void_p(self._result_buf)
            check(lib.pgstrom_kds_tupslot_init(buf, length, ncols, self.result_colmeta, nrooms))
            needed = C.c_uint32()
            status = C.c_int32()
            rc = lib.pgs_preagg_finish(self.handle, buf, 1 if reset else 0,
                                       C.byref(needed), C.byref(status))
            if rc == 301 and needed.value > nrooms:
                nrooms = needed.value + 16
                continue
            check(rc)
            break
        return buf, C.cast(buf, C.POINTER(kern_data_store)).contents

    def finish(self, reset=True, nrooms=None):
        """Partial rows of everything submitted so far, decoded."""
        buf, kds = self.finish_raw(reset, nrooms)
        return self.decode_rows(buf, kds)

    def decode_rows(self, buf, kds):
        """The rows of a TUPSLOT store finish_raw() returned, as python tuples."""
        lib = self.lib
        ncols = len(self.coltypes)
        values = (C.c_uint64 * ncols)()
        isnull = C.create_string_buffer(ncols)
        rows = []
        heap = self.key_heap()
        for r in range(kds.nitems):
            check(lib.pgstrom_fetch_data_store(buf, r, values, isnull))
            rows.append(tuple(decode_datum(values[i], isnull.raw[i] != 0, self.coltypes[i],
                                           self.typmods[i], heap)
                              for i in range(ncols)))
        return rows

    def key_heap(self):
        """(address, length) of the strings behind long text keys of the rows
        of the last finish (pgs_preagg_key_heap)."""
        heap, n = C.c_void_p(), C.c_size_t()
        check(self.lib.pgs_preagg_key_heap(self.handle, C.byref(heap), C.byref(n)))
        return heap.value, n.value

    def perfmon(self):
        return json.loads(self.lib.pgs_preagg_perfmon_json(self.handle).decode())

    def launch_count(self):
        return int(self.lib.pgs_preagg_launch_count(self.handle))

    def stream(self):
        return self.lib.pgs_preagg_stream(self.handle)

    def close(self):
        if self.handle:
            self.lib.pgs_preagg_close(self.handle)
            self.handle = None
        if getattr(self, "_result_buf", None):
            self.lib.pgs_chunk_free(self._result_buf)
            self._result_buf = None
            self._result_len = 0
        if self.program:
            self.lib.pgs_program_release(self.program)
            self.program = None
