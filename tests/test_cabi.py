"""-m "not gpu": the C-ABI library loads, exports every symbol the header
declares, and the host-only parts (GUCs, layouts, builders, planner, NVRTC
build for sm_100a) work without a GPU."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from pg_strom_b200 import _capi
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P
from pg_strom_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "pgstrom_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b((?:pgstrom|pgs|gpupreagg)_\w+)\s*\(", hdr))
    names -= {"pgs_bulk_exec_fn"}
    assert len(names) > 50
    for n in sorted(names):
        assert hasattr(lib, n), "symbol %s is declared but not exported" % n
        assert n in _capi.PROTOTYPES, "no ctypes prototype for %s" % n


def test_layout_offsets(lib):
    # opencl_common.h:375-389 / SURVEY.md section 8 a1
    k = _capi.kern_data_store
    assert (k.hostptr.offset, k.length.offset, k.usage.offset, k.ncols.offset,
            k.nitems.offset, k.nrooms.offset, k.nblocks.offset, k.maxblocks.offset,
            k.format.offset, k.tdhasoid.offset, k.tdtypeid.offset,
            k.tdtypmod.offset) == (0, 8, 12, 16, 20, 24, 28, 32, 36, 37, 40, 44)
    assert C.sizeof(k) == 48 and C.sizeof(_capi.kern_colmeta) == 8
    assert lib.pgstrom_kds_head_length(11) == 48 + 88 + 8     # STROMALIGN(136)
    # TUPSLOT stride LONGALIGN(9 * ncols)
    assert lib.pgstrom_kds_tupslot_length(13, 10) == ((48 + 104 + 15) // 16 * 16) + 120 * 10


def test_guc_table(lib):
    lib.pgstrom_guc_reset_all()
    gucs = {g["name"]: g for g in json.loads(lib.pgstrom_guc_list_json())}
    for name, boot in (("pg_strom.enabled", "on"), ("pg_strom.enabled_global", "on"),
                       ("pg_strom.perfmon", "off"), ("pg_strom.show_device_kernel", "off"),
                       ("pg_strom.chunk_size", "15"), ("pg_strom.min_async_chunks", "2"),
                       ("pg_strom.max_async_chunks", "3"), ("enable_gpupreagg", "on"),
                       ("pg_strom.debug_force_gpupreagg", "off"), ("enable_gpuscan", "on"),
                       ("pg_strom.mqueue_timeout", "60000"),
                       ("pg_strom.shmem_totalsize", "2048"),
                       ("pg_strom.opencl_devices", "any")):
        assert gucs[name]["boot"] == boot
    assert lib.pgstrom_guc_set(b"pg_strom.chunk_size", b"3") != 0      # range 4..128
    assert lib.pgstrom_guc_set(b"pg_strom.chunk_size", b"64") == 0
    assert lib.pgstrom_guc_get(b"pg_strom.chunk_size") == b"64"
    assert lib.pgstrom_guc_set(b"pg_strom.no_such", b"1") != 0
    lib.pgstrom_guc_reset_all()
    assert lib.pgstrom_strerror(2) == b"To be re-checked by CPU"
    assert lib.pgstrom_strerror(301) == b"data store has no space"


def test_enabled_off_leaves_plan_alone(lib):
    lib.pgstrom_guc_reset_all()
    plan = gp.Plan(W.nogrp_plan(), gucs={"pg_strom.enabled": "off"})
    assert plan.num_gpupreagg == 0
    assert plan.explain()[0] == "Aggregate"
    plan.free()
    lib.pgstrom_guc_reset_all()


def test_column_store_builder(lib):
    x = np.arange(1000, dtype=np.int32)
    y = np.arange(1000, dtype=np.float64) / 8
    mask = (x % 7 == 0).astype(np.uint8)
    ds = gp.DataStore(["int4", "float8", "int8"], [(x, mask), (y, None), None])
    raw = ds.bytes()
    kds = _capi.kern_data_store.from_buffer_copy(raw[:48])
    assert (kds.ncols, kds.nitems, kds.format, kds.length) == (3, 1000, 4, len(raw))
    head = lib.pgstrom_kds_head_length(3)
    pos = np.frombuffer(raw[head:head + 24], dtype=np.uint32).reshape(3, 2)
    assert pos[0, 0] % 128 == 0 and pos[0, 1] % 128 == 0 and pos[0, 1] != 0
    assert pos[1, 1] == 0 and tuple(pos[2]) == (0, 0)       # no NULLs / not loaded
    assert np.array_equal(np.frombuffer(raw, np.int32, 1000, pos[0, 0]), x)
    assert np.array_equal(np.frombuffer(raw, np.float64, 1000, pos[1, 0]), y)
    bits = np.unpackbits(np.frombuffer(raw, np.uint8, 125, pos[0, 1]), bitorder="little")
    assert np.array_equal(bits[:1000], 1 - mask)             # bit set = NOT NULL
    ds.free()


def test_numeric_varlena_roundtrip(lib):
    for text in ("0", "1", "-1", "123.450", "0.00012345678901234", "-60.67617833614347448932",
                 "1000000000000000000000000000000000000000000000000", "0." + "0" * 32 + "1",
                 "99999999.9999", "10000", "0.0001", "-0.5"):
        d = gp.numeric_datum(text)
        buf = C.create_string_buffer(4096)
        n = lib.pgstrom_numeric_to_text(d, buf, len(buf))
        assert n > 0 and buf.value.decode() == text, (text, buf.value)


def test_programs_build_for_sm_100a(lib):
    for name, w in W.WORKLOADS.items():
        plan = gp.Plan(w["plan"](), gucs={"pg_strom.enabled": "on"})
        assert plan.num_gpupreagg == 1, (name, plan.reject_reason)
        src = plan.kernel_source()
        assert "gpupreagg_qual_eval" in src and "gpupreagg_projection" in src
        prog = plan.build_program()
        n = C.c_size_t()
        p = lib.pgs_program_cubin(prog, C.byref(n))
        cubin = C.string_at(p, n.value)
        assert cubin[:4] == b"\x7fELF" and n.value > 10000
        lib.pgs_program_release(prog)
        plan.free()


def test_gather_payload_variant_builds(lib, monkeypatch):
    """GROUP BY under a WHERE clause stages only the qual's columns
    (PGSTROM_GATHER_PAYLOAD=0 switches back to staging every column); every
    other program is the same either way."""
    gucs = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}
    monkeypatch.setenv("PGSTROM_GATHER_PAYLOAD", "0")
    plan = gp.Plan(W.where_plan(), gucs=gucs)
    base = plan.kernel_source()
    plan.free()
    assert "#define GPUPREAGG_GATHER_PAYLOAD 0" in base
    monkeypatch.delenv("PGSTROM_GATHER_PAYLOAD")
    plan = gp.Plan(W.where_plan(), gucs=gucs)
    src = plan.kernel_source()
    assert "#define GPUPREAGG_GATHER_PAYLOAD 1" in src
    # f (slot 0) is the qual's column; key, v, w are fetched by row number
    i = src.index("GPUPREAGG_INCOL_STAGED(int slot)")
    body = src[i:src.index("}\n}", i)]
    assert "case 0" not in body and all("case %d: return 0;" % k in body for k in (1, 2, 3))
    prog = plan.build_program()
    lib.pgs_program_release(prog)
    plan.free()
    for mk in (W.nogrp_plan, W.hc_plan):        # no WHERE / partitioned: unchanged
        plan = gp.Plan(mk(), gucs=gucs)
        assert "#define GPUPREAGG_GATHER_PAYLOAD 0" in plan.kernel_source()
        plan.free()


def test_build_failure_reports_log(lib):
    prog = C.c_void_p()
    log = C.c_char_p()
    rc = lib.pgs_program_build(b'#include "pgstrom_kds.h"\nthis is not CUDA;\n', 0,
                               C.byref(prog), C.byref(log))
    assert rc == -11
    assert b"error" in log.value and b"this is not CUDA" in log.value


def test_device_calls_fail_loudly_without_gpu(lib):
    """No CPU fallback: on a box without a CUDA device the device layer
    refuses to start and sessions cannot be opened."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.pgs_cuda_init(None, 0) == 100        # StromError_ServerNotReady
    plan = gp.Plan(W.nogrp_plan(), gucs={"pg_strom.enabled": "on"})
    with pytest.raises(_capi.StromError):
        gp.Session(plan)
    plan.free()


def test_key_heap_calls_check_their_arguments(lib):
    """pgs_preagg_key_heap / gpupreagg_key_heap without a session, and the
    host fix-up of a key-heap word without a heap: refused, nothing is read."""
    import ctypes as C
    heap, n = C.c_void_p(), C.c_size_t()
    assert lib.pgs_preagg_key_heap(None, C.byref(heap), C.byref(n)) == 101  # StromError_BadRequestMessage
    assert lib.gpupreagg_key_heap(None, C.byref(heap), C.byref(n)) == 101
    buf = C.create_string_buffer(64)
    word = (0x80 << 56) | 16
    assert lib.pgstrom_fixup_kernel_text_heap(word, -1, None, 0, buf, len(buf)) == 0
    assert lib.pgstrom_fixup_kernel_text(word, -1, buf, len(buf)) == 0
