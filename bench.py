#!/usr/bin/env python3
"""bench.py - GpuPreAgg throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ... # CPU Agg baseline
  torchrun --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU

One "step" = one pass of the hot path (fused qual + partial aggregation +
merge of the per-GPU states + flush of the partial rows) over one batch of
the synthetic table.  The headline of the JSON line is BASELINE.json
configs[1] (nogrp_agg: 100M rows int4/float8 per GPU); the same run also
measures the other throughput configurations and reports them, each with its
own value / roofline / e2e / check, under "workloads":

  where_agg          WHERE f < 10 + GROUP BY 1000 keys, 125M rows per GPU
  high_cardinality   GROUP BY ~10M distinct int8 keys, 100M rows per GPU
  nogrp_agg_heap     nogrp_agg over KDS_FORMAT_ROW heap-page chunks (the
                     reference's own input format, de-formed on the device)

(--workload NAME runs one of them alone as the headline.)  With N ranks every
rank scans its own shard (weak scaling) and the per-GPU states are merged over
NVLink inside the step.

`value` = rows/s with the chunks resident in HBM, timed with CUDA events on
the launching stream; `e2e` = the same metric through the C ABI with pinned
HOST chunks (H2D and the D2H of the partial rows inside the timed region);
`roofline` = algorithmic bytes of the scan kernel(s) / their mean launch time
(CUDA events, live) vs the measured HBM peak; `cpu_baseline` = the oracle's C
restatement of PostgreSQL's Agg-over-SeqScan timed on this box's host cores.
The two arms print the same `config`.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "GpuPreAgg (partial GROUP BY / no-group aggregation with fused qual) throughput"
# name -> table generator / plan, rows per GPU, rows per device-resident chunk
# (KDS_FORMAT_COLUMN has no 15 MB limit: a chunk is as large as its 32-bit
# length field allows), rows per host chunk of the end-to-end run (small
# enough that the H2D copy of one chunk overlaps the kernel of the previous
# one), input format
SPECS = {
    "nogrp_agg": dict(table="nogrp_agg", rows=100_000_000, chunk=100_000_000,
                      e2e_chunk=12_500_000, fmt="column", cpu_rows=100_000_000),
    "where_agg": dict(table="where_agg", rows=125_000_000, chunk=125_000_000,
                      e2e_chunk=12_500_000, fmt="column", cpu_rows=125_000_000),
    "high_cardinality": dict(table="high_cardinality", rows=100_000_000, chunk=100_000_000,
                             e2e_chunk=25_000_000, fmt="column", cpu_rows=12_500_000),
    # (a KDS_FORMAT_ROW chunk addresses at most 65536 pages: 16-bit blk_index)
    "nogrp_agg_heap": dict(table="nogrp_agg", rows=25_000_000, chunk=6_250_000,
                           e2e_chunk=3_125_000, fmt="row", cpu_rows=25_000_000),
}
HEADLINE = "nogrp_agg"
EXTRAS = ["where_agg", "high_cardinality", "nogrp_agg_heap"]
SQL = {
    "nogrp_agg": "SELECT count(*), count(x), sum(x), avg(x), min(x), max(x), sum(y), "
                 "avg(y), min(y), max(y) FROM bench_nogrp",
    "where_agg": "SELECT key, count(*), sum(w), avg(v), min(v), max(v) FROM bench_where "
                 "WHERE f < 10 GROUP BY key",
    "high_cardinality": "SELECT key, count(*), avg(v), avg(y), variance(y) FROM bench_hc "
                        "GROUP BY key",
}
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on",
        "pg_strom.perfmon": "on"}


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(SPECS),
                    help="run this workload alone (default: %s as the headline plus %s)"
                         % (HEADLINE, ", ".join(EXTRAS)))
    ap.add_argument("--no-extras", action="store_true",
                    help="headline workload only")
    ap.add_argument("--rows", type=int, default=0,
                    help="rows per GPU of the headline workload (the others scale along)")
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--e2e-chunk-rows", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--format", default=None, choices=["column", "row"],
                    help="input chunks: KDS_FORMAT_COLUMN, or the reference's heap-page "
                         "KDS_FORMAT_ROW (de-formed on the device)")
    ap.add_argument("--selectivity", type=int, default=10,
                    help="where_agg: per cent of rows the qual `f < N` keeps (10 is the "
                         "headline configuration; 1 and 50 are its reported variants)")
    ap.add_argument("--zipf", action="store_true",
                    help="high_cardinality: Zipf(1.0) keys instead of uniform ones")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    return ap.parse_args(argv)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The region of
    the headline workload lasts a few milliseconds, far less than one period of
    `nvidia-smi -lms`, so the samples are taken in-process through NVML by a
    thread that polls as fast as the calls return (~0.1 ms each) from start()
    to stop(); `nvidia-smi -lms 100` is the fallback when NVML is not there."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []           # (sm MHz, reasons bitmask)
        self.running = False
        self.thread = None

    def _nvml_open(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml = (pynvml, h)
        except Exception:           # noqa: BLE001  (no NVML: nvidia-smi below)
            self.nvml = None
        return self.nvml is not None

    def _nvml_poll(self):
        pynvml, h = self.nvml
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.running:
            try:
                self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                     int(reasons_fn(h))))
            except Exception:       # noqa: BLE001
                break
            time.sleep(0.0002)

    def start(self):
        if self._nvml_open():
            self.running = True
            self.thread = threading.Thread(target=self._nvml_poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._reader, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _reader(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def _stop_nvml(self):
        pynvml, h = self.nvml
        self.running = False
        self.thread.join(timeout=2)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                "hw_thermal_slowdown": 0x40}
        try:
            smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:           # noqa: BLE001
            smmax = None
        source = "nvml, in-process"
        if not self.samples:
            # the region ended before the first call returned: the clocks are
            # read now, microseconds after it (they have not dropped yet)
            try:
                reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
                self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                     int(reasons_fn(h))))
                source = "nvml, in-process, right after the timed region"
            except Exception:       # noqa: BLE001
                return {"sm_mhz": None, "sm_max_mhz": smmax, "reasons": ["no samples"]}
        reasons = sorted(n for n, b in bits.items() if any(r & b for _, r in self.samples))
        return {"sm_mhz": float(np.median([c for c, _ in self.samples])), "sm_max_mhz": smmax,
                "reasons": reasons, "samples": len(self.samples), "source": source}

    def stop(self):
        if self.nvml:
            return self._stop_nvml()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smmax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smmax)),
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------ workloads
def resolve_specs(args):
    """[(name, spec)] this run measures, headline first.  --rows scales every
    workload by the same factor (CPU tests run tiny tables)."""
    names = [args.workload] if args.workload else [HEADLINE] + ([] if args.no_extras else EXTRAS)
    head = SPECS[names[0]]
    scale = (args.rows / float(head["rows"])) if args.rows else 1.0
    out = []
    for i, name in enumerate(names):
        sp = dict(SPECS[name])
        if scale != 1.0:
            for k in ("rows", "chunk", "e2e_chunk", "cpu_rows"):
                sp[k] = max(4, int(sp[k] * scale) // 4 * 4)
        if i == 0:
            if args.chunk_rows:
                sp["chunk"] = args.chunk_rows
            if args.e2e_chunk_rows:
                sp["e2e_chunk"] = args.e2e_chunk_rows
            if args.format:
                sp["fmt"] = args.format
        sp["chunk"] = min(sp["chunk"], sp["rows"])
        sp["e2e_chunk"] = min(sp["e2e_chunk"], sp["rows"], sp["chunk"] if sp["fmt"] == "row" else sp["rows"])
        sp["selectivity"] = args.selectivity if sp["table"] == "where_agg" else None
        sp["zipf"] = bool(args.zipf and sp["table"] == "high_cardinality")
        out.append((name, sp))
    return out


def table_columns(sp, rank):
    """The rank's shard of the synthetic table as [(values, nullmask|None)]:
    counter based RNG, row i of the table does not depend on the chunking."""
    from pg_strom_b200 import workloads as W
    kw = {"zipf": True} if sp["zipf"] else {}
    base = rank * ((sp["rows"] + 3) // 4 * 4)
    return W.WORKLOADS[sp["table"]]["columns"](base, sp["rows"], **kw)


def slice_columns(cols, lo, hi):
    return [(v[lo:hi], None if m is None else m[lo:hi]) for v, m in cols]


def workload_config(name, sp):
    """What both arms print as `config` (identical keys and values)."""
    sql = SQL[sp["table"]]
    if sp["selectivity"] is not None:
        sql = sql.replace("f < 10", "f < %d" % sp["selectivity"])
    return {"workload": name, "sql": sql, "rows_per_gpu": sp["rows"],
            "input_format": "KDS_FORMAT_ROW (heap pages)" if sp["fmt"] == "row"
                            else "KDS_FORMAT_COLUMN",
            "key_distribution": "zipf(1.0)" if sp["zipf"] else "uniform",
            "l2_policy": "inputs are larger than L2 (126 MB); no flush between steps"}


_HEAP_PAGES = {}        # id(cols) -> (cols, pages, npages, rows): formed once per table


def run_cpu(sp, cols, nthreads):
    """PostgreSQL-style Agg over SeqScan in C on the host cores over a bounded
    sample (the first cpu_rows rows); returns (rows/s, rows used, seconds).
    The heap-page workload scans heap pages and de-forms every tuple
    (heapgettup_pagemode + slot_deform_tuple, oracle/cpu_agg.c); forming the
    pages is not part of the timed scan - they stand in for shared buffers."""
    from oracle import cpu_agg
    used = min(sp["cpu_rows"], len(cols[0][0]))
    if sp["fmt"] == "row":
        hit = _HEAP_PAGES.get(id(cols))
        if hit is None or hit[3] != used:
            pages, npages = cpu_agg.form_heap_pages(sp["table"], slice_columns(cols, 0, used))
            hit = _HEAP_PAGES[id(cols)] = (cols, pages, npages, used)
        dt, _, _, ng = cpu_agg.run_heap(sp["table"], hit[1], hit[2], nthreads=nthreads,
                                        max_groups=1 << 24, qual_const=sp["selectivity"])
        return used / dt, used, dt
    sample = slice_columns(cols, 0, used)
    dt, _, _, ng = cpu_agg.run(sp["table"], sample, nthreads=nthreads, max_groups=1 << 24,
                               qual_const=sp["selectivity"])
    return used / dt, used, dt


def cpu_sample_text(sp):
    return ("heap pages, every tuple de-formed" if sp["fmt"] == "row" else "columnar")


# ------------------------------------------------------------------ reference arm
def bench_reference(args):
    """--impl reference: the reference's CPU implementation of the path
    (PostgreSQL Agg over SeqScan; PostgreSQL is not installable here, so the
    oracle's C restatement stands in - kind 'port') on all host cores.  The
    CUDA library is neither built nor loaded by this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge
    ge.build_oracle()
    cores = host_cores()
    results = {}
    for name, sp in resolve_specs(args):
        cols = table_columns(sp, 0)
        times, used = [], 0
        for i in range(args.warmup + args.steps):
            rps, used, dt = run_cpu(sp, cols, cores)
            if i >= args.warmup:
                times.append(dt)
        del cols
        ms = 1000.0 * sum(times) / len(times)
        value = used / (ms / 1000.0)
        results[name] = {
            "value": value, "unit": "rows/s", "ms_per_step": ms,
            "config": workload_config(name, sp),
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": "first %d rows of the %s table per step, %s "
                                       "(oracle/cpu_agg.c: Agg over SeqScan, %d threads)"
                                       % (used, sp["table"], cpu_sample_text(sp), cores)},
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
        }
    names = list(results)
    head = results[names[0]]
    line = {
        "impl": "reference", "metric": METRIC,
        "value": head["value"], "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
        "config": head["config"], "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"],
        "gpu_launches": 0,
    }
    if len(names) > 1:
        line["workloads"] = {n: results[n] for n in names[1:]}
    print(json.dumps(line))


# ------------------------------------------------------------------ result checks
def result_arrays(sess, kds_addr, kds):
    """The TUPSLOT result store as numpy views: (values[nitems, ncols] uint64,
    isnull[nitems, ncols] uint8)."""
    ncols = len(sess.coltypes)
    head = int(sess.lib.pgstrom_kds_head_length(ncols))
    stride = (9 * ncols + 7) // 8 * 8
    n = int(kds.nitems)
    addr = kds_addr.value if hasattr(kds_addr, "value") else int(kds_addr)
    raw = (C.c_char * (head + stride * max(n, 1))).from_address(addr)
    vals = np.ndarray((n, ncols), dtype=np.uint64, buffer=raw, offset=head, strides=(stride, 8))
    nuls = np.ndarray((n, ncols), dtype=np.uint8, buffer=raw, offset=head + 8 * ncols,
                      strides=(stride, 1))
    return vals, nuls


def _mix64(u):
    from pg_strom_b200.workloads import _mix64 as m
    return m(u.astype(np.uint64))


FLOAT_GRID = float(1 << 20)     # the bench tables' float columns are multiples of 2^-10


def checksum_expected(node, desc, cols):
    """Size-independent check of a GROUP BY result with millions of groups:
    partial aggregation is linear, so for every PSUM column c
        sum over input rows   of  w(key(row)) * init_c(row)
      = sum over partial rows of  w(key)      * partial_c
    (mod 2^64, w = a 64-bit mix of the key), where init_c(row) is the per-row
    initial value of gpupreagg.c:1495-1748 (nrows: 0/1, psum: x, psum_x2:
    x*x).  The same sums without the weight pin the totals, and with it every
    (key, value) association.  Float columns sit on a 2^-10 grid: they are
    compared as exact integers (value * 2^20).  Returns {resno: (sum, weighted sum)}."""
    from oracle import bench_oracle
    n = len(cols[0][0])
    tlist = node["targetlist"]
    keep = np.ones(n, bool)
    for q in node.get("outer_quals") or []:
        v, m = bench_oracle.npeval(q, cols, n)
        keep &= (v.astype(bool) & ~m)
    w = np.zeros(n, np.uint64)
    for i, c in enumerate(desc["columns"]):
        if c["role"] == 1:
            v, m = bench_oracle.npeval(tlist[i]["expr"], cols, n)
            assert not m.any(), "NULL keys are not part of the bench tables"
            with np.errstate(over="ignore"):
                w = _mix64(w ^ v.astype(np.int64).view(np.uint64))
    out = {}
    with np.errstate(over="ignore"):
        for i, c in enumerate(desc["columns"]):
            if c["role"] != 2:
                continue
            assert c["op"] == "PSUM", "checksum check covers sums only (%s)" % c["text"]
            e = tlist[i]["expr"]
            f = e["funcname"]
            if f == "nrows":
                ok = keep.copy()
                for a in e.get("args", []):
                    v, m = bench_oracle.npeval(a, cols, n)
                    ok &= (v.astype(bool) & ~m)
                val = ok.astype(np.uint64)
            else:
                v, m = bench_oracle.npeval(e["args"][0], cols, n)
                ok = keep & ~m
                if np.issubdtype(v.dtype, np.floating):
                    x = v.astype(np.float64)
                    if f == "psum_x2":
                        x = x * x
                    x = x * FLOAT_GRID
                    assert np.array_equal(x, np.rint(x)), "float column off the dyadic grid"
                    val = x.astype(np.int64).view(np.uint64)
                else:
                    assert f == "psum", f
                    val = v.astype(np.int64).view(np.uint64)
                val = np.where(ok, val, np.uint64(0))
            out[c["resno"]] = (int(val.sum(dtype=np.uint64)), int((val * w).sum(dtype=np.uint64)))
    return out, int(keep.sum())


def checksum_device(desc, vals, nuls):
    """The same sums over the partial rows the device returned."""
    w = np.zeros(len(vals), np.uint64)
    out = {}
    with np.errstate(over="ignore"):
        for i, c in enumerate(desc["columns"]):
            if c["role"] == 1:
                k = np.ascontiguousarray(vals[:, i])
                if c["type"] == "int4":
                    k = k.astype(np.uint32).view(np.int32).astype(np.int64).view(np.uint64)
                w = _mix64(w ^ k)
        for i, c in enumerate(desc["columns"]):
            if c["role"] != 2:
                continue
            col = np.ascontiguousarray(vals[:, i])
            if c["type"] == "float8":
                x = col.view(np.float64) * FLOAT_GRID
                assert np.array_equal(x, np.rint(x)), "device float sum off the dyadic grid"
                val = x.astype(np.int64).view(np.uint64)
            elif c["type"] == "int4":
                val = col.astype(np.uint32).view(np.int32).astype(np.int64).view(np.uint64)
            else:
                val = col
            val = np.where(np.ascontiguousarray(nuls[:, i]) != 0, np.uint64(0), val)
            out[c["resno"]] = (int(val.sum(dtype=np.uint64)), int((val * w).sum(dtype=np.uint64)))
    return out


# ------------------------------------------------------------------ our arm
class Context:
    pass


def open_context():
    import torch
    from pg_strom_b200 import _capi
    from pg_strom_b200 import gpupreagg as gp
    ctx = Context()
    ctx.torch = torch
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ctx.dist = None
    if ctx.world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(ctx.local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local_rank))
        ctx.dist = dist
    torch.cuda.set_device(ctx.local_rank)
    ctx.dev = torch.device("cuda", ctx.local_rank)
    ctx.lib = _capi.load()
    gp.cuda_init([ctx.local_rank])
    ctx.comm = C.c_void_p()
    if ctx.world > 1:
        from pg_strom_b200 import multigpu
        ctx.comm = multigpu.nccl_communicator(ctx.lib, ctx.dist, 0, ctx.rank, ctx.world,
                                              device=ctx.dev)
    return ctx


def barrier(ctx):
    if ctx.world > 1:
        ctx.dist.barrier()
    ctx.torch.cuda.synchronize()


def max_over_ranks(ctx, x):
    t = ctx.torch.tensor([float(x)], device=ctx.dev, dtype=ctx.torch.float64)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks_u64(ctx, xs):
    """Wrapping sum of unsigned 64-bit integers over the ranks."""
    a = np.array(xs, dtype=np.uint64).view(np.int64)
    t = ctx.torch.from_numpy(a.copy()).to(ctx.dev)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.SUM)
    return [int(v) for v in t.cpu().numpy().view(np.uint64)]


def run_workload(ctx, name, sp, args, steps, warmup, want_cpu, headline):
    """Everything for one workload: tables, parity check, the device-resident
    timed region, the end-to-end timed region, the CPU baseline.  Returns the
    dict that becomes the JSON line (headline) or an entry of "workloads"."""
    from pg_strom_b200 import _capi
    from pg_strom_b200 import gpupreagg as gp
    from pg_strom_b200 import multigpu
    from pg_strom_b200 import workloads as W
    from oracle import bench_oracle             # the checker of the pre-timing pass only
    torch, lib, world, rank = ctx.torch, ctx.lib, ctx.world, ctx.rank
    w = W.WORKLOADS[sp["table"]]
    plan_kw = {}
    if sp["selectivity"] is not None and sp["selectivity"] != 10:
        plan_kw["selectivity_pct"] = sp["selectivity"]
    plan = gp.Plan(w["plan"](**plan_kw), gucs=GUCS)
    assert plan.num_gpupreagg == 1, plan.reject_reason
    desc = plan.describe()
    node = plan.tree()["lefttree"]
    rows, chunk_rows, e2e_chunk_rows = sp["rows"], sp["chunk"], sp["e2e_chunk"]
    heap = sp["fmt"] == "row"
    fmt = gp.KDS_FORMAT_ROW if heap else gp.KDS_FORMAT_COLUMN
    coltypes = [t for _, t in w["table"].columns]

    # ---- synthetic table: resident device chunks + pinned host chunks ----
    cols = table_columns(sp, rank)
    dev_chunks, host_chunks = [], []
    for lo in range(0, rows, chunk_rows):
        part = slice_columns(cols, lo, min(rows, lo + chunk_rows))
        if heap:
            ds = gp.HeapDataStore(coltypes, part, nrows=len(part[0][0]))
            dptr, length = ds.upload(0)
        else:
            ds = gp.DataStore(coltypes, part, nrows=len(part[0][0]))
            length = ds.length
            dptr = lib.pgs_device_alloc(0, length)
            assert dptr, lib.pgs_last_error()
            _capi.check(lib.pgs_device_upload(0, dptr, ds.ptr, length))
        dev_chunks.append((dptr, length, ds.nrows))
        ds.free()
    for lo in range(0, rows, e2e_chunk_rows):
        part = slice_columns(cols, lo, min(rows, lo + e2e_chunk_rows))
        if heap:
            ds = gp.HeapDataStore(coltypes, part, nrows=len(part[0][0]))
            ds.length = ds.device_layout()[2]
        else:
            ds = gp.DataStore(coltypes, part, nrows=len(part[0][0]))
        host_chunks.append(ds)
    total_bytes = sum(d.length for d in host_chunks)
    total_dev_bytes = sum(length for _, length, _ in dev_chunks)
    nullable = sum(1 for c in cols if c[1] is not None)
    alg_bytes_per_row = desc["row_bytes"] + nullable / 8.0

    sess = gp.Session(plan, max_async_chunks=3, max_chunk_rows=max(chunk_rows, e2e_chunk_rows),
                      max_chunk_bytes=max(d.length for d in host_chunks))
    merge = multigpu.StateMerge(lib, sess, ctx) if world > 1 else None
    stream = torch.cuda.ExternalStream(sess.stream(), device=ctx.dev)

    def step(chunks, resident):
        for ch in chunks:
            if resident:
                sess.submit_device_format(ch[0], ch[1], ch[2], fmt)
            else:
                sess.submit(ch)
        if merge:
            merge.run()
        return sess.finish_raw()

    # ---- warm-up + one correctness check against the oracle ----
    nresult = 0
    for i in range(max(warmup, 3)):
        nresult = int(step(dev_chunks, True)[1].nitems)
    nresult = int(sum_over_ranks_u64(ctx, [nresult])[0])
    checked = None
    if not args.no_check:
        many = bool(desc["needs_grouping"]) and desc["num_groups"] > 100_000
        buf, kds = step(dev_chunks, True)
        if many:
            # millions of groups: linearity checksums (see checksum_expected),
            # summed over the ranks - every rank holds a disjoint set of groups
            # after the partitioned merge, or the root holds all of them
            exp, npass = checksum_expected(node, desc, cols)
            vals, nuls = result_arrays(sess, buf, kds)
            got = checksum_device(desc, vals, nuls)
            resnos = sorted(exp)
            e_all = sum_over_ranks_u64(ctx, [x for r in resnos for x in exp[r]])
            g_all = sum_over_ranks_u64(ctx, [x for r in resnos for x in got[r]])
            if rank == 0:
                assert e_all == g_all, "checksum mismatch: expected %r, device %r" % (e_all, g_all)
            checked = "linearity checksums of every partial-sum column over %d partial rows " \
                      "(sum and key-weighted sum mod 2^64, exact)" % nresult
        else:
            # every rank restates its shard with the numpy oracle, the root
            # merges the expectations and compares them with the rows the
            # device path (merge included) returns - bit-exact
            part = bench_oracle.expected_partial_node(node, cols)
            if world > 1:
                parts = [None] * world if rank == 0 else None
                ctx.dist.gather_object(part, parts, dst=0)
            else:
                parts = [part]
            lib_rows = sess.decode_rows(buf, kds)
            if rank == 0:
                keys, exp = bench_oracle.merge_expected(desc, parts)
                bench_oracle.assert_rows_equal_expected(desc, lib_rows, keys, exp)
            checked = "every partial row against the numpy oracle, bit-exact"

    # ---- timed region: device resident ----
    pm0 = sess.perfmon()
    l0 = sess.launch_count()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    barrier(ctx)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(steps):
        step(dev_chunks, True)
    ev1.record(stream)
    barrier(ctx)
    ms_per_step = max_over_ranks(ctx, ev0.elapsed_time(ev1)) / steps
    clocks = sampler.stop() if rank == 0 else None
    pm1 = sess.perfmon()
    launches = sess.launch_count() - l0
    value = rows * world / (ms_per_step / 1000.0)
    n_k = pm1["num_kern_main"] - pm0["num_kern_main"]
    k_ms = (pm1["time_kern_main_ms"] - pm0["time_kern_main_ms"]) / max(n_k, 1)
    k_rows = (pm1["rows_kern_main"] - pm0["rows_kern_main"]) / max(n_k, 1)
    peak, peak_src = measured_peaks()
    achieved = k_rows * alg_bytes_per_row / (k_ms / 1000.0) / 1e9 if k_ms > 0 else 0.0

    # ---- where a resident step spends its time (host clock, synchronised
    # between the phases; diagnostic only, not part of `value`) ----
    phase_ms = {"scan": 0.0, "merge": 0.0, "finish": 0.0}
    for i in range(3):
        barrier(ctx)
        t0 = time.perf_counter()
        for dptr, length, n in dev_chunks:
            sess.submit_device_format(dptr, length, n, fmt)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if merge:
            merge.run()
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        sess.finish_raw()
        t3 = time.perf_counter()
        phase_ms["scan"] += (t1 - t0) * 1000.0 / 3
        phase_ms["merge"] += (t2 - t1) * 1000.0 / 3
        phase_ms["finish"] += (t3 - t2) * 1000.0 / 3

    # ---- timed region: end to end (host chunks in pinned memory) ----
    e2e_steps = args.e2e_steps or max(3, min(steps, 5))
    for i in range(2):
        step(host_chunks, False)
    pm2 = sess.perfmon()
    barrier(ctx)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step(host_chunks, False)
    barrier(ctx)
    e2e_ms = max_over_ranks(ctx, (time.perf_counter() - t0) * 1000.0 / e2e_steps)
    pm3 = sess.perfmon()
    e2e_value = rows * world / (e2e_ms / 1000.0)
    h2d = (pm3["bytes_dma_send"] - pm2["bytes_dma_send"]) // e2e_steps
    d2h = (pm3["bytes_dma_recv"] - pm2["bytes_dma_recv"]) // e2e_steps
    merge_trace = merge.trace() if merge else None

    # ---- the platform's own ceiling for that copy: the same pinned chunks,
    # plain cudaMemcpy into the resident buffers, all ranks at once ----
    plain_gbs = None
    if headline and not heap:       # (a heap chunk's pages do not sit behind its head)
        scratch = lib.pgs_device_alloc(0, max(d.length for d in host_chunks))
        assert scratch, lib.pgs_last_error()
        nbytes = 0
        barrier(ctx)
        t0 = time.perf_counter()
        for rep in range(2):
            for ds in host_chunks:
                _capi.check(lib.pgs_device_upload(0, scratch, ds.ptr, ds.length))
                nbytes += ds.length
        torch.cuda.synchronize()
        barrier(ctx)
        plain_gbs = nbytes / (max_over_ranks(ctx, time.perf_counter() - t0)) / 1e9
        lib.pgs_device_free(0, scratch)

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        cores = host_cores()
        rps1, used, dt1 = run_cpu(sp, cols, 1)
        rpsn, used, dtn = run_cpu(sp, cols, cores)
        cpu = {"value": rpsn, "unit": "rows/s", "cores": cores, "kind": "port",
               "value_1core": rps1,
               "sample": "first %d rows of the %s table, %s (oracle/cpu_agg.c: "
                         "PostgreSQL-style Agg over SeqScan; %.2fs on 1 core, %.2fs on %d cores)"
                         % (used, sp["table"], cpu_sample_text(sp), dt1, dtn, cores)}

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            t = json.load(f).get(name)
        if t and t.get("rows_per_launch") and k_rows:
            # ncu dram__bytes_read+write of one profiled launch, scaled to
            # the rows one timed launch processed
            traffic = t["dram_bytes_per_launch"] * (k_rows / t["rows_per_launch"])
    kernel = "gpupreagg_heap_index + gpupreagg_main_heap_staged" if heap else (
        "gpupreagg_main + gpupreagg_partagg" if pm1.get("part_nparts") else "gpupreagg_main")
    config = workload_config(name, sp)
    result = {
        "value": value, "unit": "rows/s", "ms_per_step": ms_per_step,
        "config": config,
        "details": {"chunk_rows": chunk_rows, "chunks_per_step": len(dev_chunks),
                    "e2e_chunk_rows": e2e_chunk_rows, "input_bytes_per_gpu": total_dev_bytes,
                    "physical_bytes_per_row": total_dev_bytes / float(rows),
                    "algorithmic_bytes_per_row": alg_bytes_per_row,
                    "e2e_input_bytes_per_step": total_bytes,
                    "merge": merge.describe() if merge else "none",
                    "partial_rows": nresult,
                    "group_updates_per_s": value if desc["needs_grouping"] else None},
        "checked": checked,
        "gb_per_s": value * alg_bytes_per_row / 1e9,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "peak_source": peak_src, "kernel": kernel,
                     "launch_ms": k_ms, "bytes_per_launch": k_rows * alg_bytes_per_row,
                     "launches_timed": n_k,
                     # bytes the kernel has to read as the chunks are laid out
                     # (heap pages: tuple headers, line pointers, padding)
                     "achieved_physical": (k_rows * (total_dev_bytes / float(rows)) /
                                           (k_ms / 1000.0) / 1e9) if k_ms > 0 else 0.0},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "rows/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "h2d_gb_per_s": h2d / (e2e_ms / 1000.0) / 1e9,
                "pcie_gen5_x16_frac": h2d / (e2e_ms / 1000.0) / 1e9 / 64.0,
                "plain_copy_gb_per_s": plain_gbs,
                "plain_copy_frac": (h2d / (e2e_ms / 1000.0) / 1e9 / plain_gbs) if plain_gbs else None},
        "gpu_launches": int(launches),
        "phase_ms": phase_ms,
        "merge_trace": merge_trace,
        "clocks": clocks,
    }
    if merge:
        merge.close()
    sess.close()
    for dptr, _, _ in dev_chunks:
        lib.pgs_device_free(0, dptr)
    for ds in host_chunks:
        ds.free()
    plan.free()
    return result


def bench_ours(args):
    import __graft_entry__ as ge
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        ge.build()
    ctx = open_context()
    if ctx.world > 1:
        ctx.dist.barrier()          # the other ranks wait for rank 0's build
    specs = resolve_specs(args)
    results = {}
    for i, (name, sp) in enumerate(specs):
        steps = args.steps if i == 0 else max(3, min(args.steps, 10))
        try:
            results[name] = run_workload(ctx, name, sp, args, steps, args.warmup,
                                         not args.no_cpu_baseline, i == 0)
        except Exception as e:          # noqa: BLE001
            # the headline must run; a failing extra workload is reported as
            # such in the line instead of taking the other numbers with it
            if i == 0 or ctx.world > 1:
                raise
            import traceback
            traceback.print_exc()
            results[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300]),
                             "gpu_launches": 0}
    if ctx.rank == 0:
        names = list(results)
        head = results[names[0]]
        line = {
            "metric": METRIC, "value": head["value"], "unit": "rows/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
        }
        for k in ("config", "details", "checked", "gb_per_s", "roofline", "cpu_baseline", "e2e",
                  "phase_ms", "merge_trace", "clocks"):
            line[k] = head[k]
        line["gpu_launches"] = sum(r["gpu_launches"] for r in results.values())
        if len(names) > 1:
            line["workloads"] = {n: results[n] for n in names[1:]}
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.lib.pgs_nccl_comm_destroy(ctx.comm)
        ctx.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
