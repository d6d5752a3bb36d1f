#!/usr/bin/env python3
"""bench.py - GpuPreAgg throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ... # CPU Agg baseline
  torchrun --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU

One "step" = one pass of the hot path (fused qual + partial aggregation +
flush of the partial rows) over one batch of the synthetic table; at N=1 the
table is BASELINE.json configs[1] (nogrp_agg: 100M rows int4/float8).  With
N ranks every rank scans its own 100M-row shard (weak scaling) and the
per-GPU states are merged over NCCL into rank 0 inside the step.

Prints ONE JSON line (rank 0).  `value` = rows/s with the chunks resident in
HBM, timed with CUDA events on the launching stream; `e2e` = the same metric
through the C ABI with pinned HOST chunks (H2D and the D2H of the partial
rows inside the timed region); `roofline` = algorithmic bytes of the main
kernel / its mean launch time (CUDA events, live) vs the measured HBM peak;
`cpu_baseline` = the oracle's C restatement of PostgreSQL's Agg-over-SeqScan
timed on this box's host cores.
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

DEFAULT_ROWS = {"nogrp_agg": 100_000_000, "where_agg": 125_000_000,
                "high_cardinality": 100_000_000}
# device-resident chunks: the column format (KDS_FORMAT_COLUMN) has no 15 MB
# limit, a chunk is as large as its 32-bit length field allows
DEFAULT_CHUNK = {"nogrp_agg": 100_000_000, "where_agg": 125_000_000,
                 "high_cardinality": 50_000_000}
# host chunks of the end-to-end run: small enough that the H2D copy of one
# chunk overlaps the kernel of the previous one
DEFAULT_E2E_CHUNK = {"nogrp_agg": 12_500_000, "where_agg": 12_500_000,
                     "high_cardinality": 12_500_000}
METRIC = "GpuPreAgg (partial GROUP BY / no-group aggregation with fused qual) throughput"
SQL = {
    "nogrp_agg": "SELECT count(*), count(x), sum(x), avg(x), min(x), max(x), sum(y), "
                 "avg(y), min(y), max(y) FROM bench_nogrp",
    "where_agg": "SELECT key, count(*), sum(w), avg(v), min(v), max(v) FROM bench_where "
                 "WHERE f < 10 GROUP BY key",
    "high_cardinality": "SELECT key, count(*), avg(v), avg(y), variance(y) FROM bench_hc "
                        "GROUP BY key",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nogrp_agg", choices=sorted(DEFAULT_ROWS))
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU")
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--e2e-chunk-rows", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--format", default="column", choices=["column", "row"],
                    help="input chunks: KDS_FORMAT_COLUMN, or the reference's heap-page "
                         "KDS_FORMAT_ROW (de-formed on the device)")
    ap.add_argument("--selectivity", type=int, default=10,
                    help="where_agg: per cent of rows the qual `f < N` keeps (10 is the "
                         "headline configuration; 1 and 50 are its reported variants)")
    ap.add_argument("--zipf", action="store_true",
                    help="high_cardinality: Zipf(1.0) keys instead of uniform ones")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
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

    def stop(self):
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
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def generate_columns(workload, rank, rows, chunk_rows, **col_kw):
    from pg_strom_b200 import workloads as W
    out = []
    base = rank * ((rows + 3) // 4 * 4)
    for r0 in range(0, rows, chunk_rows):
        n = min(chunk_rows, rows - r0)
        out.append(W.WORKLOADS[workload]["columns"](base + r0, n, **col_kw))
    return out


def column_options(args):
    return {"zipf": True} if (args.zipf and args.workload == "high_cardinality") else {}


def run_cpu(workload, colchunks, nthreads, max_seconds=30.0, qual_const=None):
    """PostgreSQL-style Agg over SeqScan in C on the host cores; returns
    (rows/s, rows used, seconds)."""
    from oracle import cpu_agg
    cols = []
    ncol = len(colchunks[0])
    # bounded sample: whole chunks until roughly max_seconds of single-core work
    budget_rows = {"nogrp_agg": 100_000_000, "where_agg": 125_000_000,
                   "high_cardinality": 12_500_000}[workload]
    use, total = [], 0
    for ch in colchunks:
        use.append(ch)
        total += len(ch[0][0])
        if total >= budget_rows:
            break
    for c in range(ncol):
        v = np.concatenate([ch[c][0] for ch in use])
        if all(ch[c][1] is None for ch in use):
            m = None
        else:
            m = np.concatenate([np.zeros(len(ch[c][0]), np.uint8) if ch[c][1] is None
                                else ch[c][1] for ch in use])
        cols.append((v, m))
    dt, _, _, ng = cpu_agg.run(workload, cols, nthreads=nthreads, max_groups=1 << 24,
                               qual_const=qual_const)
    return total / dt, total, dt, ng


def bench_reference(args):
    """--impl reference: the reference's CPU implementation of the path
    (PostgreSQL Agg over SeqScan; PostgreSQL is not installable here, so the
    oracle's C restatement stands in - kind 'port') on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.rows or DEFAULT_ROWS[args.workload]
    chunk_rows = args.chunk_rows or DEFAULT_CHUNK[args.workload]
    cores = host_cores()
    colchunks = generate_columns(args.workload, 0, rows, chunk_rows, **column_options(args))
    times = []
    used = 0
    for i in range(args.warmup + args.steps):
        rps, used, dt, ng = run_cpu(args.workload, colchunks, cores,
                                    qual_const=args.selectivity)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = used / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
        "config": {"workload": args.workload,
                   "sql": SQL[args.workload].replace("f < 10", "f < %d" % args.selectivity),
                   "rows_per_step": used, "note": "CPU Agg over SeqScan on host cores"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": "%d rows of the %s table per step" % (used, args.workload)},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bench_ours(args):
    import torch
    from pg_strom_b200 import _capi
    from pg_strom_b200 import gpupreagg as gp
    from pg_strom_b200 import multigpu
    from pg_strom_b200 import workloads as W
    from oracle import bench_oracle

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    lib = _capi.load()
    gp.cuda_init([local_rank])
    workload = args.workload
    rows = args.rows or DEFAULT_ROWS[workload]
    chunk_rows = args.chunk_rows or DEFAULT_CHUNK[workload]
    w = W.WORKLOADS[workload]
    gucs = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on",
            "pg_strom.perfmon": "on"}
    plan_kw = {}
    if workload == "where_agg" and args.selectivity != 10:
        plan_kw["selectivity_pct"] = args.selectivity
    plan = gp.Plan(w["plan"](**plan_kw), gucs=gucs)
    assert plan.num_gpupreagg == 1, plan.reject_reason
    desc = plan.describe()
    node = plan.tree()["lefttree"]

    # ---- synthetic table: resident device chunks + pinned host chunks ----
    colchunks = generate_columns(workload, rank, rows, chunk_rows, **column_options(args))
    coltypes = [t for _, t in w["table"].columns]
    host_chunks, dev_chunks = [], []
    heap = (args.format == "row")
    fmt = gp.KDS_FORMAT_ROW if heap else gp.KDS_FORMAT_COLUMN
    for cols in colchunks:
        if heap:
            ds = gp.HeapDataStore(coltypes, cols, nrows=len(cols[0][0]))
            dptr, length = ds.upload(0)
        else:
            ds = gp.DataStore(coltypes, cols, nrows=len(cols[0][0]))
            length = ds.length
            dptr = lib.pgs_device_alloc(0, length)
            assert dptr, lib.pgs_last_error()
            _capi.check(lib.pgs_device_upload(0, dptr, ds.ptr, length))
        dev_chunks.append((dptr, length, ds.nrows))
        ds.free()
    e2e_chunk_rows = min(args.e2e_chunk_rows or DEFAULT_E2E_CHUNK[workload], rows)
    if heap:
        e2e_chunk_rows = min(e2e_chunk_rows, chunk_rows)
    for cols in generate_columns(workload, rank, rows, e2e_chunk_rows, **column_options(args)):
        if heap:
            host_chunks.append(gp.HeapDataStore(coltypes, cols, nrows=len(cols[0][0])))
            host_chunks[-1].length = host_chunks[-1].device_layout()[2]
        else:
            host_chunks.append(gp.DataStore(coltypes, cols, nrows=len(cols[0][0])))
    total_bytes = sum(d.length for d in host_chunks)
    nullable = sum(1 for c in colchunks[0] if c[1] is not None)
    alg_bytes_per_row = desc["row_bytes"] + nullable / 8.0

    sess = gp.Session(plan, max_async_chunks=3, max_chunk_rows=chunk_rows,
                      max_chunk_bytes=max(d.length for d in host_chunks))
    total_dev_bytes = sum(length for _, length, _ in dev_chunks)
    comm = C.c_void_p()
    if world > 1:
        comm = multigpu.nccl_communicator(lib, dist, 0, rank, world, device=dev)

    stream = torch.cuda.ExternalStream(sess.stream(), device=dev)

    def step_resident():
        for dptr, length, n in dev_chunks:
            sess.submit_device_format(dptr, length, n, fmt)
        if world > 1:
            _capi.check(lib.pgs_preagg_merge_nccl(sess.handle, comm, rank, world, 0))
        return sess.finish_raw()[1].nitems

    def step_e2e():
        for ds in host_chunks:
            sess.submit(ds)
        if world > 1:
            _capi.check(lib.pgs_preagg_merge_nccl(sess.handle, comm, rank, world, 0))
        return sess.finish_raw()[1].nitems

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + one correctness check against the oracle ----
    nresult = None
    for i in range(max(args.warmup, 3)):
        nresult = step_resident()
    if not args.no_check:
        # one checked pass: every rank restates its shard with the numpy
        # oracle, the root merges the expectations and compares them with the
        # rows the device path (NCCL merge included) returns - bit-exact
        merged = []
        for c in range(len(coltypes)):
            v = np.concatenate([ch[c][0] for ch in colchunks])
            m = None if colchunks[0][c][1] is None else np.concatenate([ch[c][1] for ch in colchunks])
            merged.append((v, m))
        part = bench_oracle.expected_partial_node(node, merged)
        del merged
        for dptr, length, n in dev_chunks:
            sess.submit_device_format(dptr, length, n, fmt)
        if world > 1:
            _capi.check(lib.pgs_preagg_merge_nccl(sess.handle, comm, rank, world, 0))
            parts = [None] * world if rank == 0 else None
            dist.gather_object(part, parts, dst=0)
        else:
            parts = [part]
        result = sess.finish()
        if rank == 0:
            keys, exp = bench_oracle.merge_expected(desc, parts)
            bench_oracle.assert_rows_equal_expected(desc, result, keys, exp)
        del part, parts

    # ---- timed region: device resident ----
    pm0 = sess.perfmon()
    l0 = sess.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step_resident()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    pm1 = sess.perfmon()
    launches = sess.launch_count() - l0
    tms = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tms.item()) / args.steps
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
        barrier()
        t0 = time.perf_counter()
        for dptr, length, n in dev_chunks:
            sess.submit_device_format(dptr, length, n, fmt)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if world > 1:
            _capi.check(lib.pgs_preagg_merge_nccl(sess.handle, comm, rank, world, 0))
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        sess.finish_raw()
        t3 = time.perf_counter()
        phase_ms["scan"] += (t1 - t0) * 1000.0 / 3
        phase_ms["merge"] += (t2 - t1) * 1000.0 / 3
        phase_ms["finish"] += (t3 - t2) * 1000.0 / 3

    # ---- timed region: end to end (host chunks in pinned memory) ----
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 5))
    for i in range(2):
        step_e2e()
    pm2 = sess.perfmon()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / e2e_steps
    pm3 = sess.perfmon()
    te = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    e2e_value = rows * world / (e2e_ms / 1000.0)
    h2d = (pm3["bytes_dma_send"] - pm2["bytes_dma_send"]) // e2e_steps
    d2h = (pm3["bytes_dma_recv"] - pm2["bytes_dma_recv"]) // e2e_steps

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        rps1, used, dt1, _ = run_cpu(workload, colchunks, 1, qual_const=args.selectivity)
        rpsn, used, dtn, _ = run_cpu(workload, colchunks, cores, qual_const=args.selectivity)
        cpu = {"value": rpsn, "unit": "rows/s", "cores": cores, "kind": "port",
               "value_1core": rps1,
               "sample": "%d rows of the %s table (oracle/cpu_agg.c: PostgreSQL-style "
                         "Agg over SeqScan; %.2fs on 1 core, %.2fs on %d cores)"
                         % (used, workload, dt1, dtn, cores)}

    if rank == 0:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                t = json.load(f).get(workload)
            if t and t.get("rows_per_launch") and k_rows and not heap:
                # ncu dram__bytes_read+write of one profiled launch, scaled to
                # the rows one timed launch processed
                traffic = t["dram_bytes_per_launch"] * (k_rows / t["rows_per_launch"])
        line = {
            "metric": METRIC,
            "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64/f64", "data": "synthetic",
            "config": {"workload": workload,
                       "key_distribution": ("zipf(1.0)" if column_options(args) else "uniform"),
                       "sql": SQL[workload].replace("f < 10", "f < %d" % args.selectivity),
                       "input_format": "KDS_FORMAT_ROW (heap pages, %.1f physical bytes per row)"
                                       % (total_dev_bytes / float(rows)) if heap
                                       else "KDS_FORMAT_COLUMN",
                       "rows_per_gpu": rows, "chunk_rows": chunk_rows,
                       "chunks_per_step": len(dev_chunks),
                       "e2e_chunk_rows": e2e_chunk_rows,
                       "input_bytes_per_gpu": total_dev_bytes,
                       "algorithmic_bytes_per_row": alg_bytes_per_row,
                       "l2_policy": "inputs (%.2f GB per step) are larger than L2 (126 MB)"
                                    % (total_bytes / 1e9),
                       "merge": "ncclAllGather of exported state blocks, import on rank 0" if world > 1 else "none",
                       "partial_rows": nresult},
            "gb_per_s": value * alg_bytes_per_row / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": ("gpupreagg_main_heap" if heap else
                                    "gpupreagg_main + gpupreagg_partagg" if pm1.get("part_nparts")
                                    else "gpupreagg_main"),
                         "launch_ms": k_ms,
                         "bytes_per_launch": k_rows * alg_bytes_per_row,
                         "launches_timed": n_k},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "rows/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "h2d_gb_per_s": h2d / (e2e_ms / 1000.0) / 1e9,
                    "pcie_gen5_x16_frac": h2d / (e2e_ms / 1000.0) / 1e9 / 64.0},
            "gpu_launches": int(launches),
            "phase_ms": phase_ms,
            "clocks": clocks,
        }
        print(json.dumps(line))
    sess.close()
    for dptr, _, _ in dev_chunks:
        lib.pgs_device_free(0, dptr)
    for ds in host_chunks:
        ds.free()
    if world > 1:
        lib.pgs_nccl_comm_destroy(comm)
        dist.destroy_process_group()


def main():
    args = parse_args()
    import __graft_entry__ as ge
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        ge.build()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
