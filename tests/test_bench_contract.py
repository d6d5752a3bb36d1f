"""CPU: the bench contract that can be checked without a GPU - the reference
arm (`bench.py --impl reference`: the CPU Agg-over-SeqScan port on the host
cores) prints one JSON line with the agreed keys for every workload, under
torchrun only rank 0 prints, and the product arm refuses to run without a
CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
        "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
        "cpu_baseline", "e2e", "gpu_launches"}


def _run(args, env=None, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT,
                          capture_output=True, text=True, timeout=timeout,
                          env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("workload", ["nogrp_agg", "where_agg", "high_cardinality", "nogrp_agg_heap"])
def test_reference_arm_line(workload):
    r = _run(["--impl", "reference", "--workload", workload, "--rows", "500000",
              "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference" and line["unit"] == "rows/s" and line["value"] > 0
    assert line["config"]["workload"] == workload and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rows/s",
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--rows", "200000", "--steps", "1", "--warmup", "0",
              "--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--rows", "200000", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())
