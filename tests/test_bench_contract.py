"""The JSON line bench.py prints keeps the driver's contract (keys, units, types).  The reference
arm runs here on the CPU with a reduced N; the GPU arm is checked on the B200 box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(args, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-1000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench(["--impl", "reference", "--steps", "2", "--warmup", "3"], {"GH_BENCH_DIRECT_N": "8192"})
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "pairwise interactions/s" and d["unit"] == "interactions/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run_bench(["--steps", "3", "--warmup", "3"], {"GH_BENCH_DIRECT_N": "65536"})
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["metric"] == "pairwise interactions/s" and d["n_gpus"] == 1 and d["dtype"] == "f32"
    assert d["data"] == "synthetic" and d["scaling"] in ("weak", "strong")
    assert d["gpu_launches"] >= d["steps"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"]
    assert 0.2 < d["roofline"]["frac"] < 1.05
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] <= d["value"] * 1.05
    if d.get("clocks"):
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


def test_tree_roofline_object():
    """bench.tree_roofline assembles the tree workloads' roofline object from measured inputs; feed it
    the numbers of the recorded N = 4M run (profiles/r01_bench_tree_N4M_1gpu.json)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    n = 1 << 22
    st = {"accepted": int(1133.7616 * n), "visited": int(1372.0795 * n), "entries": 6214160, "cells": 2019856,
          "maxlevel": 18}
    acc = {"vs": "x", "timed_fp32_walk": {}, "reference_criterion_fp64_walk": {}}
    r = bench.tree_roofline(n, 1, 4.377, 3.000, st, "group", acc, {"hbm_gbs": 6541.1}, 74.44992)
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "build_roofline", "accuracy"):
        assert k in r
    assert r["kernel"] == "walk_group_kernel" and r["unit"] == "TFLOP/s"
    assert abs(r["achieved"] - 1133.7616 * n * 20 / 3.0e-3 / 1e12) < 1e-6 and 0.40 < r["frac"] < 0.45
    b = r["build_roofline"]
    assert b["bound"] == "hbm" and b["unit"] == "GB/s" and b["peak"] == 6541.1
    assert abs(b["achieved"] - 382.0 * n / 1.377e-3 / 1e9) < 1e-6 and 0.15 < b["frac"] < 0.2
    assert "of measured" in b["how"]
    t = bench.tree_roofline(n, 2, 4.377, 3.000, st, "target", acc, {}, 74.44992)
    assert t["kernel"] == "walk_kernel" and t["traffic"] is None
    assert t["build_roofline"]["peak"] == 6650.0 and "of fallback" in t["build_roofline"]["how"]
    assert abs(t["achieved"] - r["achieved"] / 2) < 1e-9
