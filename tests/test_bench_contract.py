"""bench.py prints ONE JSON line with the keys the driver reads.  Small workloads through the
BTB_BENCH_N / BTB_BENCH_L overrides; the default (BASELINE config 4) is what the driver runs."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(args, n, L):
    env = dict(os.environ, BTB_BENCH_N=str(n), BTB_BENCH_L=str(L))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.timeout(600)
def test_reference_arm_line():
    d = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600, 500)
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "pairwise SNP distances/sec" and d["unit"] == "pairs/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_gpu_arm_line():
    d = run_bench(["--steps", "20", "--warmup", "3"], 6400, 3000)      # 20 steps: a 1 ms timed region is at the mercy of any hiccup
    assert BASE | {"gpu_launches", "clocks", "roofline"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["gpu_launches"] >= 2 * d["steps"] and d["data"] == "synthetic"
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "tensor"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 6400 * 3000 and e["d2h_bytes_per_step"] == 6400 * 6400 * 2 and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["rows_match_gpu"] is True and c["cells_compared"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
