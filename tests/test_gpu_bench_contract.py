"""bench.py prints ONE JSON line with the keys the driver reads (small --reads so it runs in seconds)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=280, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_bench_line_contract():
    j = _run("--reads", "200000", "--steps", "3", "--warmup", "3", "--recover-paths", "1", "--ref-sample", "4000",
             "--bam-reads", "50000")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in j, key
    assert j["n_gpus"] == 1 and j["steps"] == 3 and j["value"] > 0 and j["gpu_launches"] > 0
    assert j["vs_baseline"] is None and j["data"] == "synthetic" and "workload" in j["config"]
    assert j["e2e"]["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0
    assert j["e2e"]["value"] < j["value"]                       # the end-to-end number includes the copies
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = j["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert j["recovery"]["haplotypes"] >= 0
    assert j["parity_probe"]["ok"] is True and j["parity_probe"]["rows_equal"] and j["parity_probe"]["band_sum"] > 0
    b = j["e2e_bam"]
    assert b["reads_per_s"] > 0 and b["same_totals_as_packed_arrays"] and 0 < b["gpu_share"] < 1
    assert set(b["stage_seconds"]) >= {"read", "inflate", "scan", "walk", "gather", "gpu_ingest"}


def test_reference_arm_contract():
    j = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample", "3000")
    assert j["impl"] == "reference" and j["value"] > 0 and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["e2e"]["h2d_bytes_per_step"] == 0
