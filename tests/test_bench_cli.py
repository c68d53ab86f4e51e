"""CPU: bench.py's reference arm (the reference's own CPU detect path through oracle/_ref, or the plain-C port) prints ONE JSON line
with the contract's keys, on the same `config` the GPU arm reports (the driver compares the two arms' configs)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_config():
    sys.path.insert(0, ROOT)
    import bench
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.strip().split("\n") if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"] == bench.config_of("c2")          # what run_ours() emits as `config` for the default workload
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["median_ms_integral"] > 0 and cb["median_ms_scan"] > 0   # IntegralImage and the scan are timed separately
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data"):
        assert k in d
    assert set(bench.WORKLOADS) == {"c2", "c2_paper8", "c3"}
    assert all(os.path.exists(w["model"]) for w in bench.WORKLOADS.values())
