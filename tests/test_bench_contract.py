"""The bench lines committed under profiles/ (written by `bench.py` on the GPU box) carry every key the driver's contract
names, with the meaning the contract gives them.  A CPU-side regression guard for bench.py's output format."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    rows = [l for l in open(path) if l.startswith("{")]
    assert len(rows) == 1, "bench.py prints exactly one JSON line"
    return json.loads(rows[0])


def test_gpu_arm_line():
    d = _line("r2_final_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "scans/sec" and d["unit"] == "scans/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None   # BASELINE.md publishes no number
    assert d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert "500-key-frame" in d["config"]["workload"] and d["config"]["map"] == "kf500"   # BASELINE configs[3]
    # value = scans of all sequences / device time
    assert abs(d["value"] - d["config"]["batch_per_gpu"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]   # measured through the C ABI from host buffers, not a copy of `value`
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    assert d["gpu_launches"] > 0
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not any(x in d["clocks"]["reasons"] for x in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"))


def test_reference_arm_line():
    d = _line("r2_final_bench_reference.json")
    assert d["impl"] == "reference"
    assert d["metric"] == "scans/sec" and d["unit"] == "scans/s" and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["value"] == d["value"] and c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["sample"]
    g = _line("r2_final_bench.json")
    assert d["config"]["map"] == g["config"]["map"] and d["config"]["sensor"] == g["config"]["sensor"]
