"""The JSON line bench.py prints, checked on the lines recorded on the B200 (profiles/bench_r01_*.json): every key the
driver's contract names is there, with the meaning the contract gives it.  CPU only (the lines were measured on the GPU)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lines(pattern):
    out = []
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern))):
        with open(path) as f:
            rows = [l for l in f.read().splitlines() if l.startswith("{")]
        if rows:
            out.append((os.path.basename(path), json.loads(rows[-1])))
    return out


@pytest.mark.parametrize("name,d", lines("bench_r01_v3*.json"))
def test_our_arm(name, d):
    assert d["metric"] == "photons propagated/sec" and d["unit"] == "photons/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["n_gpus"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    # value: whole-job photons over the max-over-ranks kernel time
    photons_per_step = 1048576 * 200 * d["n_gpus"]
    assert abs(d["value"] - photons_per_step / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "photons/s" and e["h2d_bytes_per_step"] == 1048576 * 48 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * d["value"] < e["value"] < d["value"]            # host buffers in the timed region: below the resident number
    assert d["gpu_launches"] == d["steps"]
    r = d["roofline"]
    assert r["bound"] == "compute-fp32-issue" and r["unit"] == "Tlane-op/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.3 < r["frac"] < 1.0 and r["traffic"] > 0 and r["hbm"]["achieved_gbs"] < 0.01 * r["hbm"]["peak_gbs"]
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_mhz"] >= 0.9 * c["sm_max_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        b = d["cpu_baseline"]
        assert b["kind"] == "port" and b["cores"] >= 1 and b["unit"] == "photons/s" and b["sample"] and 0 < b["value"] < 1e-2 * d["value"]


@pytest.mark.parametrize("name,d", lines("bench_ref_r01_v3*.json"))
def test_reference_arm(name, d):
    assert d["impl"] == "reference" and d["metric"] == "photons propagated/sec" and d["unit"] == "photons/s"
    assert d["higher_is_better"] is True and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
