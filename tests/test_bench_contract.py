"""The bench.py JSON-line contract: (1) the reference arm (`--impl reference`: the oracle's C restatement on the host cores) runs here
without a GPU and prints the contract keys; (2) the committed bench lines of the product arm (profiles/r2_bench_*.json, measured on
B200) carry every key the driver reads, with self-consistent numbers."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config"}


def _line(text):
    for l in text.splitlines():
        if l.startswith("{"):
            return json.loads(l)
    raise AssertionError("no JSON line")


def test_reference_arm_runs_on_the_host_cores():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-cells", "6"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = _line(out.stdout)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "nnz/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_bench_*_n[128].json"))))
def test_committed_bench_lines_follow_the_contract(path):
    d = _line(open(path).read())
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == int(path.rsplit("_n", 1)[1].split(".")[0]) and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert abs(d["value"] - d["config"]["nnz"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1.05
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (d["ms_per_step"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    if d["e2e"] is not None:
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["d2h_bytes_per_step"] > 0
        assert d["e2e"]["value"] < d["value"]      # host buffers, copies inside the timed region
    if d["n_gpus"] == 1 and "poisson" in path:
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
        g = d["general_route"]
        assert g["route"] == "sumfact-gather" and g["invariants"] == "ok" and 0 < g["roofline"]["frac"] < 1 and g["roofline"]["traffic"] > g["roofline"]["algorithmic_bytes_per_launch"]
        assert d["roofline_symbolic"]["algorithmic_bytes"] == 4 * d["config"]["nnz"] + 4 * (16581375 + 1)
    if "poisson" in path or "elasticity" in path:
        assert d["invariants"] == "ok"                # nnz closed form, symmetry, zero-row-sum count, sum(b - A u_h) at full size
