"""The bench lines committed under profiles/ (written by bench.py on a B200 box) carry every key the measurement contract
names, and their derived numbers are consistent with their inputs.  CPU only: reads JSON, runs nothing."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config")


def load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_gpu_arm_line_c2():
    d = load("r2_bench_final.json")
    for k in BASE + ("roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "kernels", "stages"):
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["unit"] == base.get("unit", d["unit"]) and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = audio seconds of K steps / device time
    assert d["value"] == pytest.approx(1000 * 5 * d["steps"] / (d["ms_per_step"] * d["steps"] / 1e3), rel=1e-9)
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12)
    # the roofline names the kernel with the longest launch of the step, and its numbers are that kernel's row
    k = d["kernels"]
    assert r["kernel"] == max(k, key=lambda n: k[n]["ms"])
    assert r["achieved"] == pytest.approx(k[r["kernel"]]["achieved_gbs"]) and r["traffic"] == k[r["kernel"]]["ncu_dram_bytes"]
    assert sum(v["ms"] for v in k.values()) == pytest.approx(sum(v["ms"] for v in d["stages"].values()), rel=1e-6)
    for v in k.values():
        assert v["achieved_gbs"] == pytest.approx(v["algorithmic_bytes"] / (v["ms"] * 1e-3) / 1e9, rel=1e-6)   # bytes are stored truncated
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    assert d["gpu_launches"] > 0
    cl = d["clocks"]
    assert cl["sm_mhz"] > 0.9 * cl["sm_max_mhz"] and not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = load("r2_bench_ref.json")
    g = load("r2_bench_final.json")
    assert d["impl"] == "reference" and d["metric"] == g["metric"] and d["unit"] == g["unit"]
    assert d["config"]["workload"] == g["config"]["workload"] and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")


def test_c3_line_carries_the_floor_and_the_gather():
    d = load("r2_c3_n1.json")
    for k in BASE + ("roofline", "e2e", "pcie_floor", "kernels", "stages", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["e2e"]["gathered_rows_per_step"] > 0 and 0 < d["pcie_floor"]["e2e_fraction_of_floor"] <= 1.05
    assert "u mod N" in d["config"]["workload"]
