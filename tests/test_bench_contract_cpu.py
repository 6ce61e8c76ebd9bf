"""The JSON line bench.py printed on the B200 (committed under profiles/) carries every key of
the bench contract - a guard against dropping one while editing bench.py."""

import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    with open(path) as f:
        return json.load(f)


@pytest.mark.parametrize("name,n", [("r1_bench_n1.json", 1), ("r1_bench_n2.json", 2), ("r1_bench_n8.json", 8),
                                    ("r2_bench_n1.json", 1), ("r2_bench_n8.json", 8)])
def test_bench_line_has_the_contract_keys(name, n):
    d = _load(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == n and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    if n == 1:
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c, k
        assert c["kind"] in ("port", "reference") and c["cores"] >= 1


@pytest.mark.parametrize("name,n", [("r2_bench_n1.json", 1), ("r2_bench_n8.json", 8)])
def test_round2_line_has_the_round2_records(name, n):
    """what the round-1 verdict asked the line to carry: exact mode with its own e2e, the mismatch
    rate of the fast mode on the workload, the sweep sub-record, whole-path traffic; plus the
    pipelined and metrics-only e2e legs"""
    d = _load(name)
    x = d["exact_mode"]
    assert x["dtype"] == "f64" and x["e2e"]["value"] > 0 and x["e2e"]["h2d_bytes_per_step"] > 0
    assert 0 < x["value"] < d["value"] and x["sync_api"]["value"] <= x["value"] * 1.02
    mm = d["fast_mode_mismatch"]
    assert 0 <= mm["coeff"] < 1e-4 and 0 <= mm["pixel"] < 1e-3 and mm["d_psnr_y_db"] < 1e-3 and mm["d_ssim_y"] < 1e-5
    sw = d["sweep"]
    assert sw["scaling"] == "strong" and sw["points"] == 100 and sw["n_gpus"] == n
    for leg in ("single_sweep", "single_device", "pipelined"):
        assert sw[leg]["ms_per_sweep"] > 0
    assert sw["single_device"]["ms_per_sweep"] <= sw["single_sweep"]["ms_per_sweep"]
    wp = d["roofline"]["whole_path"]
    assert wp["traffic"]["dram_bytes_per_pixel"] > wp["traffic"]["algorithmic_bytes_per_pixel"] == 6.0
    e = d["e2e"]
    assert e["pipelined"]["h2d_bytes_per_step"] == e["h2d_bytes_per_step"]
    assert e["pipelined"]["d2h_bytes_per_step"] == e["d2h_bytes_per_step"]
    assert e["metrics_only"]["d2h_bytes_per_step"] < 1e5 and e["metrics_only"]["value"] > e["value"]


def test_bench_source_mentions_every_contract_key():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for k in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"',
              '"higher_is_better"', '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"e2e"',
              '"gpu_launches"', '"clocks"', '"roofline"', '"cpu_baseline"', '"impl": "reference"',
              '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"traffic"'):
        assert k in src, k
