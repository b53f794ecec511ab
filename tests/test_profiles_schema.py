"""The bench lines committed under profiles/ (what DESIGN.md section 8 quotes) carry the keys the bench contract asks for:
a cheap guard against committing a truncated or hand-edited line."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lines():
    return sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*.json")))


def test_committed_round2_bench_lines_follow_the_contract():
    paths = _lines()
    assert len(paths) >= 5, paths
    for path in paths:
        d = json.load(open(path))
        name = os.path.basename(path)
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "e2e"):
            assert key in d, (name, key)
        assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["data"] == "synthetic" and d["vs_baseline"] is None
        assert d["value"] > 0 and d["e2e"]["value"] > 0 and "workload" in d["config"]
        if d.get("impl") == "reference":
            assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
            continue
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
        assert "sm_mhz" in d["clocks"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_default_line_names_the_default_kernel_and_has_a_cpu_baseline():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1_bf16.json")))
    assert d["n_gpus"] == 1 and d["dtype"] == "bf16" and "configs[2]" in d["config"]["workload"]
    assert d["config"]["conv_kernel"].startswith("round-1")
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["launches_per_step"] > 0 and d["host_ms_per_step"] > 0
