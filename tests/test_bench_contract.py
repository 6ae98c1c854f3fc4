"""The committed bench lines (profiles/r0N_bench_*.json, written by bench.py on the B200 pool) carry every key of the
bench contract and are internally consistent.  Runs on CPU: it checks the evidence files, not the GPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not recorded yet")
    text = open(path).read().strip().splitlines()
    return json.loads([t for t in text if t.startswith("{")][-1])


@pytest.mark.parametrize("name,n", [("r01_bench_n1.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n8.json", 8)])
def test_bench_line_has_the_contract_keys(name, n):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == n and d["metric"] == "rays_per_sec" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    rays = d["config"]["rays_per_step_per_gpu"] * n
    assert abs(d["value"] - rays / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.2 < r["frac"] < 0.4 and 0.8 < r["issued_frac"] < 1.0 and r["mlp_share_of_step"] > 0.95
    e = d["e2e"]
    assert e["unit"] == "rays/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert 0.9 * d["value"] < e["value"] <= 1.02 * d["value"]
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] == "port" and b["cores"] >= 1 and b["value"] > 0 and "sample" in b
        assert r["traffic"] is not None and r["traffic"]["bytes_per_launch"] > 0
    else:
        assert d["cpu_baseline"] is None          # rank 0 at N = 1 only


def test_scaling_of_the_recorded_lines():
    one, eight = _line("r01_bench_n1.json"), _line("r01_bench_n8.json")
    assert eight["value"] / one["value"] > 6.0                     # north_star: >= 6x at 8 GPUs
    t1, t8 = one["train_step"], eight["train_step"]
    assert t8["value"] / t1["value"] > 6.0


def test_reference_arm_line():
    d = _line("r01_bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] == "rays_per_sec" and d["unit"] == "rays/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


# ------------------------------------------------------------------------------------------------ round 2
@pytest.mark.parametrize("name,n", [("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4), ("r02_bench_n8.json", 8)])
def test_round2_bench_lines(name, n):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == n and d["metric"] == "rays_per_sec" and d["warmup"] >= 3 and d["vs_baseline"] is None
    rays = d["config"]["rays_per_step_per_gpu"] * n
    assert abs(d["value"] - rays / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    assert d["value"] / n > 2.15e6
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["frac"] >= 0.36 and "mlp_tc3_kernel" in r["kernel"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"] and 0.9 * d["value"] < e["value"] <= 1.02 * d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    t = d["train_step"]
    assert t["cuda_graph"] is True and t["gpu_launches_per_step"] <= 30 and t["ms_per_step"] <= 7.5
    if n == 1:
        assert isinstance(r["traffic"], float) and r["traffic"] > 0 and r["traffic_detail"]["over_algorithmic_bytes"] < 1.1
        b = d["cpu_baseline"]
        assert b["kind"] == "reference" and b["cores"] >= 1 and b["value"] > 0 and "sample" in b
    else:
        assert d["cpu_baseline"] is None


def test_round2_strong_scaling_of_the_training_step():
    one, eight = _line("r02_bench_n1.json"), _line("r02_bench_n8.json")
    assert eight["value"] / one["value"] > 7.0
    strong = eight["train_step_strong_scaling"]
    assert one["train_step"]["ms_per_step"] / strong["ms_per_step"] >= 4.5
    assert strong["gpu_launches_per_step"] <= 30


def test_round2_reference_arm_line():
    d = _line("r02_bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] == "rays_per_sec" and d["unit"] == "rays/s"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_here_and_reports_our_config():
    """`bench.py --impl reference` is CPU-only: run it for real on a tiny sample.  It must print ONE JSON line with the contract keys of
    the reference arm, the same `config` dict our arm prints (the workload both numbers are about) and, with the reference staged
    by __graft_entry__.build(), time the unmodified reference (kind "reference"); without it, the oracle port (kind "port")."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-pixels", "8"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rays_per_sec" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["config"] == bench.workload_config(65536, 1)
    assert d["sample_rays_per_step"] == 8 * (bench.N_POSES + 2)
    b = d["cpu_baseline"]
    staged = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "model"))
    assert b["kind"] == ("reference" if staged else "port") and b["value"] == d["value"] > 0 and b["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
