"""bench.py's reference arm runs without a GPU: check the JSON contract of the line it prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4",
                          "--warmup", "2", "--batch", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "nmpc_solves_per_sec" and d["unit"] == "solves/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["n_gpus"] == 1
    assert d["steps"] == 4 and d["warmup"] == 2                 # the arm honours --steps / --warmup
    assert d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    # the same config block, key for key, as the CUDA arm prints (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_block(64)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_cpu_latency_and_reference_attempt_legs():
    """BASELINE.md section 3 items 1 and 3, as bench.py reports them inside cpu_baseline."""
    sys.path.insert(0, ROOT)
    import bench
    lat = bench.cpu_single_thread_latency()
    for algo in ("riccati_fp64", "schur_fp64"):
        for leg in ("config1_x200", "config2_1000_instances"):
            r = lat[algo][leg]
            assert 0 < r["p50"] <= r["p99"] and r["mean"] > 0
    assert lat["riccati_fp64"]["config2_1000_instances"]["n"] == 1000
    att = bench.reference_binary_attempt()
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "forces_attempt")):
        assert att["linked"] is True and att["exitflag"] == -100 and "license" in att["solver_message"].lower()
