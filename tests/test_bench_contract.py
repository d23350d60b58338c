"""CPU: the bench.py contract on its CPU-runnable arm (`--impl reference`, tiny scale) and its helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--scale", "0.005"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["data"] == "synthetic" and d["value"] > 0


def test_reference_arm_non_zero_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--scale", "0.005"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300,
                         cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_algorithmic_bytes_match_the_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    t, f, b = bench.layer_bytes(232_965, 114_615_892, 32)
    assert round((t + f + b) / 1e6) == 2663 and round(t / 1e6) == 276 and round(f / 1e6) == 1194
    assert bench.measured_peak_gbs()[0] > 1000
