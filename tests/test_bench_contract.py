"""The bench contract on the CPU: the reference arm (`bench.py --impl reference`) runs without a GPU, prints one JSON line
with the keys the driver reads, tracks the same `config` object as the GPU arm and loads nothing of the product."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, VSB_BENCH_FRAMES="6", VSB_REF_UNITS_PAIRS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    # the arm maps the oracle's libraries only
    assert all(p.startswith("oracle/") for p in d["native_libraries_of_this_repo_loaded"]), d["native_libraries_of_this_repo_loaded"]
    ru = cb["reference_units"]
    if ru is not None and "error" not in ru:          # oracle/_ref is built in this container and travels to the GPU box
        assert ru["kind"] == "reference" and ru["max_abs_pose_diff_vs_port"] == 0.0


def test_both_arms_share_the_config_object():
    sys.path.insert(0, ROOT)
    import bench
    cfg = bench.base_config(1)
    assert cfg["frames"] == 2000 and cfg["pairs_per_step"] == 1999 and cfg["workload"].startswith("configs[1]")
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("base_config(world)") >= 2            # run_reference and run_gpu both emit it
