"""The reference arm of bench.py (`--impl reference`: the reference's CPU algorithm on the host cores) runs without a
GPU; its JSON line carries the keys the driver reads. Under torchrun the other ranks exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env=None, workload="toy"):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "3", "--warmup", "1"],
                          capture_output=True, text=True, timeout=300, env=env)


def test_reference_arm_line():
    r = run()
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "rollout-steps/s" and line["unit"] == "rollout-steps/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 3 and line["dtype"] == "f64"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] and "model" not in line["config"]


def test_reference_arm_other_ranks_do_nothing():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
