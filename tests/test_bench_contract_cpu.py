"""bench.py's JSON contract, exercised on the one arm that runs without a GPU: `--impl reference` (the reference's CPU
forward, oracle/_ref when staged, else the oracle port).  Also the non-zero ranks of a torchrun launch: they must exit 0
without work or output."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--ref-clips", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "clips/sec forward" and line["unit"] == "clips/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["steps"] == 1 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
