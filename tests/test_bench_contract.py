"""bench.py contract pieces that run without a GPU: the `--impl reference` arm (the reference's op sequence on the host
cores) prints ONE JSON line with the agreed keys on rank 0 and nothing on the other ranks."""
import json
import os
import subprocess
import sys

import cases

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(cases.ROOT, "bench.py"), "--impl", "reference", *args],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run({}, "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    r = json.loads(lines[0])
    assert REQUIRED <= set(r), REQUIRED - set(r)
    assert r["impl"] == "reference" and r["higher_is_better"] is True and r["value"] > 0
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in r["config"] and r["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--steps", "1", "--warmup", "0", "--gpus", "2")
    assert p.returncode == 0 and p.stdout.strip() == ""
