"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and
the `ours` arm refuses to run without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-pairs", "20000", "--cpu-step-seconds", "0.05")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "reads_per_s_depleted" and j["unit"] == "reads/s"
    assert j["higher_is_better"] is True and j["steps"] == 2 and j["warmup"] == 1 and j["value"] > 0
    assert j["config"]["workload"].startswith("classifier: synthetic 10M 2x150 pairs")
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] == 2 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_ours_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "3", "--pairs", "1000")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
