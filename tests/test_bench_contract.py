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
    assert j["config"]["workload"].startswith("C4: 100M 2x150 pairs") and j["scaling"] == "strong"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] == 2 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_c2_line():
    r = _run("--impl", "reference", "--config", "c2", "--steps", "1", "--warmup", "0", "--cpu-pairs", "20000",
             "--cpu-step-seconds", "0.05")
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert j["config"]["workload"].startswith("classifier: synthetic 10M 2x150 pairs") and j["scaling"] == "weak"


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_ours_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "3", "--pairs", "1000")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)


def test_committed_bench_lines_carry_the_contract_keys():
    """the bench lines committed under profiles/ (the evidence of the round) carry every key of the contract, and
    their derived numbers are consistent (value = reads / time, frac = achieved / peak, e2e counted from real bytes)"""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_v36*.json")))
    assert files
    for f in files:
        j = json.load(open(f))
        if j.get("impl") == "reference":
            assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["e2e"]["h2d_bytes_per_step"] == 0
            continue
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
            assert k in j, (f, k)
        assert j["metric"] == "reads_per_s_depleted" and j["unit"] == "reads/s" and j["dtype"] == "u8"
        assert j["scaling"] == "weak" and j["vs_baseline"] is None and j["data"] == "synthetic" and j["warmup"] >= 3
        assert "workload" in j["config"] and "model" not in j["config"]
        reads = 2 * j["config"]["pairs_per_gpu"] * j["n_gpus"]
        assert abs(j["value"] - reads / (j["ms_per_step"] * 1e-3)) / j["value"] < 1e-6
        r = j["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_ms"] * 1e-3) / 1e9) / r["achieved"] < 1e-6
        e = j["e2e"]
        assert e["h2d_bytes_per_step"] > j["config"]["fastq_bytes_per_gpu"] and e["d2h_bytes_per_step"] > 0
        assert e["value"] < j["value"] and abs(e["value"] - reads / (e["ms_per_step"] * 1e-3)) / e["value"] < 1e-6
        assert j["gpu_launches"] > 0 and j["clocks"]["sm_mhz"] and not set(j["clocks"]["reasons"]) & {
            "hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        c = j["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] == 2 and c["value"] > 0 and c["sample"]


def test_committed_round2_bench_lines():
    """profiles/r02_bench_c4_n*.json (the round's evidence): C4, strong scaling, every contract key, consistent derived
    numbers, the per-phase breakdown adds up to the step, e2e counted from real bytes and below the box's copy ceiling"""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_c4_n[1248].json")))
    assert files
    for f in files:
        j = json.load(open(f))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks",
                  "phases_ms"):
            assert k in j, (f, k)
        assert j["metric"] == "reads_per_s_depleted" and j["unit"] == "reads/s" and j["dtype"] == "u8"
        assert j["scaling"] == "strong" and j["vs_baseline"] is None and j["data"] == "synthetic" and j["warmup"] >= 3
        assert j["config"]["workload"].startswith("C4: 100M 2x150 pairs") and j["config"]["pairs_total"] == 100_000_000
        assert abs(j["value"] - 2 * 100_000_000 / (j["ms_per_step"] * 1e-3)) / j["value"] < 1e-6
        ph = j["phases_ms"]
        assert set(ph) == {"evidence", "set_build", "filter", "exchange"}
        assert sum(ph.values()) <= j["ms_per_step"] * 1.15 and ph["filter"] > 0.3 * j["ms_per_step"]
        r = j["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_ms"] * 1e-3) / 1e9) / r["achieved"] < 1e-6
        e = j["e2e"]
        assert e["h2d_bytes_per_step"] > j["config"]["fastq_bytes_total"] and e["d2h_bytes_per_step"] > 0
        assert e["value"] < j["value"] and abs(e["value"] - 2 * e["pairs"] / (e["ms_per_step"] * 1e-3)) / e["value"] < 1e-6
        c = e["host_copy_ceiling_all_ranks"]
        assert e["h2d_bytes_per_step"] / (e["ms_per_step"] * 1e-3) / 1e9 <= 1.05 * c["h2d_gb_per_s"]
        assert j["gpu_launches"] > 0 and j["clocks"]["sm_mhz"] and not set(j["clocks"]["reasons"]) & {
            "hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] == 2
