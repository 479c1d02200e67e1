"""bench.py contract checks that need no GPU: the reference arm (the CPU restatement timed on the host cores) prints exactly
one JSON line with the keys the driver reads, and the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--preroll", "20")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_s" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_cpu_sample_wall_clock_limit(monkeypatch):
    """The stacks workload's CPU sample can take the LU-per-pivot oracle minutes per env-step: bench.py runs it in a forked
    child with a limit, kills that child by its PID and reports (env-steps asked for) / limit as an upper bound."""
    import time
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "_cpu_baseline", lambda *a: (time.sleep(a[0]) or (7.0, 3.0, a[0], {"env_steps": 8})))
    assert bench._cpu_baseline_bounded(5, 0.05, 0, 0, 0, 0, 4, 2, 1) == (7.0, 3.0, 0.05, {"env_steps": 8}, False)
    val, lps, el, c, capped = bench._cpu_baseline_bounded(0.5, 30.0, 0, 0, 0, 0, 4, 2, 1)
    assert capped and val == 4 * 2 / 0.5 and el == 0.5 and c == {}
    assert bench._cpu_baseline_bounded(None, 0.01, 0, 0, 0, 0, 4, 2, 1)[4] is False
    assert bench.WORKLOADS["stacks"]["cpu_limit_s"] > 0 and "cpu_limit_s" not in bench.WORKLOADS["small"]
