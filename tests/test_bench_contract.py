"""bench.py's reference arm (`--impl reference`) on the CPU: one JSON line with the contract's keys, ranks other than 0 silent.

The arm times nanogi's own CPU code (oracle/_ref) or, where that was not built, the oracle port — the one place besides `cpu_baseline`
where bench.py executes anything under oracle/. The GPU arm needs a device and is exercised on the GPU box.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True,
                          env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line_on_c1():
    r = _run({}, "--workload", "c1", "--steps", "1", "--warmup", "0", "--cpu-step-seconds", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "Mpaths/s" and j["higher_is_better"] is True and j["n_gpus"] == 1
    assert j["value"] > 0 and j["steps"] == 1 and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["config"]["workload"].startswith("C1") and j["config"]["renderer"] == "pt" and j["config"]["max_num_vertices"] == 8
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_runs_on_rank_0_only():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--workload", "c1", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_own_arm_fails_loudly_without_a_device():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the arm would run")
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "1", "--warmup", "3"], capture_output=True,
                       text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in r.stderr and not [l for l in r.stdout.splitlines() if l.startswith("{")]
