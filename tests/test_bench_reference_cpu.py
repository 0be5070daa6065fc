"""The reference arm of bench.py (`--impl reference`) needs no GPU: it times the CPU oracle port on the host cores.  Contract checked here:
one JSON line on rank 0 with the arm's keys; ranks != 0 exit 0 without work or output (the driver launches the arm under torchrun too)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *flags):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--cpu-eval-users", "64", *flags], capture_output=True, text=True, env=env, cwd=REPO, timeout=900)


def test_reference_arm_prints_the_contract_line():
    r = _run({}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "train_interactions_per_sec" and j["unit"] == "interactions/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["n_gpus"] == 2 and j["steps"] == 1 and j["warmup"] == 1
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    assert j["e2e"] == {"value": j["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "netflix" in j["config"]["workload"] and "n_gpus_note" in j["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
