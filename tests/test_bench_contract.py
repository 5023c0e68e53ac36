"""bench.py host-side contract (no GPU): the reference arm prints one JSON line with the keys the
driver reads, on the cores the container really has; non-zero ranks exit quietly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], env=env,
                          capture_output=True, text=True, timeout=600)


def test_host_cores_respects_affinity():
    sys.path.insert(0, ROOT)
    import bench
    n = bench.host_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
    if hasattr(os, "sched_getaffinity"):
        assert n <= len(os.sched_getaffinity(0))


def test_reference_arm_json_line():
    r = _run(None, "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
             "--ref-batch", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["metric"].startswith("Segmentor.fit images/sec")
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
             "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_cuda():
    r = _run(None, "--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
