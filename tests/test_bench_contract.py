"""bench.py contract checks that need no GPU: the `--impl reference` arm (the oracle on the host cores) prints exactly
one JSON line with the keys the driver reads, and the non-rank-0 processes of a torchrun launch print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                           "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line_on_rank0_only():
    r0 = _run({"RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"})
    assert r0.returncode == 0, r0.stderr[-2000:]
    lines = [l for l in r0.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 1
    assert d["unit"] == "grid points/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    if (os.cpu_count() or 1) > 1:
        assert d["cpu_baseline"]["cores"] > 1                    # torchrun's OMP_NUM_THREADS=1 is overridden
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r1 = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""
