"""bench.py contract checks that need no GPU: the reference arm (CPU restatement timed on the
host cores), its multi-rank behaviour, and the loud failure of the product arm without a device."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    e.pop("CUDA_VISIBLE_DEVICES", None)
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e,
                          cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "tiny_b1"])
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "tiny_b1" and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "tiny_b1"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "MASTER_ADDR": "127.0.0.1",
                  "MASTER_PORT": "29533"}, timeout=60)
    assert p.returncode == 0
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_product_arm_fails_loudly_without_a_gpu():
    p = _run(["--steps", "1", "--warmup", "3", "--workload", "tiny_b1"], timeout=120)
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stdout + p.stderr)
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
