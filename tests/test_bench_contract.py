"""CPU test of bench.py's reference arm (the only arm that runs without a GPU): one JSON line with the contract's
keys, timed on the unmodified reference (oracle/_ref) at a tiny scale."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref.available("pagerank"),
                               reason="oracle/_ref is built where /root/reference exists (make -C oracle ref)")


@needs_ref
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-scale", "12",
                          "--steps", "1", "--warmup", "1", "--iters", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GTEPS" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "sample" in d["config"] and d["config"]["sample_scale"] == 12
    assert d["warmup"] == 1
    # the reference arm never loads the product library (its input comes from oracle/librmat.so)
    assert d["repo_libs_loaded"] and all(l.startswith("oracle/") for l in d["repo_libs_loaded"])


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
