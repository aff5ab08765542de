"""The sharded engine on REAL GPUs, one process per GPU under torchrun (collected by `-m gpu`; skipped on a box
with fewer than two GPUs, where tests/test_multirank.py covers the same kernels with in-process ranks).
peers: the kernels store into the other GPUs' memory over NVLink and meet in the barrier kernel (gm_peer.cu);
nccl: all-gather callbacks.  Both must reproduce the single-rank oracle bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["peers", "nccl"])
def test_multigpu_parity_under_torchrun(mode):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs (this box has %d)" % ngpu)
    n = min(ngpu, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n, "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_parity_check.py"), "--exchange", mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("identical to the oracle") >= 7 and "MISMATCH" not in out.stdout
