"""The NCCL transport of the library under torch.distributed.run with one process per GPU ("what's weak" 2): the
in-process transport of test_gpu_multirank.py never touches ncclAllReduce.  Needs two GPUs; skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.timeout(600)
def test_two_ranks_over_nccl():
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540, cwd=ROOT)
    assert r.returncode == 0 and "NCCL_WORKER_OK 2" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
