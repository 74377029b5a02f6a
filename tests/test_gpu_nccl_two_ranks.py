"""The NCCL transport of the library under torch.distributed.run with one process per GPU ("what's weak" 2): the
in-process transport of test_gpu_multirank.py never touches ncclAllReduce.  Needs two GPUs; skipped on a one-GPU box."""
import os
import signal
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


@pytest.mark.timeout(300)
def test_two_ranks_over_nccl():
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_worker.py")]
    # own process group, so that a rank stuck in a collective is killed with its launcher (a plain subprocess timeout kills
    # torchrun only and then waits for the pipes the workers still hold)
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, start_new_session=True)
    try:
        out, err = p.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        pytest.fail("two-rank NCCL worker timed out\n" + out[-3000:] + "\n" + err[-5000:])
    assert p.returncode == 0 and "NCCL_WORKER_OK 2" in out, (out[-3000:], err[-5000:])
