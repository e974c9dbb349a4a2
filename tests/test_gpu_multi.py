"""The NCCL paths under a real multi-process launch: skipped on a box with one GPU."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:   # noqa: BLE001
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_sharded_msm_and_batch_verify_over_nccl(world):
    """One process per GPU (torch.distributed.run), libbpgpu's own ncclAllGather on the data path: the point-sliced MSM (resident
    handles with and without precomputed multiples, host operands) equals the oracle's full MSM on every rank, and the proof-sharded
    batch verification returns the whole batch's decisions on every rank (tests/multi_rank_check.py)."""
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_rank_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_OK world=%d" % world in out.stdout, (out.stdout[-2000:], out.stderr[-4000:])
