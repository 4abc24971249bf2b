"""Multi-GPU parity: the sharded retrieval paths -- query rows sharded over a
replicated database, database sharded and all-gathered, database sharded for good with shards travelling around a ring
-- and both all-gathers of the top-k lists (torch.distributed and the library's own NCCL communicator) must return, bit
for bit, what one GPU returns, for the bf16 and the fp32-accurate kernels; the exact path is also checked against the
fp64 ranking and the device label metrics against the oracle.  Needs two GPUs (skipped otherwise); one process per GPU."""
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_retrieval_matches_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(HERE, "_mgpu_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and "MGPU_OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-4000:]
