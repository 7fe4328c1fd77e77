"""One part per GPU over NCCL (needs 2 GPUs on the box; skipped otherwise): parity of the ghost-row reduction, the halo
mul! and the CG solve under the NCCL transport -- the single-GPU parity tests use the in-process transport."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpus_nccl_parity():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    nproc = max(n for n in (2, 4, 8) if n <= torch.cuda.device_count())   # part grids (2,1,1) / (2,2,1) / (2,2,2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert all(f"rank {k} ok" in r.stdout for k in range(nproc))
    print(r.stdout[-3000:])
