"""Two-GPU test of the NCCL plumbing behind the C ABI (cw_comm_*, cw_circuit_broadcast, cw_batch_gather_witness_packed,
cw_status_allreduce): skipped on single-GPU boxes; the world-size-2 host logic is covered on CPU by test_multi_rank_cpu.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_two_ranks_nccl_broadcast_gather_allreduce():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "nccl_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "nccl_check ok" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
