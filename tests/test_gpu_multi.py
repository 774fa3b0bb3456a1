"""scenes sharded over the GPUs of one box, one NCCL all-gather of the results (mmw_gather_nccl), and the
G-GPU results are bit-identical per scene to the 1-GPU results.  Needs >= 2 GPUs (skipped on a single-GPU box; the
sharding arithmetic and the gather layout are covered on CPU by test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_results_equal_single_gpu_results_bitwise():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29541", worker], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical=True" in r.stdout
