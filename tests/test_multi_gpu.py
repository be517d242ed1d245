"""Multi-GPU parity on hardware (needs >= 2 GPUs on the box; skipped otherwise): the target-sharded index and the
per-shard C-ABI entry (lrge_gpu_twoset_counts_shard) against the oracle, one process per GPU under torchrun."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_index_and_shard_entry_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "manual", "shard_entry_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "per-shard entry: counts identical = True" in r.stdout
