"""GPU, N = 2: subsample-sharded Step 1 + gene-sharded Step 3 with NCCL all-reduces must give the
reference's scores (same golden fixtures as the single-GPU parity tests); the multi-class branch gene-sharded
against its reference fixtures.  Skipped with < 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_fit_matches_reference():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_fit_worker.py"),
           "hbecs_like", "srbct_s200", "cells10k_m70", "multi_k3_sqrt", "multi_k3_unbalanced", "multi_k4_kruskal_counts_f32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("rank0 ok") == 6
