"""The multi-rank path on real GPUs (skips below two devices): spgemm_b200.multigpu.distribute / spgemm / concat /
gather_csr under torchrun with NCCL, compared with the oracle on rank 0 (tests/mgpu_worker.py)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                                  reason="needs two CUDA devices")]


@pytest.mark.parametrize("world", [2])
def test_multi_rank_path_matches_oracle(world):
    env = dict(os.environ, NCCL_DEBUG=os.environ.get("NCCL_DEBUG", "WARN"), TSG_ROWPLANS_MIN_ROWS="64")  # tile-row templates on the small test matrices too
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "MGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
