"""Multi-GPU parity on real GPUs: runs tools/dist_check.py under torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange,path", [("peer", "grid2d"), ("collective", "grid2d"), ("peer", "kd")])
def test_two_rank_filter_matches_oracle(exchange, path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, PF_FRAMES="25", PF_PARTICLES="8192", PF_EXCHANGE=exchange, PF_PATH=path)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(helpers.ROOT, "tools", "dist_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK bit-exact" in r.stdout
