"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/multi_gpu_parity.py
under torchrun with 2 ranks over NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["host_collective", "builtin", "p2p"])
def test_two_gpu_shard_gather_matches_oracle(mode):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", {"builtin": "29533", "host_collective": "29534", "p2p": "29535"}[mode],
           os.path.join(here, "multi_gpu_parity.py")]
    if mode != "host_collective":
        cmd.append("--" + mode)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(res.stdout[-2000:]); print(res.stderr[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
