"""2-GPU integration of the peer-memory data plane (launched through torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
HERE = os.path.dirname(os.path.abspath(__file__))


def test_peer_data_plane_two_gpus():
    import torch
    n = min(torch.cuda.device_count(), 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(HERE, "mp_device_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
