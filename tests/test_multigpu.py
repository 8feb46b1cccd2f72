"""Multi-GPU integration of the peer-memory data plane: one rank per GPU of the box (2, 4 or 8), launched through
torchrun.  Covers VMM / NVLS mappings, pull / push, the three ship-mode flavours, ship == pull after 40 steps, live
migration under racing pushes and the key guard (tests/mp_device_worker.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
HERE = os.path.dirname(os.path.abspath(__file__))


def test_peer_data_plane_all_gpus():
    import torch
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(HERE, "mp_device_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]
    out = os.path.join(os.path.dirname(HERE), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"multigpu_report_{n}gpu.json"), "w") as f:
            f.write([ln for ln in r.stdout.splitlines() if ln.startswith("MULTIGPU_REPORT ")][-1][16:])
