"""CLI scripts + one OS process per executor over the TCP transport (the reference's "multi-node
without a cluster" = REEF local runtime: driver + each evaluator a separate local process)."""
import os
import socket
import subprocess
import sys
import tempfile
import time

import pytest

from helpers import write_mlr_sample

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.integration
def test_start_submit_stop_scripts_with_executor_processes():
    tmp = tempfile.mkdtemp(prefix="hbcli")
    data = write_mlr_sample(os.path.join(tmp, "sample_mlr"), n=120)
    port = _free_port()
    env = dict(os.environ, PYTHONPATH=ROOT)
    srv = subprocess.Popen([os.path.join(ROOT, "bin", "start_jobserver.sh"), "-num_executors", "2", "-port", str(port),
                            "-executor_mode", "process", "-backend", "host", "-timeout", "0",
                            "-chkp_temp_path", tmp + "/t", "-chkp_commit_path", tmp + "/c"],
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        deadline = time.time() + 90
        line = ""
        while time.time() < deadline:
            line = srv.stdout.readline()
            if "Job Server is ready" in line:
                break
        assert "Job Server is ready" in line, line
        out = subprocess.run([os.path.join(ROOT, "bin", "submit_mlr.sh"), "-input", data, "-max_num_epochs", "2",
                              "-num_mini_batches", "4", "-num_worker_blocks", "4", "-num_server_blocks", "4",
                              "-step_size", "0.05", "-classes", "4", "-features", "64", "-port", str(port)],
                             env=env, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and "Accept. JobId: MLR-0" in out.stdout, out.stdout + out.stderr
        out = subprocess.run([os.path.join(ROOT, "bin", "stop_jobserver.sh"), "-kill", "false", "-port", str(port)],
                             env=env, capture_output=True, text=True, timeout=60)
        assert out.stdout.strip() == "true", out.stdout + out.stderr
        assert srv.wait(timeout=120) == 0
        log = srv.stdout.read()
        assert "Job finished. JobId: MLR-0" in log or "Epoch progress: [2 / 2], JobId: MLR-0" in log, log[-3000:]
    finally:
        if srv.poll() is None:
            srv.kill()
