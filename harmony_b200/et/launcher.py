"""Executor launchers: the GPU resource pool of the box.

Reference: services/evaluator-manager/.../EvaluatorManager.java:39-86 + HomogeneousEvalManager /
HeterogeneousEvalManager (request N containers with (mem, cores), route allocation events to the
requester).  Here a "container" is one OS process pinned to one B200 (`SubprocessLauncher`) or a
thread-hosted executor of this process (`InProcessLauncher` in driver.py); heterogeneous requests
are matched by the `ResourceConfiguration.device` field.
"""
from __future__ import annotations

import logging
import os
import subprocess
import sys
import threading
from typing import Dict, Tuple

from .config import ExecutorConfiguration
from .driver import ExecutorLauncher

LOG = logging.getLogger(__name__)


def detect_backend(requested: str = "auto") -> str:
    if requested in ("host", "device"):
        return requested
    try:
        import torch
        from ..ops import native_available
        return "device" if torch.cuda.is_available() and native_available() else "host"
    except Exception:  # noqa: BLE001
        return "host"


def num_gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # noqa: BLE001
        return 0


def make_executor(executor_id: str, transport, conf: ExecutorConfiguration, driver_id: str, backend: str,
                  ordinal: int):
    """Factory shared by the in-process and the subprocess launchers."""
    if backend == "device":
        from .device_executor import DeviceExecutor
        dev = conf.resource.device if conf.resource.device is not None else ordinal % max(1, num_gpus())
        return DeviceExecutor(executor_id, transport, conf, driver_id, device=dev)
    from .executor import Executor
    return Executor(executor_id, transport, conf, driver_id)


class SubprocessLauncher(ExecutorLauncher):
    """One OS process per executor (one per GPU on the device backend), TCP control plane."""

    def __init__(self, name_server: Tuple[str, int], backend: str = "auto") -> None:
        self.name_server = name_server
        self.backend = detect_backend(backend)
        self._procs: Dict[str, subprocess.Popen] = {}
        self._count = 0
        self._lock = threading.Lock()

    def launch(self, executor_id: str, conf: ExecutorConfiguration, master) -> None:
        import json
        with self._lock:
            ordinal = self._count
            self._count += 1
        env = dict(os.environ)
        env["PYTHONPATH"] = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) + \
            os.pathsep + env.get("PYTHONPATH", "")
        cmd = [sys.executable, "-m", "harmony_b200.et.executor_main", "--id", executor_id,
               "--driver", f"{self.name_server[0]}:{self.name_server[1]}", "--driver-id", master.driver_id,
               "--backend", self.backend, "--ordinal", str(ordinal), "--conf", json.dumps(conf.to_dict())]
        proc = subprocess.Popen(cmd, env=env)
        self._procs[executor_id] = proc
        # the child registers itself with an ExecutorRegister message; a monitor thread turns an
        # unexpected exit into an executor-failed event (failure detection)

        def monitor():
            rc = proc.wait()
            if executor_id in self._procs and rc != 0:
                self._procs.pop(executor_id, None)
                master.on_executor_failed(executor_id, f"process exited with code {rc}")
        threading.Thread(target=monitor, name=f"monitor-{executor_id}", daemon=True).start()

    def terminate(self, executor_id: str) -> None:
        p = self._procs.pop(executor_id, None)
        if p is not None:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
