"""Elastic Table service (host backend + HBM device backend)."""
from .config import (ExecutorConfiguration, RemoteAccessConfiguration, ResourceConfiguration,  # noqa: F401
                     TableConfiguration, TaskletConfiguration)
from .transport import LocalTransport, TcpTransport, NetworkError  # noqa: F401


def local_cluster(num_executors: int, conf: "ExecutorConfiguration | None" = None, **master_kw):
    """In-process ET cluster on a LocalTransport: returns (master, [AllocatedExecutor...], launcher)."""
    from .driver import ETMaster, InProcessLauncher
    transport = LocalTransport()
    launcher = InProcessLauncher()
    master = ETMaster(transport, launcher, **master_kw)
    execs = master.add_executors(num_executors, conf).get(timeout=30)
    return master, execs, launcher
