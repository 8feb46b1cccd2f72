"""Entry point of an executor process (spawned by SubprocessLauncher)."""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys
import threading
import time

from .config import ExecutorConfiguration
from .launcher import make_executor
from .transport import TcpTransport


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--id", required=True)
    ap.add_argument("--driver", required=True)
    ap.add_argument("--driver-id", default="driver")
    ap.add_argument("--backend", default="host")
    ap.add_argument("--ordinal", type=int, default=0)
    ap.add_argument("--conf", default="{}")
    a = ap.parse_args(argv)
    import faulthandler
    import signal
    # `kill -USR1 <pid>` dumps every thread's stack: to stderr, or to $HB_DUMP_DIR/<executor id>.stacks (several
    # executors dumping into one inherited stderr interleave their lines)
    dump_dir = os.environ.get("HB_DUMP_DIR")
    dump_file = open(os.path.join(dump_dir, f"{a.id}.stacks"), "w") if dump_dir else sys.stderr
    faulthandler.register(signal.SIGUSR1, file=dump_file, all_threads=True)
    logging.basicConfig(level=logging.INFO, format=f"%(asctime)s %(levelname)s [{a.id}] %(name)s | %(message)s")
    host, port = a.driver.rsplit(":", 1)
    transport = TcpTransport("127.0.0.1", 0, (host, int(port)))
    conf = ExecutorConfiguration.from_dict(json.loads(a.conf))
    ex = make_executor(a.id, transport, conf, a.driver_id, a.backend, a.ordinal)
    parent = os.getppid()

    def watch_parent() -> None:
        """An executor must not outlive its driver: when the job server is killed (not shut down) the process is
        re-parented -- release the GPU and exit instead of lingering as an orphan."""
        while os.getppid() == parent:
            time.sleep(1.0)
        logging.getLogger(__name__).warning("driver process %s is gone: exiting", parent)
        os._exit(0)
    threading.Thread(target=watch_parent, name="parent-watchdog", daemon=True).start()
    transport.send(a.driver_id, {"type": "ExecutorRegister", "executor_id": a.id, "src": a.id})
    ex.wait_closed()
    transport.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
