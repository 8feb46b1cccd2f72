"""Entry point of an executor process (spawned by SubprocessLauncher)."""
from __future__ import annotations

import argparse
import json
import logging
import sys

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
    logging.basicConfig(level=logging.INFO, format=f"%(asctime)s %(levelname)s [{a.id}] %(name)s | %(message)s")
    host, port = a.driver.rsplit(":", 1)
    transport = TcpTransport("127.0.0.1", 0, (host, int(port)))
    conf = ExecutorConfiguration.from_dict(json.loads(a.conf))
    ex = make_executor(a.id, transport, conf, a.driver_id, a.backend, a.ordinal)
    transport.send(a.driver_id, {"type": "ExecutorRegister", "executor_id": a.id, "src": a.id})
    ex.wait_closed()
    transport.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
