"""Job-server process + its command channel.

Reference: jobserver/.../client/JobServerClient.java:76-199 (parse flags, launch the driver),
CommandListener.java:39-63 (TCP :7008, one text line per command: `SUBMIT <serialized conf>` /
`SHUTDOWN <bool>`, replies `true|false`), CommandSender.java:49-83, JobServerCloser.java:32-45,
DriverLauncher.java:162-183.  The wire protocol is kept: a line-oriented TCP socket on port 7008.
"""
from __future__ import annotations

import logging
import socket
import socketserver
import sys
import threading
from typing import List, Optional

from ..dolphin import params as P
from ..et.driver import ETMaster, InProcessLauncher
from ..et.transport import LocalTransport, TcpTransport
from ..utils.flags import Configuration
from .driver import JobServerDriver

LOG = logging.getLogger(__name__)
DEFAULT_PORT = 7008
SUBMIT, SHUTDOWN, STATUS = "SUBMIT", "SHUTDOWN", "STATUS"


class JobServer:
    """The long-running master: ET master + job-server driver + TCP command listener."""

    def __init__(self, conf: Configuration) -> None:
        self.conf = conf
        mode = conf.get("executor_mode", "inproc")
        if mode == "process":
            from ..et.launcher import SubprocessLauncher
            self.transport = TcpTransport("127.0.0.1", 0, None)
            launcher = SubprocessLauncher(self.transport.address, backend=conf.get("backend", "auto"))
        else:
            self.transport = LocalTransport()
            launcher = InProcessLauncher(_executor_factory(conf.get("backend", "auto")))
        self.et_master = ETMaster(self.transport, launcher, chkp_temp_path=conf.get("chkp_temp_path", "/tmp/chkp"),
                                  chkp_commit_path=conf.get("chkp_commit_path", "/tmp/chkp/commit"))
        from ..et.launcher import detect_backend
        self.backend = detect_backend(conf.get("backend", "auto"))
        self.driver = JobServerDriver(self.et_master, conf.with_values(backend=self.backend))
        self._server: Optional[socketserver.ThreadingTCPServer] = None
        self._closed = threading.Event()
        self.port = int(conf.get("port", DEFAULT_PORT))

    def start(self) -> "JobServer":
        self.driver.init()
        outer = self

        class Handler(socketserver.StreamRequestHandler):
            def handle(self) -> None:
                for raw in self.rfile:
                    line = raw.decode("utf-8").rstrip("\r\n")
                    if not line:
                        continue
                    reply = outer.on_command(line)
                    self.wfile.write((reply + "\n").encode("utf-8"))
                    self.wfile.flush()

        socketserver.ThreadingTCPServer.allow_reuse_address = True
        self._server = socketserver.ThreadingTCPServer(("127.0.0.1", self.port), Handler)
        self._server.daemon_threads = True
        self.port = self._server.server_address[1]
        threading.Thread(target=self._server.serve_forever, name="command-listener", daemon=True).start()
        LOG.info("job server listening on 127.0.0.1:%d", self.port)
        return self

    def on_command(self, line: str) -> str:
        cmd, _, arg = line.partition(" ")
        if cmd == SUBMIT:
            reply = self.driver.on_submit(arg)
            return "true " + reply if reply.startswith("Accept") else "false " + reply
        if cmd == SHUTDOWN:
            kill = arg.strip().lower() == "true"
            threading.Thread(target=self._shutdown, args=(kill,), daemon=True).start()
            return "true"
        if cmd == STATUS:
            return "true running=" + ",".join(self.driver.running_jobs()) + " finished=" + \
                ",".join(f"{k}:{'ok' if v[0] else 'failed'}" for k, v in self.driver.job_results.items())
        return "false unknown command"

    def _shutdown(self, kill: bool) -> None:
        self.driver.on_shutdown(kill)
        self.et_master.close()
        if self._server is not None:
            self._server.shutdown()
        self.transport.close()
        self._closed.set()

    def wait_closed(self, timeout: Optional[float] = None) -> bool:
        return self._closed.wait(timeout)

    def stop(self, kill: bool = True) -> None:
        if not self._closed.is_set():
            self._shutdown(kill)


def _executor_factory(backend: str):
    import itertools
    from ..et.launcher import detect_backend, make_executor
    resolved = detect_backend(backend)
    counter = itertools.count()

    def factory(executor_id, transport, conf, driver_id):
        return make_executor(executor_id, transport, conf, driver_id, resolved, next(counter))
    return factory


class CommandSender:
    """Client that sends one command line to the job server's TCP port.  Reference:
    jobserver/src/main/java/edu/snu/cay/jobserver/client/CommandSender.java:49-83."""

    def __init__(self, port: int = DEFAULT_PORT, host: str = "127.0.0.1") -> None:
        self.addr = (host, port)

    def _call(self, line: str, timeout: float = 60.0) -> str:
        with socket.create_connection(self.addr, timeout=timeout) as s:
            s.sendall((line + "\n").encode("utf-8"))
            f = s.makefile("r", encoding="utf-8")
            return f.readline().rstrip("\n")

    def send_job_submit_command(self, serialized_conf: str) -> str:
        return self._call(f"{SUBMIT} {serialized_conf}")

    def send_shutdown_command(self, kill: bool) -> str:
        return self._call(f"{SHUTDOWN} {'true' if kill else 'false'}")

    def send_status_command(self) -> str:
        return self._call(STATUS)


def start_jobserver_main(argv: List[str]) -> int:
    """`start_jobserver.sh` entry point."""
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(name)s | %(message)s")
    reg = P.registry(P.COMMON, P.JOBSERVER, P.SERVICES)
    conf, _ = reg.parse(argv, allow_unknown=True)
    from ..utils import trace
    trace.configure(conf)
    server = JobServer(conf).start()
    print(f"Job Server is ready (port {server.port})", flush=True)
    timeout_ms = int(conf.get("timeout", 0))
    server.wait_closed(None if timeout_ms <= 0 else timeout_ms / 1000.0)
    server.stop(kill=True)
    where = trace.flush_to_receiver()
    if where:
        print(f"trace spans -> {where}", flush=True)
    return 0


def stop_jobserver_main(argv: List[str]) -> int:
    """`stop_jobserver.sh -kill <bool>` entry point (JobServerCloser)."""
    reg = P.registry(P.JOBSERVER)
    conf, _ = reg.parse(argv, allow_unknown=True)
    reply = CommandSender(int(conf.get("port", DEFAULT_PORT))).send_shutdown_command(bool(conf.get("kill", False)))
    print(reply)
    return 0 if reply.startswith("true") else 1


if __name__ == "__main__":
    sys.exit(start_jobserver_main(sys.argv[1:]))
