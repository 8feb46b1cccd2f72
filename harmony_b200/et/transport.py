"""Control-plane transport: named endpoints exchanging small dict messages.

This replaces REEF's NetworkConnectionService + NameServer (Netty TCP, Avro payloads;
services/et/.../common/impl/NetworkConnectionImpl.java:72-95).  Two
implementations share one interface:

  * LocalTransport  -- in-process (threads): unit tests and the single-process launcher.
  * TcpTransport    -- one listener per process, length-prefixed pickle frames over persistent
                       sockets, a name registry hosted by the driver process.

Only control messages and the HOST-backend table traffic travel here.  On GPUs the data plane
(pull / push / migration payloads) never touches a socket: it is peer loads/reductions inside
kernels (ops/csrc).
"""
from __future__ import annotations

import logging
import pickle
import queue
import socket
import struct
import threading
from typing import Any, Callable, Dict, Optional, Tuple

LOG = logging.getLogger(__name__)
Handler = Callable[[Dict[str, Any]], None]


class NetworkError(RuntimeError):
    """Raised when the destination endpoint is unknown / unreachable (NetworkException)."""


class _Inbox:
    def __init__(self, endpoint_id: str, handler: Handler) -> None:
        self.endpoint_id = endpoint_id
        self.handler = handler
        self.q: "queue.Queue[Optional[Dict[str, Any]]]" = queue.Queue()
        self.thread = threading.Thread(target=self._run, name=f"inbox-{endpoint_id}", daemon=True)
        self.thread.start()

    def _run(self) -> None:
        while True:
            msg = self.q.get()
            if msg is None:
                return
            try:
                self.handler(msg)
            except Exception:  # noqa: BLE001
                LOG.exception("[%s] handler failed for message type %s", self.endpoint_id, msg.get("type"))

    def close(self) -> None:
        self.q.put(None)


class Transport:
    """Control-plane transport SPI (register an endpoint handler, send a dict message).  Replaces REEF
    NetworkConnectionService (services/et/.../common/impl/NetworkConnectionImpl.java:72-95)."""

    def register(self, endpoint_id: str, handler: Handler) -> None:
        raise NotImplementedError

    def unregister(self, endpoint_id: str) -> None:
        raise NotImplementedError

    def send(self, dst: str, msg: Dict[str, Any]) -> None:
        raise NotImplementedError

    def close(self) -> None:
        pass


class LocalTransport(Transport):
    """In-process transport: one inbox thread per endpoint (drivers, executors of the tests and of `-executor_mode
    inproc`); has a drop filter for fault injection."""

    def __init__(self) -> None:
        self._inboxes: Dict[str, _Inbox] = {}
        self._lock = threading.Lock()
        self.sent_count = 0
        self.drop_filter: Optional[Callable[[str, Dict[str, Any]], bool]] = None  # fault injection hook

    def register(self, endpoint_id: str, handler: Handler) -> None:
        with self._lock:
            if endpoint_id in self._inboxes:
                raise ValueError(f"endpoint {endpoint_id} already registered")
            self._inboxes[endpoint_id] = _Inbox(endpoint_id, handler)

    def unregister(self, endpoint_id: str) -> None:
        with self._lock:
            ib = self._inboxes.pop(endpoint_id, None)
        if ib is not None:
            ib.close()

    def send(self, dst: str, msg: Dict[str, Any]) -> None:
        if self.drop_filter is not None and self.drop_filter(dst, msg):
            raise NetworkError(f"injected failure sending to {dst}")
        with self._lock:
            ib = self._inboxes.get(dst)
            self.sent_count += 1
        if ib is None:
            raise NetworkError(f"unknown endpoint {dst}")
        ib.q.put(msg)

    def endpoints(self):
        with self._lock:
            return list(self._inboxes)

    def close(self) -> None:
        with self._lock:
            ibs, self._inboxes = list(self._inboxes.values()), {}
        for ib in ibs:
            ib.close()


def _send_frame(sock: socket.socket, obj: Any) -> None:
    data = pickle.dumps(obj, protocol=5)
    sock.sendall(struct.pack("!Q", len(data)) + data)


def _recv_exact(sock: socket.socket, n: int) -> Optional[bytes]:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(n - len(buf), 1 << 20))
        if not chunk:
            return None
        buf.extend(chunk)
    return bytes(buf)


def _recv_frame(sock: socket.socket) -> Any:
    hdr = _recv_exact(sock, 8)
    if hdr is None:
        return None
    (n,) = struct.unpack("!Q", hdr)
    body = _recv_exact(sock, n)
    if body is None:
        return None
    return pickle.loads(body)


class TcpTransport(Transport):
    """`name_server=None` makes this process the registry host (the driver)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0,
                 name_server: Optional[Tuple[str, int]] = None) -> None:
        self._host = host
        self._srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._srv.bind((host, port))
        self._srv.listen(128)
        self.address: Tuple[str, int] = (host, self._srv.getsockname()[1])
        self._name_server = name_server
        self._inboxes: Dict[str, _Inbox] = {}
        self._registry: Dict[str, Tuple[str, int]] = {}       # authoritative on the name server
        self._addr_cache: Dict[str, Tuple[str, int]] = {}
        self._conns: Dict[Tuple[str, int], socket.socket] = {}
        self._conn_locks: Dict[Tuple[str, int], threading.Lock] = {}
        self._lock = threading.RLock()
        self._closed = False
        threading.Thread(target=self._accept_loop, name="tcp-accept", daemon=True).start()

    # ---- server side ----
    def _accept_loop(self) -> None:
        while not self._closed:
            try:
                conn, _ = self._srv.accept()
            except OSError:
                return
            conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            threading.Thread(target=self._serve, args=(conn,), daemon=True).start()

    def _serve(self, conn: socket.socket) -> None:
        try:
            while not self._closed:
                frame = _recv_frame(conn)
                if frame is None:
                    return
                kind = frame.get("k")
                if kind == "msg":
                    with self._lock:
                        ib = self._inboxes.get(frame["dst"])
                    msg = frame["msg"]
                    if ib is None:
                        LOG.warning("dropping message for unknown local endpoint %s", frame["dst"])
                    elif msg.get("inline"):
                        # the sender vouches that the handler never blocks (per-mini-batch sync / grant): run it on
                        # this connection's reader thread -- per-sender order is kept, the inbox hop is saved
                        try:
                            ib.handler(msg)
                        except Exception:  # noqa: BLE001
                            LOG.exception("[%s] handler failed for message type %s", ib.endpoint_id, msg.get("type"))
                    else:
                        ib.q.put(msg)
                elif kind == "ns_register":
                    with self._lock:
                        self._registry[frame["id"]] = tuple(frame["addr"])
                    _send_frame(conn, {"ok": True})
                elif kind == "ns_unregister":
                    with self._lock:
                        self._registry.pop(frame["id"], None)
                    _send_frame(conn, {"ok": True})
                elif kind == "ns_lookup":
                    with self._lock:
                        addr = self._registry.get(frame["id"])
                    _send_frame(conn, {"addr": addr})
        except (OSError, EOFError):
            return
        finally:
            try:
                conn.close()
            except OSError:
                pass

    # ---- name service ----
    def _ns_call(self, req: Dict[str, Any]) -> Dict[str, Any]:
        assert self._name_server is not None
        with socket.create_connection(self._name_server, timeout=30) as s:
            _send_frame(s, req)
            return _recv_frame(s)

    def register(self, endpoint_id: str, handler: Handler) -> None:
        with self._lock:
            if endpoint_id in self._inboxes:
                raise ValueError(f"endpoint {endpoint_id} already registered")
            self._inboxes[endpoint_id] = _Inbox(endpoint_id, handler)
        if self._name_server is None:
            with self._lock:
                self._registry[endpoint_id] = self.address
        else:
            self._ns_call({"k": "ns_register", "id": endpoint_id, "addr": self.address})

    def unregister(self, endpoint_id: str) -> None:
        with self._lock:
            ib = self._inboxes.pop(endpoint_id, None)
        if ib is not None:
            ib.close()
        try:
            if self._name_server is None:
                with self._lock:
                    self._registry.pop(endpoint_id, None)
            else:
                self._ns_call({"k": "ns_unregister", "id": endpoint_id})
        except OSError:
            pass

    def _lookup(self, dst: str) -> Optional[Tuple[str, int]]:
        with self._lock:
            if dst in self._addr_cache:
                return self._addr_cache[dst]
            if self._name_server is None:
                addr = self._registry.get(dst)
            else:
                addr = None
        if addr is None and self._name_server is not None:
            res = self._ns_call({"k": "ns_lookup", "id": dst}).get("addr")
            addr = tuple(res) if res else None
        if addr is not None:
            with self._lock:
                self._addr_cache[dst] = addr
        return addr

    # ---- send ----
    def send(self, dst: str, msg: Dict[str, Any]) -> None:
        with self._lock:
            ib = self._inboxes.get(dst)
        if ib is not None:
            ib.q.put(msg)
            return
        addr = self._lookup(dst)
        if addr is None:
            raise NetworkError(f"unknown endpoint {dst}")
        with self._lock:
            lk = self._conn_locks.setdefault(addr, threading.Lock())
        with lk:
            for attempt in (0, 1):
                try:
                    sock = self._conns.get(addr)
                    if sock is None:
                        sock = socket.create_connection(addr, timeout=30)
                        sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                        sock.settimeout(None)
                        self._conns[addr] = sock
                    _send_frame(sock, {"k": "msg", "dst": dst, "msg": msg})
                    return
                except OSError as e:
                    self._conns.pop(addr, None)
                    with self._lock:
                        self._addr_cache.pop(dst, None)
                    if attempt == 1:
                        raise NetworkError(f"cannot reach {dst} at {addr}: {e}") from e

    def close(self) -> None:
        self._closed = True
        try:
            self._srv.close()
        except OSError:
            pass
        with self._lock:
            ibs, self._inboxes = list(self._inboxes.values()), {}
            conns, self._conns = list(self._conns.values()), {}
        for ib in ibs:
            ib.close()
        for c in conns:
            try:
                c.close()
            except OSError:
                pass
