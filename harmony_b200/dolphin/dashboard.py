"""Optional metric dashboard: the driver POSTs every metric to an HTTP endpoint that stores it in
sqlite and serves a small JSON / HTML view.

Reference: dolphin/dashboard/DashboardConnector.java:112-217 (async HTTP client posting form data),
DashboardLauncher.java, jobserver/src/main/resources/dashboard/dashboard.py:47-151 (Flask + sqlite +
plotly).  No Flask in this image: the server is http.server from the standard library, the wire
format (POST form fields `id`, `metrics`, `time`) is kept.
"""
from __future__ import annotations

import json
import logging
import sqlite3
import threading
import time
import urllib.parse
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Any, Dict

LOG = logging.getLogger(__name__)


class DashboardServer:
    """Metric dashboard: HTTP endpoint receiving every metric as form data, sqlite store, plot page.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/dashboard/DashboardConnector.java:112-217,
    DashboardLauncher.java, jobserver/bin/dashboard.py."""

    def __init__(self, port: int = 9005, db_path: str = ":memory:") -> None:
        self.port = port
        self._db = sqlite3.connect(db_path, check_same_thread=False)
        self._db.execute("CREATE TABLE IF NOT EXISTS metrics (time REAL, id TEXT, metrics TEXT)")
        self._lock = threading.Lock()
        outer = self

        class H(BaseHTTPRequestHandler):
            def log_message(self, *a):  # quiet
                pass

            def do_POST(self):
                n = int(self.headers.get("Content-Length", 0))
                form = urllib.parse.parse_qs(self.rfile.read(n).decode("utf-8"))
                with outer._lock:
                    outer._db.execute("INSERT INTO metrics VALUES (?, ?, ?)",
                                      (float(form.get("time", [time.time()])[0]), form.get("id", [""])[0],
                                       form.get("metrics", ["{}"])[0]))
                    outer._db.commit()
                self.send_response(200); self.end_headers(); self.wfile.write(b"ok")

            def do_GET(self):
                rows = outer.rows()
                if self.path.startswith("/json"):
                    body, ctype = json.dumps(rows).encode(), "application/json"
                else:
                    trs = "".join(f"<tr><td>{t:.3f}</td><td>{i}</td><td><code>{m}</code></td></tr>" for t, i, m in rows[-200:])
                    body = (f"<html><body><h3>harmony_b200 dashboard ({len(rows)} metrics)</h3>"
                            f"<table border=1><tr><th>time</th><th>id</th><th>metrics</th></tr>{trs}</table>"
                            "</body></html>").encode()
                    ctype = "text/html"
                self.send_response(200); self.send_header("Content-Type", ctype); self.end_headers()
                self.wfile.write(body)

        self._srv = ThreadingHTTPServer(("127.0.0.1", port), H)
        self.port = self._srv.server_address[1]
        threading.Thread(target=self._srv.serve_forever, daemon=True).start()

    def rows(self):
        with self._lock:
            return list(self._db.execute("SELECT time, id, metrics FROM metrics ORDER BY time"))

    def close(self) -> None:
        self._srv.shutdown()
        self._db.close()


class DashboardConnector:
    """Metric sink plugged into DolphinMetricManager.sinks (fire-and-forget POSTs)."""

    def __init__(self, port: int, host: str = "127.0.0.1") -> None:
        self.url = f"http://{host}:{port}/"

    def post_metric(self, source_id: str, metric: Dict[str, Any]) -> None:
        data = urllib.parse.urlencode({"id": source_id, "metrics": json.dumps(metric, default=str),
                                       "time": time.time()}).encode()

        def go():
            try:
                urllib.request.urlopen(urllib.request.Request(self.url, data=data), timeout=2).read()
            except Exception:  # noqa: BLE001
                LOG.debug("dashboard unreachable")
        threading.Thread(target=go, daemon=True).start()
