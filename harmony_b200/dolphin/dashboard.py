"""Optional metric dashboard: the driver POSTs every metric to an HTTP endpoint that stores it in
sqlite and serves a small JSON / HTML view.

Reference: dolphin/dashboard/DashboardConnector.java:112-217 (async HTTP client posting form data),
DashboardLauncher.java, jobserver/src/main/resources/dashboard/dashboard.py:47-151 (Flask + sqlite +
plotly).  No Flask / plotly in this image (and no network for a CDN): the server is http.server from the
standard library, the plots are inline-SVG line charts drawn by ~60 lines of page script.  The wire format
(POST form fields `id`, `metrics`, `time`) and the reference's URL surface are kept:
    POST /            store one metric               GET /           plot page (source + y-axis selectors, live)
    POST|GET /plot    {time: value} of one (id, y)   POST|GET /selectors   {ids, y_axis} of what has arrived
    GET /json         every stored row
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

            def _query(self):
                q = urllib.parse.urlparse(self.path).query
                if self.command == "POST":
                    n = int(self.headers.get("Content-Length", 0))
                    q = self.rfile.read(n).decode("utf-8") or q
                return {k: v[0] for k, v in urllib.parse.parse_qs(q).items()}

            def _reply(self, body: bytes, ctype: str) -> None:
                self.send_response(200); self.send_header("Content-Type", ctype); self.end_headers()
                self.wfile.write(body)

            def do_POST(self):
                path = urllib.parse.urlparse(self.path).path
                if path == "/plot":
                    q = self._query()
                    return self._reply(json.dumps(outer.series(q.get("id", ""), q.get("y", ""))).encode(), "application/json")
                if path == "/selectors":
                    return self._reply(json.dumps(outer.selectors()).encode(), "application/json")
                n = int(self.headers.get("Content-Length", 0))
                form = urllib.parse.parse_qs(self.rfile.read(n).decode("utf-8"))
                with outer._lock:
                    outer._db.execute("INSERT INTO metrics VALUES (?, ?, ?)",
                                      (float(form.get("time", [time.time()])[0]), form.get("id", [""])[0],
                                       form.get("metrics", ["{}"])[0]))
                    outer._db.commit()
                self._reply(b"ok", "text/plain")

            def do_GET(self):
                path = urllib.parse.urlparse(self.path).path
                if path.startswith("/json"):
                    return self._reply(json.dumps(outer.rows()).encode(), "application/json")
                if path == "/plot":
                    q = self._query()
                    return self._reply(json.dumps(outer.series(q.get("id", ""), q.get("y", ""))).encode(), "application/json")
                if path == "/selectors":
                    return self._reply(json.dumps(outer.selectors()).encode(), "application/json")
                self._reply(_PAGE.encode(), "text/html")

        self._srv = ThreadingHTTPServer(("127.0.0.1", port), H)
        self.port = self._srv.server_address[1]
        threading.Thread(target=self._srv.serve_forever, daemon=True).start()

    def rows(self):
        with self._lock:
            return list(self._db.execute("SELECT time, id, metrics FROM metrics ORDER BY time"))

    @staticmethod
    def _flatten(obj: Any, prefix: str = "") -> Dict[str, float]:
        """Numeric leaves of a metric blob, nested keys joined with dots (batch / epoch / server metrics alike)."""
        out: Dict[str, float] = {}
        if isinstance(obj, dict):
            for k, v in obj.items():
                out.update(DashboardServer._flatten(v, f"{prefix}{k}."))
        elif isinstance(obj, bool):
            pass
        elif isinstance(obj, (int, float)):
            out[prefix[:-1]] = float(obj)
        return out

    def selectors(self) -> Dict[str, Any]:
        """{ids: [...], y_axis: {id: [numeric metric names]}} of everything received so far (dashboard.py /selectors)."""
        ids: Dict[str, set] = {}
        for _t, i, m in self.rows():
            try:
                ids.setdefault(i, set()).update(self._flatten(json.loads(m)))
            except ValueError:
                continue
        return {"ids": sorted(ids), "y_axis": {i: sorted(v) for i, v in ids.items()}}

    def series(self, source_id: str, y: str) -> Dict[str, float]:
        """{seconds since the first metric: value} of metric `y` reported by `source_id` (dashboard.py /plot)."""
        rows = self.rows()
        t0 = rows[0][0] if rows else 0.0
        out: Dict[str, float] = {}
        for t, i, m in rows:
            if i != source_id:
                continue
            try:
                v = self._flatten(json.loads(m)).get(y)
            except ValueError:
                v = None
            if v is not None:
                out[f"{t - t0:.3f}"] = v
        return out

    def close(self) -> None:
        self._srv.shutdown()
        self._db.close()


_PAGE = """<!doctype html><html><head><meta charset="utf-8"><title>harmony_b200 dashboard</title>
<style>body{font-family:sans-serif;margin:1.2em}select{margin-right:1em}svg{border:1px solid #bbb;background:#fcfcfc}
.ax{stroke:#888;stroke-width:1}.ln{fill:none;stroke:#1f77b4;stroke-width:1.6}.pt{fill:#1f77b4}text{font-size:11px;fill:#444}</style>
</head><body><h3>harmony_b200 dashboard</h3>
<div>source <select id="sid"></select> metric <select id="y"></select> <span id="n"></span></div>
<svg id="c" width="900" height="360"></svg>
<script>
const W=900,H=360,L=60,R=15,T=15,Bm=30; let sel={ids:[],y_axis:{}};
function get(u){return fetch(u).then(r=>r.json());}
function fill(el,items,keep){const old=keep?el.value:null; el.innerHTML=items.map(i=>'<option>'+i+'</option>').join('');
  if(old&&items.includes(old)) el.value=old;}
function refreshSelectors(){return get('/selectors').then(s=>{sel=s; const sid=document.getElementById('sid');
  fill(sid,s.ids,true); fill(document.getElementById('y'),s.y_axis[sid.value]||[],true);});}
function draw(d){const c=document.getElementById('c'); const xs=Object.keys(d).map(Number), ys=xs.map(x=>d[x.toFixed(3)]);
  document.getElementById('n').textContent=xs.length+' points';
  if(!xs.length){c.innerHTML='<text x="20" y="30">no data yet</text>';return;}
  const x0=Math.min(...xs),x1=Math.max(...xs)||1,y0=Math.min(...ys),y1=Math.max(...ys);
  const sx=x=>L+(W-L-R)*(x1>x0?(x-x0)/(x1-x0):0.5), sy=y=>H-Bm-(H-T-Bm)*(y1>y0?(y-y0)/(y1-y0):0.5);
  let s='<line class="ax" x1="'+L+'" y1="'+(H-Bm)+'" x2="'+(W-R)+'" y2="'+(H-Bm)+'"/><line class="ax" x1="'+L+'" y1="'+T+'" x2="'+L+'" y2="'+(H-Bm)+'"/>';
  for(let k=0;k<=4;k++){const yv=y0+(y1-y0)*k/4, xv=x0+(x1-x0)*k/4;
    s+='<text x="4" y="'+(sy(yv)+4)+'">'+yv.toPrecision(4)+'</text><text x="'+(sx(xv)-10)+'" y="'+(H-8)+'">'+xv.toFixed(1)+'s</text>';}
  s+='<polyline class="ln" points="'+xs.map((x,i)=>sx(x)+','+sy(ys[i])).join(' ')+'"/>';
  s+=xs.map((x,i)=>'<circle class="pt" r="2" cx="'+sx(x)+'" cy="'+sy(ys[i])+'"/>').join(''); c.innerHTML=s;}
function refresh(){const sid=document.getElementById('sid').value, y=document.getElementById('y').value;
  if(sid&&y) get('/plot?id='+encodeURIComponent(sid)+'&y='+encodeURIComponent(y)).then(draw);}
document.getElementById('sid').onchange=()=>{fill(document.getElementById('y'),sel.y_axis[document.getElementById('sid').value]||[],false);refresh();};
document.getElementById('y').onchange=refresh;
function tick(){refreshSelectors().then(refresh);} tick(); setInterval(tick,2000);
</script></body></html>"""


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
