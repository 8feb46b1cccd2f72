"""Pregel on the new runtime: PageRank and SSSP complete and are numerically right (reference:
jobserver/src/test/.../pregel/integration/ExampleTest.java:36-46 only checks completion)."""
import os
import socket
import tempfile

import numpy as np
import pytest

from harmony_b200.pregel import apps
from harmony_b200.pregel.runtime import PregelJobLauncher, PregelLauncher


def _graph(tmp, weighted):
    # 0->1,2 ; 1->2 ; 2->0 ; 3->2 ; 4->3
    adj = {0: [(1, 4), (2, 1)], 1: [(2, 2)], 2: [(0, 7)], 3: [(2, 3)], 4: [(3, 1)]}
    path = os.path.join(tmp, "adj_list")
    with open(path, "w") as f:
        f.write("# graph\n")
        for v, es in adj.items():
            f.write(f"{v} " + " ".join(f"{t} {w}" if weighted else f"{t}" for t, w in es) + "\n")
    return path, adj


@pytest.mark.parametrize("mode", ["inproc", "process"])
def test_pagerank_matches_power_iteration(mode):
    """`process`: one OS process per Pregel worker, vertex / message tables and superstep control over TCP."""
    tmp = tempfile.mkdtemp(prefix="hbpr")
    path, adj = _graph(tmp, weighted=False)
    ok, results, pm = PregelLauncher.launch(apps.PAGERANK, ["-input", path, "-num_pregel_workers", "2",
                                                            "-executor_mode", mode])
    assert ok
    vals = {}
    for r in results:
        vals.update(r["values"])
    n = 5
    pr = np.ones(n)
    for _ in range(10):
        nxt = np.full(n, 0.15)
        for v, es in adj.items():
            for t, _ in es:
                nxt[t] += 0.85 * pr[v] / len(es)
        pr = nxt
    np.testing.assert_allclose([vals[i] for i in range(n)], pr, rtol=1e-9)
    assert pm.num_supersteps >= 11


def test_shortest_path_through_job_server():
    from harmony_b200.jobserver.client import JobServer
    from harmony_b200.utils.flags import Configuration
    tmp = tempfile.mkdtemp(prefix="hbsp")
    path, adj = _graph(tmp, weighted=True)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    js = JobServer(Configuration({"num_executors": 2, "port": port})).start()
    try:
        r = PregelJobLauncher.submit_job(apps.SHORTEST_PATH_PATH, apps.SHORTEST_PATH,
                                         ["-input", path, "-source", "4", "-port", str(port)], port=port)
        assert r.startswith("true Accept. JobId: ShortestPath-0"), r
        ok, results = js.driver.wait_job("ShortestPath-0", timeout=120)
        assert ok, results
        vals = {}
        for res in results:
            vals.update(res["values"])
        assert vals == {4: 0, 3: 1, 2: 4, 0: 11, 1: 15}
    finally:
        js.stop(kill=True)
