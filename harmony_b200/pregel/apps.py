"""Graph apps: PageRank (10 supersteps, damping 0.85) and single-source shortest path.

Reference: pregel/graphapps/pagerank/PagerankComputation.java:45-69,
pregel/graphapps/shortestpath/ShortestPathComputation.java:42-66, SourceId.java.
"""
from __future__ import annotations

from typing import Iterable, List

from ..dolphin import params as P
from .graph import Computation, Vertex
from .runtime import PregelApp, PregelJobLauncher, PregelLauncher

DAMPING_FACTOR, NUM_TOTAL_SUPERSTEP = 0.85, 10
LONG_MAX = 2 ** 63 - 1


class PagerankComputation(Computation):
    """PageRank vertex program (damping 0.85, fixed number of supersteps).  Reference:
    jobserver/src/main/java/edu/snu/cay/pregel/graphapps/pagerank/PagerankComputation.java:45-69."""

    def compute(self, vertex: Vertex, messages: Iterable[float]) -> None:
        if self.get_superstep() == 0:
            vertex.set_value(1.0)
        else:
            vertex.set_value((1 - DAMPING_FACTOR) + DAMPING_FACTOR * sum(messages))
        if self.get_superstep() < NUM_TOTAL_SUPERSTEP:
            if vertex.get_num_edges():
                self.send_messages_to_adjacents(vertex, vertex.get_value() / vertex.get_num_edges())
        else:
            vertex.vote_to_halt()


class ShortestPathComputation(Computation):
    """Single-source shortest paths vertex program.  Reference:
    jobserver/src/main/java/edu/snu/cay/pregel/graphapps/shortestpath/ShortestPathComputation.java:42-66."""

    def compute(self, vertex: Vertex, messages: Iterable[int]) -> None:
        source = int(self.params.get("source", 0))
        if self.get_superstep() == 0:
            vertex.set_value(LONG_MAX)
        min_dist = 0 if vertex.get_id() == source else LONG_MAX
        for m in messages:
            if m < min_dist:
                min_dist = m
        if min_dist < vertex.get_value():
            vertex.set_value(min_dist)
            for e in vertex.get_edges():
                self.send_message(e.target_vertex_id, min_dist + int(e.value))
        vertex.vote_to_halt()


PAGERANK = PregelApp("Pagerank", "harmony_b200.pregel.apps:PagerankComputation", "graph_none_edge_value",
                     "harmony_b200.pregel.graph:DoubleSumMessageCombiner")
PAGERANK_PATH = "harmony_b200.pregel.apps:PAGERANK"
SHORTEST_PATH = PregelApp("ShortestPath", "harmony_b200.pregel.apps:ShortestPathComputation", "graph",
                          "harmony_b200.pregel.graph:MinimumLongMessageCombiner", P.PREGEL)
SHORTEST_PATH_PATH = "harmony_b200.pregel.apps:SHORTEST_PATH"


def _run(app: PregelApp, argv: List[str]) -> int:
    ok, results, _ = PregelLauncher.launch(app, argv)
    print(f"LauncherStatus: {'COMPLETED' if ok else 'FAILED'}")
    return 0 if ok else 1


def main_run_pagerank(argv: List[str]) -> int:
    return _run(PAGERANK, argv)


def main_run_shortest_path(argv: List[str]) -> int:
    return _run(SHORTEST_PATH, argv)


def main_submit_pagerank(argv: List[str]) -> int:
    r = PregelJobLauncher.submit_job(PAGERANK_PATH, PAGERANK, argv)
    print(r)
    return 0 if r.startswith("true") else 1


def main_submit_shortest_path(argv: List[str]) -> int:
    r = PregelJobLauncher.submit_job(SHORTEST_PATH_PATH, SHORTEST_PATH, argv)
    print(r)
    return 0 if r.startswith("true") else 1
