"""Generic DAG used for reconfiguration plans (reference: utils/.../DAG.java:31-101, DAGImpl.java)."""
from __future__ import annotations

import json
import threading
from typing import Dict, Generic, Hashable, Set, TypeVar

V = TypeVar("V", bound=Hashable)


class DAG(Generic[V]):
    """Generic DAG (roots, neighbours, in-degree, JSON) used by reconfiguration plans.  Reference:
    utils/src/main/java/edu/snu/cay/utils/DAG.java:31-101, DAGImpl.java."""

    def __init__(self) -> None:
        self._adj: Dict[V, Set[V]] = {}
        self._indeg: Dict[V, int] = {}
        self._roots: Set[V] = set()
        self._lock = threading.RLock()

    def get_root_vertices(self) -> Set[V]:
        with self._lock:
            return set(self._roots)

    def is_adjacent(self, v: V, w: V) -> bool:
        with self._lock:
            return v in self._adj and w in self._adj[v]

    def get_neighbors(self, v: V) -> Set[V]:
        with self._lock:
            if v not in self._adj:
                raise KeyError(f"No such vertex {v}")
            return set(self._adj[v])

    def add_vertex(self, v: V) -> bool:
        with self._lock:
            if v in self._adj:
                return False
            self._adj[v] = set()
            self._indeg[v] = 0
            self._roots.add(v)
            return True

    def remove_vertex(self, v: V) -> bool:
        with self._lock:
            if v not in self._adj:
                return False
            for w in self._adj.pop(v):
                self._indeg[w] -= 1
                if self._indeg[w] == 0:
                    self._roots.add(w)
            # drop incoming edges too
            for u, ns in self._adj.items():
                ns.discard(v)
            self._indeg.pop(v, None)
            self._roots.discard(v)
            return True

    def add_edge(self, v: V, w: V) -> bool:
        with self._lock:
            if v not in self._adj or w not in self._adj:
                raise KeyError(f"No such vertex in edge {v} -> {w}")
            if w in self._adj[v]:
                return False
            if self._reachable(w, v):
                raise ValueError(f"Edge {v} -> {w} would create a cycle")
            self._adj[v].add(w)
            self._indeg[w] += 1
            self._roots.discard(w)
            return True

    def remove_edge(self, v: V, w: V) -> bool:
        with self._lock:
            if v not in self._adj or w not in self._adj[v]:
                return False
            self._adj[v].discard(w)
            self._indeg[w] -= 1
            if self._indeg[w] == 0:
                self._roots.add(w)
            return True

    def _reachable(self, src: V, dst: V) -> bool:
        stack, seen = [src], set()
        while stack:
            x = stack.pop()
            if x == dst:
                return True
            if x in seen:
                continue
            seen.add(x)
            stack.extend(self._adj.get(x, ()))
        return False

    def get_in_degree(self, v: V) -> int:
        with self._lock:
            return self._indeg[v]

    def get_num_vertices(self) -> int:
        with self._lock:
            return len(self._adj)

    def vertices(self):
        with self._lock:
            return list(self._adj)

    def to_json(self) -> str:
        with self._lock:
            return json.dumps({"vertices": [str(v) for v in self._adj],
                               "edges": [[str(v), str(w)] for v, ns in self._adj.items() for w in ns]})
