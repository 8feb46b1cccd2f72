"""Graph data model, parsers, message manager and the Computation SPI.

Reference (jobserver/src/main/java/edu/snu/cay/pregel/): graph/api/{Vertex,Edge,Computation}.java,
graph/impl/{DefaultVertex,NoneValueVertex,DefaultEdge,NoneValueEdge,AbstractComputation}.java,
graph/impl/MessageManager.java:61-110 (two message tables flip-flopped per superstep; outgoing
messages are buffered per target vertex, optionally combined, then flushed with multiUpdate),
graph/impl/ComputationCallable.java:58-88 (remove the vertex's messages, wake it up, compute),
combiner/*.java, common/{DefaultGraphParser,NoneEdgeValueGraphParser}.java (`v t w t w ...` and
`v t t t ...` adjacency lists), common/MessageUpdateFunction.java (append to the message list).
"""
from __future__ import annotations

import threading
from typing import Any, Dict, Iterable, List, Optional

from ..common.dataloader import DataParser, register_parser


class Edge:
    """(target vertex, value).  Reference: pregel/graph/api/Edge.java."""

    __slots__ = ("target_vertex_id", "value")

    def __init__(self, target_vertex_id: int, value: Any = None) -> None:
        self.target_vertex_id, self.value = target_vertex_id, value


class Vertex:
    """Vertex id, value, edges, halted flag.  Reference: pregel/graph/api/Vertex.java."""

    def __init__(self, vertex_id: int, edges: List[Edge], value: Any = None) -> None:
        self.id, self.edges, self.value = vertex_id, edges, value
        self._halt = False

    def get_id(self) -> int:
        return self.id

    def get_value(self) -> Any:
        return self.value

    def set_value(self, v: Any) -> None:
        self.value = v

    def get_edges(self) -> List[Edge]:
        return self.edges

    def get_num_edges(self) -> int:
        return len(self.edges)

    def vote_to_halt(self) -> None:
        self._halt = True

    def wake_up(self) -> None:
        self._halt = False

    def is_halted(self) -> bool:
        return self._halt


class DefaultGraphParser(DataParser):
    """`vertexId target weight target weight ...` (weighted adjacency list)."""
    has_key = True

    def parse_line(self, line: str):
        parts = line.split()
        vid = int(parts[0])
        edges = [Edge(int(parts[i]), int(parts[i + 1])) for i in range(1, len(parts) - 1, 2)]
        return vid, Vertex(vid, edges)


class NoneEdgeValueGraphParser(DataParser):
    """`vertexId target target ...` (unweighted adjacency list)."""
    has_key = True

    def parse_line(self, line: str):
        parts = line.split()
        vid = int(parts[0])
        return vid, Vertex(vid, [Edge(int(t)) for t in parts[1:]])


register_parser("graph", DefaultGraphParser)
register_parser("graph_none_edge_value", NoneEdgeValueGraphParser)


class MessageCombiner:
    """combine(vertex, old, new) SPI applied at the receiving message table.  Reference:
    pregel/combiner/MessageCombiner.java."""

    def combine(self, vertex_id: int, original: Any, message: Any) -> Any:
        raise NotImplementedError


class DoubleSumMessageCombiner(MessageCombiner):
    """Sum combiner (PageRank).  Reference: pregel/combiner/DoubleSumMessageCombiner.java."""

    def combine(self, vertex_id, original, message):
        return original + message


class MinimumLongMessageCombiner(MessageCombiner):
    """Min combiner (shortest path).  Reference: pregel/combiner/MinimumLongMessageCombiner.java."""

    def combine(self, vertex_id, original, message):
        return min(original, message)


class MessageManager:
    """Current (incoming) / next (outgoing) message tables, swapped every superstep."""

    def __init__(self, table1, table2, combiner: Optional[MessageCombiner] = None) -> None:
        self._current, self._next = table1, table2
        self._combiner = combiner
        self._buf: Dict[int, List[Any]] = {}
        self._lock = threading.Lock()

    def get_current_message_table(self):
        return self._current

    def get_next_message_table(self):
        return self._next

    def add_message(self, target_vertex_id: int, message: Any) -> None:
        with self._lock:
            if self._combiner is not None:
                cur = self._buf.get(target_vertex_id)
                self._buf[target_vertex_id] = [message] if cur is None else \
                    [self._combiner.combine(target_vertex_id, cur[0], message)]
            else:
                self._buf.setdefault(target_vertex_id, []).append(message)

    def flush_all_messages(self) -> bool:
        """Send the buffered messages to the next table; returns whether any message was sent."""
        with self._lock:
            buf, self._buf = self._buf, {}
        if not buf:
            return False
        keys = list(buf)
        self._next.multi_update(keys, [buf[k] for k in keys]).get(timeout=600)
        return True

    def prepare_for_next_superstep(self) -> None:
        self._current, self._next = self._next, self._current


class Computation:
    """User logic: compute(vertex, messages) once per active vertex per superstep."""

    def __init__(self, message_manager: MessageManager, params: Optional[Dict[str, Any]] = None) -> None:
        self._mm = message_manager
        self._superstep = 0
        self.params = params or {}

    def initialize(self, superstep: int) -> None:
        self._superstep = superstep

    def get_superstep(self) -> int:
        return self._superstep

    def send_message(self, target_vertex_id: int, message: Any) -> None:
        self._mm.add_message(target_vertex_id, message)

    def send_messages_to_adjacents(self, vertex: Vertex, message: Any) -> None:
        for e in vertex.get_edges():
            self._mm.add_message(e.target_vertex_id, message)

    def compute(self, vertex: Vertex, messages: Iterable[Any]) -> None:
        raise NotImplementedError


def run_partition(computation: Computation, vertices: List[Vertex], current_msg_table) -> int:
    """ComputationCallable: returns the number of vertices still active after this superstep."""
    active = 0
    for v in vertices:
        msgs = current_msg_table.remove(v.get_id()).get(timeout=600) or []
        if msgs:
            v.wake_up()
        if not v.is_halted():
            computation.compute(v, msgs)
        if not v.is_halted():
            active += 1
    return active
