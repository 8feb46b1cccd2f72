"""GPU-resident Pregel supersteps (K13): vertex state and messages live in sharded device tables,
the message COMBINE is the update function of the push kernel, so messages sent to one vertex from
any GPU are reduced in the owner's memory system (sum for PageRank = red.add.f32, min for SSSP =
red.min.s32) while they arrive -- no message lists, no per-vertex host loop.

Reference: pregel/graph/impl/MessageManager.java:61-110 (two message tables flip-flopped per
superstep, combiner, multiUpdate to the next table), ComputationCallable.java:58-88,
graphapps/pagerank/PagerankComputation.java:45-69, shortestpath/ShortestPathComputation.java:42-66,
PregelMaster.java:126-171 (superstep control = AND-reduce of halt / no-messages; here a device clock
barrier + one all-reduced flag).
SPMD: rank r owns the vertices whose id % world == r (their adjacency lists and values).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from ..et.config import TableConfiguration
from ..et.device_table import DeviceTable
from ..parallel.device_clock import DeviceClock
from ..parallel.peer import PeerArena, PeerGroup

INT_INF = 2 ** 31 - 1


class DevicePregel:
    """GPU-resident supersteps: vertex values and flip-flopped message tables in device tables, the combiner is the
    push kernel's update function at the owner (sum / min).  Reference:
    pregel/graph/impl/MessageManager.java:61-110."""

    def __init__(self, arena: PeerArena, group: PeerGroup, num_vertices: int,
                 local_adj: Dict[int, List[Tuple[int, int]]]) -> None:
        """`local_adj`: vertex id -> [(target, weight)] for the vertices this rank owns."""
        from ..ops import native
        self._n = native()
        self.arena, self.group = arena, group
        self.rank, self.world = group.rank, group.world
        self.V = num_vertices
        self.device = torch.device("cuda", arena.device)
        vs = sorted(local_adj)
        self.local_vertices = torch.tensor(vs, dtype=torch.int64, device=self.device)
        deg = [len(local_adj[v]) for v in vs]
        self.deg = torch.tensor(deg, dtype=torch.float32, device=self.device)
        src_pos = np.repeat(np.arange(len(vs)), deg)
        self.edge_src = torch.tensor(src_pos, dtype=torch.int64, device=self.device)          # position in local_vertices
        self.edge_dst = torch.tensor([t for v in vs for t, _ in local_adj[v]], dtype=torch.int64, device=self.device)
        self.edge_w = torch.tensor([w for v in vs for _, w in local_adj[v]], dtype=torch.int32, device=self.device)
        self.clock = DeviceClock(arena, group)
        self._tables: List[DeviceTable] = []

    def _msg_table(self, tag: str, dtype: str) -> DeviceTable:
        nb = max(self.world, min(64, self.V))
        nb -= nb % self.world or 0
        conf = TableConfiguration(table_id=f"msg_table-{tag}", dense=True, capacity=self.V, row_dim=1, dtype=dtype,
                                  num_total_blocks=max(self.world, nb), is_ordered=False, backend="device",
                                  update_function="add" if dtype == "float32" else None)
        t = DeviceTable(conf, self.arena, self.group)
        self._tables.append(t)
        return t

    def _local_rows(self, t: DeviceTable) -> torch.Tensor:
        return t.pull(self.local_vertices)[:, 0]

    def _any(self, flag: bool) -> bool:
        if self.world == 1:
            return flag
        import torch.distributed as dist
        x = torch.tensor([1 if flag else 0], device=self.device)
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return bool(x.item())

    # ---- PageRank: sum combiner ----
    def pagerank(self, supersteps: int = 10, damping: float = 0.85) -> Dict[int, float]:
        cur, nxt = self._msg_table("pr-1", "float32"), self._msg_table("pr-2", "float32")
        rank = torch.ones(len(self.local_vertices), dtype=torch.float32, device=self.device)
        for step in range(supersteps + 1):
            if step > 0:
                msgs = self._local_rows(cur)                      # combined messages of my vertices
                rank = (1 - damping) + damping * msgs
                cur.put_rows(self.local_vertices, torch.zeros((len(rank), 1), device=self.device))   # consumed
            if step < supersteps and self.edge_dst.numel():
                contrib = (rank / self.deg.clamp(min=1))[self.edge_src]
                nxt.push(self.edge_dst, contrib[:, None].contiguous())   # scatter + COMBINE at the owners
            self.clock.barrier()                                   # every message of this superstep has landed
            cur, nxt = nxt, cur
        torch.cuda.synchronize(self.device)
        self.clock.check()
        return dict(zip(self.local_vertices.tolist(), rank.tolist()))

    # ---- single-source shortest path: min combiner ----
    def shortest_path(self, source: int, max_supersteps: int = 10_000) -> Dict[int, int]:
        from ..ops import count_launch
        cur, nxt = self._msg_table("sp-1", "int32"), self._msg_table("sp-2", "int32")
        inf_rows = lambda t: t.put_rows(self.local_vertices,  # noqa: E731
                                        torch.full((len(self.local_vertices), 1), float(INT_INF), device=self.device))
        for t in (cur, nxt):
            for b in t.local_block_ids():
                t.local_block(b).fill_(INT_INF)
        self.clock.barrier()
        dist_v = torch.full((len(self.local_vertices),), INT_INF, dtype=torch.int32, device=self.device)
        for step in range(max_supersteps):
            msgs = self._local_rows(cur).to(torch.int32)
            if step == 0:
                msgs = torch.where(self.local_vertices == source, torch.zeros_like(msgs), msgs)
            improved = msgs < dist_v
            dist_v = torch.minimum(dist_v, msgs)
            for b in cur.local_block_ids():                        # consumed: reset to +inf
                cur.local_block(b).fill_(INT_INF)
            active = improved[self.edge_src] if self.edge_dst.numel() else improved[:0]
            sent = bool(active.any().item()) if active.numel() else False
            if sent:
                keys = self.edge_dst[active]
                vals = (dist_v[self.edge_src[active]] + self.edge_w[active]).to(torch.int32)
                self._n.table_push_scalar_i32(*nxt.view(), keys, vals.contiguous(), 1)    # red.min at the owners
                count_launch()
            self.clock.barrier()
            cur, nxt = nxt, cur
            if not self._any(sent):                                # all halted and no messages in flight
                break
        torch.cuda.synchronize(self.device)
        self.clock.check()
        return dict(zip(self.local_vertices.tolist(), dist_v.tolist()))

    def close(self) -> None:
        for t in self._tables:
            t.close()


def partition_adjacency(adj: Dict[int, List[Tuple[int, int]]], num_vertices: int, rank: int, world: int):
    """Vertices without outgoing edges still exist (they receive messages)."""
    return {v: adj.get(v, []) for v in range(num_vertices) if v % world == rank}
