"""Dolphin reconfiguration plans and their compilation into ET op DAGs.

Reference: dolphin/plan/api/{Plan,TransferStep}.java, plan/impl/PlanImpl.java, EmptyPlan.java,
plan/impl/PlanCompiler.java:159-232 (add/del pairs across namespaces become a SWITCH: a server turns
into a worker or vice versa without re-allocation), :234-274 (op bookkeeping), :304-380 (switch),
:382-423 (delete), :425-472 (add), :276-302 (deallocate-before-allocate edges when resources are
short), :474-523 (move ops + their dependencies).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

from ..et.config import ExecutorConfiguration, TaskletConfiguration
from ..et.plan import (AllocateOp, AssociateOp, DeallocateOp, ETPlan, MoveOp, StartOp, StopOp, SubscribeOp,
                       UnassociateOp, UnsubscribeOp)

NAMESPACE_SERVER, NAMESPACE_WORKER = "SERVER", "WORKER"
HOMOGENEOUS, HETEROGENEOUS = "HOMOGENEOUS", "HETEROGENEOUS"


@dataclass(frozen=True)
class TransferStep:
    """Move `num_blocks` blocks from `src` to `dst` inside one namespace.  Reference:
    dolphin/plan/api/TransferStep.java."""

    src_id: str
    dst_id: str
    num_blocks: int


@dataclass
class Plan:
    """Optimizer output: evaluators to add / delete and transfer steps per namespace (SERVER / WORKER).  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/plan/impl/PlanImpl.java."""

    evaluators_to_add: Dict[str, List[str]] = field(default_factory=lambda: {NAMESPACE_SERVER: [], NAMESPACE_WORKER: []})
    evaluators_to_delete: Dict[str, List[str]] = field(default_factory=lambda: {NAMESPACE_SERVER: [], NAMESPACE_WORKER: []})
    transfer_steps: Dict[str, List[TransferStep]] = field(default_factory=lambda: {NAMESPACE_SERVER: [], NAMESPACE_WORKER: []})
    optimizer_type: str = HOMOGENEOUS
    num_available_extra_evals: int = 0

    def get_evaluators_to_add(self, ns: str) -> List[str]:
        return self.evaluators_to_add.get(ns, [])

    def get_evaluators_to_delete(self, ns: str) -> List[str]:
        return self.evaluators_to_delete.get(ns, [])

    def get_transfer_steps(self, ns: str) -> List[TransferStep]:
        return self.transfer_steps.get(ns, [])

    def is_empty(self) -> bool:
        return not any(self.evaluators_to_add.values()) and not any(self.evaluators_to_delete.values()) \
            and not any(self.transfer_steps.values())

    def validate(self) -> "Plan":
        for ns in (NAMESPACE_SERVER, NAMESPACE_WORKER):
            add, dele = set(self.get_evaluators_to_add(ns)), set(self.get_evaluators_to_delete(ns))
            if add & dele:
                raise ValueError(f"{add & dele} are both added and deleted in {ns}")
            for t in self.get_transfer_steps(ns):
                if t.src_id in add:
                    raise ValueError(f"a new evaluator {t.src_id} cannot be the source of a transfer")
                if t.dst_id in dele:
                    raise ValueError(f"a deleted evaluator {t.dst_id} cannot be the destination of a transfer")
        return self


def empty_plan(optimizer_type: str = HOMOGENEOUS) -> Plan:
    return Plan(optimizer_type=optimizer_type)


class PlanCompiler:
    """Plan -> ETPlan DAG: switch (server <-> worker without reallocation), add, delete, move, with the dependency
    edges between allocate / associate / subscribe / move / start / stop ops.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/plan/impl/PlanCompiler.java:159-232,234-274,304-380,382-423,425-472,474-523."""

    def __init__(self, model_table_id: str, input_table_id: str,
                 worker_tasklet_conf: Callable[[str], TaskletConfiguration],
                 server_tasklet_conf: Callable[[str], TaskletConfiguration],
                 worker_executor_conf: Optional[ExecutorConfiguration] = None,
                 server_executor_conf: Optional[ExecutorConfiguration] = None,
                 msg_handler: Optional[Callable[[Any], None]] = None, metric_flush_ms: int = 0,
                 local_model_table_id: Optional[str] = None) -> None:
        self.model_table_id, self.input_table_id = model_table_id, input_table_id
        self.local_model_table_id = local_model_table_id
        self._wt, self._st = worker_tasklet_conf, server_tasklet_conf
        self._wec, self._sec = worker_executor_conf, server_executor_conf
        self._msg_handler, self._flush = msg_handler, metric_flush_ms

    # ---- switch translation ----
    @staticmethod
    def _translate_to_switch(dels: List[str], adds: List[str], steps: List[TransferStep], n: int):
        """Pair `n` (delete-in-src-namespace, add-in-dst-namespace) ops: the deleted executor id
        replaces the added virtual id everywhere in the destination namespace's transfer steps."""
        switched: List[str] = []
        for _ in range(n):
            d, a = dels.pop(0), adds.pop(0)
            switched.append(d)
            steps = [TransferStep(d if s.src_id == a else s.src_id, d if s.dst_id == a else s.dst_id, s.num_blocks)
                     for s in steps]
        return switched, steps

    def compile(self, plan: Plan, num_available_extra_evals: int = 0) -> ETPlan:
        servers_del = list(plan.get_evaluators_to_delete(NAMESPACE_SERVER))
        workers_del = list(plan.get_evaluators_to_delete(NAMESPACE_WORKER))
        servers_add = list(plan.get_evaluators_to_add(NAMESPACE_SERVER))
        workers_add = list(plan.get_evaluators_to_add(NAMESPACE_WORKER))
        s_steps = list(plan.get_transfer_steps(NAMESPACE_SERVER))
        w_steps = list(plan.get_transfer_steps(NAMESPACE_WORKER))
        switch: Dict[str, List[str]] = {NAMESPACE_SERVER: [], NAMESPACE_WORKER: []}
        if plan.optimizer_type == HETEROGENEOUS:
            same = [e for e in servers_del if e in workers_add]
            for e in same:
                servers_del.remove(e); workers_add.remove(e); switch[NAMESPACE_SERVER].append(e)
            same = [e for e in workers_del if e in servers_add]
            for e in same:
                workers_del.remove(e); servers_add.remove(e); switch[NAMESPACE_WORKER].append(e)
        else:
            n_s2w = min(len(workers_add), len(servers_del))
            n_w2s = min(len(servers_add), len(workers_del))
            if n_s2w > 0:
                switch[NAMESPACE_SERVER], w_steps = self._translate_to_switch(servers_del, workers_add, w_steps, n_s2w)
            else:
                switch[NAMESPACE_WORKER], s_steps = self._translate_to_switch(workers_del, servers_add, s_steps, n_w2s)
        return self._build(switch, {NAMESPACE_SERVER: servers_add, NAMESPACE_WORKER: workers_add},
                           {NAMESPACE_SERVER: servers_del, NAMESPACE_WORKER: workers_del},
                           {NAMESPACE_SERVER: s_steps, NAMESPACE_WORKER: w_steps}, num_available_extra_evals)

    # ---- DAG construction ----
    def _start_op(self, executor: str, worker: bool) -> StartOp:
        conf = self._wt(executor) if worker else self._st(executor)
        return StartOp(executor, conf, self._flush, self._msg_handler if worker else None)

    def _build(self, switch, adds, dels, steps, extra: int) -> ETPlan:
        p = ETPlan()
        alloc: Dict[str, AllocateOp] = {}
        dealloc: Dict[str, DeallocateOp] = {}
        start: Dict[str, StartOp] = {}
        stop: Dict[str, StopOp] = {}
        assoc: Dict[str, AssociateOp] = {}
        unassoc: Dict[str, UnassociateOp] = {}
        sub: Dict[str, SubscribeOp] = {}
        unsub: Dict[str, UnsubscribeOp] = {}
        # apps with a worker-local model table (NMF's L, LDA's assignments): its blocks are key-aligned with the input
        # table's, so every worker-side associate / unassociate / move has a twin on that table
        local = self.local_model_table_id
        assoc_l: Dict[str, AssociateOp] = {}
        unassoc_l: Dict[str, UnassociateOp] = {}
        # switch
        for e in switch[NAMESPACE_SERVER]:        # server -> worker
            stop[e] = p.add_op(StopOp(e))
            unassoc[e] = p.add_op(UnassociateOp(e, self.model_table_id), [stop[e]])
            sub[e] = p.add_op(SubscribeOp(e, self.model_table_id), [unassoc[e]])
            assoc[e] = p.add_op(AssociateOp(e, self.input_table_id))
            start[e] = p.add_op(self._start_op(e, True), [assoc[e], sub[e]])
            if local:
                assoc_l[e] = p.add_op(AssociateOp(e, local))
                p.dag.add_edge(assoc_l[e], start[e])
        for e in switch[NAMESPACE_WORKER]:        # worker -> server
            stop[e] = p.add_op(StopOp(e))
            unsub[e] = p.add_op(UnsubscribeOp(e, self.model_table_id), [stop[e]])
            unassoc[e] = p.add_op(UnassociateOp(e, self.input_table_id), [stop[e]])
            if local:
                unassoc_l[e] = p.add_op(UnassociateOp(e, local), [stop[e]])
            assoc[e] = p.add_op(AssociateOp(e, self.model_table_id), [unsub[e]])
            start[e] = p.add_op(self._start_op(e, False), [assoc[e]])
        # delete
        for ns, evs in dels.items():
            table = self.input_table_id if ns == NAMESPACE_WORKER else self.model_table_id
            for e in evs:
                dealloc[e] = p.add_op(DeallocateOp(e))
                unassoc[e] = p.add_op(UnassociateOp(e, table))
                p.dag.add_edge(unassoc[e], dealloc[e])
                stop[e] = p.add_op(StopOp(e))
                if ns == NAMESPACE_WORKER:
                    unsub[e] = p.add_op(UnsubscribeOp(e, self.model_table_id), [stop[e]])
                    p.dag.add_edge(unsub[e], dealloc[e])
                    if local:
                        unassoc_l[e] = p.add_op(UnassociateOp(e, local))
                        p.dag.add_edge(unassoc_l[e], dealloc[e])
                else:
                    p.dag.add_edge(stop[e], dealloc[e])
        # add
        for ns, evs in adds.items():
            table = self.input_table_id if ns == NAMESPACE_WORKER else self.model_table_id
            for e in evs:
                alloc[e] = p.add_op(AllocateOp(e, self._wec if ns == NAMESPACE_WORKER else self._sec))
                assoc[e] = p.add_op(AssociateOp(e, table), [alloc[e]])
                if ns == NAMESPACE_WORKER:
                    sub[e] = p.add_op(SubscribeOp(e, self.model_table_id), [alloc[e]])
                    start[e] = p.add_op(self._start_op(e, True), [sub[e]])
                    if local:
                        assoc_l[e] = p.add_op(AssociateOp(e, local), [alloc[e]])
                        p.dag.add_edge(assoc_l[e], start[e])
                else:
                    start[e] = p.add_op(self._start_op(e, False), [alloc[e]])
        # add/del dependency when resources are short
        required = len(alloc) - len(dealloc)
        if required > extra:
            raise RuntimeError("Infeasible plan; it tries to use more resources than allowed.")
        must_follow = len(alloc) - extra
        if must_follow > 0:
            for a, d in zip(list(alloc.values())[:must_follow], list(dealloc.values())[:must_follow]):
                p.dag.add_edge(d, a)
        # move
        moves: List[MoveOp] = []
        for ns, ts in steps.items():
            table = self.input_table_id if ns == NAMESPACE_WORKER else self.model_table_id
            for t in ts:
                moves.append(p.add_op(MoveOp(t.src_id, t.dst_id, table, t.num_blocks)))
        if local:
            # the twin move follows the input move and picks the same (lowest) block ids: the two tables keep
            # identical partitions, so a record and its local-model row always live on the same worker
            for m in list(moves):
                if m.table_id == self.input_table_id:
                    tw = p.add_op(MoveOp(m.src_executor_id, m.dst_executor_id, local, m.num_blocks), [m])
                    if m.dst_executor_id in assoc_l:
                        p.dag.add_edge(assoc_l[m.dst_executor_id], tw)
                    if m.dst_executor_id in start:
                        p.dag.add_edge(tw, start[m.dst_executor_id])
                    if m.src_executor_id in unassoc_l:
                        p.dag.add_edge(tw, unassoc_l[m.src_executor_id])
                    if m.src_executor_id in stop:
                        p.dag.add_edge(stop[m.src_executor_id], tw)
        for m in moves:
            if m.dst_executor_id in assoc:
                p.dag.add_edge(assoc[m.dst_executor_id], m)
            if m.dst_executor_id in start:
                p.dag.add_edge(m, start[m.dst_executor_id])
            if m.src_executor_id in unassoc:
                p.dag.add_edge(m, unassoc[m.src_executor_id])
            if m.src_executor_id in stop:
                p.dag.add_edge(stop[m.src_executor_id], m)
        return p
