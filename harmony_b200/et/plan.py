"""Reconfiguration plans: a DAG of ops executed by the plan executor.

Reference: services/et/.../plan/api/Op.java:35-71 (11 op types), plan/impl/ETPlan.java:44-87
(DAG with ready-op tracking), plan/impl/PlanExecutorImpl.java:43,81-160,176-198 (16 threads, one
plan at a time, virtual -> actual executor id map filled by ALLOCATE, per-op listeners),
plan/impl/op/*.java (12 files), LoggingPlanExecutor.
"""
from __future__ import annotations

import itertools
import logging
import threading
import time
from typing import Any, Callable, Dict, List, Optional, Set

from ..utils.concurrent import CompletedFuture, ResultFuture, new_fixed_thread_pool
from ..utils.dag import DAG
from .config import ExecutorConfiguration, TableConfiguration, TaskletConfiguration
from .exceptions import PlanAlreadyExecutingError, PlanOpExecutionError

LOG = logging.getLogger(__name__)
ALLOCATE, DEALLOCATE, ASSOCIATE, UNASSOCIATE, SUBSCRIBE, UNSUBSCRIBE, MOVE, CREATE, DROP, START, STOP = (
    "ALLOCATE", "DEALLOCATE", "ASSOCIATE", "UNASSOCIATE", "SUBSCRIBE", "UNSUBSCRIBE", "MOVE", "CREATE", "DROP",
    "START", "STOP")
_op_ids = itertools.count()


class OpResult:
    """Result of one plan op.  Reference: services/et/src/main/java/edu/snu/cay/services/et/plan/api/Op.java:35-71."""

    def __init__(self, op: "Op", payload: Any = None) -> None:
        self.op, self.payload = op, payload

    @property
    def op_type(self) -> str:
        return self.op.op_type


class Op:
    """One node of an ETPlan: type, execute(context) -> future, virtual executor ids.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/plan/api/Op.java:35-71."""

    op_type = "?"

    def __init__(self) -> None:
        self.op_id = next(_op_ids)

    def execute(self, master, vmap: Dict[str, str]) -> ResultFuture:
        raise NotImplementedError

    @staticmethod
    def actual(vmap: Dict[str, str], eid: str) -> str:
        return vmap.get(eid, eid)

    def __hash__(self) -> int:
        return hash(self.op_id)

    def __eq__(self, other) -> bool:
        return isinstance(other, Op) and other.op_id == self.op_id

    def __repr__(self) -> str:
        fields = ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items() if k != "op_id")
        return f"{type(self).__name__}({fields})"


def _chain(src: ResultFuture, op: Op, payload_fn: Callable[[Any], Any] = lambda r: r) -> ResultFuture:
    out: ResultFuture = ResultFuture()
    src.add_listener(lambda r: out.on_completed(OpResult(op, payload_fn(r))))
    return out


class AllocateOp(Op):
    """ALLOCATE an executor (virtual id -> actual id).  Reference: services/et/.../plan/impl/op/AllocateOp.java."""

    op_type = ALLOCATE

    def __init__(self, virtual_id: str, executor_conf: Optional[ExecutorConfiguration] = None) -> None:
        super().__init__()
        self.virtual_id, self.executor_conf = virtual_id, executor_conf

    def execute(self, master, vmap):
        def done(execs):
            vmap[self.virtual_id] = execs[0].id
            return execs[0]
        return _chain(master.add_executors(1, self.executor_conf), self, done)


class DeallocateOp(Op):
    """DEALLOCATE an executor.  Reference: plan/impl/op/DeallocateOp.java."""

    op_type = DEALLOCATE

    def __init__(self, executor_id: str) -> None:
        super().__init__()
        self.executor_id = executor_id

    def execute(self, master, vmap):
        return _chain(master.get_executor(self.actual(vmap, self.executor_id)).close(), self)


class AssociateOp(Op):
    """ASSOCIATE an executor with a table.  Reference: plan/impl/op/AssociateOp.java."""

    op_type = ASSOCIATE

    def __init__(self, executor_id: str, table_id: str) -> None:
        super().__init__()
        self.executor_id, self.table_id = executor_id, table_id

    def execute(self, master, vmap):
        ex = master.get_executor(self.actual(vmap, self.executor_id))
        return _chain(master.get_table(self.table_id).associate([ex]), self)


class UnassociateOp(Op):
    """UNASSOCIATE an executor from a table (after its blocks moved away).  Reference:
    plan/impl/op/UnassociateOp.java."""

    op_type = UNASSOCIATE

    def __init__(self, executor_id: str, table_id: str) -> None:
        super().__init__()
        self.executor_id, self.table_id = executor_id, table_id

    def execute(self, master, vmap):
        return _chain(master.get_table(self.table_id).unassociate(self.actual(vmap, self.executor_id)), self)


class SubscribeOp(Op):
    """SUBSCRIBE an executor to a table's ownership updates.  Reference: plan/impl/op/SubscribeOp.java."""

    op_type = SUBSCRIBE

    def __init__(self, executor_id: str, table_id: str) -> None:
        super().__init__()
        self.executor_id, self.table_id = executor_id, table_id

    def execute(self, master, vmap):
        ex = master.get_executor(self.actual(vmap, self.executor_id))
        return _chain(master.get_table(self.table_id).subscribe([ex]), self)


class UnsubscribeOp(Op):
    """UNSUBSCRIBE.  Reference: plan/impl/op/UnsubscribeOp.java."""

    op_type = UNSUBSCRIBE

    def __init__(self, executor_id: str, table_id: str) -> None:
        super().__init__()
        self.executor_id, self.table_id = executor_id, table_id

    def execute(self, master, vmap):
        return _chain(master.get_table(self.table_id).unsubscribe(self.actual(vmap, self.executor_id)), self)


class MoveOp(Op):
    """MOVE n blocks of a table between executors.  Reference: plan/impl/op/MoveOp.java."""

    op_type = MOVE

    def __init__(self, src_executor_id: str, dst_executor_id: str, table_id: str, num_blocks: int) -> None:
        super().__init__()
        self.src_executor_id, self.dst_executor_id = src_executor_id, dst_executor_id
        self.table_id, self.num_blocks = table_id, num_blocks

    def execute(self, master, vmap):
        return _chain(master.get_table(self.table_id).move_blocks(
            self.actual(vmap, self.src_executor_id), self.actual(vmap, self.dst_executor_id), self.num_blocks), self)


class CreateOp(Op):
    """CREATE a table.  Reference: plan/impl/op/CreateOp.java."""

    op_type = CREATE

    def __init__(self, executor_ids: List[str], table_conf: TableConfiguration) -> None:
        super().__init__()
        self.executor_ids, self.table_conf = list(executor_ids), table_conf

    def execute(self, master, vmap):
        execs = [master.get_executor(self.actual(vmap, e)) for e in self.executor_ids]
        return _chain(master.create_table(self.table_conf, execs), self)


class DropOp(Op):
    """DROP a table.  Reference: plan/impl/op/DropOp.java."""

    op_type = DROP

    def __init__(self, table_id: str) -> None:
        super().__init__()
        self.table_id = table_id

    def execute(self, master, vmap):
        return _chain(master.get_table(self.table_id).drop(), self)


class StartOp(Op):
    """START a tasklet on an executor.  Reference: plan/impl/op/StartOp.java."""

    op_type = START

    def __init__(self, executor_id: str, tasklet_conf: TaskletConfiguration, metric_flush_period_ms: int = 0,
                 custom_handler: Optional[Callable[[Any], None]] = None) -> None:
        super().__init__()
        self.executor_id, self.tasklet_conf = executor_id, tasklet_conf
        self.metric_flush_period_ms, self.custom_handler = metric_flush_period_ms, custom_handler

    def execute(self, master, vmap):
        actual = self.actual(vmap, self.executor_id)
        master.metrics.start_metric_collection(actual, self.metric_flush_period_ms)
        rt = master.get_executor(actual).submit_tasklet(self.tasklet_conf, self.custom_handler)
        return CompletedFuture(OpResult(self, rt))


class StopOp(Op):
    """STOP a tasklet.  Reference: plan/impl/op/StopOp.java."""

    op_type = STOP

    def __init__(self, executor_id: str, tasklet_id: Optional[str] = None) -> None:
        super().__init__()
        self.executor_id, self.tasklet_id = executor_id, tasklet_id

    def execute(self, master, vmap):
        actual = self.actual(vmap, self.executor_id)
        ex = master.get_executor(actual)
        out: ResultFuture = ResultFuture()
        tasklets = ex.get_running_tasklets()
        targets = [t for tid, t in tasklets.items() if self.tasklet_id is None or tid == self.tasklet_id]

        def run():
            for t in targets:
                t.stop()
            for t in targets:
                try:
                    t.get_task_result(timeout=120)
                except TimeoutError:
                    LOG.warning("tasklet %s did not stop in time", t.tasklet_id)
            master.metrics.stop_metric_collection(actual)
            out.on_completed(OpResult(self, [t.tasklet_id for t in targets]))
        master.pool.submit(run)
        return out


class ETPlan:
    """DAG of ops; an op becomes ready when all its predecessors completed."""

    def __init__(self, dag: Optional[DAG] = None) -> None:
        self.dag: DAG = dag or DAG()
        self._initial = 0

    def add_op(self, op: Op, depends_on: Optional[List[Op]] = None) -> Op:
        self.dag.add_vertex(op)
        for d in depends_on or []:
            self.dag.add_vertex(d)
            self.dag.add_edge(d, op)
        return op

    def get_num_total_ops(self) -> int:
        return self.dag.get_num_vertices()

    def get_ready_ops(self) -> Set[Op]:
        return self.dag.get_root_vertices()

    def on_complete(self, op: Op) -> Set[Op]:
        """Remove `op`; returns the ops that became ready."""
        before = self.dag.get_root_vertices()
        self.dag.remove_vertex(op)
        return self.dag.get_root_vertices() - before

    def ops(self) -> List[Op]:
        return self.dag.vertices()

    def __repr__(self) -> str:
        return f"ETPlan({self.dag.to_json()})"


class PlanExecutor:
    """Executes one plan at a time on a pool of 16 threads."""
    NUM_THREADS = 16

    def __init__(self, master) -> None:
        self._m = master
        self._pool = new_fixed_thread_pool(self.NUM_THREADS, "plan-exec")
        self._listeners: Dict[str, Callable[[OpResult], None]] = {}
        self._busy = threading.Lock()
        self.last_op_times: Dict[str, float] = {}

    def register_listener(self, ident: str, cb: Callable[[OpResult], None]) -> None:
        self._listeners[ident] = cb

    def unregister_listener(self, ident: str) -> None:
        self._listeners.pop(ident, None)

    def execute(self, plan: ETPlan) -> ResultFuture:
        if not self._busy.acquire(blocking=False):
            raise PlanAlreadyExecutingError("a plan is already executing")
        result: ResultFuture = ResultFuture()
        total = plan.get_num_total_ops()
        if total == 0:
            self._busy.release()
            result.on_completed(None)
            return result
        vmap: Dict[str, str] = {}
        state = {"done": 0, "failed": None}
        lock = threading.Lock()
        t_begin = time.time()

        def finish(err: Optional[BaseException]) -> None:
            self._busy.release()
            LOG.info("plan finished in %.3fs (%d ops)", time.time() - t_begin, total)
            if err is None:
                result.on_completed(vmap)
            else:
                result.on_failed(err)

        def run_op(op: Op) -> None:
            t0 = time.time()
            try:
                fut = op.execute(self._m, vmap)
            except BaseException as e:  # noqa: BLE001
                LOG.exception("plan op %r failed", op)
                with lock:
                    first = state["failed"] is None
                    state["failed"] = e
                if first:
                    finish(PlanOpExecutionError(f"{op!r}: {e}"))
                return

            def on_done(res: OpResult) -> None:
                self.last_op_times[f"{op.op_type}-{op.op_id}"] = time.time() - t0
                for cb in list(self._listeners.values()):
                    try:
                        cb(res)
                    except Exception:  # noqa: BLE001
                        LOG.exception("plan listener failed")
                with lock:
                    if state["failed"] is not None:
                        return
                    nxt = plan.on_complete(op)
                    state["done"] += 1
                    all_done = state["done"] == total
                for n in nxt:
                    self._pool.submit(run_op, n)
                if all_done:
                    finish(None)
            fut.add_listener(on_done)

        for op in plan.get_ready_ops():
            self._pool.submit(run_op, op)
        return result


class LoggingPlanExecutor(PlanExecutor):
    """Logs the plan instead of executing it (reference: LoggingPlanExecutor.java)."""

    def execute(self, plan: ETPlan) -> ResultFuture:
        LOG.info("plan (not executed): %r", plan)
        return CompletedFuture({})
