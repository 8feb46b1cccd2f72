"""Tasklets, the tasklet runtime and the per-executor task-unit scheduler.

Reference: evaluator/api/Tasklet.java:23-35 (run/close), evaluator/impl/TaskletRuntime.java:69-130
(thread pool of `num_tasklets`, start/stop by driver messages, Running/Done/Failed reports,
per-tasklet custom message handler), evaluator/impl/LocalTaskUnitScheduler.java:36-145
(resource slots: 1 CPU, 2 NET, VOID = no wait; waitSchedule = Wait to the driver, block until the
driver's Ready -- all executors of the job reached the same unit -- then queue for a local slot).

On a GPU executor the COMP unit is the SM-bound phase and NET the NVLink-bound pull/push phase of
a job; the slots arbitrate co-located jobs so that job A's COMP overlaps job B's PULL/PUSH.
"""
from __future__ import annotations

import importlib
import logging
import threading
import traceback
from dataclasses import dataclass
from typing import Any, Dict, Optional

from .config import TaskletConfiguration

LOG = logging.getLogger(__name__)

CPU, NET, VOID = "CPU", "NET", "VOID"
SLOTS = {CPU: 1, NET: 2}


@dataclass(frozen=True)
class TaskUnitInfo:
    """(job, unit name, resource type, sequence number) of one schedulable phase.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/TaskUnitInfo.java:26-30."""

    job_id: str
    unit_name: str        # e.g. PULL / COMP / PUSH / SYNC
    resource: str         # CPU | NET | VOID
    seq: int = 0


class Tasklet:
    """User code executed on an executor. Subclasses override run() (and optionally close())."""

    def __init__(self, context: "TaskletContext") -> None:
        self.context = context

    def run(self) -> Any:
        raise NotImplementedError

    def close(self) -> None:
        pass

    def on_msg(self, payload: Any) -> None:   # TaskletCustomMsgHandler
        pass


class TaskletContext:
    """What a tasklet sees: its id and parameters, the executor's table accessor, metric collector, task-unit
    scheduler and a channel to the driver."""

    def __init__(self, executor, conf: TaskletConfiguration) -> None:
        self.executor = executor
        self.tasklet_id = conf.tasklet_id
        self.params: Dict[str, Any] = dict(conf.user_params)
        self.table_accessor = executor.table_accessor
        self.task_units = executor.task_unit_scheduler
        self.metric_collector = executor.metric_collector

    def send_to_driver(self, payload: Any, inline: bool = False) -> None:
        """TaskletCustomMsgSender.send (evaluator/impl/TaskletCustomMsgSender.java:53-70).  `inline`: the driver-side
        handler of this payload never blocks, the driver may run it on its inbox thread."""
        m = {"type": "TaskletCustom", "tasklet_id": self.tasklet_id, "executor_id": self.executor.executor_id,
             "payload": payload}
        if inline:
            m["inline"] = True
        self.executor.send_to_driver(m)


def load_class(path: str):
    mod, _, name = path.partition(":")
    return getattr(importlib.import_module(mod), name)


class TaskletRuntime:
    """Runs tasklets on daemon threads bounded by the tasklet slots (each under its own CUDA stream on a device
    executor); start / stop by driver messages, status reports, custom message routing.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/TaskletRuntime.java:69-130."""

    def __init__(self, executor, num_tasklets: int) -> None:
        self._ex = executor
        # daemon threads bounded by a semaphore (a parked tasklet must never block interpreter exit)
        self._slots = threading.Semaphore(max(1, num_tasklets))
        self._tasklets: Dict[str, Tasklet] = {}
        self._lock = threading.Lock()

    def start_tasklet(self, conf: TaskletConfiguration) -> None:
        ctx = TaskletContext(self._ex, conf)
        cls = load_class(conf.tasklet_class)
        tasklet = cls(ctx)
        with self._lock:
            if conf.tasklet_id in self._tasklets:
                raise RuntimeError(f"tasklet {conf.tasklet_id} already running")
            self._tasklets[conf.tasklet_id] = tasklet
        threading.Thread(target=self._run_slot, args=(conf.tasklet_id, tasklet), daemon=True,
                         name=f"{self._ex.executor_id}-tasklet-{conf.tasklet_id}").start()

    def _run_slot(self, tasklet_id: str, tasklet: Tasklet) -> None:
        if getattr(tasklet, "passive", False):
            # a tasklet that only parks (the servers' idle latch) must not hold an execution slot: with one server
            # tasklet per job on every executor, a handful of jobs would starve every worker tasklet
            self._run(tasklet_id, tasklet)
            return
        with self._slots:
            self._run(tasklet_id, tasklet)

    def _status(self, tasklet_id: str, status: str, result: Any = None, error: Optional[str] = None) -> None:
        self._ex.send_to_driver({"type": "TaskletStatus", "tasklet_id": tasklet_id,
                                 "executor_id": self._ex.executor_id, "status": status, "result": result,
                                 "error": error})

    def _run(self, tasklet_id: str, tasklet: Tasklet) -> None:
        self._status(tasklet_id, "Running")
        try:
            result = self._run_on_device_stream(tasklet)
            self._status(tasklet_id, "Done", result=result)
        except BaseException as e:  # noqa: BLE001
            LOG.error("tasklet %s failed:\n%s", tasklet_id, traceback.format_exc())
            self._status(tasklet_id, "Failed", error=f"{type(e).__name__}: {e}")
        finally:
            with self._lock:
                self._tasklets.pop(tasklet_id, None)

    def _run_on_device_stream(self, tasklet: Tasklet) -> Any:
        ctx = getattr(self._ex, "device_ctx", None)
        if ctx is None:
            return tasklet.run()
        import torch
        params = getattr(tasklet.context, "params", None) or {}
        high = str(params.get("job_priority", "normal")).lower() in ("high", "-1")
        stream = ctx.tasklet_stream(-1 if high else 0)
        with torch.cuda.device(ctx.device), torch.cuda.stream(stream):
            try:
                return tasklet.run()
            finally:
                stream.synchronize()
                release = getattr(ctx, "release_stream", None)
                if release is not None:
                    release(stream)
                # the trainer's cycles (trainer <-> model <-> streams / events / HBM tensors) die HERE, on the thread
                # and device that created them, instead of inside whichever thread happens to trigger the next
                # automatic collection (a GC run inside torch.cuda.memory_stats of another worker segfaulted once)
                tasklet = None   # noqa: F841
                import gc
                gc.collect()

    def stop_tasklet(self, tasklet_id: str) -> None:
        with self._lock:
            t = self._tasklets.get(tasklet_id)
        if t is not None:
            t.close()

    def on_custom_msg(self, tasklet_id: str, payload: Any) -> None:
        with self._lock:
            t = self._tasklets.get(tasklet_id)
        if t is not None:
            t.on_msg(payload)
        else:
            LOG.debug("custom message for unknown tasklet %s (already finished)", tasklet_id)

    def running(self):
        with self._lock:
            return list(self._tasklets)

    def close(self) -> None:
        with self._lock:
            ts = list(self._tasklets.values())
        for t in ts:
            try:
                t.close()
            except Exception:  # noqa: BLE001
                pass


class LocalTaskUnitScheduler:
    """Per-executor slots + the global per-job unit barrier hosted by the driver."""

    def __init__(self, executor) -> None:
        self._ex = executor
        self._sems = {r: threading.Semaphore(n) for r, n in SLOTS.items()}
        self._ready: Dict[Any, threading.Event] = {}
        self._lock = threading.Lock()
        self.enabled = True

    def _key(self, tasklet_id: str, u: TaskUnitInfo):
        return (tasklet_id, u.job_id, u.unit_name, u.seq)

    def wait_schedule(self, tasklet_id: str, unit: TaskUnitInfo) -> None:
        if not self.enabled or unit.resource == VOID:
            return
        key = self._key(tasklet_id, unit)
        ev = threading.Event()
        with self._lock:
            self._ready[key] = ev
        self._ex.send_to_driver({"type": "TaskUnitWait", "tasklet_id": tasklet_id,
                                 "executor_id": self._ex.executor_id, "job_id": unit.job_id,
                                 "unit_name": unit.unit_name, "seq": unit.seq, "inline": True})
        ev.wait()
        self._sems[unit.resource].acquire()

    def abort(self, tasklet_id: str) -> None:
        """Wake every unit wait of `tasklet_id` (its job is being aborted)."""
        with self._lock:
            keys = [k for k in self._ready if k[0] == tasklet_id]
            evs = [self._ready.pop(k) for k in keys]
        for ev in evs:
            ev.set()

    def on_ready(self, m: Dict[str, Any]) -> None:
        key = (m["tasklet_id"], m["job_id"], m["unit_name"], m["seq"])
        with self._lock:
            ev = self._ready.pop(key, None)
        if ev is not None:
            ev.set()

    def leave(self, job_id: str) -> None:
        """This executor's worker of `job_id` is done: stop counting it in the global unit barrier."""
        if self.enabled:
            self._ex.send_to_driver({"type": "TaskUnitLeave", "job_id": job_id,
                                     "executor_id": self._ex.executor_id})

    def join(self, job_id: str) -> None:
        """The worker takes part in the job's unit barriers again (it had left while it held no data)."""
        if self.enabled:
            self._ex.send_to_driver({"type": "TaskUnitJoin", "job_id": job_id,
                                     "executor_id": self._ex.executor_id})

    def on_task_unit_finished(self, tasklet_id: str, unit: TaskUnitInfo) -> None:
        if not self.enabled or unit.resource == VOID:
            return
        self._sems[unit.resource].release()
