"""Job-server driver: job entity / master SPI, pluggable scheduler, resource pool, dispatcher.

Reference (jobserver/src/main/java/edu/snu/cay/jobserver/driver/):
  JobServerDriver.java:80-299 (FSM NOT_INIT -> INIT -> CLOSED; SUBMIT -> JobEntity -> scheduler ->
  "Accept/Reject. JobId"; SHUTDOWN(kill?) waits for running jobs, runs deferred model evaluations,
  closes the pool), JobScheduler.java:22-44 + SchedulerImpl.java:42-66 (FIFO-immediate: every job
  gets ALL executors, never rejects), ResourcePool.java:48-105, JobDispatcher.java:59-84,
  JobEntity.java:31-64, JobMaster.java:26-40, DriverSideMsgHandler.java:40-47, JobLogger.java.

Additional schedulers for the 8-GPU box: `PackingScheduler` (gives a job the least-loaded GPUs it
asks for) and `FifoExclusiveScheduler` (queues jobs when the pool is busy).
"""
from __future__ import annotations

import importlib
import collections
import itertools
import logging
import threading
import time
import traceback
from typing import Any, Callable, Dict, List, Optional, Sequence, Set

from ..et.config import ExecutorConfiguration, RemoteAccessConfiguration, ResourceConfiguration
from ..et.driver import AllocatedExecutor, ETMaster
from ..utils.flags import Configuration
from ..utils.state_machine import StateMachine

LOG = logging.getLogger(__name__)


class JobLogger:
    """Prefixes records with [JobId: x] (JobLogger.java:38-72)."""

    def __init__(self, job_id: str, sink: Optional[Callable[[str], None]] = None) -> None:
        self.job_id, self._sink = job_id, sink

    def log(self, msg: str, *args: Any) -> None:
        text = f"[JobId: {self.job_id}] " + (msg % args if args else msg)
        LOG.info(text)
        if self._sink is not None:
            self._sink(text)


class JobMaster:
    """Per-job master SPI: `on_msg` receives worker messages, `start` runs the job (blocking)."""

    def on_msg(self, source_id: str, payload: Any) -> None:
        raise NotImplementedError

    def start(self) -> Any:
        raise NotImplementedError


class JobEntity:
    """Immutable description of a submitted job."""
    job_id: str

    def get_job_id(self) -> str:
        return self.job_id

    def get_job_master(self) -> JobMaster:
        raise NotImplementedError

    def setup_executors_and_tables(self, executors: Sequence[AllocatedExecutor]) -> None:
        raise NotImplementedError

    def cleanup(self) -> None:
        pass

    def num_executors_wanted(self) -> int:
        return 0   # 0 => all


class JobScheduler:
    """Pluggable global scheduling policy."""

    exclusive = False      # True: a running job never shares executors with another one

    def bind(self, driver: "JobServerDriver") -> None:
        self.driver = driver

    def on_job_arrival(self, job: JobEntity) -> bool:
        raise NotImplementedError

    def on_job_finish(self, num_released: int, job: Optional[JobEntity] = None) -> None:
        raise NotImplementedError

    def on_resource_change(self, delta: int) -> None:
        raise NotImplementedError("resource change is not supported by this scheduler")


class SchedulerImpl(JobScheduler):
    """FIFO-immediate, every job shares the WHOLE pool (the reference default)."""

    def on_job_arrival(self, job: JobEntity) -> bool:
        self.driver.dispatcher.execute_job(job, self.driver.resource_pool.get_executors())
        return True

    def on_job_finish(self, num_released: int, job: Optional[JobEntity] = None) -> None:
        pass


class FifoExclusiveScheduler(JobScheduler):
    """One job at a time on the whole pool; later arrivals queue."""

    exclusive = True

    def __init__(self) -> None:
        self._queue: List[JobEntity] = []
        self._busy = False
        self._lock = threading.Lock()

    def on_job_arrival(self, job: JobEntity) -> bool:
        with self._lock:
            if self._busy:
                self._queue.append(job)
                return True
            self._busy = True
        self.driver.dispatcher.execute_job(job, self.driver.resource_pool.get_executors())
        return True

    def on_job_finish(self, num_released: int, job: Optional[JobEntity] = None) -> None:
        with self._lock:
            nxt = self._queue.pop(0) if self._queue else None
            self._busy = nxt is not None
        if nxt is not None:
            self.driver.dispatcher.execute_job(nxt, self.driver.resource_pool.get_executors())

    def on_resource_change(self, delta: int) -> None:
        pass


class PackingScheduler(JobScheduler):
    """Packs jobs onto the least-loaded executors: a job asking for k executors (0 = all) gets the k
    executors currently running the fewest jobs; phases of co-located jobs interleave through the
    task-unit scheduler."""

    def __init__(self) -> None:
        self._load: Dict[str, int] = {}
        self._assigned: Dict[str, List[str]] = {}
        self._lock = threading.Lock()

    def on_job_arrival(self, job: JobEntity) -> bool:
        pool = self.driver.resource_pool.get_executors()
        k = job.num_executors_wanted() or len(pool)
        if k > len(pool):
            return False
        with self._lock:
            ranked = sorted(pool, key=lambda e: (self._load.get(e.id, 0), e.id))[:k]
            for e in ranked:
                self._load[e.id] = self._load.get(e.id, 0) + 1
            self._assigned[job.get_job_id()] = [e.id for e in ranked]
        self.driver.dispatcher.execute_job(job, ranked)
        return True

    def on_job_finish(self, num_released: int, job: Optional[JobEntity] = None) -> None:
        if job is None:
            return
        with self._lock:
            for e in self._assigned.pop(job.get_job_id(), []):
                self._load[e] = max(0, self._load.get(e, 0) - 1)

    def on_resource_change(self, delta: int) -> None:
        pass


_SCHEDULERS = {
    "edu.snu.cay.jobserver.driver.SchedulerImpl": SchedulerImpl, "SchedulerImpl": SchedulerImpl,
    "FifoExclusiveScheduler": FifoExclusiveScheduler, "PackingScheduler": PackingScheduler,
}


def make_scheduler(name: str) -> JobScheduler:
    if name in _SCHEDULERS:
        return _SCHEDULERS[name]()
    mod, _, cls = name.partition(":")
    return getattr(importlib.import_module(mod), cls)()


class ResourcePool:
    """Fixed pool of `num_executors` ET executors created when the job server starts."""

    def __init__(self, master: ETMaster, conf: Configuration) -> None:
        self._m = master
        self._conf = conf
        self._executors: Dict[str, AllocatedExecutor] = {}

    def init(self) -> None:
        c = self._conf
        ec = ExecutorConfiguration(
            resource=ResourceConfiguration(int(c.get("executor_num_cores", 1)), int(c.get("executor_mem_size", 128))),
            remote_access=RemoteAccessConfiguration(int(c.get("handler_queue_size", 1024)),
                                                    int(c.get("sender_queue_size", 1024)),
                                                    int(c.get("handler_num_threads", 2)),
                                                    int(c.get("sender_num_threads", 2))),
            num_tasklets=max(4, int(c.get("executor_num_tasklets", 4))),
            chkp_temp_path=c.get("chkp_temp_path", "/tmp/chkp"),
            chkp_commit_path=c.get("chkp_commit_path", "/tmp/chkp/commit"))
        execs = self._m.add_executors(int(c.get("num_executors", 1)), ec).get(timeout=300)
        for e in execs:
            self._executors[e.id] = e

    def get_executors(self) -> List[AllocatedExecutor]:
        return [self._executors[k] for k in sorted(self._executors)]

    def add(self, e: AllocatedExecutor) -> None:
        self._executors[e.id] = e

    def remove(self, executor_id: str) -> None:
        self._executors.pop(executor_id, None)

    def close(self) -> None:
        for e in list(self._executors.values()):
            try:
                e.close().get(timeout=30)
            except Exception:  # noqa: BLE001
                pass
        self._executors.clear()


class JobDispatcher:
    """One thread per job: set up tables -> register master -> run -> clean up -> finish hook."""

    def __init__(self, driver: "JobServerDriver") -> None:
        self._d = driver
        # admission: a job starts only when EVERY executor it runs on has a free tasklet slot.  The slots are taken
        # all-or-nothing -- two jobs that each hold a slot on one executor and wait for the other's (their workers
        # parked at the INIT barrier) would deadlock the pool.
        self._slots_cv = threading.Condition()
        self._jobs_on: Dict[str, int] = {}
        self.slots_per_executor = max(4, int(driver.conf.get("executor_num_tasklets", 4)))

    def _admit(self, executors: Sequence[AllocatedExecutor], job_id: str = "") -> None:
        ids = [e.id for e in executors]
        free = lambda: all(self._jobs_on.get(i, 0) < self.slots_per_executor for i in ids)  # noqa: E731
        with self._slots_cv:
            if not free() and job_id:
                self._d.send_to_client(f"Waiting for a free tasklet slot on every executor. JobId: {job_id}")
            self._slots_cv.wait_for(free)
            for i in ids:
                self._jobs_on[i] = self._jobs_on.get(i, 0) + 1

    def _release(self, executors: Sequence[AllocatedExecutor]) -> None:
        with self._slots_cv:
            for e in executors:
                self._jobs_on[e.id] = max(0, self._jobs_on.get(e.id, 0) - 1)
            self._slots_cv.notify_all()

    def execute_job(self, job: JobEntity, executors: Sequence[AllocatedExecutor]) -> None:
        self._d.mark_job_active(job.get_job_id())
        threading.Thread(target=self._run, args=(job, list(executors)), name=f"job-{job.get_job_id()}",
                         daemon=True).start()

    def _run(self, job: JobEntity, executors: List[AllocatedExecutor]) -> None:
        d = self._d
        job_id = job.get_job_id()
        self._admit(executors, job_id)
        d.send_to_client(f"Start executing a job. JobId: {job_id}")
        ok, result = False, None
        try:
            job.setup_executors_and_tables(executors)
            master = job.get_job_master()
            d.put_job_master(job_id, master)
            d.et_master.task_units.on_job_start(job_id, len(executors))
            result = master.start()
            ok = True
        except BaseException as e:  # noqa: BLE001
            LOG.error("job %s failed:\n%s", job_id, traceback.format_exc())
            result = e
        finally:
            d.et_master.task_units.on_job_finish(job_id)
            try:
                job.cleanup()
            except Exception:  # noqa: BLE001
                LOG.exception("cleanup of %s failed", job_id)
            self._release(executors)
            d.remove_job_master(job_id, ok, result)
            d.send_to_client(f"Job {'finished' if ok else 'FAILED'}. JobId: {job_id}")
            d.scheduler.on_job_finish(len(executors), job)


class JobServerDriver:
    """The long-running driver: FSM NOT_INIT -> INIT -> CLOSED, resource pool, SUBMIT / SHUTDOWN / STATUS handling,
    scheduler hook-up, job dispatch and bookkeeping.  Reference:
    jobserver/src/main/java/edu/snu/cay/jobserver/driver/JobServerDriver.java:80-109,125-145,178-214,222-266."""

    MAX_CLIENT_MESSAGES = 20000
    MAX_FINISHED_JOBS = 2000

    def __init__(self, et_master: ETMaster, conf: Configuration,
                 client_sink: Optional[Callable[[str], None]] = None) -> None:
        self.et_master = et_master
        self.conf = conf
        self.sm = (StateMachine.new_builder().add_state("NOT_INIT").add_state("INIT").add_state("CLOSED")
                   .add_transition("NOT_INIT", "INIT").add_transition("INIT", "CLOSED")
                   .set_initial_state("NOT_INIT").build())
        self.resource_pool = ResourcePool(et_master, conf)
        self.dispatcher = JobDispatcher(self)
        self.scheduler = make_scheduler(conf.get("scheduler", "SchedulerImpl"))
        self.scheduler.bind(self)
        self._job_masters: Dict[str, JobMaster] = {}
        self._active_jobs: Set[str] = set()
        self.job_results: Dict[str, Any] = {}
        self.job_finish_times: Dict[str, float] = {}
        self._job_counter = itertools.count()
        self._lock = threading.Condition()
        self._client_sink = client_sink
        # a server runs for weeks: progress lines and finished-job records are bounded histories, not logs
        self.client_messages: "collections.deque[str]" = collections.deque(maxlen=self.MAX_CLIENT_MESSAGES)
        self.deferred_evaluations: List[Callable[[], None]] = []
        self.job_entity_builders: Dict[str, Callable[..., JobEntity]] = {}
        et_master.custom_handlers["TaskletCustom"] = self._on_tasklet_custom

    # ---- lifecycle ----
    def init(self) -> None:
        self.sm.check_state("NOT_INIT")
        self.resource_pool.init()
        self.sm.set_state("INIT")
        self.send_to_client("Job Server is ready")

    def next_job_count(self) -> int:
        return next(self._job_counter)

    def send_to_client(self, msg: str) -> None:
        LOG.info("[client] %s", msg)
        self.client_messages.append(msg)
        if self._client_sink is not None:
            try:
                self._client_sink(msg)
            except Exception:  # noqa: BLE001
                pass

    # ---- routing of worker -> job master messages (DriverSideMsgHandler) ----
    def _on_tasklet_custom(self, m: Dict[str, Any]) -> None:
        payload = m.get("payload") or {}
        job_id = payload.get("job_id")
        with self._lock:
            jm = self._job_masters.get(job_id)
        if jm is None:
            LOG.warning("message for unknown job %s", job_id)
            return
        jm.on_msg(m.get("executor_id"), payload)

    def mark_job_active(self, job_id: str) -> None:
        with self._lock:
            self._active_jobs.add(job_id)

    def put_job_master(self, job_id: str, master: JobMaster) -> None:
        with self._lock:
            self._job_masters[job_id] = master

    def remove_job_master(self, job_id: str, ok: bool, result: Any) -> None:
        with self._lock:
            self._job_masters.pop(job_id, None)
            self._active_jobs.discard(job_id)
            self.job_results[job_id] = (ok, result)
            self.job_finish_times[job_id] = time.time()
            while len(self.job_results) > self.MAX_FINISHED_JOBS:      # dicts keep insertion order: oldest first
                old_id = next(iter(self.job_results))
                self.job_results.pop(old_id)
                self.job_finish_times.pop(old_id, None)
            self._lock.notify_all()

    def running_jobs(self) -> List[str]:
        with self._lock:
            return sorted(self._active_jobs)

    def wait_job(self, job_id: str, timeout: Optional[float] = None):
        with self._lock:
            if not self._lock.wait_for(lambda: job_id in self.job_results, timeout):
                raise TimeoutError(f"job {job_id} did not finish")
            return self.job_results[job_id]

    # ---- commands ----
    def on_submit(self, serialized_conf: str) -> str:
        if self.sm.current_state != "INIT":
            return "Reject. Job server is not ready"
        try:
            job_conf = Configuration.deserialize(serialized_conf)
            builder_path = job_conf["job_entity_builder"]
            mod, _, name = builder_path.partition(":")
            builder = getattr(importlib.import_module(mod), name)
            job = builder(self, job_conf).build()
        except Exception as e:  # noqa: BLE001
            LOG.error("cannot build job entity:\n%s", traceback.format_exc())
            self.send_to_client(f"Reject. {type(e).__name__}: {e}")
            return f"Reject. {e}"
        accepted = self.scheduler.on_job_arrival(job)
        msg = f"{'Accept' if accepted else 'Reject'}. JobId: {job.get_job_id()}"
        self.send_to_client(msg)
        return msg

    def on_shutdown(self, kill: bool) -> None:
        if not self.sm.compare_and_set_state("INIT", "CLOSED"):
            return
        if not kill:
            with self._lock:
                self._lock.wait_for(lambda: not self._active_jobs, timeout=3600)
            for fn in self.deferred_evaluations:
                try:
                    fn()
                except Exception:  # noqa: BLE001
                    LOG.exception("deferred model evaluation failed")
        self.resource_pool.close()
        self.send_to_client("Job Server is closed")
