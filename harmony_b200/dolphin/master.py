"""Master side of Dolphin (one per job, hosted in the job-server process).

Reference (jobserver/src/main/java/edu/snu/cay/dolphin/core/master/):
  WorkerStateManager.java:94-346 (global barrier FSM INIT -> RUN <-> OPTIMIZE, RUN -> RUN_FINISHING ->
  CLEANUP; late-added workers skip the INIT barrier), MiniBatchController.java:51-117 (SSP clock over a
  GLOBAL mini-batch budget = epochs x num_mini_batches), BatchProgressTracker.java:84-105 ("Epoch
  progress: [e / E]" lines; per-epoch model checkpoint when -offline_model_eval),
  ModelChkpManager.java:114-272, ETTaskRunner.java:82-189, DolphinMaster.java:108-220,
  MasterSideMsgHandler.java:57-75 / MasterSideMsgSender.java:66-111.
"""
from __future__ import annotations

import logging
import threading
import time
from typing import Any, Callable, Dict, List, Optional, Sequence, Set

from ..et.config import TaskletConfiguration
from ..et.driver import AllocatedExecutor, ETMaster, MetricReceiver, RunningTasklet
from ..utils.concurrent import CountDownLatch
from ..utils.state_machine import StateMachine
from .metrics import DolphinMetricManager, ServerMetrics

LOG = logging.getLogger(__name__)
WORKER_TASKLET = "harmony_b200.dolphin.worker:WorkerTasklet"
SERVER_TASKLET = "harmony_b200.dolphin.worker:ServerTasklet"
EVAL_TASKLET = "harmony_b200.dolphin.worker:ModelEvaluationTasklet"


class MasterSideMsgSender:
    """Master -> worker Dolphin messages (release, mini-batch control, model-eval answers) wrapped as tasklet
    custom messages.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/core/master/MasterSideMsgSender.java:66-111."""

    def __init__(self, task_runner: "ETTaskRunner") -> None:
        self._runner = task_runner

    def _send(self, worker_id: str, body: Dict[str, Any], inline: bool = False) -> None:
        rt = self._runner.get_worker_tasklet(worker_id)
        if rt is None:
            # a worker started by a plan op talks to the master before the optimizer has handed its tasklet over:
            # keep the message until the tasklet is registered
            self._runner.defer(worker_id, body)
            return
        if inline:
            rt.send(body, inline=True)           # barrier releases: handled on the executor's inbox thread
        else:
            rt.send(body)

    def send_release_msg(self, worker_id: str) -> None:
        self._send(worker_id, {"type": "ReleaseMsg"}, inline=True)

    def send_mini_batch_control_msg(self, worker_id: str, stop: bool, granted: int = 1) -> None:
        body = {"type": "MiniBatchControlMsg", "stop": stop}
        if granted != 1:
            body["granted"] = granted            # a lease: the worker runs this many mini-batches before its next sync
        self._send(worker_id, body, inline=True)

    def send_model_eval_ans_msg(self, worker_id: str, do_next: bool) -> None:
        self._send(worker_id, {"type": "ModelEvalAnsMsg", "do_next": do_next})


class WorkerStateManager:
    """Global worker barrier FSM INIT -> RUN <-> OPTIMIZE -> RUN_FINISHING -> CLEANUP: workers send SyncMsg(state),
    the master releases all of them once every running worker is blocked; workers added or deleted by a plan are
    accounted for.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/core/master/WorkerStateManager.java:94-116,152-199,204-262,269-346."""

    def __init__(self, num_workers: int, sender: MasterSideMsgSender, log: Callable[[str], None] = LOG.info) -> None:
        self.num_workers = num_workers
        self._sender = sender
        self._log = log
        self.sm = (StateMachine.new_builder()
                   .add_state("INIT").add_state("RUN").add_state("OPTIMIZE").add_state("RUN_FINISHING")
                   .add_state("CLEANUP")
                   .add_transition("INIT", "RUN").add_transition("RUN", "OPTIMIZE").add_transition("OPTIMIZE", "RUN")
                   .add_transition("RUN", "RUN_FINISHING").add_transition("RUN_FINISHING", "CLEANUP")
                   .set_initial_state("INIT").build())
        self.running: Set[str] = set()
        self.blocked: Set[str] = set()
        self._lock = threading.RLock()
        self.finish_init_latch = CountDownLatch(1)
        self.finish_run_latch = CountDownLatch(1)

    def wait_workers_to_finish_init_stage(self, timeout: Optional[float] = None) -> bool:
        return self.finish_init_latch.await_(timeout)

    def wait_workers_to_finish_run_stage(self, timeout: Optional[float] = None) -> bool:
        return self.finish_run_latch.await_(timeout)

    def try_enter_optimization(self) -> bool:
        with self._lock:
            if self.sm.current_state != "RUN":
                return False
            self.sm.set_state("OPTIMIZE")
            return True

    def on_optimization_finished(self, added: Set[str], deleted: Set[str]) -> None:
        with self._lock:
            self.sm.check_state("OPTIMIZE")
            if not deleted <= self.running or added & deleted or added & self.running:
                raise RuntimeError("The change by optimization is invalid")
            self.running |= added
            self.running -= deleted
            self.blocked -= deleted
            self.sm.set_state("RUN")
            if self.blocked:
                self._transit()          # someone already finished RUN: go to RUN_FINISHING
            self._try_releasing()

    def _transit(self) -> None:
        st = self.sm.current_state
        if st == "INIT":
            self.finish_init_latch.count_down()
            self.sm.set_state("RUN")
        elif st == "RUN":
            self.finish_run_latch.count_down()
            self.sm.set_state("RUN_FINISHING")
        elif st == "RUN_FINISHING":
            self.sm.set_state("CLEANUP")
        else:
            raise RuntimeError(f"No more transition is allowed after {st} state")
        self._log(f"State transition: {st} -> {self.sm.current_state}")

    def _release(self) -> None:
        for w in list(self.blocked):
            self._sender.send_release_msg(w)
        self.blocked.clear()

    def _try_releasing(self) -> None:
        if self.running and self.blocked >= self.running:
            self._transit()
            self._release()

    def on_sync_msg(self, worker_id: str, local_state: str) -> None:
        with self._lock:
            g = self.sm.current_state
            if g == "INIT":
                if local_state != "INIT":
                    raise RuntimeError(f"Worker {worker_id} is in invalid state: {local_state}")
                self.blocked.add(worker_id)
                self.running.add(worker_id)
                if len(self.running) == self.num_workers:
                    self._transit()
                    self._release()
            elif g == "RUN":
                if local_state == "RUN":
                    self._transit()
                    self.blocked.add(worker_id)
                    self._try_releasing()
                elif local_state == "INIT":
                    self._sender.send_release_msg(worker_id)     # late-added worker skips the INIT barrier
                else:
                    raise RuntimeError(f"Worker {worker_id} is in invalid state: {local_state}")
            elif g == "RUN_FINISHING":
                if local_state == "RUN":
                    self.blocked.add(worker_id)
                    self._try_releasing()
                elif local_state == "INIT":
                    self._sender.send_release_msg(worker_id)
                else:
                    raise RuntimeError(f"Worker {worker_id} is in invalid state: {local_state}")
            elif g == "OPTIMIZE":
                if local_state == "RUN":
                    self.blocked.add(worker_id)
                elif local_state == "INIT":
                    self._sender.send_release_msg(worker_id)
                else:
                    raise RuntimeError(f"Worker {worker_id} is in invalid state: {local_state}")
            else:
                raise RuntimeError(f"Workers should not send message in {g} state")


class MiniBatchController:
    """SSP clock: a worker is blocked while its progress > min progress + slack; everyone is told
    to stop when the GLOBAL budget epochs x num_mini_batches is used up.

    A sync may ask for a lease of `want` mini-batches (-mini_batch_grant_size): the grant is counted against the budget
    and the clock at once, tapers to single batches as the budget runs out (so the last batches still spread over the
    workers), and the SSP check applies to the FIRST batch of the lease -- the staleness bound widens by want - 1."""

    def __init__(self, num_epochs: int, num_mini_batches_in_epoch: int, slack: int,
                 sender: MasterSideMsgSender) -> None:
        self.slack = slack
        self._sender = sender
        self.total = num_epochs * num_mini_batches_in_epoch
        self._count = 0
        self._progress: Dict[str, int] = {}
        self._blocked: Set[str] = set()
        self._min = 0
        self._closed = False
        self._lock = threading.Lock()
        self._grant: Dict[str, int] = {}          # size of each worker's latest grant

    def _reply(self, worker_id: str, stop: bool) -> None:
        g = self._grant.get(worker_id, 1)
        if g == 1 or stop:
            self._sender.send_mini_batch_control_msg(worker_id, stop)
        else:
            self._sender.send_mini_batch_control_msg(worker_id, stop, g)

    def _release_blocked(self, everyone: bool = False) -> None:
        """Release the blocked requests whose first mini-batch is within the SSP bound again (with single-batch grants
        that is all of them whenever the clock advances); `everyone`: the budget is used up."""
        for w in list(self._blocked):
            start = self._progress.get(w, 0) - self._grant.get(w, 1) + 1
            if everyone or start <= self._min + self.slack:
                self._blocked.discard(w)
                self._reply(w, False)

    def on_idle_sync(self, worker_id: str) -> None:
        """A worker that currently holds no input data (all its blocks migrated away) asks whether to stop: it
        neither consumes the global budget nor holds the SSP clock back while it idles."""
        with self._lock:
            self._progress.pop(worker_id, None)
            self._blocked.discard(worker_id)
            if self._progress:
                new_min = min(self._progress.values())
                if new_min > self._min:
                    self._min = new_min
                    self._release_blocked()
            self._sender.send_mini_batch_control_msg(worker_id, self._closed)

    def on_sync(self, worker_id: str, want: int = 1) -> None:
        with self._lock:
            if self._closed:
                self._sender.send_mini_batch_control_msg(worker_id, True)
                return
            g = 1
            if want > 1:
                # taper: never lease more than half of an even share of what is left
                share = (self.total - self._count) // (2 * max(1, len(self._progress)))
                g = max(1, min(want, share))
            self._grant[worker_id] = g
            self._count += g
            # a worker that joins mid-run (added by a plan, or back from idling) starts level with the slowest
            # worker (its grant ends where the slowest one's ends) instead of at 0, which would stall everybody else
            # until it has caught up alone
            prog = self._progress.get(worker_id, max(self._min - g, 0)) + g
            self._progress[worker_id] = prog
            new_min = min(self._progress.values())
            if new_min > self._min:
                self._min = new_min
                self._release_blocked()
            if self._count + 1 > self.total:
                self._closed = True     # let everyone run their last mini-batch(es), then stop
                self._reply(worker_id, False)
                self._release_blocked(everyone=True)
            elif prog - g + 1 > self._min + self.slack:
                self._blocked.add(worker_id)
            else:
                self._reply(worker_id, False)

    def on_return(self, worker_id: str, count: Optional[int] = None) -> None:
        """A stopped worker hands back mini-batches it was granted but did not run (`count` None: its whole latest
        grant -- it was stopped while waiting for the reply and never saw the size)."""
        with self._lock:
            n = self._grant.get(worker_id, 1) if count is None else count
            if n > 0:
                self._count = max(0, self._count - n)
                # handed back after the budget had been used up (the stopped worker had asked for the very last
                # batches): reopen, so that a worker that has not been told to stop yet runs them
                if self._closed and self._count < self.total:
                    self._closed = False

    def on_worker_deleted(self, worker_id: str) -> None:
        with self._lock:
            self._progress.pop(worker_id, None)
            self._blocked.discard(worker_id)
            if self._progress:               # the clock no longer waits for the deleted worker
                new_min = min(self._progress.values())
                if new_min > self._min:
                    self._min = new_min
                    self._release_blocked()

    @property
    def num_mini_batches_done(self) -> int:
        with self._lock:
            return self._count


class BatchProgressTracker:
    """Counts ProgressMsgs, prints `Epoch progress: [e / E]` for the client and triggers the per-epoch model
    checkpoint when -offline_model_eval is on.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/core/master/BatchProgressTracker.java:84-105."""

    def __init__(self, job_id: str, num_epochs: int, num_mini_batches_in_epoch: int,
                 report: Callable[[str], None], on_epoch: Optional[Callable[[int], None]] = None) -> None:
        self.job_id, self.num_epochs, self.per_epoch = job_id, num_epochs, max(1, num_mini_batches_in_epoch)
        self.total = num_epochs * self.per_epoch
        self._report, self._on_epoch = report, on_epoch
        self._count = 0
        self._per_worker: Dict[str, int] = {}
        self._lock = threading.Lock()

    def on_progress_msg(self, worker_id: str, progress: int, count: int = 1) -> None:
        """`count` > 1: one report for the mini-batches of a lease (`progress` is the index of the last one)."""
        with self._lock:
            self._per_worker[worker_id] = progress
            first = self._count + 1
            self._count += max(1, count)
            last = self._count
        for c in range(first, last + 1):
            if c % self.per_epoch == 0 and c <= self.total:
                e = c // self.per_epoch
                self._report(f"Epoch progress: [{e} / {self.num_epochs}], JobId: {self.job_id}")
                if self._on_epoch is not None:
                    self._on_epoch(e - 1)

    def get_progress(self) -> float:
        with self._lock:
            return self._count / self.total if self.total else 1.0   # (the reference truncates to 0)


class ModelChkpManager:
    """Per-epoch checkpoints of the model (+ local model) tables; replayed oldest-first for the
    deferred (offline) model evaluation."""

    def __init__(self, master: ETMaster, model_table_id: str, local_model_table_id: Optional[str],
                 input_table_id: str) -> None:
        self._m = master
        self.model_table_id, self.local_model_table_id, self.input_table_id = \
            model_table_id, local_model_table_id, input_table_id
        self.model_chkps: List[str] = []
        self.local_chkps: List[str] = []
        self.input_chkp: Optional[str] = None
        self._lock = threading.Lock()
        # ONE thread takes the per-epoch checkpoints: they complete -- and are recorded -- in issue order, i.e.
        # chronologically (ModelChkpManager.checkpointIdFutures keeps the futures in issue order for the same
        # reason), and a slow checkpoint never parks a thread of the driver pool that SyncMsg & co. need.
        from concurrent.futures import ThreadPoolExecutor
        self._exec = ThreadPoolExecutor(max_workers=1, thread_name_prefix="model-chkp")
        self._pending: List[Any] = []

    def create_checkpoint_async(self):
        """Queue the checkpoint of the current epoch (returns a concurrent.futures.Future)."""
        f = self._exec.submit(self.create_checkpoint)
        with self._lock:
            self._pending.append(f)
        return f

    def wait_pending(self, timeout: float = 600.0) -> None:
        with self._lock:
            pending, self._pending = self._pending, []
        for f in pending:
            try:
                f.result(timeout=timeout)
            except Exception as e:  # noqa: BLE001
                LOG.warning("a per-epoch model checkpoint failed: %s", e)

    def create_checkpoint(self) -> None:
        mid = self._m.get_table(self.model_table_id).checkpoint().get(timeout=600)
        lid = None
        if self.local_model_table_id:
            lid = self._m.get_table(self.local_model_table_id).checkpoint().get(timeout=600)
        with self._lock:
            self.model_chkps.append(mid)
            if lid:
                self.local_chkps.append(lid)

    def checkpoint_input(self) -> None:
        self.input_chkp = self._m.get_table(self.input_table_id).checkpoint().get(timeout=600)

    def num_chkps(self) -> int:
        with self._lock:
            return len(self.model_chkps)


class ETTaskRunner:
    """Submits the server tasklets (idle latch) and the worker tasklets, waits for the run stage, collects the
    results and follows executor-set changes made by the optimizer; a failed tasklet fails the job fast.
    Reference: jobserver/src/main/java/edu/snu/cay/dolphin/core/master/ETTaskRunner.java:82-106,121-172,174-189."""

    def __init__(self, master: ETMaster, job_id: str) -> None:
        self._m = master
        self.job_id = job_id
        self.worker_tasklets: Dict[str, RunningTasklet] = {}
        self.server_tasklets: Dict[str, RunningTasklet] = {}
        self._deferred: Dict[str, List[Dict[str, Any]]] = {}     # master -> worker messages awaiting registration
        self._ever_started: Dict[str, RunningTasklet] = {}
        self._lock = threading.Lock()
        self.msg_handler: Optional[Callable[[Any], None]] = None

    def get_worker_tasklet(self, worker_id: str) -> Optional[RunningTasklet]:
        with self._lock:
            return self.worker_tasklets.get(worker_id)

    def defer(self, worker_id: str, body: Dict[str, Any]) -> None:
        with self._lock:
            rt = self.worker_tasklets.get(worker_id)
            if rt is None:
                self._deferred.setdefault(worker_id, []).append(body)
                return
        rt.send(body)

    def register_worker(self, rt: RunningTasklet) -> None:
        """A worker tasklet started outside `submit_worker` (plan START op): route master messages to it, including
        the ones that were addressed to it before it was known."""
        with self._lock:
            self.worker_tasklets[rt.executor_id] = rt
            self._ever_started.setdefault(rt.tasklet_id + "@" + rt.executor_id + str(id(rt)), rt)
            pending = self._deferred.pop(rt.executor_id, [])
        for body in pending:
            rt.send(body)

    def submit_worker(self, ex: AllocatedExecutor, params: Dict[str, Any], cls: str = WORKER_TASKLET,
                      suffix: str = "worker") -> RunningTasklet:
        conf = TaskletConfiguration(f"{self.job_id}-{suffix}-{ex.id}", cls, None, params)
        rt = ex.submit_tasklet(conf, self.msg_handler)
        with self._lock:
            self.worker_tasklets[ex.id] = rt
        return rt

    def submit_server(self, ex: AllocatedExecutor) -> RunningTasklet:
        conf = TaskletConfiguration(f"{self.job_id}-server-{ex.id}", SERVER_TASKLET, None, {"job_id": self.job_id})
        rt = ex.submit_tasklet(conf)
        with self._lock:
            self.server_tasklets[ex.id] = rt
        return rt

    def update_executor_entry(self, added_workers: Dict[str, RunningTasklet], deleted_workers: Set[str],
                              added_servers: Dict[str, RunningTasklet], deleted_servers: Set[str]) -> None:
        for rt in added_workers.values():
            self.register_worker(rt)
        with self._lock:
            for w in deleted_workers:
                if w not in added_workers:
                    self.worker_tasklets.pop(w, None)
                self._deferred.pop(w, None)
            self.server_tasklets.update(added_servers)
            for s in deleted_servers:
                self.server_tasklets.pop(s, None)

    def wait_and_get_results(self, timeout: Optional[float] = None, on_failure=None):
        """Waits for every worker; when one FAILS the others are aborted (they would otherwise wait
        for it at a barrier forever) and the job fails fast."""
        failed = threading.Event()
        watched: Dict[str, RunningTasklet] = {}

        def snapshot() -> List[RunningTasklet]:
            """Every worker tasklet the job has had so far: the initial ones, the ones plans started later, and the
            ones plans stopped (they end `Done` with a partial result)."""
            with self._lock:
                for rt in self.worker_tasklets.values():
                    self._ever_started.setdefault(rt.tasklet_id + "@" + rt.executor_id + str(id(rt)), rt)
                current = list(self._ever_started.items())
            for key, rt in current:
                if key not in watched:
                    watched[key] = rt
                    rt._result.add_listener(lambda r: failed.set() if (r is not None and not r.is_success) else None)
            return [rt for _, rt in current]
        deadline = None if timeout is None else time.time() + timeout
        workers = snapshot()
        while True:
            workers = snapshot()
            if all(w.is_done() for w in workers):
                break
            if failed.wait(0.05):
                if on_failure is not None:
                    on_failure()
                for w in workers:
                    if not w.is_done():
                        try:
                            w.stop()
                        except Exception:  # noqa: BLE001
                            pass
                break
            if deadline is not None and time.time() > deadline:
                raise TimeoutError("workers did not finish in time")
        grace = 15.0 if failed.is_set() else timeout
        results = []
        for w in workers:
            try:
                results.append(w.get_task_result(grace))
            except TimeoutError:
                from ..et.driver import TaskletResult
                results.append(TaskletResult(w.tasklet_id, w.executor_id, "Failed", None, "aborted after a peer failed"))
        with self._lock:
            servers = list(self.server_tasklets.values())
        for s in servers:
            s.stop()
        for s in servers:
            s.get_task_result(timeout)
        return results


class ETDolphinMetricReceiver(MetricReceiver):
    """Converts ET metric reports (+ custom Batch/Epoch metrics) into Worker/ServerMetrics."""

    def __init__(self) -> None:
        self.jobs: Dict[str, DolphinMetricManager] = {}

    def register(self, job_id: str, mm: DolphinMetricManager) -> None:
        self.jobs[job_id] = mm

    def unregister(self, job_id: str) -> None:
        self.jobs.pop(job_id, None)

    def on_metric_report(self, report: Dict[str, Any]) -> None:
        ex = report["executor_id"]
        t2b = report.get("table_to_num_blocks", {})
        for cm in report.get("custom_metrics", []):
            mm = self.jobs.get(cm.get("job_id"))
            if mm is None:
                continue
            cm = dict(cm)
            cm["hostname"] = report.get("hostname")
            cm["bytes_received_get_resp"] = report.get("bytes_received_get_resp", 0)
            cm["total_pull_count"] = report.get("count_sent_get_req", 0)
            mm.store_worker_metric(ex, sum(t2b.values()), cm)
        for job_id, mm in list(self.jobs.items()):
            model_table = getattr(mm, "model_table_id", None)
            if model_table in t2b:
                mm.store_server_metric(ServerMetrics(ex, report.get("hostname", ""), t2b[model_table]))


class DolphinMaster:
    """Builds the tasklet confs, runs the workers/servers, routes worker messages."""

    def __init__(self, master: ETMaster, job_id: str, params: Dict[str, Any],
                 report: Callable[[str], None] = LOG.info) -> None:
        self._m = master
        self.job_id = job_id
        self.params = dict(params)
        self.report = report
        self.task_runner = ETTaskRunner(master, job_id)
        self.task_runner.msg_handler = self.on_worker_msg
        self.sender = MasterSideMsgSender(self.task_runner)
        from .group import MasterSideRendezvous
        self.rendezvous = MasterSideRendezvous(lambda w, body: self.sender._send(w, body))
        self.metric_manager = DolphinMetricManager(int(params.get("num_initial_batch_metrics_to_skip", 0)))
        self.metric_manager.model_table_id = params.get("model_table_id")
        self.state_manager: Optional[WorkerStateManager] = None
        self.mini_batch_controller: Optional[MiniBatchController] = None
        self.progress_tracker: Optional[BatchProgressTracker] = None
        self.chkp_manager = ModelChkpManager(master, params["model_table_id"], params.get("local_model_table_id"),
                                             params["input_table_id"])
        self._eval_lock = threading.Lock()
        self._eval_asked: Set[str] = set()
        self._eval_iter: Optional[Any] = None
        self._eval_workers = 0

    # ---- worker -> master ----
    def on_worker_msg(self, payload: Any) -> None:
        body = payload["dolphin"] if "dolphin" in payload else payload
        t = body["type"]
        if t == "SyncMsg":
            self.state_manager.on_sync_msg(body["executor_id"], body["state"])
        elif t == "MiniBatchSyncMsg":
            if body.get("progress") is not None:           # piggybacked ProgressMsg of the previous mini-batch(es)
                self.progress_tracker.on_progress_msg(body["worker_id"], body["progress"],
                                                      int(body.get("progress_count", 1)))
            if body.get("idle"):
                self.mini_batch_controller.on_idle_sync(body["worker_id"])
            elif int(body.get("want", 1)) > 1:
                self.mini_batch_controller.on_sync(body["worker_id"], int(body["want"]))
            else:
                self.mini_batch_controller.on_sync(body["worker_id"])
        elif t == "MiniBatchReturnMsg":
            if "count" in body:
                self.mini_batch_controller.on_return(body["worker_id"], int(body["count"]))
            else:
                self.mini_batch_controller.on_return(body["worker_id"])
        elif t == "ProgressMsg":
            self.progress_tracker.on_progress_msg(body["executor_id"], body["progress"], int(body.get("count", 1)))
        elif t == "ModelEvalAskMsg":
            self._on_model_eval_ask(body["executor_id"])
        elif t == "GroupGather":
            self.rendezvous.on_msg(body)
        else:
            LOG.warning("unknown dolphin message %s", t)

    def worker_params(self) -> Dict[str, Any]:
        p = dict(self.params)
        p["job_id"] = self.job_id
        return p

    # ---- train ----
    def start(self, servers: Sequence[AllocatedExecutor], workers: Sequence[AllocatedExecutor]) -> List[Any]:
        p = self.params
        epochs, nmb = int(p.get("max_num_epochs", 1)), int(p.get("num_mini_batches", 10))
        self.state_manager = WorkerStateManager(len(workers), self.sender,
                                                lambda s: LOG.info("[%s] %s", self.job_id, s))
        self.mini_batch_controller = MiniBatchController(epochs, nmb, int(p.get("clock_slack", 0)), self.sender)
        on_epoch = None
        if p.get("offline_model_eval"):
            on_epoch = lambda e: self.chkp_manager.create_checkpoint_async()  # noqa: E731  (issue order == epoch order)
        self.progress_tracker = BatchProgressTracker(self.job_id, epochs, nmb, self.report, on_epoch)
        self.metric_manager.start_metric_collection()
        for ex in workers:
            self._m.metrics.start_metric_collection(ex.id, int(p.get("server_metric_flush_period_ms", 0)))
        for ex in servers:
            self.task_runner.submit_server(ex)
        for ex in workers:
            self.task_runner.submit_worker(ex, self.worker_params())
        def release_all():       # unblock every barrier a surviving worker may be parked in
            for ex in workers:
                try:
                    self.sender.send_mini_batch_control_msg(ex.id, True)
                    self.sender.send_release_msg(ex.id)
                    self.sender.send_release_msg(ex.id)
                except Exception:  # noqa: BLE001
                    pass
        results = self.task_runner.wait_and_get_results(timeout=float(p.get("job_timeout_sec", 3600)),
                                                        on_failure=release_all)
        for ex in workers:
            try:
                self._m.metrics.stop_metric_collection(ex.id)
            except Exception:  # noqa: BLE001
                pass
        self.metric_manager.stop_metric_collection()
        failed = [r for r in results if not r.is_success]
        if failed:
            raise RuntimeError(f"job {self.job_id}: worker tasklets failed: " +
                               "; ".join(f"{r.executor_id}: {r.error}" for r in failed))
        if p.get("offline_model_eval"):
            # the per-epoch checkpoints run on the driver pool: let the last ones finish before anybody evaluates
            # (or drops the tables they are reading)
            self.chkp_manager.wait_pending()
            self.chkp_manager.checkpoint_input()
        return [r.result for r in results]

    # ---- offline model evaluation ----
    def evaluate(self, workers: Sequence[AllocatedExecutor], restore_fn: Callable[[int], bool]) -> List[Any]:
        """`restore_fn(i)` restores the i-th checkpoint into the eval tables; False when exhausted."""
        self._eval_workers = len(workers)
        self._eval_idx = 0
        self._restore_fn = restore_fn
        rts = [self.task_runner.submit_worker(ex, self.worker_params(), EVAL_TASKLET, "eval") for ex in workers]
        results = [rt.get_task_result(timeout=3600) for rt in rts]
        return [r.result for r in results]

    def _on_model_eval_ask(self, worker_id: str) -> None:
        with self._eval_lock:
            self._eval_asked.add(worker_id)
            if len(self._eval_asked) < self._eval_workers:
                return
            asked, self._eval_asked = self._eval_asked, set()
            do_next = self._restore_fn(self._eval_idx)
            self._eval_idx += 1
            if do_next:
                self.report(f"Model eval progress: [{self._eval_idx}], JobId: {self.job_id}")
        for w in asked:
            self.sender.send_model_eval_ans_msg(w, do_next)
