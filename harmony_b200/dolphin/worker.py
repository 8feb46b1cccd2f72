"""Worker side of Dolphin: Trainer SPI, model accessor, training data provider, barriers and the
training loop.

Reference (jobserver/src/main/java/edu/snu/cay/dolphin/core/worker/):
  Trainer.java:44-92, ModelAccessor.java:29-83, ETModelAccessor.java:60-159 (push = updateNoReply /
  multiUpdate().get(), pull = getOrInit / multiGetOrInit().get(), pull/push timers),
  CachedModelAccessor.java:37-199, TrainingDataProvider.java:26-51, ETTrainingDataProvider.java:54-108
  (one local input block == one mini-batch, shuffled per epoch), WorkerGlobalBarrier.java:49-127,
  MiniBatchBarrier.java:47-64, WorkerTasklet.java:96-261 (SYNC -> PULL(net) -> COMP(cpu) -> PUSH(net)
  each wrapped in a task unit; Batch/Epoch metrics), ProgressReporter, ModelEvaluator.java:69-147,
  core/server/ServerTasklet.java:43-55.
"""
from __future__ import annotations

import logging
import random
import threading
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

from ..et.tasklet import CPU, NET, TaskUnitInfo, Tasklet, load_class
from ..utils import trace
from .metrics import BatchMetrics, EpochMetrics

LOG = logging.getLogger(__name__)


# ------------------------------------------------------------------------------------------
# SPI
# ------------------------------------------------------------------------------------------
class Trainer:
    """A machine-learning algorithm plugged into the PS training loop."""

    def init_global_settings(self) -> None:
        pass

    def set_mini_batch_data(self, mini_batch: Sequence[Tuple[Any, Any]]) -> None:
        raise NotImplementedError

    def pull_model(self) -> None:
        raise NotImplementedError

    def local_compute(self) -> None:
        raise NotImplementedError

    def push_update(self) -> None:
        raise NotImplementedError

    def on_epoch_finished(self, epoch_idx: int) -> None:
        pass

    def evaluate_model(self, input_data: Sequence[Tuple[Any, Any]], test_data: Sequence[Any]) -> Dict[str, float]:
        return {}

    def cleanup(self) -> None:
        pass

    def on_worker_stopped(self) -> None:
        """A plan removed this worker mid-training (deleted / switched to a server).  Unlike `cleanup()` this is not
        the end of the job: publish worker-local state that must survive (its blocks migrate next), validate nothing."""

    def summary(self) -> Dict[str, float]:
        """Optional end-of-training numbers returned with the tasklet result."""
        return {}


class ModelAccessor:
    """Trainer-facing pull / push SPI over the model table.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/core/worker/ModelAccessor.java:29-83."""

    def push(self, key, delta) -> None: raise NotImplementedError
    def push_many(self, keys, deltas) -> None: raise NotImplementedError
    def pull(self, key): raise NotImplementedError
    def pull_many(self, keys): raise NotImplementedError
    def get_and_reset_metrics(self) -> Dict[str, float]: return {}


class ETModelAccessor(ModelAccessor):
    """push(k, d) = updateNoReply; push_many = multiUpdate().get() (blocking, reply required);
    pull(k) = getOrInit; pull_many = multiGetOrInit().get(); wall-clock timers for both."""

    def __init__(self, table) -> None:
        self.table = table
        self._pull = trace.Tracer()
        self._push = trace.Tracer()

    def push(self, key, delta) -> None:
        self._push.start_timer()
        self.table.update_no_reply(key, delta)
        self._push.record_time(1)

    def push_many(self, keys, deltas) -> None:
        self._push.start_timer()
        self.table.multi_update(keys, deltas).get(timeout=600)
        self._push.record_time(len(keys))

    def pull(self, key):
        self._pull.start_timer()
        v = self.table.get_or_init(key).get(timeout=600)
        self._pull.record_time(1)
        return v

    def pull_many(self, keys):
        self._pull.start_timer()
        v = self.table.multi_get_or_init(keys).get(timeout=600)
        self._pull.record_time(len(keys))
        return v

    def get_and_reset_metrics(self) -> Dict[str, float]:
        out = {"pull_time_sec": self._pull.total_elapsed_time(), "push_time_sec": self._push.total_elapsed_time()}
        self._pull.reset_trace()
        self._push.reset_trace()
        return out


class CachedModelAccessor(ETModelAccessor):
    """Worker-side model cache refreshed every `refresh_sec`, write-through local apply
    (stand-alone `-model_cache_enabled`).  Unlike the reference (whose refresh task builds an
    empty key list) the refresh really re-pulls the cached keys."""

    def __init__(self, table, update_fn, refresh_sec: float = 10.0) -> None:
        super().__init__(table)
        self._fn = update_fn
        self._cache: Dict[Any, Any] = {}
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._refresh_loop, args=(refresh_sec,), daemon=True)
        self._thread.start()

    def _refresh_loop(self, period: float) -> None:
        while not self._stop.wait(period):
            with self._lock:
                keys = list(self._cache)
            if keys:
                try:
                    vals = self.table.multi_get_or_init(keys).get(timeout=600)
                except Exception as e:  # noqa: BLE001
                    # the table went away under a refresh (job teardown): the cache keeps its last values
                    LOG.debug("model cache refresh failed: %s", e)
                    continue
                with self._lock:
                    for k, v in zip(keys, vals):
                        self._cache[k] = v

    def pull(self, key):
        with self._lock:
            if key in self._cache:
                return self._cache[key]
        v = super().pull(key)
        with self._lock:
            self._cache[key] = v
        return v

    def pull_many(self, keys):
        return [self.pull(k) for k in keys]

    def push(self, key, delta) -> None:
        super().push(key, delta)
        with self._lock:
            if key in self._cache:
                self._cache[key] = self._fn.update_value(key, self._cache[key], delta)

    def push_many(self, keys, deltas) -> None:
        for k, d in zip(keys, deltas):
            self.push(k, d)

    def close(self) -> None:
        self._stop.set()


class TrainingDataProvider:
    """Mini-batch source SPI.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/core/worker/TrainingDataProvider.java:26-51."""

    def prepare_data_for_epoch(self) -> None: raise NotImplementedError
    def get_next_batch_data(self) -> List[Tuple[Any, Any]]: raise NotImplementedError
    def get_epoch_data(self) -> List[Tuple[Any, Any]]: raise NotImplementedError
    def get_num_batches_per_epoch(self) -> int: raise NotImplementedError


class ETTrainingDataProvider(TrainingDataProvider):
    """One local block of the input table is one mini-batch; blocks are shuffled every epoch and
    the provider tolerates blocks that migrate away in the middle of an epoch."""

    def __init__(self, input_table, shuffle: bool = True) -> None:
        self._table = input_table
        self._blocks: List[int] = []
        self.shuffle = shuffle

    def _tablet(self):
        return self._table.get_local_tablet()

    def prepare_data_for_epoch(self) -> None:
        self._blocks = sorted(self._tablet().get_block_ids(), reverse=True)   # popped from the end: ascending
        if self.shuffle:
            random.shuffle(self._blocks)

    def get_num_batches_per_epoch(self) -> int:
        return len(self._tablet().get_block_ids())

    def get_next_batch_data(self) -> List[Tuple[Any, Any]]:
        while self._blocks:
            b = self._blocks.pop()
            data = self._tablet().get_data_in_block(b)
            if data:
                return data
        return []

    def get_epoch_data(self) -> List[Tuple[Any, Any]]:
        return list(self._tablet().get_data_iterator())


# ------------------------------------------------------------------------------------------
# barriers / messaging
# ------------------------------------------------------------------------------------------
class WorkerSideMsgSender:
    """Worker -> master Dolphin messages (progress, sync, mini-batch sync, model-eval ask, group gather).
    Reference: jobserver/src/main/java/edu/snu/cay/dolphin/core/worker/WorkerSideMsgSender.java:64-127."""

    def __init__(self, context, job_id: str) -> None:
        self._ctx = context
        self.job_id = job_id
        self.worker_id = context.executor.executor_id

    def _send(self, body: Dict[str, Any], inline: bool = False) -> None:
        payload = {"job_id": self.job_id, "dolphin": body}
        if inline:
            # per-mini-batch traffic whose master-side handling never blocks (SSP bookkeeping + a reply)
            self._ctx.send_to_driver(payload, inline=True)
        else:
            self._ctx.send_to_driver(payload)

    def send_progress_msg(self, progress: int, count: int = 1) -> None:
        body = {"type": "ProgressMsg", "executor_id": self.worker_id, "progress": progress}
        if count != 1:
            body["count"] = count
        self._send(body, inline=True)

    def send_sync_msg(self, state: str) -> None:
        self._send({"type": "SyncMsg", "executor_id": self.worker_id, "state": state})

    def send_mini_batch_sync_msg(self, progress: Optional[int] = None, idle: bool = False, want: int = 1,
                                 progress_count: int = 1) -> None:
        """`progress` piggybacks the ProgressMsg of the mini-batch(es) finished since the last sync (one message less
        per batch); `idle` = this worker has no data right now: ask for the stop flag without consuming the budget;
        `want` > 1 asks for a lease of up to that many mini-batches (-mini_batch_grant_size)."""
        body = {"type": "MiniBatchSyncMsg", "worker_id": self.worker_id}
        if progress is not None:
            body["progress"] = progress
            if progress_count != 1:
                body["progress_count"] = progress_count
        if idle:
            body["idle"] = True
        if want > 1:
            body["want"] = want
        self._send(body, inline=True)

    def send_mini_batch_return_msg(self, count: Optional[int] = None) -> None:
        """`count` unused granted mini-batches go back to the budget; None = the whole latest grant (the worker was
        stopped while waiting for the reply and never saw its size)."""
        body = {"type": "MiniBatchReturnMsg", "worker_id": self.worker_id}
        if count is not None:
            body["count"] = count
        self._send(body, inline=True)

    def send_model_eval_ask_msg(self) -> None:
        self._send({"type": "ModelEvalAskMsg", "executor_id": self.worker_id})


class WorkerGlobalBarrier:
    """Local FSM INIT -> RUN -> CLEANUP; await() = SyncMsg(state) then block until ReleaseMsg."""

    def __init__(self, sender: WorkerSideMsgSender) -> None:
        self._sender = sender
        self.state = "INIT"
        self._released = threading.Semaphore(0)

    def await_(self) -> None:
        self._sender.send_sync_msg(self.state)
        self._released.acquire()
        self.state = {"INIT": "RUN", "RUN": "CLEANUP"}.get(self.state, self.state)

    def on_release_msg(self) -> None:
        self._released.release()


class MiniBatchBarrier:
    """await() returns True when the master says the global mini-batch budget is exhausted."""

    def __init__(self, sender: WorkerSideMsgSender) -> None:
        self._sender = sender
        self._ev = threading.Semaphore(0)
        self._stop = False
        self.granted = 1                 # mini-batches covered by the last reply (a lease when > 1)

    def await_(self, progress: Optional[int] = None, idle: bool = False, want: int = 1,
               progress_count: int = 1) -> bool:
        if want > 1 or progress_count != 1:
            self._sender.send_mini_batch_sync_msg(progress, idle, want, progress_count)
        else:
            self._sender.send_mini_batch_sync_msg(progress, idle)
        self._ev.acquire()
        return self._stop

    def on_control_msg(self, stop: bool, granted: int = 1) -> None:
        self._stop = stop
        self.granted = max(1, int(granted))
        self._ev.release()


class ModelEvalBarrier:
    """Worker side of the offline model evaluation loop: asks the master whether another checkpoint is to be
    evaluated.  Reference: jobserver/src/main/java/edu/snu/cay/dolphin/core/worker/ModelEvaluator.java:69-147."""

    def __init__(self, sender: WorkerSideMsgSender) -> None:
        self._sender = sender
        self._ev = threading.Semaphore(0)
        self._do_next = False

    def ask(self) -> bool:
        self._sender.send_model_eval_ask_msg()
        self._ev.acquire()
        return self._do_next

    def on_answer(self, do_next: bool) -> None:
        self._do_next = do_next
        self._ev.release()


# ------------------------------------------------------------------------------------------
# tasklets
# ------------------------------------------------------------------------------------------
def build_trainer(context, params: Dict[str, Any], model_table, local_model_table, input_table):
    has_gpu = getattr(getattr(context, "executor", None), "device_ctx", None) is not None
    is_device = bool(params.get("device_trainer_class")) and (
        hasattr(model_table, "device_table") or (has_gpu and bool(params.get("device_trainer_host_model"))))
    cls = load_class(params["device_trainer_class"] if is_device else params["trainer_class"])
    accessor_cls = CachedModelAccessor if params.get("model_cache_enabled") else ETModelAccessor
    if accessor_cls is CachedModelAccessor:
        accessor = CachedModelAccessor(model_table, model_table.components.update_fn)
    else:
        accessor = ETModelAccessor(model_table)
    return cls(params=params, model_accessor=accessor, model_table=model_table,
               local_model_table=local_model_table, input_table=input_table, context=context), accessor


class WorkerTasklet(Tasklet):
    """THE training loop."""

    def __init__(self, context) -> None:
        super().__init__(context)
        p = context.params
        self.job_id = p["job_id"]
        self.sender = WorkerSideMsgSender(context, self.job_id)
        self.global_barrier = WorkerGlobalBarrier(self.sender)
        self.mini_batch_barrier = MiniBatchBarrier(self.sender)
        self._abort = threading.Event()
        acc = context.table_accessor
        self.model_table = acc.get_table(p["model_table_id"])
        self.input_table = acc.get_table(p["input_table_id"])
        self.local_model_table = acc.get_table(p["local_model_table_id"]) if p.get("local_model_table_id") else None
        # rendezvous group of this job's workers (rank = position among the job's executors)
        ex_ids = list(p.get("job_executor_ids") or [context.executor.executor_id])
        from .group import WorkerSideGroup
        context.job_group = WorkerSideGroup(
            ex_ids.index(context.executor.executor_id) if context.executor.executor_id in ex_ids else 0, len(ex_ids),
            context.executor.executor_id, lambda body: context.send_to_driver({"job_id": self.job_id, "dolphin": body}))
        self._trainer_args = (context, p, self.model_table, self.local_model_table, self.input_table)
        self.trainer = self.accessor = self.data_provider = None
        self.max_epochs = int(p.get("max_num_epochs", 1))
        self.use_task_units = bool(p.get("task_units_enabled", True))

    # master -> worker messages
    def on_msg(self, payload: Any) -> None:
        t = payload["type"]
        if t == "ReleaseMsg":
            self.global_barrier.on_release_msg()
        elif t == "MiniBatchControlMsg":
            self.mini_batch_barrier.on_control_msg(payload["stop"], payload.get("granted", 1))
        elif t == "GroupGatherRes":
            grp = getattr(self.context, "job_group", None)
            if grp is not None:
                grp.on_msg(payload)

    def _unit(self, name: str, resource: str, seq: int, fn) -> float:
        info = TaskUnitInfo(self.job_id, name, resource, seq)
        sched = self.context.task_units
        # a stopping worker (plan op / job teardown) finishes the mini-batch it has already been granted -- the global
        # budget counted it -- but no longer takes part in the job's unit barriers
        if self.use_task_units and not self._abort.is_set():
            sched.wait_schedule(self.context.tasklet_id, info)
        dctx = self._device_ctx
        if dctx is not None:            # GPU slot: this job's kernels take 1 / (jobs active on the GPU) of the SMs
            dctx.apply_sm_share(self._sm_share)
        t0 = time.perf_counter()
        try:
            with trace.span(name, job=self.job_id):
                fn()
        finally:
            if self.use_task_units:
                sched.on_task_unit_finished(self.context.tasklet_id, info)
        return time.perf_counter() - t0

    def run(self) -> Any:
        self._device_ctx = getattr(self.context.executor, "device_ctx", None)
        self._sm_share = float(self.context.params.get("gpu_sm_share", 0.0) or 0.0)
        if self._device_ctx is not None:
            self._device_ctx.job_enter(self.job_id)
        try:
            return self._train()
        finally:
            if self._device_ctx is not None:
                self._device_ctx.job_leave(self.job_id)
            close = getattr(getattr(self, "accessor", None), "close", None)
            if close is not None:
                close()                   # the cached accessor's refresh thread ends with its tasklet

    def _train(self) -> Any:
        ex_id = self.context.executor.executor_id
        # the trainer is built on the tasklet's own thread, AFTER the runtime registered the tasklet: a trainer
        # that rendezvouses with its peers in the constructor (ship mode) needs the master's replies routed to it
        self.trainer, self.accessor = build_trainer(*self._trainer_args)
        self.data_provider = ETTrainingDataProvider(self.input_table,
                                                    shuffle=not getattr(self.trainer, "wants_sorted_batches", False))
        self.trainer.init_global_settings()
        self.global_barrier.await_()                       # INIT barrier
        seq = 0
        stop = False
        total_batches = 0
        epoch = 0
        from ..common.metric import MemoryMetricTracker
        # host RSS + HBM sampled every -memory_measure_interval ms; max / mean ride on the epoch metrics
        mem = MemoryMetricTracker(max(0.01, int(self.context.params.get("memory_measure_interval", 1000)) / 1000.0))
        pending_progress: Optional[int] = None
        pending_n = 0              # mini-batches finished since the last sync (> 1 only inside a lease)
        # -mini_batch_grant_size k: one sync leases up to k mini-batches (k - 1 driver round trips less; the SSP
        # staleness bound widens by k - 1); 1 = the reference's sync per mini-batch
        grant_size = int(self.context.params.get("mini_batch_grant_size", 0) or 0)
        if grant_size <= 0:        # auto: device trainers keep the stream busy across a lease, host trainers sync per batch
            grant_size = 4 if (getattr(self.trainer, "fused_step", False) and hasattr(self.trainer, "flush")
                               and not self.use_task_units) else 1
        if grant_size > 1 and self.use_task_units:
            # the unit barriers need the job's workers in lock-step (same unit, same sequence number); a lease lets
            # one worker run ahead while a peer is SSP-blocked at its sync -- the two waits would deadlock
            LOG.warning("[%s] -mini_batch_grant_size %d ignored: task-unit co-scheduling is on for this job "
                        "(-task_unit_scheduling off, or an exclusive scheduler, enables leases)", self.job_id, grant_size)
            grant_size = 1
        credits = 0                # granted mini-batches not started yet
        idle_left = False          # True while this worker has left the unit barriers because it holds no data
        flush_every = float(self.context.params.get("batch_metric_flush_interval_ms", 50)) / 1000.0
        last_flush = 0.0
        flush = getattr(self.trainer, "flush", None)       # device trainers: stream sync, once per lease (not per batch)
        while not stop and not self._abort.is_set():
            if self.input_table.get_local_tablet().get_num_data_items() == 0:
                # no input data here (all blocks migrated away, or none was loaded): ask the master whether to stop,
                # without consuming the mini-batch budget, holding the SSP clock back or counting in the job's unit
                # barriers; no epoch accounting while idle
                if self.use_task_units and not idle_left:
                    self.context.task_units.leave(self.job_id)
                    idle_left = True
                if credits:                                  # its data left in the middle of a lease
                    self.sender.send_mini_batch_return_msg(credits)
                    credits = 0
                done, n_done, pending_progress, pending_n = pending_progress, pending_n, None, 0
                if self.mini_batch_barrier.await_(done, idle=True, progress_count=max(1, n_done)):
                    stop = True
                    break
                time.sleep(0.05)
                continue
            if idle_left:
                self.context.task_units.join(self.job_id)       # it has data again
                idle_left = False
            epoch_t0 = time.perf_counter()
            mem.start()
            self.data_provider.prepare_data_for_epoch()
            n_batches, n_inst = 0, 0
            while True:
                batch = self.data_provider.get_next_batch_data()
                if not batch:
                    break
                if credits > 0:
                    if self._abort.is_set():                 # graceful stop inside a lease: the rest is handed back
                        stop = True
                        break
                    credits -= 1                             # this mini-batch is already granted
                else:
                    if self._abort.is_set():                 # stopped during the last batch: nothing more to ask for
                        stop = True
                        break
                    # SYNC (void unit): SSP clock + stop flag; carries the progress of the previous mini-batch(es)
                    if flush is not None:
                        flush()        # device trainers: the lease's kernels have finished before progress is reported
                    done, n_done, pending_progress, pending_n = pending_progress, pending_n, None, 0
                    if self.mini_batch_barrier.await_(done, want=grant_size, progress_count=max(1, n_done)):
                        stop = True
                        if self._abort.is_set():
                            # stopped while asking for mini-batches: the controller counted the request against the
                            # global budget but they will not run here -- hand back whatever it granted, so that
                            # another worker trains on them
                            self.sender.send_mini_batch_return_msg()
                        break
                    credits = self.mini_batch_barrier.granted - 1
                self.trainer.set_mini_batch_data(batch)
                self.accessor.get_and_reset_metrics()
                t0 = time.perf_counter()
                if getattr(self.trainer, "fused_step", False):
                    # device trainers fuse pull/compute/push into back-to-back kernels: ONE task unit
                    # (one driver round trip) per mini-batch instead of three
                    def fused():
                        self.trainer.pull_model(); self.trainer.local_compute(); self.trainer.push_update()
                    self._unit("STEP", NET, seq, fused)
                else:
                    self._unit("PULL", NET, seq, self.trainer.pull_model)
                    self._unit("COMP", CPU, seq, self.trainer.local_compute)
                    self._unit("PUSH", NET, seq, self.trainer.push_update)
                seq += 1
                bt = time.perf_counter() - t0
                am = self.accessor.get_and_reset_metrics()
                pending_progress = total_batches
                pending_n += 1
                total_batches += 1
                n_batches += 1
                n_inst += len(batch)
                bm = BatchMetrics(self.job_id, ex_id, epoch, n_batches - 1, len(batch), bt,
                                  am.get("pull_time_sec", 0.0), max(0.0, bt - am.get("pull_time_sec", 0.0)
                                                                    - am.get("push_time_sec", 0.0)),
                                  am.get("push_time_sec", 0.0), len(batch) / bt if bt > 0 else 0.0)
                self.context.metric_collector.add_custom_metric(bm.to_dict())
                now = time.perf_counter()
                if now - last_flush >= flush_every:        # batch metrics ride together (the epoch end flushes too)
                    self.context.metric_collector.flush()
                    last_flush = now
            if n_batches == 0 and not stop:
                time.sleep(0.01)                 # its blocks left in the middle of the epoch: idle from the next pass
            if flush is not None:
                flush()
            if not stop:
                self.trainer.on_epoch_finished(epoch)
            em = EpochMetrics(self.job_id, ex_id, epoch, n_batches, n_inst, time.perf_counter() - epoch_t0,
                              epoch_custom_metrics=mem.stop())
            self.context.metric_collector.add_custom_metric(em.to_dict())
            self.context.metric_collector.flush()
            epoch += 1
        mem.close()
        if credits:                                        # stopped inside a lease: the rest goes back to the budget
            self.sender.send_mini_batch_return_msg(credits)
        if pending_progress is not None:                   # stopped between a batch and its next sync
            self.sender.send_progress_msg(pending_progress, max(1, pending_n))
        if self.use_task_units and not idle_left:
            self.context.task_units.leave(self.job_id)
        result = {"num_batches": total_batches, "num_epochs": epoch, "trainer": type(self.trainer).__name__}
        if self._abort.is_set():
            # stopped by a plan op (worker deleted / switched to a server) or by the teardown of a failed job: the
            # job goes on (or is over) without this worker -- no CLEANUP barrier, no end-of-training validation
            on_stopped = getattr(self.trainer, "on_worker_stopped", None)
            if on_stopped is not None:
                on_stopped()
            return {**result, "stopped": True, **self._summary(total_batches)}
        self.global_barrier.await_()                       # RUN -> CLEANUP barrier
        summary = self._summary(total_batches)
        self.trainer.cleanup()
        return {**result, **summary}

    def _summary(self, total_batches: int) -> Dict[str, float]:
        # "last batch" numbers exist only on a worker that ran a batch: with fewer mini-batches than workers (or
        # leases) some worker of a short job legitimately trains on none
        if total_batches > 0:
            return self.trainer.summary()
        try:
            return self.trainer.summary()
        except Exception as e:  # noqa: BLE001
            LOG.debug("[%s] no summary from a worker that ran no mini-batch: %s", self.job_id, e)
            return {}

    def close(self) -> None:
        """Graceful stop: the loop ends at the next mini-batch boundary (a batch in flight completes)."""
        self._abort.set()
        self.mini_batch_barrier.on_control_msg(True)
        self.global_barrier.on_release_msg()
        self.global_barrier.on_release_msg()
        self.context.task_units.abort(self.context.tasklet_id)


class ServerTasklet(Tasklet):
    """Servers are passive shard owners: the tasklet idles until closed."""

    passive = True          # parks on a latch: takes no tasklet slot (et/tasklet.py)

    def __init__(self, context) -> None:
        super().__init__(context)
        self._latch = threading.Event()

    def run(self) -> Any:
        self._latch.wait()
        return None

    def close(self) -> None:
        self._latch.set()


class ModelEvaluationTasklet(Tasklet):
    """Offline model evaluation: for every restored checkpoint the master answers `do_next`."""

    def __init__(self, context) -> None:
        super().__init__(context)
        p = context.params
        self.job_id = p["job_id"]
        self.sender = WorkerSideMsgSender(context, self.job_id)
        self.barrier = ModelEvalBarrier(self.sender)
        self.params = p

    def on_msg(self, payload: Any) -> None:
        if payload["type"] == "ModelEvalAnsMsg":
            self.barrier.on_answer(payload["do_next"])

    def run(self) -> Any:
        from ..common.dataloader import get_splits, make_parser, read_split
        acc = self.context.table_accessor
        p = self.params
        results = []
        test_data: List[Any] = []
        if p.get("test_data_path"):
            parser = make_parser(p.get("data_parser"))
            for s in get_splits(p["test_data_path"], 1):
                test_data.extend(parser.parse(read_split(s)))
        while self.barrier.ask():
            model_table = acc.get_table(p["model_table_id"])
            input_table = acc.get_table(p["input_table_id"])
            local_table = acc.get_table(p["local_model_table_id"]) if p.get("local_model_table_id") else None
            trainer, _ = build_trainer(self.context, p, model_table, local_table, input_table)
            data = list(input_table.get_local_tablet().get_data_iterator())
            m = trainer.evaluate_model(data, test_data)
            LOG.info("[%s] model evaluation: %s", self.job_id, m)
            results.append(m)
            trainer.cleanup()
        return results
