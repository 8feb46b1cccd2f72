"""Pregel master / worker / launchers / job-server glue.

Reference: pregel/PregelMaster.java:88-172 (superstep controller: AND-reduce of `allHalt` and
`noMsgs` over the workers, broadcast Start/Stop), PregelWorkerTask.java:100-161 (partition local
vertices over threads -> compute -> flush messages -> sync), WorkerMsgManager.java,
PregelDriver/PregelLauncher (stand-alone), jobserver/PregelJobEntity.java:76-111 (vertex table
hashed+mutable, bulk loaded with keys; two message tables `msg_table<n>-1/-2` whose update
function appends to the list), PregelParameters.java:33-89.
"""
from __future__ import annotations

import logging
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, List, Optional, Sequence

from ..dolphin import params as P
from ..et import ExecutorConfiguration, LocalTransport, TableConfiguration
from ..et.config import TaskletConfiguration
from ..et.driver import AllocatedExecutor, ETMaster, InProcessLauncher
from ..et.tasklet import CPU, NET, TaskUnitInfo, Tasklet, load_class
from ..jobserver.client import CommandSender, DEFAULT_PORT
from ..jobserver.driver import JobEntity, JobMaster, JobServerDriver
from ..utils.concurrent import CountDownLatch
from ..utils.flags import Configuration
from .graph import MessageManager, run_partition

LOG = logging.getLogger(__name__)
WORKER_TASKLET = "harmony_b200.pregel.runtime:PregelWorkerTask"
VERTEX_TABLE_ID, MSG_TABLE_ID = "vertex_table", "msg_table"


class PregelMaster:
    """Superstep control: gathers every worker's (halted, sent-messages) flags, AND-reduces them, starts the next
    superstep or stops.  Reference: jobserver/src/main/java/edu/snu/cay/pregel/PregelMaster.java:88-172."""

    def __init__(self, master: ETMaster, job_id: str, params: Dict[str, Any]) -> None:
        self._m, self.job_id, self.params = master, job_id, params
        self._tasklets: List[Any] = []
        self._lock = threading.Lock()
        self._all_halt, self._no_msgs = True, True
        self._latch: Optional[CountDownLatch] = None
        self.num_supersteps = 0

    def on_worker_msg(self, payload: Dict[str, Any]) -> None:
        body = payload.get("pregel", payload)
        with self._lock:
            self._all_halt &= bool(body["is_all_vertices_halt"])
            self._no_msgs &= bool(body["is_no_ongoing_msgs"])
        self._latch.count_down()

    def _control_loop(self, n: int) -> None:
        while True:
            self._latch.await_()
            with self._lock:
                stop = self._all_halt and self._no_msgs
                self._all_halt = self._no_msgs = True
                self._latch = CountDownLatch(n)
            self.num_supersteps += 1
            for t in self._tasklets:
                t.send({"type": "SuperstepControlMsg", "control": "Stop" if stop else "Start"})
            if stop:
                return

    def start(self, executors: Sequence[AllocatedExecutor]) -> List[Any]:
        n = len(executors)
        self._latch = CountDownLatch(n)
        for ex in executors:
            conf = TaskletConfiguration(f"{self.job_id}-pregel-worker-{ex.id}", WORKER_TASKLET, None,
                                        dict(self.params, job_id=self.job_id))
            self._tasklets.append(ex.submit_tasklet(conf, self.on_worker_msg))
        threading.Thread(target=self._control_loop, args=(n,), daemon=True).start()
        results = [t.get_task_result(timeout=3600) for t in self._tasklets]
        failed = [r for r in results if not r.is_success]
        if failed:
            raise RuntimeError("pregel workers failed: " + "; ".join(str(r.error) for r in failed))
        return [r.result for r in results]


class PregelWorkerTask(Tasklet):
    """Worker loop: per superstep run the computation over the local vertices with the current message table, flush
    outgoing messages, report to the master.  Reference:
    jobserver/src/main/java/edu/snu/cay/pregel/PregelWorkerTask.java:100-161."""

    def __init__(self, context) -> None:
        super().__init__(context)
        p = context.params
        self.job_id = p["job_id"]
        acc = context.table_accessor
        self.vertex_table = acc.get_table(p["vertex_table_id"])
        combiner = load_class(p["combiner_class"])() if p.get("combiner_class") else None
        self.mm = MessageManager(acc.get_table(p["msg_table_id"] + "-1"), acc.get_table(p["msg_table_id"] + "-2"),
                                 combiner)
        self.computation = load_class(p["computation_class"])(self.mm, p)
        self.num_threads = int(p.get("num_worker_threads", 0) or 0) or 2
        self._ctl = threading.Semaphore(0)
        self._continue = True
        self.use_task_units = bool(p.get("task_units_enabled", False))

    def on_msg(self, payload: Any) -> None:
        if payload.get("type") == "SuperstepControlMsg":
            self._continue = payload["control"] == "Start"
            self._ctl.release()

    def _unit(self, name: str, res: str, seq: int):
        return TaskUnitInfo(self.job_id, name, res, seq)

    def run(self) -> Any:
        pool = ThreadPoolExecutor(max_workers=self.num_threads)
        superstep = 0
        sched, tid = self.context.task_units, self.context.tasklet_id
        try:
            while True:
                self.computation.initialize(superstep)
                vertices = list(self.vertex_table.get_local_tablet().get_data_map().values())
                parts = [vertices[i::self.num_threads] for i in range(self.num_threads)]
                u = self._unit("COMP", CPU, superstep)
                if self.use_task_units:
                    sched.wait_schedule(tid, u)
                futs = [pool.submit(run_partition, self.computation, part, self.mm.get_current_message_table())
                        for part in parts]
                active = sum(f.result() for f in futs)
                if self.use_task_units:
                    sched.on_task_unit_finished(tid, u)
                u = self._unit("SEND", NET, superstep)
                if self.use_task_units:
                    sched.wait_schedule(tid, u)
                msg_exist = self.mm.flush_all_messages()
                if self.use_task_units:
                    sched.on_task_unit_finished(tid, u)
                self.context.send_to_driver({"job_id": self.job_id, "pregel": {
                    "type": "SuperstepResultMsg", "is_all_vertices_halt": active == 0,
                    "is_no_ongoing_msgs": not msg_exist}})
                self._ctl.acquire()
                if not self._continue:
                    break
                self.mm.prepare_for_next_superstep()
                superstep += 1
        finally:
            pool.shutdown(wait=False)
        return {"supersteps": superstep + 1,
                "values": {v.get_id(): v.get_value() for v in
                           self.vertex_table.get_local_tablet().get_data_map().values()}}


# ------------------------------------------------------------------------------------------
# app descriptor, tables, launchers
# ------------------------------------------------------------------------------------------
class PregelApp:
    """App descriptor of a Pregel job (computation, combiner, parser, value types)."""

    def __init__(self, app_id: str, computation_class: str, parser: str, combiner_class: Optional[str] = None,
                 parameters=()) -> None:
        self.app_id, self.computation_class, self.parser = app_id, computation_class, parser
        self.combiner_class, self.parameters = combiner_class, list(parameters)


def table_confs(app: PregelApp, conf: Configuration, n: int):
    vt = TableConfiguration(table_id=f"{VERTEX_TABLE_ID}{n}", num_total_blocks=int(conf.get("num_worker_blocks", 64)),
                            is_mutable=True, is_ordered=False, update_function="void", input_path=conf.get("input"),
                            data_parser=app.parser, bulk_loader="exist_key")
    mts = [TableConfiguration(table_id=f"{MSG_TABLE_ID}{n}-{i}", num_total_blocks=int(conf.get("num_worker_blocks", 64)),
                              is_mutable=True, is_ordered=False, update_function="append") for i in (1, 2)]
    return vt, mts


def worker_params(app: PregelApp, conf: Configuration, n: int) -> Dict[str, Any]:
    p = dict(conf)
    p.update({"vertex_table_id": f"{VERTEX_TABLE_ID}{n}", "msg_table_id": f"{MSG_TABLE_ID}{n}",
              "computation_class": app.computation_class, "combiner_class": app.combiner_class})
    return p


def parse_conf(app: PregelApp, argv: Sequence[str]) -> Configuration:
    reg = P.registry(P.COMMON, P.JOBSERVER, P.PREGEL, [P.Param("num_worker_blocks", int, 64, "", "worker")],
                     app.parameters)
    conf, _ = reg.parse(list(argv), allow_unknown=True)
    from ..common.dataloader import resolve_data_paths
    return resolve_data_paths(conf)


class PregelJobMaster(JobMaster):
    """JobMaster of a Pregel job inside the job server.  Reference: pregel/jobserver/PregelJobMaster.java."""

    def __init__(self, entity: "PregelJobEntity") -> None:
        self.entity = entity
        self.master: Optional[PregelMaster] = None

    def on_msg(self, source_id: str, payload: Any) -> None:
        if self.master is not None:
            self.master.on_worker_msg(payload)

    def start(self) -> Any:
        e = self.entity
        self.master = PregelMaster(e.driver.et_master, e.job_id, worker_params(e.app, e.conf, e.n))
        return self.master.start(e.executors)


class PregelJobEntity(JobEntity):
    """A submitted Pregel job: vertex table + two message tables on all executors.  Reference:
    jobserver/src/main/java/edu/snu/cay/pregel/jobserver/PregelJobEntity.java:76-111."""

    def __init__(self, driver: JobServerDriver, conf: Configuration, app: PregelApp, n: int) -> None:
        self.driver, self.conf, self.app, self.n = driver, conf, app, n
        self.job_id = f"{app.app_id}-{n}"
        self.executors: List[AllocatedExecutor] = []
        self._master = PregelJobMaster(self)
        self._tables: List[str] = []

    def get_job_master(self) -> JobMaster:
        return self._master

    def setup_executors_and_tables(self, executors: Sequence[AllocatedExecutor]) -> None:
        m = self.driver.et_master
        self.executors = list(executors)
        vt, mts = table_confs(self.app, self.conf, self.n)
        futs = [m.create_table(vt, self.executors)] + [m.create_table(t, self.executors) for t in mts]
        for f in futs:
            f.get(timeout=3600)
        self._tables = [vt.table_id] + [t.table_id for t in mts]

    def cleanup(self) -> None:
        m = self.driver.et_master
        for t in self._tables:
            if m.has_table(t):
                m.get_table(t).drop().get(timeout=120)


class PregelJobEntityBuilder:
    """Builds a PregelJobEntity from the job configuration.  Reference:
    pregel/jobserver/PregelJobEntityBuilder.java."""

    def __init__(self, driver: JobServerDriver, conf: Configuration) -> None:
        self.driver, self.conf = driver, conf

    def build(self) -> PregelJobEntity:
        app = load_class(self.conf["app"])
        return PregelJobEntity(self.driver, self.conf, app, self.driver.next_job_count())


class PregelJobLauncher:
    """Client side of submit_pagerank.sh / submit_shortest_path.sh.  Reference:
    pregel/jobserver/PregelJobLauncher.java."""

    @staticmethod
    def submit_job(app_path: str, app: PregelApp, argv: Sequence[str], port: Optional[int] = None) -> str:
        conf = parse_conf(app, argv).with_values(
            job_entity_builder="harmony_b200.pregel.runtime:PregelJobEntityBuilder", app=app_path)
        return CommandSender(port or int(conf.get("port", DEFAULT_PORT))).send_job_submit_command(conf.serialize())


class PregelLauncher:
    """Stand-alone run (run_pagerank.sh / run_shortest_path.sh)."""

    @staticmethod
    def launch(app: PregelApp, argv: Sequence[str]):
        conf = parse_conf(app, argv)
        if conf.get("executor_mode", "inproc") == "process":
            # one OS process per Pregel worker over the TCP control plane (the Dolphin launcher's `-executor_mode process`)
            from ..et.launcher import SubprocessLauncher
            from ..et.transport import TcpTransport
            transport = TcpTransport("127.0.0.1", 0, None)
            master = ETMaster(transport, SubprocessLauncher(transport.address, backend="host"))
        else:
            transport = LocalTransport()
            master = ETMaster(transport, InProcessLauncher())
        ok, results, pm = False, None, None
        try:
            execs = master.add_executors(int(conf.get("num_pregel_workers", 1)), ExecutorConfiguration()).get(120)
            vt, mts = table_confs(app, conf, 0)
            master.create_table(vt, execs).get(3600)
            for t in mts:
                master.create_table(t, execs).get(600)
            pm = PregelMaster(master, f"{app.app_id}-0", worker_params(app, conf, 0))
            master.custom_handlers["TaskletCustom"] = lambda m: pm.on_worker_msg(m["payload"])
            results = pm.start(execs)
            ok = True
        except BaseException:  # noqa: BLE001
            LOG.exception("pregel job failed")
        finally:
            master.close()
            transport.close()
        return ok, results, pm
