"""Dolphin launchers.

`DolphinJobLauncher.submit_job`  -- submit to a running job server (submit_<app>.sh;
                                    reference dolphin/jobserver/DolphinJobLauncher.java:75-215).
`ETDolphinLauncher.launch`       -- stand-alone job with its own driver (run_<app>.sh;
                                    reference dolphin/core/client/ETDolphinLauncher.java:111-378 +
                                    core/driver/DolphinDriver.java:97-346): separate worker / server
                                    executors (or co-located with -num_servers 0), workers subscribe
                                    to the model table, optional optimizer orchestration and
                                    offline model evaluation.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Sequence

from ..et import ExecutorConfiguration, LocalTransport
from ..et.driver import ETMaster, InProcessLauncher
from ..jobserver.client import CommandSender, DEFAULT_PORT
from ..utils.flags import Configuration
from . import params as P
from .conf import ETDolphinConfiguration
from .jobserver import build_tables_confs
from .master import DolphinMaster, ETDolphinMetricReceiver

LOG = logging.getLogger(__name__)


def parse_job_conf(app: ETDolphinConfiguration, argv: Sequence[str]) -> Configuration:
    reg = P.dolphin_registry(app.parameters, P.EXAMPLES if app.app_id.startswith("Add") else [])
    conf, _ = reg.parse(list(argv), allow_unknown=False)
    from ..common.dataloader import resolve_data_paths
    return resolve_data_paths(conf)


class DolphinJobLauncher:
    """Client side of `submit_<app>.sh`: flags -> one serialisable job configuration -> SUBMIT command to the job
    server.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/jobserver/DolphinJobLauncher.java:75-144,147-196."""

    @staticmethod
    def submit_job(app_path: str, app: ETDolphinConfiguration, argv: Sequence[str],
                   port: Optional[int] = None) -> str:
        conf = parse_job_conf(app, argv).with_values(
            job_entity_builder="harmony_b200.dolphin.jobserver:DolphinJobEntityBuilder", app=app_path)
        reply = CommandSender(port or int(conf.get("port", DEFAULT_PORT))).send_job_submit_command(conf.serialize())
        return reply


class LauncherResult:
    """Outcome of a stand-alone run (`run_<app>.sh`): COMPLETED / FAILED plus the worker results."""

    def __init__(self, ok: bool, results: Any, messages: List[str], master: Optional[DolphinMaster]) -> None:
        self.ok, self.results, self.messages, self.dolphin_master = ok, results, messages, master
        self.eval_results: Optional[List[Any]] = None     # -offline_model_eval: one entry per evaluation worker

    @property
    def status(self) -> str:
        return "COMPLETED" if self.ok else "FAILED"


class ETDolphinLauncher:
    """Stand-alone launcher (own driver per app, `run_<app>.sh`): worker / server executors, tables, bulk load,
    training, optional optimizer, dashboard and offline evaluation.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/core/client/ETDolphinLauncher.java:111-211,226-327,
    core/driver/DolphinDriver.java:97-162,262-346."""

    @staticmethod
    def launch(app: ETDolphinConfiguration, argv: Sequence[str],
               extra_conf: Optional[Dict[str, Any]] = None) -> LauncherResult:
        conf = parse_job_conf(app, argv)
        if extra_conf:
            conf = conf.merge(extra_conf)
        num_workers = int(conf.get("num_dolphin_workers", 1))
        num_servers = int(conf.get("num_servers", 1))
        colocated = num_servers == 0
        # -backend auto|host|device: on a GPU box the stand-alone job also runs on HBM tables + fused kernels
        import itertools
        from ..et.launcher import SubprocessLauncher, detect_backend, make_executor
        backend = detect_backend(conf.get("backend", "auto"))
        conf = conf.with_values(backend=backend)
        if conf.get("executor_mode", "inproc") == "process":
            # one OS process per executor (per GPU on the device backend), TCP control plane -- the layout of the job
            # server's `-executor_mode process`; the optimizer's plans (add / delete / switch executors) run over it
            from ..et.transport import TcpTransport
            transport = TcpTransport("127.0.0.1", 0, None)
            launcher = SubprocessLauncher(transport.address, backend=backend)
        else:
            transport = LocalTransport()
            ordinal = itertools.count()
            launcher = InProcessLauncher(
                lambda eid, tr, ec_, did: make_executor(eid, tr, ec_, did, backend, next(ordinal)))
        master = ETMaster(transport, launcher, chkp_temp_path=conf.get("chkp_temp_path", "/tmp/chkp"),
                          chkp_commit_path=conf.get("chkp_commit_path", "/tmp/chkp/commit"))
        messages: List[str] = []
        receiver = ETDolphinMetricReceiver()
        master.metrics.receivers.append(receiver)
        ok, results, dm, eval_results = False, None, None, None
        try:
            ec = ExecutorConfiguration(num_tasklets=4, chkp_temp_path=conf.get("chkp_temp_path", "/tmp/chkp"),
                                       chkp_commit_path=conf.get("chkp_commit_path", "/tmp/chkp/commit"))
            workers = master.add_executors(num_workers, ec).get(timeout=120)
            servers = workers if colocated else master.add_executors(num_servers, ec).get(timeout=120)
            server_tc, worker_tc, local_tc = build_tables_confs(app, conf, 0)
            master.create_table(server_tc, servers).get(timeout=600)
            model_table = master.get_table(server_tc.table_id)
            if not colocated:
                model_table.subscribe(workers).get(timeout=120)
            if local_tc is not None:
                master.create_table(local_tc, workers).get(timeout=600)
            master.create_table(worker_tc, workers).get(timeout=3600)
            params = dict(conf)
            params.update({"trainer_class": app.trainer_class, "device_trainer_class": app.device_trainer_class,
                           "device_trainer_host_model": app.device_trainer_host_model,
                           "model_table_id": server_tc.table_id, "input_table_id": worker_tc.table_id,
                           "local_model_table_id": local_tc.table_id if local_tc else None,
                           "data_parser": worker_tc.data_parser, "app_id": app.app_id, "task_units_enabled": False})
            job_id = f"{app.app_id}-0"
            dm = DolphinMaster(master, job_id, params, report=lambda s: (messages.append(s), LOG.info(s)))
            receiver.register(job_id, dm.metric_manager)
            dashboard = None
            if conf.get("dashboard_enabled"):
                from .dashboard import DashboardConnector, DashboardServer
                dashboard = DashboardServer(int(conf.get("dashboard_port", 9005)))
                dm.metric_manager.sinks.append(DashboardConnector(dashboard.port))
                messages.append(f"Dashboard: http://127.0.0.1:{dashboard.port}/")
            master.custom_handlers["TaskletCustom"] = lambda m: dm.on_worker_msg(m["payload"])
            orchestrator = None
            opt_name = conf.get("optimizer", "")
            if opt_name and not opt_name.endswith("EmptyPlanOptimizer"):
                from .optimizer import ETOptimizationOrchestrator
                orchestrator = ETOptimizationOrchestrator(master, dm, conf, server_tc.table_id, worker_tc.table_id,
                                                          servers, workers, local_tc.table_id if local_tc else None)
                orchestrator.start()
            try:
                results = dm.start(servers if not colocated else [], workers) if not colocated \
                    else dm.start([], workers)
            finally:
                if orchestrator is not None:
                    orchestrator.stop()
            if conf.get("offline_model_eval"):
                eval_results = _offline_evaluation(master, dm, workers, server_tc, worker_tc, local_tc)
            ok = True
        except BaseException:  # noqa: BLE001
            LOG.exception("stand-alone job failed")
        finally:
            master.close()
            transport.close()
        res = LauncherResult(ok, results, messages, dm)
        res.eval_results = eval_results
        return res


def _offline_evaluation(master: ETMaster, dm: DolphinMaster, execs, server_tc, worker_tc, local_tc) -> List[Any]:
    """-offline_model_eval on a stand-alone job: after training, replay the per-epoch model checkpoints oldest first
    and evaluate each against the input checkpoint / -test_data_path (ETDolphinLauncher.java:342-378,
    ModelChkpManager.java:221-272; same replay as the job server's deferred evaluation)."""
    ck = dm.chkp_manager
    for tc in (server_tc, local_tc, worker_tc):          # the training tables make room for the restored ones
        if tc is not None and master.has_table(tc.table_id):
            master.get_table(tc.table_id).drop().get(timeout=120)
    restored: List[str] = []

    def drop_restored() -> None:
        for tid in restored:
            if master.has_table(tid):
                master.get_table(tid).drop().get(timeout=120)
        restored.clear()

    def restore(i: int) -> bool:
        drop_restored()
        if i >= ck.num_chkps():
            return False
        master.create_table_from_chkp(ck.model_chkps[i], execs, table_id=server_tc.table_id).get(timeout=600)
        restored.append(server_tc.table_id)
        if local_tc is not None and i < len(ck.local_chkps):
            master.create_table_from_chkp(ck.local_chkps[i], execs, table_id=local_tc.table_id).get(timeout=600)
            restored.append(local_tc.table_id)
        if not master.has_table(worker_tc.table_id) and ck.input_chkp:
            master.create_table_from_chkp(ck.input_chkp, execs, table_id=worker_tc.table_id).get(timeout=600)
        return True
    out = dm.evaluate(execs, restore)
    drop_restored()
    if master.has_table(worker_tc.table_id):
        master.get_table(worker_tc.table_id).drop().get(timeout=120)
    return out


def main_submit(app_path: str, app: ETDolphinConfiguration, argv: Sequence[str]) -> int:
    reply = DolphinJobLauncher.submit_job(app_path, app, argv)
    print(reply)
    return 0 if reply.startswith("true") else 1


def main_run(app: ETDolphinConfiguration, argv: Sequence[str]) -> int:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(name)s | %(message)s")
    from ..utils import trace
    trace.configure(parse_job_conf(app, argv))
    res = ETDolphinLauncher.launch(app, argv)
    where = trace.flush_to_receiver()
    if where:
        print(f"trace spans -> {where}")
    print(f"LauncherStatus: {res.status}")
    return 0 if res.ok else 1
