"""Dolphin <-> job server glue.

Reference: dolphin/jobserver/DolphinJobEntityBuilder.java:57-205 (jobId = <App>-<n>; table ids
model_table<n>, local_model_table<n>, <inputFileName><n>; model table mutable+hashed with
-num_server_blocks; input table immutable+ordered with -num_worker_blocks, bulk loaded; local
model table mutable+ordered), DolphinJobEntity.java:76-121 (co-location: servers = workers = the
given executors; input table re-used when it already exists), DolphinJobMaster.java:65-87.
"""
from __future__ import annotations

import importlib
import logging
import os
from typing import Any, Dict, List, Optional, Sequence

from ..et.config import TableConfiguration
from ..et.driver import AllocatedExecutor
from ..jobserver.driver import JobEntity, JobMaster, JobServerDriver
from ..utils.flags import Configuration
from .conf import ETDolphinConfiguration
from .master import DolphinMaster

LOG = logging.getLogger(__name__)
MODEL_TABLE_ID, LOCAL_MODEL_TABLE_ID = "model_table", "local_model_table"


def load_app(path: str) -> ETDolphinConfiguration:
    mod, _, name = path.partition(":")
    return getattr(importlib.import_module(mod), name or "APP")


def build_tables_confs(app: ETDolphinConfiguration, conf: Configuration, job_count: int):
    model_id = f"{MODEL_TABLE_ID}{job_count}"
    local_id = f"{LOCAL_MODEL_TABLE_ID}{job_count}"
    input_path = conf.get("input") or ""
    input_id = (os.path.basename(input_path.rstrip("/")) or "input") + str(job_count)
    load_model = bool(conf.get("load_model", False))
    server_tc = app.model_table_conf(conf, model_id)
    server_tc.num_total_blocks = int(conf.get("num_server_blocks", 1024))
    server_tc.is_mutable, server_tc.is_ordered = True, False
    if server_tc.user_params.get("dense_input"):
        # the tensor-core path TMA-loads K tiles of 32 consecutive features: range-partitioned blocks whose
        # size is a multiple of 32 rows
        server_tc.is_ordered = True
        nb = max(1, min(server_tc.num_total_blocks, server_tc.capacity // 32))
        while nb > 1 and (-(-server_tc.capacity // nb)) % 32 != 0:
            nb -= 1
        server_tc.num_total_blocks = nb
    server_tc.chkp_path = conf.get("model_path") if load_model and conf.get("model_path") else None
    parser_spec = app.input_parser
    if isinstance(parser_spec, str):      # the parser validates labels / indices against the job's dimensions
        bounds = {k: conf[k] for k in ("classes", "features", "num_vocabs") if conf.get(k)}
        parser_spec = {"name": parser_spec, "params": bounds}
    worker_tc = TableConfiguration(
        table_id=input_id, num_total_blocks=int(conf.get("num_worker_blocks", 1024)), is_mutable=False,
        is_ordered=True, update_function="void", input_path=input_path or None, data_parser=parser_spec,
        bulk_loader="exist_key" if app.has_input_data_key else "none_key",
        chkp_path=(conf.get("input_chkp_path") or None) if load_model and app.has_local_model_table else None)
    local_tc = None
    if app.has_local_model_table:
        local_tc = app.local_model_table_conf(conf, local_id)
        local_tc.num_total_blocks = int(conf.get("num_worker_blocks", 1024))
        local_tc.is_mutable, local_tc.is_ordered = True, True
        local_tc.chkp_path = conf.get("local_model_path") if load_model and conf.get("local_model_path") else None
    return server_tc, worker_tc, local_tc


class DolphinJobMaster(JobMaster):
    """JobMaster of a Dolphin job inside the job server: routes tasklet messages to the DolphinMaster, runs it,
    returns the results.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/jobserver/DolphinJobMaster.java:65-87."""

    def __init__(self, entity: "DolphinJobEntity") -> None:
        self.entity = entity
        self.dolphin_master: Optional[DolphinMaster] = None

    def on_msg(self, source_id: str, payload: Any) -> None:
        if self.dolphin_master is not None:
            self.dolphin_master.on_worker_msg(payload)

    def start(self) -> Any:
        e = self.entity
        self.dolphin_master = DolphinMaster(e.driver.et_master, e.job_id, e.worker_params(),
                                            report=e.driver.send_to_client)
        e.driver.metric_receiver.register(e.job_id, self.dolphin_master.metric_manager)
        try:
            if e.conf.get("model_eval"):
                return self._evaluate_only()
            result = self.dolphin_master.start(e.executors, e.executors)
            if e.conf.get("offline_model_eval"):
                dm = self.dolphin_master
                e.driver.deferred_evaluations.append(lambda: e.run_offline_evaluation(dm))
            return result
        finally:
            e.driver.metric_receiver.unregister(e.job_id)

    def _evaluate_only(self) -> Any:
        """-model_eval: the model table was created from -model_path; evaluate it once."""
        dm = self.dolphin_master
        done = {"n": 0}

        def restore(i: int) -> bool:
            done["n"] += 1
            return i == 0
        return dm.evaluate(self.entity.executors, restore)


class DolphinJobEntity(JobEntity):
    """A submitted Dolphin job: its configuration, executors (co-located workers / servers) and tables (model,
    local model, input).  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/jobserver/DolphinJobEntity.java:76-121."""

    def __init__(self, driver: JobServerDriver, conf: Configuration, app: ETDolphinConfiguration, job_id: str,
                 server_tc: TableConfiguration, worker_tc: TableConfiguration,
                 local_tc: Optional[TableConfiguration]) -> None:
        self.driver, self.conf, self.app, self.job_id = driver, conf, app, job_id
        self.server_tc, self.worker_tc, self.local_tc = server_tc, worker_tc, local_tc
        self.executors: List[AllocatedExecutor] = []
        self._master = DolphinJobMaster(self)
        self._created_input = False

    def get_job_master(self) -> JobMaster:
        return self._master

    def num_executors_wanted(self) -> int:
        return int(self.conf.get("num_job_executors", 0) or 0)

    def worker_params(self) -> Dict[str, Any]:
        p = dict(self.conf)
        p.update({"trainer_class": self.app.trainer_class, "device_trainer_class": self.app.device_trainer_class,
                  "device_trainer_host_model": self.app.device_trainer_host_model,
                  "model_table_id": self.server_tc.table_id, "input_table_id": self.worker_tc.table_id,
                  "local_model_table_id": self.local_tc.table_id if self.local_tc else None,
                  "data_parser": self.worker_tc.data_parser, "app_id": self.app.app_id,
                  "num_job_workers": max(1, len(self.executors)),
                  "job_executor_ids": [e.id for e in self.executors]})
        # the task-unit barriers exist to interleave the phases of jobs that SHARE executors: a scheduler that gives
        # each job the pool for itself makes them pure overhead (one driver round trip per phase per mini-batch)
        mode = str(self.conf.get("task_unit_scheduling", "auto")).lower()
        if mode not in ("auto", "on", "off"):
            raise ValueError(f"-task_unit_scheduling must be auto, on or off (got {mode!r})")
        exclusive = bool(getattr(getattr(self.driver, "scheduler", None), "exclusive", False))
        p["task_units_enabled"] = mode == "on" or (mode == "auto" and not exclusive)
        return p

    def setup_executors_and_tables(self, executors: Sequence[AllocatedExecutor]) -> None:
        m = self.driver.et_master
        self.executors = list(executors)        # co-location: every executor is server AND worker
        futs = [m.create_table(self.server_tc, self.executors)]
        if self.local_tc is not None:
            futs.append(m.create_table(self.local_tc, self.executors))
        if not m.has_table(self.worker_tc.table_id):
            futs.append(m.create_table(self.worker_tc, self.executors))
            self._created_input = True
        for f in futs:
            f.get(timeout=3600)

    def cleanup(self) -> None:
        m = self.driver.et_master
        keep = bool(self.conf.get("offline_model_eval"))
        for tc, created in ((self.server_tc, True), (self.local_tc, True), (self.worker_tc, self._created_input)):
            if tc is None or not created or not m.has_table(tc.table_id):
                continue
            if keep and tc is self.worker_tc:
                continue
            try:
                m.get_table(tc.table_id).drop().get(timeout=120)
            except Exception:  # noqa: BLE001
                LOG.exception("drop of %s failed", tc.table_id)

    def run_offline_evaluation(self, dm: DolphinMaster) -> List[Any]:
        """Replay every per-epoch checkpoint (oldest first) and evaluate (ModelChkpManager.java:221-272)."""
        m = self.driver.et_master
        ck = dm.chkp_manager
        execs = self.driver.resource_pool.get_executors()
        state = {"tables": []}

        def drop_prev():
            for tid in state["tables"]:
                if m.has_table(tid):
                    m.get_table(tid).drop().get(timeout=120)
            state["tables"] = []

        def restore(i: int) -> bool:
            drop_prev()
            if i >= ck.num_chkps():
                return False
            m.create_table_from_chkp(ck.model_chkps[i], execs, table_id=self.server_tc.table_id).get(timeout=600)
            state["tables"].append(self.server_tc.table_id)
            if self.local_tc is not None and i < len(ck.local_chkps):
                m.create_table_from_chkp(ck.local_chkps[i], execs, table_id=self.local_tc.table_id).get(timeout=600)
                state["tables"].append(self.local_tc.table_id)
            if not m.has_table(self.worker_tc.table_id) and ck.input_chkp:
                m.create_table_from_chkp(ck.input_chkp, execs, table_id=self.worker_tc.table_id).get(timeout=600)
            return True
        out = dm.evaluate(execs, restore)
        drop_prev()
        if m.has_table(self.worker_tc.table_id):
            m.get_table(self.worker_tc.table_id).drop().get(timeout=120)
        self.driver.job_results[self.job_id + "/eval"] = (True, out)
        return out


class DolphinJobEntityBuilder:
    """Instantiated by JobServerDriver.on_submit with the deserialised job configuration."""

    def __init__(self, driver: JobServerDriver, conf: Configuration) -> None:
        self.driver, self.conf = driver, conf

    def build(self) -> DolphinJobEntity:
        app = load_app(self.conf["app"])
        if self.conf.get("backend", "auto") not in ("host", "device"):
            self.conf = self.conf.with_values(backend=self.driver.conf.get("backend", "host"))
        n = self.driver.next_job_count()
        job_id = f"{app.app_id}-{n}"
        server_tc, worker_tc, local_tc = build_tables_confs(app, self.conf, n)
        if not hasattr(self.driver, "metric_receiver"):
            from .master import ETDolphinMetricReceiver
            self.driver.metric_receiver = ETDolphinMetricReceiver()
            self.driver.et_master.metrics.receivers.append(self.driver.metric_receiver)
        return DolphinJobEntity(self.driver, self.conf, app, job_id, server_tc, worker_tc, local_tc)
