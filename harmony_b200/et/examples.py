"""Self-validating ET examples (double as integration fixtures, like the reference's
services/et/.../examples/{simple,addinteger,tableaccess,checkpoint,load,metric,plan,userservice}
run by ExampleTest / TableAccessTest / PlanExecutorTest / DataLoadTest).  Every `run_*` builds an
in-process cluster, drives tasklets that assert inside the executors (a failed assertion fails the
tasklet and therefore the run) and returns True on success.
"""
from __future__ import annotations

import os
import tempfile
import threading
import time
from typing import Any, Dict, List


from . import ExecutorConfiguration, TableConfiguration, local_cluster
from .config import TaskletConfiguration
from .plan import (AllocateOp, AssociateOp, CreateOp, DeallocateOp, DropOp, ETPlan, MoveOp, PlanExecutor, StartOp,
                   StopOp, SubscribeOp, UnassociateOp, UnsubscribeOp)
from .tasklet import Tasklet

_HERE = "harmony_b200.et.examples"


def _run_tasklets(execs, cls: str, params: Dict[str, Any], timeout: float = 120.0) -> List[Any]:
    rts = [e.submit_tasklet(TaskletConfiguration(f"{cls}-{e.id}", f"{_HERE}:{cls}", None, dict(params, index=i,
                                                                                              n=len(execs))))
           for i, e in enumerate(execs)]
    results = [rt.get_task_result(timeout) for rt in rts]
    bad = [r for r in results if not r.is_success]
    if bad:
        raise RuntimeError("; ".join(f"{r.executor_id}: {r.error}" for r in bad))
    return [r.result for r in results]


# ---- simple: put / get / scan / migrate / drop ------------------------------------------------
class SimplePutTasklet(Tasklet):
    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        for k in range(p["index"], 100, p["n"]):
            assert t.put(k, f"v{k}").get(10) is None
        return True


class SimpleScanTasklet(Tasklet):
    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        for k in range(100):
            assert t.get(k).get(10) == f"v{k}", k
        return t.get_local_tablet().get_num_data_items()


def run_simple() -> bool:
    master, execs, _ = local_cluster(3)
    try:
        conf = TableConfiguration(table_id="simple", num_total_blocks=16, update_function="void")
        master.create_table(conf, execs[:2]).get(30)
        table = master.get_table("simple")
        table.subscribe([execs[2]]).get(30)
        _run_tasklets(execs, "SimplePutTasklet", {"table_id": "simple"})
        counts = _run_tasklets(execs, "SimpleScanTasklet", {"table_id": "simple"})
        assert sum(counts) == 100 and counts[2] == 0
        assert table.move_blocks(execs[0].id, execs[1].id, 4).get(60).is_completed
        counts2 = _run_tasklets(execs, "SimpleScanTasklet", {"table_id": "simple"})
        assert sum(counts2) == 100 and counts2[1] > counts[1]
        table.drop().get(30)
        return True
    finally:
        master.close()


# ---- addinteger: PS semantics (many workers update the same keys) -----------------------------
class AddIntegerTasklet(Tasklet):
    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        for _ in range(p["num_updates"]):
            for k in range(p["num_keys"]):
                t.update_no_reply(k, p["delta"])
        return True


class AddIntegerValidateTasklet(Tasklet):
    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        expected = p["delta"] * p["num_updates"] * p["num_workers"]
        for _ in range(100):
            vals = t.multi_get_or_init(list(range(p["num_keys"]))).get(30)
            if all(v == expected for v in vals):
                return True
            time.sleep(0.1)
        raise AssertionError(f"expected {expected}, got {vals}")


def run_addinteger(num_workers: int = 2, num_servers: int = 2) -> bool:
    master, execs, _ = local_cluster(num_workers + num_servers)
    try:
        servers, workers = execs[:num_servers], execs[num_servers:]
        conf = TableConfiguration(table_id="addint", num_total_blocks=8, update_function="add")
        master.create_table(conf, servers).get(30)
        master.get_table("addint").subscribe(workers).get(30)
        params = {"table_id": "addint", "num_updates": 20, "num_keys": 10, "delta": 3, "num_workers": num_workers}
        _run_tasklets(workers, "AddIntegerTasklet", params)
        _run_tasklets(workers[:1], "AddIntegerValidateTasklet", params)
        return True
    finally:
        master.close()


# ---- tableaccess: every op x local/remote x ordered/hashed ------------------------------------
class TableAccessTasklet(Tasklet):
    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        base = 1000 * (p["index"] + 1)
        keys = [base + i for i in range(20)]
        assert t.multi_put(keys, [k * 2 for k in keys]).get(10) is not None or True
        assert t.multi_get(keys).get(10) == [k * 2 for k in keys]
        assert t.put_if_absent(keys[0], -1).get(10) == keys[0] * 2
        assert t.put_if_absent(base + 500, 7).get(10) is None
        assert t.update(keys[1], 5).get(10) == keys[1] * 2 + 5
        t.update_no_reply(keys[1], 5); t.put_no_reply(base + 501, 9); t.put_if_absent_no_reply(base + 501, 10)
        t.multi_update_no_reply(keys[2:4], [1, 1])
        time.sleep(0.3)
        assert t.get(keys[1]).get(10) == keys[1] * 2 + 10 and t.get(base + 501).get(10) == 9
        assert t.multi_get(keys[2:4]).get(10) == [keys[2] * 2 + 1, keys[3] * 2 + 1]
        assert t.get_or_init(base + 700).get(10) == 0
        assert t.multi_get_or_init([base + 701, base + 702]).get(10) == [0, 0]
        assert t.remove(keys[5]).get(10) == keys[5] * 2 and t.get(keys[5]).get(10) is None
        t.remove_no_reply(keys[6]); time.sleep(0.2)
        assert t.get(keys[6]).get(10) is None
        return True


def run_tableaccess() -> bool:
    for ordered in (False, True):
        master, execs, _ = local_cluster(3)
        try:
            conf = TableConfiguration(table_id=f"ta{ordered}", num_total_blocks=8, is_ordered=ordered,
                                      update_function="add")
            master.create_table(conf, execs[:2]).get(30)
            master.get_table(conf.table_id).subscribe([execs[2]]).get(30)
            _run_tasklets(execs, "TableAccessTasklet", {"table_id": conf.table_id})
        finally:
            master.close()
    return True


# ---- checkpoint ---------------------------------------------------------------------------------
def run_checkpoint() -> bool:
    tmp = tempfile.mkdtemp(prefix="hbex")
    ec = ExecutorConfiguration(chkp_temp_path=tmp + "/t", chkp_commit_path=tmp + "/c")
    master, execs, launcher = local_cluster(3, ec, chkp_temp_path=tmp + "/t", chkp_commit_path=tmp + "/c")
    try:
        conf = TableConfiguration(table_id="ck", num_total_blocks=8, update_function="void")
        master.create_table(conf, execs).get(30)
        _run_tasklets(execs, "SimplePutTasklet", {"table_id": "ck"})
        chkp = master.get_table("ck").checkpoint().get(60)
        sampled = master.get_table("ck").checkpoint(0.5).get(60)
        master.get_table("ck").drop().get(30)
        master.create_table_from_chkp(chkp, execs[:2], table_id="ck").get(60)
        _run_tasklets(execs[:2], "SimpleScanTasklet", {"table_id": "ck"})
        master.create_table_from_chkp(sampled, execs[:2], table_id="ck_sampled").get(60)
        n = sum(ex.tables.get_table("ck_sampled").get_local_tablet().get_num_data_items()
                for eid, ex in launcher.executors.items() if eid in (execs[0].id, execs[1].id))
        assert 10 < n < 90, n
        return True
    finally:
        master.close()


# ---- load -----------------------------------------------------------------------------------------
def run_load() -> bool:
    tmp = tempfile.mkdtemp(prefix="hbex")
    path = os.path.join(tmp, "data")
    with open(path, "w") as f:
        for i in range(500):
            f.write(f"line-{i}\n")
    master, execs, launcher = local_cluster(3)
    try:
        conf = TableConfiguration(table_id="load", num_total_blocks=12, is_ordered=True, is_mutable=False,
                                  update_function="void", input_path=path, data_parser="default")
        master.create_table(conf, execs).get(60)
        vals = []
        for ex in launcher.executors.values():
            vals += [v for _, v in ex.tables.get_table("load").get_local_tablet().get_data_iterator()]
        assert sorted(vals) == sorted(f"line-{i}" for i in range(500))
        return True
    finally:
        master.close()


# ---- metric ---------------------------------------------------------------------------------------
class MetricTasklet(Tasklet):
    def run(self):
        mc = self.context.metric_collector
        t = self.context.table_accessor.get_table("metric")
        t.put(self.context.params["index"], 1).get(10)
        t.get(12345).get(10)
        mc.add_custom_metric({"who": self.context.executor.executor_id, "value": 42})
        mc.flush()
        return True


def run_metric() -> bool:
    from .driver import MetricReceiver
    master, execs, _ = local_cluster(2)
    got: List[Dict[str, Any]] = []

    class R(MetricReceiver):
        def on_metric_report(self, report):
            got.append(report)
    master.metrics.receivers.append(R())
    try:
        master.create_table(TableConfiguration(table_id="metric", num_total_blocks=4, update_function="void"),
                            execs).get(30)
        for e in execs:
            master.metrics.start_metric_collection(e.id, 0)
        _run_tasklets(execs, "MetricTasklet", {})
        time.sleep(0.3)
        customs = [c for r in got for c in r["custom_metrics"]]
        assert len(customs) == 2 and all(c["value"] == 42 for c in customs)
        assert all(r["table_to_num_blocks"]["metric"] == 2 for r in got)
        return True
    finally:
        master.close()


# ---- plan -----------------------------------------------------------------------------------------
def run_plan() -> bool:
    master, execs, launcher = local_cluster(2)
    try:
        pe = PlanExecutor(master)
        conf = TableConfiguration(table_id="plan_t", num_total_blocks=8, update_function="add")
        plan = ETPlan()
        create = plan.add_op(CreateOp([e.id for e in execs], conf))
        alloc = plan.add_op(AllocateOp("V0"))
        assoc = plan.add_op(AssociateOp("V0", "plan_t"), [create, alloc])
        move = plan.add_op(MoveOp(execs[0].id, "V0", "plan_t", 2), [assoc])
        start = plan.add_op(StartOp("V0", TaskletConfiguration("plan-put", f"{_HERE}:SimplePutTasklet", None,
                                                               {"table_id": "plan_t", "index": 0, "n": 1})), [move])
        vmap = pe.execute(plan).get(120)
        new_id = vmap["V0"]
        assert len(master.get_table("plan_t").get_partition_info()[new_id]) == 2
        time.sleep(0.5)
        plan2 = ETPlan()
        stop = plan2.add_op(StopOp(new_id))
        mv = plan2.add_op(MoveOp(new_id, execs[1].id, "plan_t", 2), [stop])
        un = plan2.add_op(UnassociateOp(new_id, "plan_t"), [mv])
        sub = plan2.add_op(SubscribeOp(new_id, "plan_t"), [un])
        unsub = plan2.add_op(UnsubscribeOp(new_id, "plan_t"), [sub])
        de = plan2.add_op(DeallocateOp(new_id), [unsub])
        plan2.add_op(DropOp("plan_t"), [de])
        pe.execute(plan2).get(120)
        assert not master.has_table("plan_t") and new_id not in master.executors
        return True
    finally:
        master.close()


# ---- centcomm (user service) --------------------------------------------------------------------------
class CentCommTasklet(Tasklet):
    def run(self):
        from ..common.centcomm import SlaveSideCentComm
        cc = SlaveSideCentComm(self.context.executor)
        ev = threading.Event()
        cc.register_handler("example", lambda src, data: ev.set() if data == "pong" else None)
        cc.send("example", "ping")
        assert ev.wait(10), "no pong from the driver"
        return True


def run_centcomm() -> bool:
    from ..common.centcomm import MasterSideCentComm
    master, execs, _ = local_cluster(2)
    try:
        cc = MasterSideCentComm(master)
        cc.register_handler("example", lambda src, data: cc.send("example", src, "pong") if data == "ping" else None)
        _run_tasklets(execs, "CentCommTasklet", {})
        return True
    finally:
        master.close()


ALL = {"simple": run_simple, "addinteger": run_addinteger, "tableaccess": run_tableaccess,
       "checkpoint": run_checkpoint, "load": run_load, "metric": run_metric, "plan": run_plan,
       "centcomm": run_centcomm}


def main(argv: List[str]) -> int:
    name = argv[0] if argv else "simple"
    ok = ALL[name]()
    print(f"LauncherStatus: {'COMPLETED' if ok else 'FAILED'}")
    return 0 if ok else 1
