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
    """simple example: every executor puts its share of 100 values."""

    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        for k in range(p["index"], 100, p["n"]):
            assert t.put(k, f"v{k}").get(10) is None
        return True


class SimpleScanTasklet(Tasklet):
    """simple example: count the locally stored values."""

    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        for k in range(100):
            assert t.get(k).get(10) == f"v{k}", k
        return t.get_local_tablet().get_num_data_items()


def run_simple(table_input: str = "") -> bool:
    master, execs, launcher = local_cluster(3)
    try:
        if table_input:          # SimpleET -table_input: bulk load a file into an (ordered, immutable) table first
            master.create_table(TableConfiguration(table_id="simple_input", num_total_blocks=16, is_ordered=True,
                                                   is_mutable=False, update_function="void", input_path=table_input,
                                                   data_parser="default"), execs[:2]).get(60)
            loaded = sum(ex.tables.get_table("simple_input").get_local_tablet().get_num_data_items()
                         for ex in list(launcher.executors.values())[:2])
            with open(table_input) as f:
                assert loaded == sum(1 for ln in f if ln.strip() and not ln.startswith("#")), loaded
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
    """addinteger example: num_updates round-robin updates of delta (UpdaterTask)."""

    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        if p.get("metric_flush_period_ms"):
            self.context.metric_collector.start(int(p["metric_flush_period_ms"]))
        for i in range(p["num_updates"]):          # UpdaterTask: round-robin over [start_key, start_key + num_keys)
            t.update_no_reply(p["start_key"] + i % p["num_keys"], p["delta"])
        return True


class AddIntegerValidateTasklet(Tasklet):
    """addinteger example: waits until every key holds the expected sum (ValidatorTask.java:74)."""

    def run(self):
        p = self.context.params
        t = self.context.table_accessor.get_table(p["table_id"])
        # ValidatorTask.java:74
        expected = (p["num_workers"] * p["num_updates"] // p["num_keys"]) * (p["delta"] * p["coefficient"])
        for _ in range(100):
            vals = t.multi_get_or_init([p["start_key"] + k for k in range(p["num_keys"])]).get(30)
            if all(v == expected for v in vals):
                return True
            time.sleep(0.1)
        raise AssertionError(f"expected {expected}, got {vals}")


def run_addinteger(num_workers: int = 2, num_servers: int = 2, num_updates: int = 120, num_keys: int = 8,
                   start_key: int = 0, delta_value: int = 3, coefficient: int = 2,
                   metric_flush_period_ms: int = 0) -> bool:
    if num_updates % num_keys:
        raise ValueError("num_updates must be a multiple of num_keys (every key gets the same number of updates)")
    master, execs, _ = local_cluster(num_workers + num_servers)
    try:
        servers, workers = execs[:num_servers], execs[num_servers:]
        conf = TableConfiguration(table_id="addint", num_total_blocks=8,
                                  update_function={"name": "add", "params": {"coefficient": coefficient}})
        master.create_table(conf, servers).get(30)
        master.get_table("addint").subscribe(workers).get(30)
        params = {"table_id": "addint", "num_updates": num_updates, "num_keys": num_keys, "start_key": start_key,
                  "delta": delta_value, "coefficient": coefficient, "num_workers": num_workers,
                  "metric_flush_period_ms": metric_flush_period_ms}
        _run_tasklets(workers, "AddIntegerTasklet", params)
        _run_tasklets(workers[:1], "AddIntegerValidateTasklet", params)
        return True
    finally:
        master.close()


# ---- tableaccess: every op x local/remote x ordered/hashed ------------------------------------
class TableAccessTasklet(Tasklet):
    """tableaccess example: every op, local and remote, ordered and hashed tables."""

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
def run_load(kv_data_path: str = "", none_key_data_path: str = "") -> bool:
    """LoadET: one table bulk-loaded from `key value` lines (exist-key loader), one from value lines (none-key
    loader, keys generated from the local blocks' ranges)."""
    from ..common.dataloader import DataParser, register_parser

    class KVLineParser(DataParser):
        has_key = True

        def parse_line(self, line: str):
            k, _, v = line.partition(" ")
            return int(k), v
    register_parser("kv_line", KVLineParser)
    tmp = tempfile.mkdtemp(prefix="hbex")
    if not none_key_data_path:
        none_key_data_path = os.path.join(tmp, "data")
        with open(none_key_data_path, "w") as f:
            for i in range(500):
                f.write(f"line-{i}\n")
    if not kv_data_path:
        kv_data_path = os.path.join(tmp, "kv")
        with open(kv_data_path, "w") as f:
            for i in range(300):
                f.write(f"{i * 7} value-{i}\n")

    def lines(path):
        with open(path) as f:
            return [ln.rstrip("\n") for ln in f if ln.strip() and not ln.startswith("#")]
    master, execs, launcher = local_cluster(3)
    try:
        conf = TableConfiguration(table_id="load", num_total_blocks=12, is_ordered=True, is_mutable=False,
                                  update_function="void", input_path=none_key_data_path, data_parser="default")
        master.create_table(conf, execs).get(60)
        vals = []
        for ex in launcher.executors.values():
            vals += [v for _, v in ex.tables.get_table("load").get_local_tablet().get_data_iterator()]
        assert sorted(vals) == sorted(lines(none_key_data_path))
        kv = TableConfiguration(table_id="load_kv", num_total_blocks=12, update_function="void",
                                input_path=kv_data_path, data_parser="kv_line", bulk_loader="exist_key")
        master.create_table(kv, execs).get(60)
        got = {}
        for ex in launcher.executors.values():
            got.update(dict(ex.tables.get_table("load_kv").get_local_tablet().get_data_iterator()))
        want = dict((int(ln.partition(" ")[0]), ln.partition(" ")[2]) for ln in lines(kv_data_path))
        assert got == want, (len(got), len(want))
        return True
    finally:
        master.close()


# ---- metric ---------------------------------------------------------------------------------------
class MetricTasklet(Tasklet):
    """MetricET's task: record a custom timestamp metric every `record_ms`, flush manually every `manual_ms`,
    for `duration_ms` (the collector also flushes by itself every automatic period)."""

    def run(self):
        p = self.context.params
        mc = self.context.metric_collector
        t = self.context.table_accessor.get_table("metric")
        t.put(p["index"], 1).get(10)
        t.get(12345).get(10)
        t0 = last_flush = time.time()
        n = 0
        while True:
            mc.add_custom_metric({"who": self.context.executor.executor_id, "value": 42, "timestamp": time.time()})
            n += 1
            now = time.time()
            if (now - last_flush) * 1000.0 >= p["manual_ms"]:
                mc.flush()
                last_flush = now
            if (now - t0) * 1000.0 >= p["duration_ms"]:
                break
            time.sleep(p["record_ms"] / 1000.0)
        mc.flush()
        return n


def run_metric(metric_automatic_flush_period_ms: int = 200, metric_manual_flush_period_ms: int = 100,
               custom_metric_record_period_ms: int = 20, task_duration_ms: int = 400) -> bool:
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
            master.metrics.start_metric_collection(e.id, metric_automatic_flush_period_ms)
        recorded = _run_tasklets(execs, "MetricTasklet",
                                 {"manual_ms": metric_manual_flush_period_ms, "record_ms": custom_metric_record_period_ms,
                                  "duration_ms": task_duration_ms}, timeout=120 + task_duration_ms / 1000.0)
        time.sleep(0.3)
        for e in execs:
            master.metrics.stop_metric_collection(e.id)
        customs = [c for r in got for c in r["custom_metrics"]]
        assert len(customs) == sum(recorded) and all(c["value"] == 42 for c in customs), (len(customs), recorded)
        assert len(got) >= 2 * len(execs)                       # several flushes per executor
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
    """centcomm example: ping the driver, wait for pong."""

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


def run_linecount(inputs: str = "") -> bool:
    """The dataloader example (common/dataloader/examples/LineCountingDriver): every executor counts the lines of
    its exact byte-range splits; the total must equal the number of lines of the inputs."""
    from ..common.dataloader import get_splits, read_split
    if not inputs:
        tmp = tempfile.mkdtemp(prefix="hbex")
        inputs = os.path.join(tmp, "lines")
        with open(inputs, "w") as f:
            for i in range(1234):
                f.write(f"{i} {'x' * (i % 17)}\n")
    total = want = 0
    for path in [p for p in inputs.split(",") if p]:
        with open(path) as f:
            want += sum(1 for _ in f)
        for split in get_splits(path, 7):
            total += sum(1 for _ in read_split(split))
    assert total == want, (total, want)
    print(f"Total {total} lines")
    return True


ALL = {"simple": run_simple, "addinteger": run_addinteger, "tableaccess": run_tableaccess,
       "checkpoint": run_checkpoint, "load": run_load, "metric": run_metric, "plan": run_plan,
       "centcomm": run_centcomm, "linecount": run_linecount}


def main(argv: List[str]) -> int:
    """`<example> [-flag value ...]` with the reference examples' flag names (dolphin/params.py: ET_EXAMPLES)."""
    import inspect
    from ..dolphin import params as P
    name = argv[0] if argv and not argv[0].startswith("-") else "simple"
    rest = argv[1:] if argv and not argv[0].startswith("-") else argv
    conf = P.registry(P.ET_EXAMPLES, [p for p in P.EXAMPLES if p.name in ("num_keys", "start_key")],
                      [p for p in P.DOLPHIN if p.name == "num_servers"]).parse(rest, allow_unknown=False)[0]
    fn = ALL[name]
    given = {a.lstrip("-") for a in rest[0::2]}
    kwargs = {k: conf[k] for k in inspect.signature(fn).parameters if k in given}
    ok = fn(**kwargs)
    print(f"LauncherStatus: {'COMPLETED' if ok else 'FAILED'}")
    return 0 if ok else 1
