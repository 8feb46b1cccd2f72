"""Dolphin features beyond the plain training loop: offline model evaluation from per-epoch
checkpoints, warm start from a model checkpoint (-load_model / -model_path), SSP clock_slack,
schedulers, dashboard sink, cached model accessor, metric manager validation."""
import json
import os
import socket
import tempfile
import time
import urllib.request

import numpy as np
import pytest

from harmony_b200.dolphin.launcher import DolphinJobLauncher, ETDolphinLauncher
from harmony_b200.dolphin.master import MiniBatchController
from harmony_b200.dolphin.metrics import DolphinMetricManager
from harmony_b200.jobserver.client import JobServer
from harmony_b200.mlapps import mlr
from harmony_b200.utils.flags import Configuration

from helpers import write_mlr_sample


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _args(data, port, extra=()):
    return ["-input", data, "-max_num_epochs", "3", "-num_mini_batches", "4", "-num_worker_blocks", "4",
            "-num_server_blocks", "8", "-step_size", "0.05", "-classes", "4", "-features", "64",
            "-port", str(port)] + list(extra)


def test_offline_model_eval_and_warm_start():
    tmp = tempfile.mkdtemp(prefix="hbeval")
    data = write_mlr_sample(os.path.join(tmp, "sample_mlr"), n=160)
    test = write_mlr_sample(os.path.join(tmp, "sample_mlr_test"), n=40, seed=9)
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 2, "port": port, "chkp_temp_path": tmp + "/t",
                                  "chkp_commit_path": tmp + "/c"})).start()
    try:
        r = DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, _args(data, port, ["-offline_model_eval", "true",
                                                                                   "-test_data_path", test]), port=port)
        assert r.startswith("true")
        assert js.driver.wait_job("MLR-0", timeout=120)[0]
        # a checkpoint of the model table was taken every epoch
        chkps = sorted(os.listdir(tmp + "/t/executor-0")) if os.path.isdir(tmp + "/t/executor-0") else []
        assert sum(c.startswith("model_table0-") for c in chkps) >= 3, chkps
        # shutdown(kill=False) runs the deferred evaluation over all checkpoints, oldest first
        js.stop(kill=False)
        ok, evals = js.driver.job_results["MLR-0/eval"]
        per_worker = evals[0]
        assert len(per_worker) >= 3
        losses = [m["loss"] for m in per_worker]
        assert losses[-1] < losses[0], losses               # later epochs fit better
        assert all("testAvgAccuracy" in m for m in per_worker)
        # committed checkpoints can warm-start a new job
        committed = sorted(d for d in os.listdir(tmp + "/c") if d.startswith("model_table0-"))
        assert committed
        port2 = _free_port()
        js2 = JobServer(Configuration({"num_executors": 2, "port": port2, "chkp_temp_path": tmp + "/t2",
                                       "chkp_commit_path": tmp + "/c2"})).start()
        try:
            r = DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, _args(
                data, port2, ["-load_model", "true", "-model_path", os.path.join(tmp, "c", committed[-1]),
                              "-max_num_epochs", "1"]), port=port2)
            assert r.startswith("true")
            ok, res = js2.driver.wait_job("MLR-0", timeout=120)
            assert ok
            warm = np.mean([x["last_batch_loss"] for x in res])
        finally:
            js2.stop(kill=True)
        assert warm < 1.2 * np.mean(losses[-1:]) / 2 + 60, (warm, losses)
    finally:
        js.stop(kill=True)


def test_minibatch_controller_ssp_and_budget():
    sent = []

    class S:
        def send_mini_batch_control_msg(self, w, stop):
            sent.append((w, stop))
    c = MiniBatchController(num_epochs=1, num_mini_batches_in_epoch=6, slack=1, sender=S())
    c.on_sync("a"); c.on_sync("b")
    assert sent == [("a", False), ("b", False)]
    c.on_sync("a")                       # progress 2 <= min(1) + slack(1): released
    c.on_sync("a")                       # progress 3 > 1 + 1: blocked
    assert sent[-1] == ("a", False) and len(sent) == 3
    c.on_sync("b")                       # min progress rises to 2 -> a is released
    assert ("a", False) in sent[3:] and ("b", False) in sent[3:]
    n = len(sent)
    c.on_sync("b")                       # 6th sync = budget reached: runs its last batch
    assert sent[n] == ("b", False)
    c.on_sync("a")
    assert sent[-1] == ("a", True)       # budget exhausted: stop


@pytest.mark.parametrize("scheduler", ["FifoExclusiveScheduler", "PackingScheduler"])
def test_alternative_schedulers(scheduler):
    tmp = tempfile.mkdtemp(prefix="hbsched")
    data = write_mlr_sample(os.path.join(tmp, "sample_mlr"), n=80)
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 3, "port": port, "scheduler": scheduler})).start()
    try:
        extra = ["-max_num_epochs", "1"]
        for _ in range(3):
            assert DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, _args(data, port, extra), port=port).startswith("true")
        for j in ("MLR-0", "MLR-1", "MLR-2"):
            assert js.driver.wait_job(j, timeout=180)[0]
    finally:
        js.stop(kill=True)


def test_dashboard_receives_metrics_and_cached_accessor():
    from harmony_b200.dolphin.dashboard import DashboardServer
    tmp = tempfile.mkdtemp(prefix="hbdash")
    data = write_mlr_sample(os.path.join(tmp, "sample_mlr"), n=80)
    port = _free_port()
    res = ETDolphinLauncher.launch(mlr.APP, ["-input", data, "-max_num_epochs", "2", "-num_mini_batches", "4",
                                             "-num_worker_blocks", "4", "-num_server_blocks", "4", "-step_size", "0.05",
                                             "-classes", "4", "-features", "64", "-num_dolphin_workers", "1",
                                             "-num_servers", "1", "-dashboard_enabled", "true",
                                             "-dashboard_port", str(port), "-model_cache_enabled", "true"])
    assert res.ok and any(m.startswith("Dashboard:") for m in res.messages)


def test_dashboard_plot_endpoints():
    """The reference dashboard's URL surface (dashboard.py:47-151): POST / stores, /selectors lists what can be
    plotted, /plot returns {time: value}; the page is a self-contained SVG plotter (no plotly / CDN here)."""
    import json
    import urllib.parse
    import urllib.request
    from harmony_b200.dolphin.dashboard import DashboardConnector, DashboardServer
    srv = DashboardServer(port=0)
    try:
        c = DashboardConnector(srv.port)
        for i in range(6):
            c.post_metric("Worker-1" if i % 2 else "Server-0",
                          {"epoch": i, "loss": 2.0 / (i + 1), "metrics": {"data": {"acc": 0.1 * i}}, "name": "x"})
            time.sleep(0.02)
        deadline = time.time() + 5
        while time.time() < deadline and len(srv.rows()) < 6:
            time.sleep(0.05)
        base = f"http://127.0.0.1:{srv.port}"
        sel = json.loads(urllib.request.urlopen(base + "/selectors").read())
        assert sel["ids"] == ["Server-0", "Worker-1"] and "metrics.data.acc" in sel["y_axis"]["Worker-1"]
        assert "name" not in sel["y_axis"]["Worker-1"]                          # only numeric leaves are plottable
        series = json.loads(urllib.request.urlopen(base + "/plot?id=Worker-1&y=loss").read())
        assert sorted(series.values(), reverse=True) == [2.0 / 2, 2.0 / 4, 2.0 / 6]
        req = urllib.request.Request(base + "/plot", data=urllib.parse.urlencode({"id": "Server-0", "y": "epoch"}).encode())
        assert sorted(json.loads(urllib.request.urlopen(req).read()).values()) == [0.0, 2.0, 4.0]
        page = urllib.request.urlopen(base + "/").read().decode()
        assert "<svg" in page and "/selectors" in page and "/plot" in page
    finally:
        srv.close()


def test_metric_manager_validation_rules():
    mm = DolphinMetricManager(num_initial_to_skip=2)
    m = {"kind": "batch", "batch_time_sec": 1.0}
    assert not mm.store_worker_metric("w", 4, m)          # collection not started
    mm.start_metric_collection()
    assert not mm.store_worker_metric("w", 4, m) and not mm.store_worker_metric("w", 4, m)   # first N skipped
    assert mm.store_worker_metric("w", 4, m)
    mm.load_metric_validation_info({"w": 5}, {})
    assert not mm.store_worker_metric("w", 5, m)          # first metric after a reconfiguration is dropped
    assert not mm.store_worker_metric("w", 4, m)          # block count disagrees with the driver's view
    assert mm.store_worker_metric("w", 5, m)
    assert len(mm.get_worker_batch_metrics()["w"]) == 2


def test_job_fails_fast_when_one_worker_fails():
    """A failing worker must fail the JOB quickly (the others are parked at barriers), not the server."""
    from harmony_b200.dolphin.conf import ETDolphinConfiguration
    from harmony_b200.mlapps import examples
    import helpers  # noqa: F401
    app = ETDolphinConfiguration(app_id="Failing", trainer_class="helpers:FailingTrainer", input_parser="example",
                                 model_table_conf=examples._vec_table, parameters=examples.P.EXAMPLES)
    import sys
    sys.modules["failing_app"] = type(sys)("failing_app")
    sys.modules["failing_app"].APP = app
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 3, "port": port})).start()
    try:
        data = examples.write_dummy_input(30)
        conf_args = ["-input", data, "-max_num_epochs", "3", "-num_mini_batches", "6", "-num_worker_blocks", "6",
                     "-num_server_blocks", "4", "-num_keys", "4", "-port", str(port)]
        from harmony_b200.dolphin.launcher import parse_job_conf
        conf = parse_job_conf(app, conf_args).with_values(
            job_entity_builder="harmony_b200.dolphin.jobserver:DolphinJobEntityBuilder", app="failing_app:APP",
            victim="executor-1")
        t0 = time.time()
        from harmony_b200.jobserver.client import CommandSender
        assert CommandSender(port).send_job_submit_command(conf.serialize()).startswith("true")
        ok, result = js.driver.wait_job("Failing-0", timeout=60)
        assert not ok and "injected trainer failure" in str(result)
        assert time.time() - t0 < 40
        # the server survives and runs the next job
        assert DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, _args(
            write_mlr_sample(os.path.join(tempfile.mkdtemp(), "s"), n=60), port, ["-max_num_epochs", "1"]), port=port
        ).startswith("true")
        assert js.driver.wait_job("MLR-1", timeout=120)[0]
    finally:
        js.stop(kill=True)


def test_control_plane_group_rendezvous():
    """WorkerSideGroup / MasterSideRendezvous: rank-ordered all_gather over the job's message channel."""
    import threading
    from harmony_b200.dolphin.group import MasterSideRendezvous, WorkerSideGroup
    groups = {}
    rdv = MasterSideRendezvous(lambda w, body: groups[w].on_msg(body))
    for r in range(3):
        groups[f"w{r}"] = WorkerSideGroup(r, 3, f"w{r}", rdv.on_msg)
    out = {}

    def run(r):
        g = groups[f"w{r}"]
        out[r] = (g.all_gather_object({"r": r}), g.all_gather_object(r * 10))
        g.barrier()
    ts = [threading.Thread(target=run, args=(r,)) for r in (2, 0, 1)]
    [t.start() for t in ts]; [t.join(10) for t in ts]
    for r in range(3):
        assert out[r] == ([{"r": 0}, {"r": 1}, {"r": 2}], [0, 10, 20])


def test_trainer_constructor_can_rendezvous_with_its_peers():
    """Regression: the trainer is built after the tasklet is registered, so the master's group-gather replies reach
    a trainer that blocks in its constructor (ship-mode MLR exchanges peer buffer addresses there)."""
    from harmony_b200.dolphin.conf import ETDolphinConfiguration
    from harmony_b200.dolphin.launcher import parse_job_conf
    from harmony_b200.jobserver.client import CommandSender
    from harmony_b200.mlapps import examples
    import helpers  # noqa: F401
    import sys
    app = ETDolphinConfiguration(app_id="Rdv", trainer_class="helpers:RendezvousTrainer", input_parser="example",
                                 model_table_conf=examples._vec_table, parameters=examples.P.EXAMPLES)
    sys.modules["rdv_app"] = type(sys)("rdv_app")
    sys.modules["rdv_app"].APP = app
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 3, "port": port})).start()
    try:
        data = examples.write_dummy_input(30)
        conf = parse_job_conf(app, ["-input", data, "-max_num_epochs", "2", "-num_mini_batches", "6",
                                    "-num_worker_blocks", "6", "-num_server_blocks", "4", "-num_keys", "4",
                                    "-port", str(port)]).with_values(
            job_entity_builder="harmony_b200.dolphin.jobserver:DolphinJobEntityBuilder", app="rdv_app:APP")
        assert CommandSender(port).send_job_submit_command(conf.serialize()).startswith("true")
        ok, result = js.driver.wait_job("Rdv-0", timeout=60)
        assert ok, result
        for r in result:
            assert sorted(x[0] for x in r["group"]) == [0, 1, 2]
            assert sorted(x[1] for x in r["group"]) == ["executor-0", "executor-1", "executor-2"]
    finally:
        js.stop(kill=True)


def test_bad_input_fails_the_job_not_the_server(tmp_path):
    """A missing input file (driver-side split computation) and a corrupt record (executor-side bulk load) must
    fail the JOB with the cause; the control-plane future chain propagates the error instead of hanging."""
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 2, "port": port})).start()
    try:
        bad = tmp_path / "corrupt"
        bad.write_text("1 0:0.5 3:1.0\n2 notanumber:1\n" * 40)
        wide = write_mlr_sample(str(tmp_path / "wide"), n=40, features=64)       # feature ids up to 63, job says 8
        for i, (inp, needle) in enumerate([("/nonexistent/input", "cannot read input"), (str(bad), "TableLoad failed"),
                                           (wide, "index out of range")]):
            t0 = time.time()
            args = ["-input", inp, "-classes", "4", "-features", "8", "-max_num_epochs", "1", "-port", str(port)]
            assert DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, args, port=port).startswith("true")
            ok, result = js.driver.wait_job(f"MLR-{i}", timeout=30)
            assert not ok and needle in str(result), result
            assert time.time() - t0 < 20
        good = write_mlr_sample(str(tmp_path / "ok"), n=60)
        assert DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, _args(good, port, ["-max_num_epochs", "1"]),
                                             port=port).startswith("true")
        assert js.driver.wait_job("MLR-3", timeout=120)[0]
    finally:
        js.stop(kill=True)


def test_worker_without_data_idles_without_blocking_the_job(tmp_path):
    """One record, two executors: the worker that received no data must neither consume the mini-batch budget nor
    make the other worker wait at the job's task-unit barriers (job-server path: task units on)."""
    one = tmp_path / "one"
    one.write_text("1 0:0.5 3:1.0\n")
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 2, "port": port})).start()
    try:
        args = ["-input", str(one), "-classes", "4", "-features", "8", "-max_num_epochs", "3", "-num_mini_batches", "2",
                "-num_worker_blocks", "2", "-step_size", "0.1", "-port", str(port)]
        assert DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, args, port=port).startswith("true")
        ok, result = js.driver.wait_job("MLR-0", timeout=60)
        assert ok, result
        assert sorted(r["num_batches"] for r in result) == [0, 6]          # the whole budget ran where the data is
    finally:
        js.stop(kill=True)


@pytest.mark.parametrize("slack", [0, 1, 3])
def test_minibatch_controller_random_schedules_keep_the_ssp_invariants(slack):
    """Random interleavings of syncs, idle syncs, deletions, late joiners and hand-backs: the controller never grants
    more than the budget, never lets a worker run more than `slack` batches ahead of the slowest active one, and never
    leaves everybody blocked (liveness)."""
    import random
    from harmony_b200.dolphin.master import MiniBatchController

    class Sender:
        def __init__(self):
            self.msgs = []

        def send_mini_batch_control_msg(self, w, stop):
            self.msgs.append((w, stop))
    for seed in range(25):
        rng = random.Random(seed)
        snd = Sender()
        total_epochs, nmb = 4, 6
        c = MiniBatchController(total_epochs, nmb, slack, snd)
        workers = [f"w{i}" for i in range(4)]
        waiting, granted, stopped = set(), {w: 0 for w in workers}, set()
        idle_req = set()                                         # pending requests that were idle syncs (not grants)
        joined_late = False
        ever_closed = False
        for _ in range(400):
            # deliver control messages
            for w, stop in snd.msgs:
                if w in waiting:
                    waiting.discard(w)
                    if stop:
                        stopped.add(w)
                    elif w not in idle_req:
                        granted[w] += 1
                    idle_req.discard(w)
            snd.msgs.clear()
            active = [w for w in workers if w not in stopped]
            runnable = [w for w in active if w not in waiting]
            if not active:
                break
            assert runnable or c._closed, "every active worker is blocked: SSP deadlock"
            if not runnable:
                continue
            w = rng.choice(runnable)
            r = rng.random()
            if r < 0.04 and len(active) > 1:                 # a plan deletes the worker
                c.on_worker_deleted(w); stopped.add(w)
            elif r < 0.08:                                   # it idles for a while
                c.on_idle_sync(w); waiting.add(w); idle_req.add(w)
            elif r < 0.10 and not joined_late:               # a plan adds a worker
                workers.append("late"); granted["late"] = 0; joined_late = True
            elif r < 0.12 and granted[w] > 0:                # stopped right after being granted: hand the batch back
                c.on_return(w); granted[w] -= 1
            else:
                c.on_sync(w); waiting.add(w)
            prog = {x: p for x, p in c._progress.items()}
            ever_closed = ever_closed or c._closed           # the close releases every blocked request; a late
            if len(prog) > 1 and not ever_closed:            # hand-back may reopen the budget after that
                assert max(prog.values()) - min(prog.values()) <= slack + 1, prog
            assert c.num_mini_batches_done <= total_epochs * nmb + len(workers)
        assert sum(granted.values()) <= total_epochs * nmb


@pytest.mark.parametrize("scheduler,extra,expect_units", [
    ("SchedulerImpl", [], True),                                  # jobs share executors: phases are co-scheduled
    ("FifoExclusiveScheduler", [], False),                        # a job has the pool for itself: no barrier round trips
    ("SchedulerImpl", ["-task_unit_scheduling", "off"], False),
    ("FifoExclusiveScheduler", ["-task_unit_scheduling", "on"], True)])
def test_task_unit_scheduling_follows_the_scheduler_unless_overridden(scheduler, extra, expect_units):
    tmp = tempfile.mkdtemp(prefix="hbunits")
    data = write_mlr_sample(os.path.join(tmp, "sample_mlr"), n=80)
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 2, "port": port, "scheduler": scheduler})).start()
    try:
        assert DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, _args(data, port, ["-max_num_epochs", "1"] + extra),
                                             port=port).startswith("true")
        assert js.driver.wait_job("MLR-0", timeout=120)[0]
        assert (js.et_master.task_units.num_waits > 0) == expect_units
    finally:
        js.stop(kill=True)


@pytest.mark.parametrize("slack,want", [(0, 3), (1, 4), (2, 8)])
def test_minibatch_leases_keep_budget_clock_and_liveness(slack, want):
    """-mini_batch_grant_size: random interleavings of lease requests, idle syncs, deletions and partial hand-backs.
    Grants never exceed the budget, the budget is used up exactly (batches handed back after the close are granted
    again while somebody still asks), the
    first batch of a lease respects the SSP bound (a running worker is at most slack + want - 1 ahead of the slowest),
    leases taper at the end, and somebody can always run."""
    import random
    from harmony_b200.dolphin.master import MiniBatchController

    class Sender:
        def __init__(self):
            self.msgs = []

        def send_mini_batch_control_msg(self, w, stop, granted=1):
            self.msgs.append((w, stop, granted))
    for seed in range(25):
        rng = random.Random(1000 + seed)
        snd = Sender()
        total = 6 * 20
        c = MiniBatchController(6, 20, slack, snd)
        workers = [f"w{i}" for i in range(4)]
        waiting, stopped, idle_req = set(), set(), set()
        credits = {w: 0 for w in workers}
        ran = 0
        ever_closed = False
        handed_back_after_close = 0
        last_grants = []
        for _ in range(3000):
            for w, stop, g in snd.msgs:
                if w in waiting:
                    waiting.discard(w)
                    if stop:
                        stopped.add(w)
                    elif w not in idle_req:
                        assert 1 <= g <= want
                        credits[w] += g
                        last_grants.append(g)
                    idle_req.discard(w)
            snd.msgs.clear()
            active = [w for w in workers if w not in stopped]
            if not active:
                break
            runnable = [w for w in active if w not in waiting]
            assert runnable or c._closed, "every active worker is blocked: SSP deadlock"
            if not runnable:
                continue
            w = rng.choice(runnable)
            r = rng.random()
            if credits[w] > 0 and r < 0.9:                   # run one leased mini-batch
                credits[w] -= 1
                ran += 1
            elif credits[w] > 0 and r < 0.95:                # its data migrated away in the middle of the lease
                if c._closed:
                    handed_back_after_close += credits[w]
                c.on_return(w, credits[w]); credits[w] = 0
                c.on_idle_sync(w); waiting.add(w); idle_req.add(w)
            elif credits[w] > 0 and len(active) > 1:         # a plan deletes the worker in the middle of the lease
                if c._closed:
                    handed_back_after_close += credits[w]
                c.on_return(w, credits[w]); credits[w] = 0
                c.on_worker_deleted(w); stopped.add(w)
            elif credits[w] == 0:
                c.on_sync(w, want); waiting.add(w)
            prog = dict(c._progress)
            ever_closed = ever_closed or c._closed
            if len(prog) > 1 and not ever_closed:            # (the close releases every blocked request)
                lo = min(prog.values())
                # a running lease started within the SSP bound; a blocked request (its lease is counted already) may
                # reach one lease further
                assert all(p - lo <= slack + want - 1 for x, p in prog.items() if x not in c._blocked), (prog, c._blocked)
                assert max(prog.values()) - lo <= slack + 2 * want - 1, prog
            assert c.num_mini_batches_done <= total
        assert not [w for w in workers if w not in stopped], "job did not finish"
        # batches handed back after the close are granted again if somebody still asks (the budget reopens)
        assert total - handed_back_after_close <= ran <= total, (ran, handed_back_after_close, total)
        assert last_grants[-1] == 1                           # the tail of the budget is spread batch by batch


def test_return_message_distinguishes_one_batch_from_the_whole_grant():
    """Regression: a worker stopped with ONE unused leased batch must hand back 1, not 'my whole latest grant'."""
    from harmony_b200.dolphin.worker import WorkerSideMsgSender

    class Ctx:
        class executor:
            executor_id = "e0"
        sent = []

        def send_to_driver(self, payload, inline=False):
            self.sent.append(payload["dolphin"])
    ctx = Ctx()
    s = WorkerSideMsgSender(ctx, "job")
    s.send_mini_batch_return_msg(1)
    s.send_mini_batch_return_msg(2)
    s.send_mini_batch_return_msg()
    assert [m.get("count") for m in ctx.sent] == [1, 2, None]


def test_budget_reopens_when_the_last_batch_is_handed_back():
    """Regression (flaky live test): the worker that asked for the very last mini-batch is removed by a plan before
    running it -- the hand-back arrives after the close and must let a remaining worker run that batch."""
    from harmony_b200.dolphin.master import MiniBatchController

    class Sender:
        def __init__(self):
            self.msgs = []

        def send_mini_batch_control_msg(self, w, stop):
            self.msgs.append((w, stop))
    snd = Sender()
    c = MiniBatchController(1, 2, 10, snd)
    c.on_sync("w0")
    c.on_sync("w1")                       # the last unit of the budget: closes
    assert c._closed and snd.msgs == [("w0", False), ("w1", False)]
    c.on_return("w1")                     # w1 was stopped before running it
    c.on_worker_deleted("w1")
    assert not c._closed and c.num_mini_batches_done == 1
    c.on_sync("w0")                       # w0 finished its batch and asks again: it gets the handed-back one
    assert snd.msgs[-1] == ("w0", False) and c._closed and c.num_mini_batches_done == 2
    c.on_sync("w0")
    assert snd.msgs[-1] == ("w0", True)


def test_killed_executor_process_fails_the_job_fast():
    """Failure detection with real processes: SIGKILL one executor of a running stand-alone job (`-executor_mode
    process`); the launcher's monitor turns the exit into an executor failure, the job fails within seconds instead
    of hanging on its barriers, and no executor process is left behind."""
    import signal
    import subprocess
    import threading
    import time
    from harmony_b200.mlapps import examples
    data = examples.write_dummy_input(48)
    args = ["-input", data, "-num_dolphin_workers", "2", "-num_servers", "1", "-max_num_epochs", "200",
            "-num_mini_batches", "8", "-num_worker_blocks", "8", "-num_server_blocks", "8", "-num_keys", "20",
            "-vector_size", "4", "-compute_time_ms", "40", "-executor_mode", "process", "-backend", "host"]
    killed = {}

    def killer():
        deadline = time.time() + 60
        while time.time() < deadline:
            kids = subprocess.run(["pgrep", "-P", str(os.getpid()), "-f", "executor_main"], capture_output=True,
                                  text=True).stdout.split()
            if len(kids) >= 3:
                break
            time.sleep(0.2)
        time.sleep(2.0)                                   # let the training loop start
        killed["pid"], killed["at"] = int(kids[0]), time.time()
        os.kill(killed["pid"], signal.SIGKILL)            # an exact PID: one of this test's own executor children
    threading.Thread(target=killer, daemon=True).start()
    res = ETDolphinLauncher.launch(examples.ADD_VECTOR, args)
    done = time.time()
    assert killed, "the executors never came up"
    assert not res.ok
    assert done - killed["at"] < 20, "the job did not fail fast"
    time.sleep(0.5)
    left = subprocess.run(["pgrep", "-P", str(os.getpid()), "-f", "executor_main"], capture_output=True,
                          text=True).stdout.split()
    assert not left, left


def test_more_concurrent_jobs_than_tasklet_slots_never_deadlock():
    """Regression (rare hang with >= 4 concurrent jobs): every job puts a server tasklet (idle latch) and a worker
    tasklet on every executor.  The idle servers take no execution slot, and a job is admitted only when all of its
    executors have a free slot (all-or-nothing), so jobs beyond the slot count queue instead of deadlocking at their
    INIT barriers."""
    tmp = tempfile.mkdtemp(prefix="hbslots")
    data = write_mlr_sample(os.path.join(tmp, "sample_mlr"), n=60)
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 2, "port": port})).start()
    try:
        assert js.driver.dispatcher.slots_per_executor == 4
        n = 7
        for _ in range(n):
            assert DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, _args(data, port, ["-max_num_epochs", "2"]),
                                                 port=port).startswith("true")
        for j in range(n):
            ok, res = js.driver.wait_job(f"MLR-{j}", timeout=120)
            assert ok, res
        assert all(v == 0 for v in js.driver.dispatcher._jobs_on.values())
    finally:
        js.stop(kill=True)


def test_parked_server_tasklets_take_no_execution_slot_and_admission_is_all_or_nothing():
    """The two halves of the fix, deterministically: (1) with every slot's worth of idle server tasklets parked on an
    executor, a worker tasklet still starts; (2) the dispatcher never lets a job hold a slot on one executor while it
    waits for another."""
    import threading
    import time
    from harmony_b200.et import local_cluster
    from harmony_b200.et.config import TaskletConfiguration
    master, execs, launcher = local_cluster(1)
    try:
        ex = execs[0]
        slots = launcher.executors[ex.id].conf.num_tasklets
        parked = [ex.submit_tasklet(TaskletConfiguration(f"srv{i}", "harmony_b200.dolphin.worker:ServerTasklet", None, {}))
                  for i in range(slots + 2)]
        time.sleep(0.3)
        rt = ex.submit_tasklet(TaskletConfiguration("work", "helpers:EchoTasklet", None, {"value": 7}))
        assert rt.get_task_result(timeout=10).result == 7
        for p in parked:
            p.stop()
    finally:
        master.close()

    from harmony_b200.jobserver.driver import JobDispatcher

    class Drv:
        conf = {"executor_num_tasklets": 4}

    class Ex:
        def __init__(self, i):
            self.id = i
    d = JobDispatcher(Drv())
    d.slots_per_executor = 1
    a, b = Ex("a"), Ex("b")
    d._admit([a])                                   # job 1 runs on a
    got = []
    t = threading.Thread(target=lambda: (d._admit([a, b]), got.append(1)), daemon=True)   # job 2 wants a AND b
    t.start()
    time.sleep(0.2)
    assert not got and d._jobs_on.get("b", 0) == 0  # it waits WITHOUT holding b
    d._admit([b])                                   # so job 3 (only b) can start
    d._release([a])
    time.sleep(0.2)
    assert not got                                  # b is taken now: job 2 still waits, still holds nothing
    d._release([b])
    t.join(5)
    assert got and d._jobs_on == {"a": 1, "b": 1}


def test_worker_that_trains_on_no_batch_does_not_fail_the_job_and_leases_yield_to_unit_barriers(tmp_path):
    """Two soak regressions in one job: (1) fewer mini-batches than workers -- the worker that ran no batch has no
    'last batch' summary to report (Lasso / GBT summaries used to raise); (2) -mini_batch_grant_size with task-unit
    co-scheduling on (shared-pool scheduler): leases are ignored (they would deadlock the per-unit barriers), the job
    still completes."""
    from harmony_b200.common.samples import write_lasso_sample
    from harmony_b200.mlapps import lasso
    data = write_lasso_sample(str(tmp_path / "sample_lasso"))
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 3, "port": port})).start()
    try:
        args = ["-input", data, "-features", "10", "-features_per_partition", "5", "-lambda", "0.5", "-step_size", "1.0",
                "-max_num_epochs", "3", "-num_mini_batches", "2", "-num_worker_blocks", "2", "-mini_batch_grant_size", "3",
                "-port", str(port)]
        assert DolphinJobLauncher.submit_job(lasso.APP_PATH, lasso.APP, args, port=port).startswith("true")
        ok, res = js.driver.wait_job("Lasso-0", timeout=120)
        assert ok, res
        assert sorted(r["num_batches"] for r in res)[0] == 0          # somebody trained on nothing ...
        assert sum(r["num_batches"] for r in res) == 6                # ... and the budget was still used up exactly
        assert js.et_master.task_units.num_waits > 0                  # unit barriers were on (leases were not)
    finally:
        js.stop(kill=True)


def test_standalone_offline_model_evaluation_replays_every_epoch_checkpoint():
    """`run_mlr.sh ... -offline_model_eval true`: after training the stand-alone launcher replays the per-epoch
    checkpoints oldest first and evaluates each (one "Model eval progress" line and one metric set per epoch)."""
    tmp = tempfile.mkdtemp(prefix="hbeval")
    data = write_mlr_sample(os.path.join(tmp, "sample_mlr"), n=80)
    test = write_mlr_sample(os.path.join(tmp, "sample_mlr_test"), n=40, seed=7)
    res = ETDolphinLauncher.launch(mlr.APP, ["-input", data, "-test_data_path", test, "-max_num_epochs", "3",
                                             "-num_mini_batches", "4", "-num_worker_blocks", "4", "-num_server_blocks", "8",
                                             "-step_size", "0.05", "-classes", "4", "-features", "64",
                                             "-num_dolphin_workers", "2", "-num_servers", "1",
                                             "-offline_model_eval", "true",
                                             "-chkp_temp_path", tmp + "/t", "-chkp_commit_path", tmp + "/c"])
    assert res.ok, res.messages
    assert [m for m in res.messages if m.startswith("Model eval progress")] == \
        [f"Model eval progress: [{i}], JobId: MLR-0" for i in (1, 2, 3)]
    assert res.eval_results is not None and len(res.eval_results) == 2


def test_fused_device_trainers_get_leases_by_default_and_one_flush_per_lease():
    """Job-server GPU path (round 2): device trainers queue a whole lease of mini-batches on their stream; the worker
    loop syncs the stream (`trainer.flush()`) once before it reports progress -- not once per mini-batch -- and
    `-mini_batch_grant_size 0` (auto) leases 4 mini-batches to a fused trainer that can flush."""
    from harmony_b200.dolphin.conf import ETDolphinConfiguration
    from harmony_b200.mlapps import examples
    import helpers  # noqa: F401
    app = ETDolphinConfiguration(app_id="Flush", trainer_class="helpers:FlushCountingTrainer", input_parser="example",
                                 model_table_conf=examples._vec_table, parameters=examples.P.EXAMPLES)
    data = examples.write_dummy_input(32)
    res = ETDolphinLauncher.launch(app, ["-input", data, "-max_num_epochs", "2", "-num_mini_batches", "8",
                                         "-num_worker_blocks", "8", "-num_server_blocks", "4", "-num_keys", "4",
                                         "-num_dolphin_workers", "1", "-num_servers", "1",
                                         "-task_unit_scheduling", "off"])
    assert res.ok, res.messages
    ev = res.results[0]["events"]
    assert ev.count("batch") == 16
    # leases of 4: a flush before every sync that follows a lease, and one at each epoch end -- never between the
    # mini-batches of a lease
    runs = [len(r) for r in "".join("b" if e == "batch" else "|" for e in ev).split("|") if r]
    assert max(runs) == 4 and all(r in (1, 2, 3, 4) for r in runs), ev
    assert 4 <= ev.count("flush") <= 10, ev
