"""Control-plane ceiling of the job-server / stand-alone Dolphin path: W workers run a trainer that does nothing, so
what is timed is one mini-batch's host messaging (SSP sync + grant, piggybacked progress, throttled batch metrics,
task-unit bookkeeping) through ONE Python driver.  CPU only -- the SPMD path (bench.py) has no per-step host messages.

    python benchmarks/bench_control_plane.py --workers 8 --epochs 20 --batches 50 [--profile]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from harmony_b200.dolphin.conf import ETDolphinConfiguration  # noqa: E402
from harmony_b200.dolphin.launcher import ETDolphinLauncher  # noqa: E402
from harmony_b200.dolphin.worker import Trainer  # noqa: E402
from harmony_b200.mlapps import examples  # noqa: E402


class NoOpTrainer(Trainer):
    def __init__(self, params, model_accessor, **kw):
        pass

    def init_global_settings(self): pass
    def set_mini_batch_data(self, batch): pass
    def pull_model(self): pass
    def local_compute(self): pass
    def push_update(self): pass
    def on_epoch_finished(self, epoch): pass
    def evaluate_model(self, *a, **k): return {}


APP = ETDolphinConfiguration(app_id="NoOp", trainer_class="bench_control_plane:NoOpTrainer", input_parser="example",
                             model_table_conf=examples._vec_table, parameters=examples.P.EXAMPLES)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--batches", type=int, default=50, help="mini-batches per worker per epoch")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--executor_mode", default="inproc", choices=["inproc", "process"],
                    help="with --jobserver: executors as threads of the driver process or as one process each (TCP)")
    ap.add_argument("--jobserver", type=int, default=0, metavar="J",
                    help="submit J concurrent copies of the job to one job server (task-unit co-scheduling on)")
    ap.add_argument("--extra", default="", help="comma-separated extra job flags, e.g. --extra=-clock_slack,2")
    a = ap.parse_args()
    sys.modules.setdefault("bench_control_plane", sys.modules[__name__])
    data = examples.write_dummy_input(a.workers * a.batches * 2)
    argv = ["-input", data, "-max_num_epochs", str(a.epochs), "-num_mini_batches", str(a.batches),
            "-num_dolphin_workers", str(a.workers), "-num_servers", "1", "-num_worker_blocks", str(a.batches),
            "-num_server_blocks", "4", "-num_keys", "4", "-backend", "host"] + [x for x in a.extra.split(",") if x]
    if a.jobserver:
        return run_jobserver(a, argv)
    prof = None
    if a.profile:
        import cProfile
        import threading
        prof = cProfile.Profile()
        threading.setprofile(lambda *x: None)
    t0 = time.perf_counter()
    if prof:
        prof.enable()
    res = ETDolphinLauncher.launch(APP, argv)
    if prof:
        prof.disable()
    dt = time.perf_counter() - t0
    assert res.ok, res
    n = sum(r["num_batches"] for r in res.results)
    print(json.dumps({"metric": "control_plane_mini_batches_per_sec", "value": round(n / dt, 1), "workers": a.workers,
                      "mini_batches": n, "wall_s": round(dt, 3), "note": "includes executor / table setup"}))
    if prof:
        import pstats
        pstats.Stats(prof).sort_stats("cumulative").print_stats(25)


def run_jobserver(a, argv):
    import socket
    from harmony_b200.dolphin.launcher import parse_job_conf
    from harmony_b200.jobserver.client import CommandSender, JobServer
    from harmony_b200.utils.flags import Configuration
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    sys.modules["bench_control_plane"].APP = APP
    os.environ["PYTHONPATH"] = os.pathsep.join(
        [os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
         os.environ.get("PYTHONPATH", "")])          # executor processes import the no-op trainer from this file
    js = JobServer(Configuration({"num_executors": a.workers, "port": port, "executor_mode": a.executor_mode,
                                  "backend": "host"})).start()
    try:
        conf = parse_job_conf(APP, argv + ["-port", str(port)]).with_values(
            job_entity_builder="harmony_b200.dolphin.jobserver:DolphinJobEntityBuilder", app="bench_control_plane:APP")
        t0 = time.perf_counter()
        for _ in range(a.jobserver):
            assert CommandSender(port).send_job_submit_command(conf.serialize()).startswith("true")
        n = 0
        for j in range(a.jobserver):
            ok, result = js.driver.wait_job("NoOp-%d" % j, timeout=600)
            assert ok, result
            n += sum(r["num_batches"] for r in result)
        dt = time.perf_counter() - t0
    finally:
        js.stop(kill=True)
    print(json.dumps({"metric": "control_plane_mini_batches_per_sec", "value": round(n / dt, 1), "workers": a.workers,
                      "jobs": a.jobserver, "executor_mode": a.executor_mode, "mini_batches": n, "wall_s": round(dt, 3), "path": "job server"}))


if __name__ == "__main__":
    main()
