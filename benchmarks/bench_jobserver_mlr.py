#!/usr/bin/env python
"""MLR 1M-feature x 1000-class THROUGH THE JOB SERVER (submit over the TCP command port, global scheduler, Dolphin
master / workers, device executors) next to the SPMD number of bench.py: what the per-mini-batch host path costs.

The job runs `-mlr_mode ship` (the workers form a rendezvous group through their master and step in lock-step on
device flags), mini-batch leases on (`-mini_batch_grant_size`), no per-batch stream sync (the worker loop calls
`trainer.flush()` once per lease).  The input file is synthetic text of the bench shape, written once.
Throughput is taken from the workers' epoch metrics (epochs 2.. : steady state, the first epoch pays parsing-cache
and kernel-load costs), whole box.
  python benchmarks/bench_jobserver_mlr.py --executors 8 --epochs 6 --batches_per_worker 8 [--executor_mode process]
"""
import argparse
import json
import os
import socket
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_input(path, n, features, nnz, classes, seed=0):
    import numpy as np
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        for lo in range(0, n, 4096):
            m = min(4096, n - lo)
            cols = np.sort(rng.integers(0, features, size=(m, nnz)), axis=1)
            vals = rng.random((m, nnz), dtype=np.float32)
            labels = (cols[:, 0].astype(np.int64) * 2654435761 % classes)
            lines = []
            for i in range(m):
                lines.append(str(int(labels[i])) + " " + " ".join(f"{c}:{v:.4f}" for c, v in zip(cols[i], vals[i])))
            f.write("\n".join(lines) + "\n")
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--executors", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--batches_per_worker", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--features", type=int, default=1_000_000)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--nnz", type=int, default=100)
    ap.add_argument("--grant", type=int, default=8)
    ap.add_argument("--mode", default="ship", choices=["ship", "pull"])
    ap.add_argument("--executor_mode", default="inproc", choices=["inproc", "process"])
    ap.add_argument("--spmd_value", type=float, default=0.0, help="bench.py value at the same N, for the ratio")
    ap.add_argument("--backend", default="device")
    a = ap.parse_args()
    from harmony_b200.dolphin.launcher import DolphinJobLauncher
    from harmony_b200.jobserver.client import JobServer
    from harmony_b200.mlapps import mlr
    from harmony_b200.utils.flags import Configuration
    tmp = tempfile.mkdtemp(prefix="hbjsmlr")
    W = a.executors
    n = W * a.batches_per_worker * a.batch
    t0 = time.perf_counter()
    data = write_input(os.path.join(tmp, "mlr_big"), n, a.features, a.nnz, a.classes)
    t_write = time.perf_counter() - t0
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    js = JobServer(Configuration({"num_executors": W, "port": port, "backend": a.backend,
                                  "executor_mode": a.executor_mode, "scheduler": "FifoExclusiveScheduler",
                                  "chkp_temp_path": tmp + "/t", "chkp_commit_path": tmp + "/c"})).start()
    try:
        nb = W * a.batches_per_worker
        args = ["-input", data, "-max_num_epochs", str(a.epochs), "-num_mini_batches", str(nb),
                "-num_worker_blocks", str(nb), "-num_server_blocks", "1024", "-step_size", "1e-4",
                "-classes", str(a.classes), "-features", str(a.features), "-port", str(port),
                "-mini_batch_grant_size", str(a.grant), "-task_unit_scheduling", "off"]
        if a.mode == "ship" and a.backend == "device":
            args += ["-mlr_mode", "ship"]
        captured = {}
        from harmony_b200.dolphin import master as djs
        orig_register = djs.ETDolphinMetricReceiver.register

        def register(self, job_id, mm, *aa, **kw):      # keep a handle on the job's metric manager
            captured[job_id] = mm
            return orig_register(self, job_id, mm, *aa, **kw)
        djs.ETDolphinMetricReceiver.register = register
        t0 = time.perf_counter()
        assert DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, args, port=port).startswith("true")
        ok, result = js.driver.wait_job("MLR-0", timeout=1500)
        wall = time.perf_counter() - t0
        assert ok, result
        # steady state from the master's epoch metrics: per worker, epochs >= 2
        mm = captured.get("MLR-0")
        per_epoch = []
        if mm is not None:
            for wid, ms in mm.get_worker_epoch_metrics().items():
                for m in ms:
                    d = m if isinstance(m, dict) else m.to_dict()
                    d = d.get("metrics", d)
                    if d.get("epoch_idx", 0) >= 2 and d.get("epoch_time_sec"):
                        per_epoch.append(d["num_batches_for_epoch"] / max(1e-9, d["epoch_time_sec"]))
        steady = sum(per_epoch) / max(1, len(per_epoch)) * W if per_epoch else None
        total_batches = sum(r["num_batches"] for r in result)
        out = {"metric": "mini-batches/sec through the job server (whole box), MLR 1M-feature x 1000-class",
               "value": round(steady, 1) if steady else round(total_batches / wall, 1), "unit": "mini-batches/s",
               "n_gpus": W, "executor_mode": a.executor_mode, "mode": a.mode, "grant": a.grant,
               "steady_state_from_epoch_metrics": steady is not None, "mini_batches": total_batches,
               "job_wall_s_incl_load_and_init": round(wall, 2), "input_write_s": round(t_write, 1),
               "trainer": result[0].get("trainer"), "last_batch_loss": result[0].get("last_batch_loss")}
        if a.spmd_value:
            out["fraction_of_spmd_bench"] = round(out["value"] / a.spmd_value, 3)
        print(json.dumps(out), flush=True)
    finally:
        js.stop(kill=True)


if __name__ == "__main__":
    main()
