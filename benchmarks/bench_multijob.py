#!/usr/bin/env python
"""BASELINE config 5: 4 concurrent jobs (MLR + NMF + LDA + MLR) under the global scheduler on one
box, with ONE elastic re-partition of a live model table mid-run.

Reports the makespan, per-job wall time / mini-batches per second and the migration time + GB/s.
Jobs are submitted over the TCP command port exactly like `submit_<app>.sh` does.  On a GPU box the
executors are device executors (one per GPU, tables in HBM, fused kernels); `--backend host` runs
the same scenario on the CPU backend.
  python benchmarks/bench_multijob.py --executors 2 --backend auto
"""
import argparse
import json
import os
import socket
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--executors", type=int, default=2)
    ap.add_argument("--backend", default="auto")
    ap.add_argument("--scheduler", default="SchedulerImpl")
    ap.add_argument("--scale", type=float, default=1.0, help="dataset size multiplier")
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--repartition_after_s", type=float, default=0.2)
    ap.add_argument("--solo", action="store_true",
                    help="first run every job ALONE on the same server (sequentially) and report the concurrent makespan "
                         "over the slowest solo job")
    ap.add_argument("--extra", default="", help="comma-separated flags added to every job, e.g. "
                                                "--extra=-mini_batch_grant_size,4,-task_unit_scheduling,off")
    a = ap.parse_args()
    import numpy as np
    from helpers import write_lda_sample, write_mlr_sample, write_nmf_sample
    from harmony_b200.dolphin.launcher import DolphinJobLauncher
    from harmony_b200.jobserver.client import JobServer
    from harmony_b200.mlapps import lda, mlr, nmf
    from harmony_b200.utils.flags import Configuration

    tmp = tempfile.mkdtemp(prefix="hbmulti")
    s = int(4000 * a.scale)
    F, C = 8192, 256
    d_mlr = write_mlr_sample(os.path.join(tmp, "mlr_a"), n=s, features=F, classes=C, nnz=32, seed=1)
    d_mlr2 = write_mlr_sample(os.path.join(tmp, "mlr_b"), n=s, features=F, classes=C, nnz=32, seed=2)
    d_nmf = write_nmf_sample(os.path.join(tmp, "nmf"), rows=s, cols=2000, rank=8, density=0.01, seed=3)
    d_lda = write_lda_sample(os.path.join(tmp, "lda"), docs=max(200, s // 8), vocab=2000, topics=8, doc_len=60, seed=4)
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    js = JobServer(Configuration({"num_executors": a.executors, "port": port, "backend": a.backend,
                                  "scheduler": a.scheduler, "chkp_temp_path": tmp + "/t",
                                  "chkp_commit_path": tmp + "/c"})).start()
    nb = 8
    common = ["-max_num_epochs", str(a.epochs), "-num_mini_batches", str(nb * a.executors),
              "-num_worker_blocks", str(nb * a.executors), "-num_server_blocks", "64", "-port", str(port)] + [x for x in a.extra.split(",") if x]
    jobs = [
        ("MLR-0", mlr, ["-input", d_mlr, "-classes", str(C), "-features", str(F), "-step_size", "0.01"] + common),
        ("NMF-1", nmf, ["-input", d_nmf, "-rank", "64", "-features", "2000", "-step_size", "0.001", "-decay_rate", "1.0"] + common),
        ("LDA-2", lda, ["-input", d_lda, "-num_topics", "64", "-num_vocabs", "2000"] + common),
        ("MLR-3", mlr, ["-input", d_mlr2, "-classes", str(C), "-features", str(F), "-step_size", "0.01"] + common),
    ]
    solo = {}
    n_prev = 0
    if a.solo:                       # job ids are <App>-<n> with one counter per server: solo runs take the first ids
        for name, app, args in jobs:
            ts = time.perf_counter()
            r = DolphinJobLauncher.submit_job(app.APP_PATH, app.APP, args, port=port)
            assert r.startswith("true"), r
            jid = f"{name.split('-')[0]}-{n_prev}"
            ok, _res = js.driver.wait_job(jid, timeout=1800)
            solo[name] = {"ok": ok, "seconds": round(time.perf_counter() - ts, 3)}
            n_prev += 1
        jobs = [(f"{name.split('-')[0]}-{n_prev + i}", app, args) for i, (name, app, args) in enumerate(jobs)]
    t0 = time.perf_counter()
    wall0 = time.time()
    for name, app, args in jobs:
        r = DolphinJobLauncher.submit_job(app.APP_PATH, app.APP, args, port=port)
        assert r.startswith("true"), r
    # one elastic re-partition of MLR-0's live model table
    mig = {}

    def repartition():
        m = js.et_master
        deadline = time.time() + 120
        tid = "model_table%d" % n_prev                # the first concurrent job's model table
        while time.time() < deadline and not (m.has_table(tid) and m.get_table(tid).sm.current_state == "INITIALIZED"):
            time.sleep(0.02)
        time.sleep(a.repartition_after_s)
        if not m.has_table(tid) or a.executors < 2:
            return
        table = m.get_table(tid)
        info = table.get_partition_info()
        src = max(sorted(info), key=lambda e: len(info[e]))
        dst = min((e for e in sorted(info) if e != src), key=lambda e: len(info[e]))
        nblk = max(1, len(info[src]) // 2)
        t = time.perf_counter()
        try:
            res = table.move_blocks(src, dst, nblk).get(timeout=120)
            dt = time.perf_counter() - t
            rows_per_block = -(-F // 64)
            stride = (C + 31) // 32 * 32
            nbytes = len(res.migrated_blocks) * rows_per_block * stride * 4
            st = getattr(res, "stats", {}) or {}
            mig.update({"blocks": len(res.migrated_blocks), "wall_seconds_alloc_copy_flip_drain": round(dt, 4),
                        "bytes": nbytes, "src": src, "dst": dst, "accessors_stopped": False,
                        "device_copy_ms": st.get("copy_ms"), "late_updates_forwarded": st.get("forwarded"),
                        "device_copy_GBps": (round(nbytes / (st["copy_ms"] / 1e3) / 1e9, 2)
                                             if st.get("copy_ms") else None),
                        "note": "a 2 MB table: the move is 4 control-plane round trips; the 1 GiB move of "
                                "tests/mp_device_worker.py measures the copy engine (>= 600 GB/s)"})
        except Exception as e:  # noqa: BLE001
            mig.update({"error": str(e)})
    th = threading.Thread(target=repartition); th.start()
    per_job = {}
    for name, _, _ in jobs:
        ok, result = js.driver.wait_job(name, timeout=1800)
        per_job[name] = {"ok": ok, "finished_at_s": round(js.driver.job_finish_times[name] - wall0, 3),
                         "mini_batches": sum(r["num_batches"] for r in result) if ok else None}
    makespan = time.perf_counter() - t0
    th.join(10)
    for name in per_job:
        if per_job[name]["ok"]:
            per_job[name]["mini_batches_per_s"] = round(per_job[name]["mini_batches"] / per_job[name]["finished_at_s"], 2)
    out = {"metric": "makespan of 4 concurrent jobs (MLR+NMF+LDA+MLR) with one elastic re-partition",
           "value": round(makespan, 3), "unit": "s", "higher_is_better": False, "n_executors": a.executors,
           "backend": js.backend, "scheduler": a.scheduler, "extra_flags": a.extra, "jobs": per_job, "migration": mig,
           "total_mini_batches_per_s": round(sum(j["mini_batches"] or 0 for j in per_job.values()) / makespan, 2)}
    if solo:
        slowest = max(v["seconds"] for v in solo.values())
        out["solo"] = solo
        out["makespan_over_slowest_solo"] = round(makespan / slowest, 3)
        out["makespan_over_sum_of_solos"] = round(makespan / sum(v["seconds"] for v in solo.values()), 3)
    print(json.dumps(out), flush=True)
    js.stop(kill=True)


if __name__ == "__main__":
    main()
