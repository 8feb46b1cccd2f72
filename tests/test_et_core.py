"""Elastic Table on the host backend: op semantics, ownership, migration, checkpoint, fallback.

Mirrors the reference's RemoteAccessOpHandlerTest / OwnershipCacheTest / BlockManagerTest /
MigrationManagerTest / FallbackManagerTest and the ET example integration tests.
"""
import tempfile
import threading
import time

import numpy as np
import pytest

from harmony_b200.et import ExecutorConfiguration, TableConfiguration, local_cluster
from harmony_b200.et.driver import BlockManager
from harmony_b200.et.ownership import OwnershipCache


@pytest.fixture()
def cluster():
    tmp = tempfile.mkdtemp(prefix="hbchk")
    conf = ExecutorConfiguration(chkp_temp_path=tmp + "/temp", chkp_commit_path=tmp + "/commit")
    master, execs, launcher = local_cluster(3, conf, chkp_temp_path=tmp + "/temp", chkp_commit_path=tmp + "/commit")
    yield master, execs, launcher
    master.close()


def _tables(launcher, table_id):
    return [ex.tables.get_table(table_id) for _, ex in sorted(launcher.executors.items())]


@pytest.mark.parametrize("ordered", [False, True])
def test_kv_table_op_sequence(cluster, ordered):
    master, execs, launcher = cluster
    conf = TableConfiguration(table_id=f"kv{ordered}", num_total_blocks=16, is_ordered=ordered,
                              update_function={"name": "add", "params": {}})
    master.create_table(conf, execs).get(timeout=10)
    t0, t1, t2 = _tables(launcher, conf.table_id)
    keys = list(range(0, 400, 7)) if not ordered else [k * (2 ** 63 // 400) for k in range(0, 400, 7)]
    for k in keys:
        assert t0.put(k, k * 10).get(5) is None            # PUT returns the old value (None)
    for k in keys:
        assert t1.get(k).get(5) == k * 10                  # GET from another executor
    assert t2.put(keys[0], 1).get(5) == keys[0] * 10        # PUT returns the previous value
    assert t1.update(keys[0], 5).get(5) == 6               # UPDATE applies the update function at the owner
    assert t0.put_if_absent(keys[0], 99).get(5) == 6       # PUT_IF_ABSENT keeps the existing value
    assert t0.put_if_absent(-1 if not ordered else 3, 99).get(5) is None
    assert t2.remove(keys[0]).get(5) == 6                  # REMOVE returns the removed value
    assert t1.get(keys[0]).get(5) is None
    assert t1.get_or_init(keys[0]).get(5) == 0             # GET_OR_INIT materialises initValue
    got = t2.multi_get(keys[1:10]).get(5)
    assert got == [k * 10 for k in keys[1:10]]
    t0.multi_update(keys[1:10], [1] * 9).get(5)
    assert t1.multi_get(keys[1:10]).get(5) == [k * 10 + 1 for k in keys[1:10]]
    # local tablets partition the data: every key lives in exactly one executor
    total = sum(t.get_local_tablet().get_num_data_items() for t in (t0, t1, t2))
    assert total == len(keys) + 1
    master.get_table(conf.table_id).drop().get(10)
    assert not master.has_table(conf.table_id)


def test_dense_table_pull_push(cluster):
    master, execs, launcher = cluster
    conf = TableConfiguration(table_id="dense", dense=True, capacity=1000, row_dim=8, num_total_blocks=16,
                              update_function={"name": "add", "params": {"init": "gaussian", "std": 0.1, "seed": 1}})
    master.create_table(conf, execs).get(timeout=10)
    t0, t1, t2 = _tables(launcher, "dense")
    keys = np.arange(0, 1000, 3)
    W0 = t0.multi_get_or_init(keys).get(10)
    assert W0.shape == (len(keys), 8)
    W0b = t1.multi_get_or_init(keys).get(10)               # same values from any executor
    np.testing.assert_array_equal(W0, W0b)
    d = np.ones((len(keys), 8), dtype=np.float32)
    t1.multi_update(keys, d).get(10)
    t2.multi_update_no_reply(keys, d)
    time.sleep(0.3)
    np.testing.assert_allclose(t0.multi_get_or_init(keys).get(10), W0 + 2, rtol=1e-6)
    # duplicate keys in one push are all applied
    t0.multi_update(np.array([5, 5, 5]), np.ones((3, 8), dtype=np.float32)).get(10)
    base = t0.get_or_init(5).get(5)
    t0.multi_update(np.array([5, 5]), np.ones((2, 8), dtype=np.float32)).get(10)
    np.testing.assert_allclose(t0.get_or_init(5).get(5), base + 2, rtol=1e-6)


def test_sgd_clamp_and_lda_update_functions(cluster):
    master, execs, launcher = cluster
    c1 = TableConfiguration(table_id="nmf", dense=True, capacity=100, row_dim=4, num_total_blocks=8,
                            update_function={"name": "sgd_clamp", "params": {"step": 0.5, "lo": 0.0, "hi": 1.0,
                                                                             "init": "uniform"}})
    master.create_table(c1, execs).get(10)
    t = _tables(launcher, "nmf")[0]
    w = t.get_or_init(3).get(5)
    assert ((w >= 0) & (w <= 1)).all()
    t.update(3, np.full(4, 10.0, dtype=np.float32)).get(5)
    np.testing.assert_array_equal(t.get(3).get(5), np.zeros(4, dtype=np.float32))    # clamped at 0
    c2 = TableConfiguration(table_id="lda", num_total_blocks=8,
                            update_function={"name": "lda_sparse", "params": {"num_topics": 5}})
    master.create_table(c2, execs).get(10)
    t = _tables(launcher, "lda")[1]
    t.update(7, np.array([1, 2, 3, 1], dtype=np.int32)).get(5)
    row = t.get(7).get(5)
    assert list(row) == [0, 2, 0, 1, 0, 2]
    t.update(7, np.array([1, -5], dtype=np.int32)).get(5)                              # clamp + nnz
    assert list(t.get(7).get(5)) == [0, 0, 0, 1, 0, 1]


def test_ownership_cache_blocks_access_until_allowed():
    acks = []
    oc = OwnershipCache("t", 4, "me", "driver", sync_ack=lambda *a: acks.append(a))
    oc.init(["me", "other", "other", "me"])
    assert oc.resolve_executor(0) is None and oc.resolve_executor(1) == "other"
    oc.update(1, "other", "me")          # becomes incoming: access blocked until allow_access
    got = []

    def reader():
        owner, unlock = oc.resolve_executor_with_lock(1)
        got.append(owner)
        unlock()
    th = threading.Thread(target=reader); th.start()
    time.sleep(0.15)
    assert got == []
    oc.allow_access_to_block(1)
    th.join(2)
    assert got == [None]
    # update() waits for readers (write lock)
    owner, unlock = oc.resolve_executor_with_lock(0)
    done = []
    th = threading.Thread(target=lambda: (oc.update(0, "me", "other"), done.append(1))); th.start()
    time.sleep(0.15)
    assert done == []
    unlock(); th.join(2)
    assert done == [1] and oc.resolve_executor(0) == "other"
    # unassociation sync acks once the leaver owns nothing
    oc.sync_unassociation(7, "other")
    assert acks == []
    oc.update(0, "other", "me"); oc.update(2, "other", "me")
    assert acks == [(7, "t", "other")]


def test_block_manager_round_robin_and_moves():
    bm = BlockManager(12)
    bm.init(["a", "b", "c"])
    assert all(len(v) == 4 for v in bm.get_partition_info().values())
    assert bm.get_owner_id(4) == "b"
    blocks = bm.choose_blocks_to_move("a", 3)
    assert len(blocks) == 3 and bm.choose_blocks_to_move("a", 3) == [9]    # moving blocks are not re-chosen
    for b in blocks:
        bm.update_owner(b, "a", "c"); bm.release_block_from_move(b)
    assert bm.get_num_blocks("a") == 1 and bm.get_num_blocks("c") == 7
    with pytest.raises(RuntimeError):
        bm.update_owner(1, "b", "c")        # not chosen for migration


@pytest.mark.parametrize("mutable", [True, False])
def test_migration_conserves_data_and_updates_caches(cluster, mutable):
    master, execs, launcher = cluster
    conf = TableConfiguration(table_id=f"mig{mutable}", num_total_blocks=12, is_mutable=mutable,
                              update_function={"name": "add", "params": {}}, chunk_size=5)
    master.create_table(conf, execs[:2]).get(10)          # executor-2 is not yet part of the table
    table = master.get_table(conf.table_id)
    table.subscribe([execs[2]]).get(10)                    # late subscriber gets the ownership map
    t0 = launcher.executors[execs[0].id].tables.get_table(conf.table_id)
    for k in range(300):
        t0.put(k, k).get(5)
    stop = threading.Event()

    def writer():                                          # live traffic during the migration
        i = 0
        while not stop.is_set():
            t0.update(i % 300, 0).get(5)
            i += 1
    th = threading.Thread(target=writer); th.start()
    res = table.move_blocks(execs[0].id, execs[1].id, 4).get(20)
    stop.set(); th.join(5)
    assert res.is_completed and len(res.migrated_blocks) == 4
    info = table.get_partition_info()
    assert len(info[execs[0].id]) == 2 and len(info[execs[1].id]) == 10
    # every executor (incl. the pure subscriber) converged to the driver's ownership truth
    time.sleep(0.3)
    truth = table.get_ownership_status()
    for ex in launcher.executors.values():
        assert ex.tables.get_components(conf.table_id).ownership.owners() == truth
    t2 = launcher.executors[execs[2].id].tables.get_table(conf.table_id)
    assert t2.multi_get(list(range(300))).get(10) == list(range(300))
    # associate a new executor and move blocks onto it, then unassociate the first one
    table.associate([execs[2]]).get(10)
    assert table.move_blocks(execs[0].id, execs[2].id, 2).get(20).is_completed
    table.unassociate(execs[0].id).get(20)
    assert execs[0].id not in table.get_partition_info()
    assert t2.multi_get(list(range(300))).get(10) == list(range(300))


def test_checkpoint_restore_with_different_blocking(cluster):
    master, execs, launcher = cluster
    conf = TableConfiguration(table_id="ck", dense=True, capacity=500, row_dim=4, num_total_blocks=10,
                              update_function={"name": "add", "params": {"init": "gaussian", "std": 1.0, "seed": 2}})
    master.create_table(conf, execs).get(10)
    t = _tables(launcher, "ck")[0]
    keys = np.arange(500)
    W = t.multi_get_or_init(keys).get(10)
    chkp_id = master.get_table("ck").checkpoint().get(20)
    t.multi_update(keys, np.ones((500, 4), dtype=np.float32)).get(10)
    # restore into a NEW table on 2 executors (re-blocking through multi_put)
    master.create_table_from_chkp(chkp_id, execs[:2], table_id="ck_restored").get(20)
    tr = launcher.executors[execs[0].id].tables.get_table("ck_restored")
    np.testing.assert_array_equal(np.stack(tr.multi_get(keys).get(10)), W)
    # commit + restore by path with another block count
    master.chkp_master.commit(chkp_id)
    time.sleep(0.3)
    import os
    path = os.path.join(master.chkp_commit_path, chkp_id)
    assert os.path.isdir(path)
    conf2 = TableConfiguration(table_id="ck_path", dense=True, capacity=500, row_dim=4, num_total_blocks=7,
                               update_function="add", chkp_path=path)
    master.create_table(conf2, execs).get(20)
    tp = _tables(launcher, "ck_path")[2]
    np.testing.assert_array_equal(np.stack(tp.multi_get(keys).get(10)), W)


def test_fallback_reroutes_after_stale_ownership(cluster):
    master, execs, launcher = cluster
    conf = TableConfiguration(table_id="fb", num_total_blocks=6, update_function="add")
    master.create_table(conf, execs[:2]).get(10)
    # executor-2 knows nothing about the table: a request sent to it bounces to the driver,
    # which forwards it to the real owner (FallbackManager)
    ex0 = launcher.executors[execs[0].id]
    t0 = ex0.tables.get_table("fb")
    t0.put(1, 41).get(5)
    comps = ex0.tables.get_components("fb")
    block = comps.partitioner.get_block_id(1)
    owner = master.get_table("fb").get_owner_id(block)
    stale = execs[2].id
    comps.ownership._owners[block] = stale if owner != ex0.executor_id else comps.ownership._owners[block]
    if owner != ex0.executor_id:
        assert t0.get(1).get(5) == 41


def test_migration_manager_concurrent_bidirectional_moves_and_late_subscriber(cluster):
    """Reference MigrationManagerTest: concurrent moves in both directions conserve the blocks and
    every subscriber (incl. one added while moves are in flight) converges to the driver's truth."""
    master, execs, launcher = cluster
    conf = TableConfiguration(table_id="mm", num_total_blocks=24, update_function="add", chunk_size=3)
    master.create_table(conf, execs[:2]).get(10)
    table = master.get_table("mm")
    t0 = launcher.executors[execs[0].id].tables.get_table("mm")
    for k in range(240):
        t0.put(k, k).get(5)
    f1 = table.move_blocks(execs[0].id, execs[1].id, 5)
    f2 = table.move_blocks(execs[1].id, execs[0].id, 3)
    sub = table.subscribe([execs[2]])
    r1, r2 = f1.get(30), f2.get(30)
    sub.get(30)
    assert r1.is_completed and r2.is_completed
    info = table.get_partition_info()
    assert len(info[execs[0].id]) + len(info[execs[1].id]) == 24
    assert len(info[execs[0].id]) == 12 - 5 + 3 and table.block_manager.num_moving() == 0
    time.sleep(0.3)
    truth = table.get_ownership_status()
    for ex in launcher.executors.values():
        assert ex.tables.get_components("mm").ownership.owners() == truth
    t2 = launcher.executors[execs[2].id].tables.get_table("mm")
    assert t2.multi_get(list(range(240))).get(10) == list(range(240))


def test_executor_failure_fails_running_tasklets(cluster):
    from harmony_b200.et.config import TaskletConfiguration
    master, execs, launcher = cluster
    seen = []
    master.failure_listeners.append(lambda e, r: seen.append(e))
    rt = execs[0].submit_tasklet(TaskletConfiguration("idle", "harmony_b200.dolphin.worker:ServerTasklet", None, {}))
    time.sleep(0.2)
    master.on_executor_failed(execs[0].id, "injected fault")
    res = rt.get_task_result(timeout=5)
    assert not res.is_success and "injected fault" in res.error
    assert seen == [execs[0].id] and execs[0].id not in master.executors


def test_task_unit_barriers_with_uneven_workers_never_deadlock():
    """Three jobs share four executors; the workers of a job run DIFFERENT numbers of mini-batches (global budget) and
    leave the job's unit barriers when they finish.  Units: PULL(NET) -> COMP(CPU) -> PUSH(NET) per batch with the
    per-executor slots CPU x1 / NET x2.  Every worker must finish (no barrier waits for a departed peer, no slot is
    held across a barrier)."""
    import random
    import threading
    from harmony_b200.et import local_cluster
    from harmony_b200.et.tasklet import CPU, NET, TaskUnitInfo
    master, execs, launcher = local_cluster(4)
    try:
        jobs = ["J0", "J1", "J2"]
        for j in jobs:
            master.task_units.on_job_start(j, 4)
        done, errors = [], []

        def worker(job, eid, n_batches, seed):
            rng = random.Random(seed)
            sched = launcher.executors[eid].task_unit_scheduler
            tid = f"{job}-worker-{eid}"
            try:
                for seq in range(n_batches):
                    for name, res in (("PULL", NET), ("COMP", CPU), ("PUSH", NET)):
                        info = TaskUnitInfo(job, name, res, seq)
                        sched.wait_schedule(tid, info)
                        time.sleep(rng.random() * 0.002)
                        sched.on_task_unit_finished(tid, info)
                sched.leave(job)
                done.append((job, eid))
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))
        ths = []
        for ji, job in enumerate(jobs):
            for ei, e in enumerate(execs):
                ths.append(threading.Thread(target=worker, args=(job, e.id, 3 + (ei * 2 + ji) % 6, 10 * ji + ei),
                                            daemon=True))
        [t.start() for t in ths]
        [t.join(60) for t in ths]
        assert not errors, errors
        assert len(done) == 12, f"{12 - len(done)} workers are stuck at a unit barrier"
    finally:
        master.close()


def test_block_snapshot_is_shared_until_the_block_changes():
    """The training loop reads every input block once per epoch: the item list is rebuilt only after a mutation."""
    from harmony_b200.et.block_store import BlockSnapshot, DictBlock
    from harmony_b200.et.update_function import make_update_function
    b = DictBlock(3, make_update_function("add"))
    for k in (7, 5, 9):
        b.put(k, k * 10)
    s1 = b.items_snapshot()
    assert isinstance(s1, BlockSnapshot) and sorted(s1) == [(5, 50), (7, 70), (9, 90)]
    assert s1.block_id == 3 and s1.min_key() == 5
    assert b.items_snapshot() is s1                           # no copy while nothing changes
    b.get(7); b.get_or_init(7); b.put_if_absent(7, 1)         # reads and no-op writes keep it valid
    assert b.items_snapshot() is s1
    for mutate in (lambda: b.put(11, 1), lambda: b.update(5, 1), lambda: b.remove(9), lambda: b.get_or_init(42),
                   lambda: b.put_if_absent(43, 1), lambda: b.import_items([1], [2]), lambda: b.clear()):
        before = b.items_snapshot()
        mutate()
        after = b.items_snapshot()
        assert after is not before and sorted(after) == sorted(b.items())
    assert b.items_snapshot().min_key() is None               # empty block


def test_failing_checkpoint_write_and_restore_fail_the_driver_future_fast(cluster, monkeypatch):
    """an exception in ChkpManagerSlave.checkpoint / load must fail checkpoint().get() /
    create_table_from_chkp().get() with its cause instead of letting them run into their 600 s timeouts."""
    master, execs, launcher = cluster
    conf = TableConfiguration(table_id="ckf", dense=True, capacity=100, row_dim=2, num_total_blocks=6,
                              update_function="add")
    master.create_table(conf, execs).get(10)
    good = master.get_table("ckf").checkpoint().get(20)
    victim = launcher.executors[execs[1].id]

    def boom(*a, **k):
        raise OSError("disk full (injected)")
    monkeypatch.setattr(victim.chkp, "checkpoint", boom)
    t0 = time.time()
    with pytest.raises(RuntimeError, match="disk full"):
        master.get_table("ckf").checkpoint().get(30)
    assert time.time() - t0 < 10
    monkeypatch.setattr(victim.chkp, "load", boom)
    with pytest.raises(RuntimeError, match="disk full"):
        master.create_table_from_chkp(good, execs, table_id="ckf_restored").get(30)


def test_model_chkp_manager_records_checkpoints_in_issue_order():
    """per-epoch model checkpoints are taken by ONE thread, so they complete -- and are replayed -- in
    epoch order even when a later one would finish faster."""
    from harmony_b200.dolphin.master import ModelChkpManager
    from harmony_b200.utils.concurrent import ResultFuture
    delays = iter([0.30, 0.01, 0.10])
    ids = iter(["e0", "e1", "e2"])

    class FakeTable:
        def checkpoint(self):
            f, d, i = ResultFuture(), next(delays), next(ids)
            threading.Timer(d, lambda: f.on_completed(i)).start()
            return f

    class FakeMaster:
        def get_table(self, _):
            return table
    table = FakeTable()
    mgr = ModelChkpManager(FakeMaster(), "model", None, "input")
    for _ in range(3):
        mgr.create_checkpoint_async()          # epochs end faster than a checkpoint takes
    mgr.wait_pending()
    assert mgr.model_chkps == ["e0", "e1", "e2"]
