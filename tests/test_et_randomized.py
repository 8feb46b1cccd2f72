"""Randomized model-based test of the Elastic Table: a few hundred random single / multi-key ops issued from random
executors, interleaved with block migrations, a subscriber joining late and a checkpoint / restore with a different
blocking, checked against a plain dict after every read and in a final full scan."""
import random

import numpy as np
import pytest

from harmony_b200.et import TableConfiguration, local_cluster


@pytest.mark.parametrize("seed,ordered,mutable", [(0, False, True), (1, True, True), (2, False, True), (3, True, False)])
def test_random_ops_with_migrations_match_a_dict_model(seed, ordered, mutable, tmp_path):
    rng = random.Random(seed)
    master, execs, launcher = local_cluster(4, chkp_temp_path=str(tmp_path / "t"), chkp_commit_path=str(tmp_path / "c"))
    try:
        conf = TableConfiguration(table_id="r", num_total_blocks=12, is_ordered=ordered, is_mutable=mutable,
                                  update_function="add")
        master.create_table(conf, execs[:2]).get(30)
        table = master.get_table("r")
        table.subscribe([execs[2]]).get(30)
        users = [e.id for e in execs[:3]]
        h = lambda eid: launcher.executors[eid].tables.get_table("r")  # noqa: E731
        model = {}
        keyspace = list(range(0, 4000, 7)) if ordered else list(range(200))
        for step in range(400):
            t = h(rng.choice(users))
            k = rng.choice(keyspace)
            op = rng.random()
            if op < 0.18:
                v = rng.randint(-50, 50); t.put(k, v).get(10); model[k] = v
            elif op < 0.30:
                v = rng.randint(-50, 50)
                prev = t.put_if_absent(k, v).get(10)
                assert prev == model.get(k)
                model.setdefault(k, v)
            elif op < 0.45:
                assert t.get(k).get(10) == model.get(k), (step, k)
            elif op < 0.60 and mutable:
                d = rng.randint(1, 9)
                new = t.update(k, d).get(10)
                model[k] = model.get(k, 0) + d
                assert new == model[k], (step, k)
            elif op < 0.68:
                assert t.remove(k).get(10) == model.pop(k, None)
            elif op < 0.80:
                ks = rng.sample(keyspace, 6)
                if mutable:
                    t.multi_update(ks, [1] * 6).get(10)
                    for kk in ks:
                        model[kk] = model.get(kk, 0) + 1
                else:
                    t.multi_put(ks, [step] * 6).get(10)
                    for kk in ks:
                        model[kk] = step
            elif op < 0.90:
                ks = rng.sample(keyspace, 6)
                assert list(t.multi_get(ks).get(10)) == [model.get(kk) for kk in ks], step
            elif op < 0.96:
                info = table.get_partition_info()
                src = rng.choice([e for e in info if info[e]])
                dst = rng.choice([e for e in info if e != src])
                assert table.move_blocks(src, dst, rng.randint(1, 2)).get(60).is_completed
            elif op < 0.98 and execs[3].id not in users:
                table.subscribe([execs[3]]).get(30)            # a late subscriber learns the current ownership
                users.append(execs[3].id)
        # full scan over the associators == model
        scan = {}
        for e in table.get_partition_info():
            scan.update(launcher.executors[e].tables.get_table("r").get_local_tablet().get_data_map())
        assert scan == model
        # checkpoint, restore with a different blocking on other executors, compare again
        chkp = table.checkpoint().get(60)
        t2 = master.create_table_from_chkp(chkp, execs[2:], table_id="r2").get(60)
        scan2 = {}
        for e in t2.get_partition_info():
            scan2.update(launcher.executors[e].tables.get_table("r2").get_local_tablet().get_data_map())
        assert scan2 == model
    finally:
        master.close()


@pytest.mark.parametrize("dense", [False, True])
def test_concurrent_updates_survive_live_migrations(dense):
    """Three updater threads (random multi-key `+1`s from different executors) race with a migrator that keeps moving
    blocks between the associators (ownership-first protocol: latch at the receiver, redirect at the old owner,
    driver fallback): every update must land exactly once."""
    import threading
    import time
    master, execs, launcher = local_cluster(3)
    try:
        conf = TableConfiguration(table_id="c", num_total_blocks=10, update_function="add", dense=dense,
                                  **({"capacity": 120, "row_dim": 2} if dense else {}))
        master.create_table(conf, execs).get(30)
        table = master.get_table("c")
        keys = list(range(120))
        counts = [np.zeros(120, dtype=np.int64) for _ in range(3)]
        stop = threading.Event()
        errors = []

        def updater(i):
            rng = random.Random(100 + i)
            t = launcher.executors[execs[i].id].tables.get_table("c")
            try:
                while not stop.is_set():
                    ks = rng.sample(keys, 8)
                    vals = np.ones((8, 2), dtype=np.float32) if dense else [1] * 8
                    t.multi_update(np.asarray(ks, dtype=np.int64) if dense else ks, vals).get(30)
                    counts[i][ks] += 1
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        def migrator():
            rng = random.Random(7)
            try:
                while not stop.is_set():
                    info = table.get_partition_info()
                    src = rng.choice([e for e in info if info[e]])
                    dst = rng.choice([e for e in info if e != src])
                    table.move_blocks(src, dst, 1).get(60)
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))
        ths = [threading.Thread(target=updater, args=(i,)) for i in range(3)] + [threading.Thread(target=migrator)]
        [t.start() for t in ths]
        time.sleep(1.5)
        stop.set()
        [t.join(60) for t in ths]
        assert not errors, errors
        want = sum(counts)
        t0 = launcher.executors[execs[0].id].tables.get_table("c")
        got = t0.multi_get_or_init(np.asarray(keys, dtype=np.int64) if dense else keys).get(30)
        got = np.asarray([g[0] if dense else (g or 0) for g in got], dtype=np.int64)
        np.testing.assert_array_equal(got, want)
        assert want.sum() > 0
    finally:
        master.close()
