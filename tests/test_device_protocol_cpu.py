"""The driver side of the device-table protocol, on CPU, against a host-memory stand-in of DeviceExecutor
(tests/emu_device.py).  The GPU twins are tests/test_gpu_jobserver.py::test_device_table_migration_* and
test_addvector_elastic_reconfiguration_on_device_backend."""
import threading
import time

import numpy as np
import pytest

from emu_device import EmuDeviceExecutor
from harmony_b200.et import LocalTransport, TableConfiguration
from harmony_b200.et.driver import ETMaster, InProcessLauncher


@pytest.fixture()
def cluster():
    EmuDeviceExecutor.fail_recv_on = set()
    transport = LocalTransport()
    launcher = InProcessLauncher(EmuDeviceExecutor)
    master = ETMaster(transport, launcher)
    execs = master.add_executors(3).get(timeout=30)
    yield master, execs, launcher
    master.close()


def _conf(table_id="dev"):
    return TableConfiguration(table_id=table_id, dense=True, capacity=256, row_dim=4, num_total_blocks=16,
                              update_function="add", backend="device")


def test_device_protocol_init_move_and_late_subscriber(cluster):
    master, execs, launcher = cluster
    master.create_table(_conf(), execs[:2]).get(30)
    table = master.get_table("dev")
    assert table.is_device and set(table.device_infos) == {execs[0].id, execs[1].id}
    # every associator received the full slab map (16 blocks) in the peer-map phase
    for e in execs[:2]:
        infos = launcher.executors[e.id].peer_maps["dev"]
        assert sum(len(i["blocks"]) for i in infos) == 16
    t0 = launcher.executors[execs[0].id].tables.get_table("dev")
    keys = np.arange(256, dtype=np.int64)
    t0.multi_update(keys, np.ones((256, 4), dtype=np.float32)).get(30)
    before = table.get_partition_info()
    res = table.move_blocks(execs[0].id, execs[1].id, 4).get(60)
    assert res.is_completed and len(res.migrated_blocks) == 4
    after = table.get_partition_info()
    assert len(after[execs[1].id]) == len(before[execs[1].id]) + 4
    # ownership flipped on both executors, data intact and still updatable through either handle
    for e in execs[:2]:
        owners = launcher.executors[e.id].tables.get_components("dev").ownership.owners()
        assert [o for o in owners] == table.get_ownership_status()
    t1 = launcher.executors[execs[1].id].tables.get_table("dev")
    t1.multi_update(keys, np.full((256, 4), 2.0, dtype=np.float32)).get(30)
    np.testing.assert_array_equal(np.asarray(t0.multi_get_or_init(keys).get(30)), np.full((256, 4), 3.0))
    # the driver's slab map follows the move, so a late subscriber is told the CURRENT locations
    moved = set(res.migrated_blocks)
    assert moved <= set(table.device_infos[execs[1].id]["blocks"])
    assert not (moved & set(table.device_infos[execs[0].id]["blocks"]))
    table.subscribe([execs[2]]).get(30)
    infos = launcher.executors[execs[2].id].peer_maps["dev"]
    loc_owner = {int(b): loc[0] for i in infos for b, loc in i["blocks"].items()}
    assert all(loc_owner[b] == execs[1].id for b in moved) and len(loc_owner) == 16
    t2 = launcher.executors[execs[2].id].tables.get_table("dev")
    np.testing.assert_array_equal(np.asarray(t2.multi_get_or_init(keys[:32]).get(30)), np.full((32, 4), 3.0))


def test_device_move_excludes_accessors_and_keeps_sums_exact(cluster):
    """Updates racing with a LIVE device migration (nobody is stopped): late updates are forwarded, nothing is lost."""
    master, execs, launcher = cluster
    master.create_table(_conf("race"), execs[:2]).get(30)
    table = master.get_table("race")
    handles = [launcher.executors[e.id].tables.get_table("race") for e in execs[:2]]
    keys = np.arange(256, dtype=np.int64)
    stop, counts = threading.Event(), [0, 0]

    def pound(i):
        while not stop.is_set():
            handles[i].multi_update(keys, np.ones((256, 4), dtype=np.float32)).get(30)
            counts[i] += 1
    ths = [threading.Thread(target=pound, args=(i,)) for i in range(2)]
    [t.start() for t in ths]
    time.sleep(0.1)
    for src, dst in ((0, 1), (1, 0), (0, 1)):
        assert table.move_blocks(execs[src].id, execs[dst].id, 3).get(60).is_completed
    stop.set()
    [t.join(30) for t in ths]
    got = np.asarray(handles[0].multi_get_or_init(keys).get(30))
    np.testing.assert_array_equal(got, np.full((256, 4), float(sum(counts))))


def test_device_move_aborts_cleanly_when_the_receiver_fails(cluster):
    master, execs, launcher = cluster
    master.create_table(_conf("abort"), execs[:2]).get(30)
    table = master.get_table("abort")
    t0 = launcher.executors[execs[0].id].tables.get_table("abort")
    keys = np.arange(256, dtype=np.int64)
    t0.multi_update(keys, np.ones((256, 4), dtype=np.float32)).get(30)
    before = table.get_partition_info()
    EmuDeviceExecutor.fail_recv_on = {execs[1].id}
    fut = table.move_blocks(execs[0].id, execs[1].id, 4)
    with pytest.raises(RuntimeError, match="DeviceMoveAlloc failed"):
        fut.get(30)
    # nothing moved, the gate is open again (accessors work), and a later move succeeds
    assert table.get_partition_info() == before
    t0.multi_update(keys, np.ones((256, 4), dtype=np.float32)).get(10)
    EmuDeviceExecutor.fail_recv_on = set()
    assert table.move_blocks(execs[0].id, execs[1].id, 4).get(60).is_completed
    np.testing.assert_array_equal(np.asarray(t0.multi_get_or_init(keys).get(30)), np.full((256, 4), 2.0))


def test_device_move_aborts_cleanly_when_the_sender_fails(cluster):
    """The copy (DeviceMoveSend) fails after the receiver allocated its slabs: the receiver is told to free them
    (DeviceMoveAbort), nothing was flipped, the blocks are released and a later move succeeds."""
    master, execs, launcher = cluster
    master.create_table(_conf("abort2"), execs[:2]).get(30)
    table = master.get_table("abort2")
    t0 = launcher.executors[execs[0].id].tables.get_table("abort2")
    keys = np.arange(256, dtype=np.int64)
    t0.multi_update(keys, np.ones((256, 4), dtype=np.float32)).get(30)
    before = table.get_partition_info()
    EmuDeviceExecutor.aborted = []
    EmuDeviceExecutor.fail_send_on = {execs[0].id}
    try:
        with pytest.raises(RuntimeError, match="DeviceMoveSend failed"):
            table.move_blocks(execs[0].id, execs[1].id, 4).get(30)
    finally:
        EmuDeviceExecutor.fail_send_on = set()
    assert table.get_partition_info() == before
    assert len(EmuDeviceExecutor.aborted) == 1 and EmuDeviceExecutor.aborted[0][0] == execs[1].id
    assert len(EmuDeviceExecutor.aborted[0][1]) == 4
    t0.multi_update(keys, np.ones((256, 4), dtype=np.float32)).get(10)          # accessors were never stopped
    assert table.move_blocks(execs[0].id, execs[1].id, 4).get(60).is_completed
    np.testing.assert_array_equal(np.asarray(t0.multi_get_or_init(keys).get(30)), np.full((256, 4), 2.0))


@pytest.mark.parametrize("optimizer", ["DeleteOneServerOptimizer", "AddOneServerOptimizer", "MoveModelBlocksOptimizer",
                                       "ExchangeOneOptimizer", "DeleteOneWorkerOptimizer"])
def test_addvector_job_reconfigures_over_the_device_protocol(optimizer, monkeypatch):
    """The self-validating AddVector job with a live re-partition, `-backend device`, executors emulated on the host:
    plan compiler -> plan executor -> MOVE op -> quiesce / copy / flip / resume, while workers keep pushing."""
    from harmony_b200.dolphin.launcher import ETDolphinLauncher
    from harmony_b200.et import launcher as et_launcher
    from harmony_b200.mlapps import examples
    monkeypatch.setattr(et_launcher, "make_executor",
                        lambda eid, tr, conf, did, backend, ordinal: EmuDeviceExecutor(eid, tr, conf, did))
    path = examples.write_dummy_input(48)
    args = ["-input", path, "-num_dolphin_workers", "2", "-num_servers", "2", "-max_num_epochs", "6",
            "-num_mini_batches", "8", "-num_worker_blocks", "8", "-num_server_blocks", "8", "-num_keys", "20",
            "-vector_size", "4", "-delta", "2", "-compute_time_ms", "40", "-optimizer", optimizer,
            "-optimization_interval_ms", "300", "-delay_after_optimization_ms", "100",
            "-min_num_required_batch_metrics", "1", "-num_extra_resources", "1", "-backend", "device"]
    res = ETDolphinLauncher.launch(examples.ADD_VECTOR, args)
    assert res.ok, res.messages
    if optimizer == "DeleteOneServerOptimizer":
        assert any("(S: 2, W: 2) -> (S: 1, W: 2)" in m for m in res.messages), res.messages
