"""Multi-GPU worker (launched by torchrun from test_multigpu.py): validates the peer-memory data
plane -- IPC mapping, pull / push across GPUs, device clocks, block migration, fused MLR step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.models.mlr import MLRDeviceModel, mlr_reference_step, synthetic_batch
    from harmony_b200.parallel.device_clock import DeviceClock
    from harmony_b200.parallel.peer import PeerArena, PeerGroup

    group = PeerGroup.from_env()
    arena = PeerArena(lr, group, segment_bytes=256 << 20)
    cap, dim = 10000, 100
    conf = TableConfiguration(table_id="t", dense=True, capacity=cap, row_dim=dim, num_total_blocks=32,
                              is_ordered=False, backend="device",
                              update_function={"name": "add", "params": {"init": "gaussian", "std": 0.01, "seed": 9}})
    table = DeviceTable(conf, arena, group)
    clock = DeviceClock(arena, group)
    W0 = table.to_dense().clone()
    # every rank must see the same (key-deterministic) initial model
    ref0 = W0.clone(); dist.broadcast(ref0, 0)
    assert torch.equal(W0, ref0), "ranks disagree on the initial table"

    # 1. cross-GPU push: every rank adds (rank+1) to every element of every row
    keys = torch.arange(cap, device=dev)
    table.push(keys, torch.full((cap, dim), float(rank + 1), device=dev))
    clock.barrier(); torch.cuda.synchronize(); clock.check()
    expect = W0 + sum(r + 1 for r in range(world))
    got = table.to_dense()
    assert torch.allclose(got, expect, atol=1e-5), f"push mismatch {(got - expect).abs().max()}"
    clock.barrier()

    # 2. fused MLR step from every rank at once == sum of the per-rank reference deltas
    classes = dim
    model = MLRDeviceModel(table, classes, step_size=0.01)
    batch = synthetic_batch(200, cap, classes, 11, seed=100 + rank).to(dev)
    W1 = table.to_dense().clone()
    clock.barrier(); torch.cuda.synchronize()
    model.step(batch)
    clock.barrier(); torch.cuda.synchronize(); clock.check()
    Wn, _, _ = mlr_reference_step(W1, batch, classes, 0.01)
    delta = Wn - W1
    dist.all_reduce(delta)
    got = table.to_dense()
    err = float((got - (W1 + delta)).abs().max())
    assert err < 1e-4, f"mlr multi-gpu mismatch {err}"
    clock.barrier()

    # 2b. "ship the activations" mode computes the same update (dot at the owners, only [B, C] moves)
    from harmony_b200.models.mlr_ship import MLRShipModel, route_batches
    all_batches = [synthetic_batch(200, cap, classes, 11, seed=300 + r) for r in range(world)]
    rb = route_batches(table, all_batches, rank).to(dev)
    ship = MLRShipModel(table, arena, group, clock, classes, 0.01, 200)
    W2 = table.to_dense().clone()
    clock.barrier(); torch.cuda.synchronize()
    ship.step(rb)
    clock.barrier(); torch.cuda.synchronize(); clock.check()
    delta = torch.zeros_like(W2)
    loss_ref = 0.0
    for r in range(world):
        Wn, l, _ = mlr_reference_step(W2, all_batches[r].to(dev), classes, 0.01)
        delta += Wn - W2
        if r == rank:
            loss_ref = l
    err = float((table.to_dense() - (W2 + delta)).abs().max())
    assert err < 1e-4, f"ship-mode mismatch {err}"
    assert abs(float(ship.stats[0]) - loss_ref) / max(1.0, loss_ref) < 2e-3, (float(ship.stats[0]), loss_ref)
    ship.step(rb); clock.barrier(); torch.cuda.synchronize(); clock.check()      # buffers are reusable
    clock.barrier()

    # 3. migration: move block 3 to the next rank; content and subsequent pushes stay correct
    before = table.to_dense().clone()
    src = table.owners[3]; dst = (src + 1) % world
    table.migrate_block(3, dst)
    assert table.owners[3] == dst
    after = table.to_dense()
    assert torch.equal(before, after), "migration changed the table content"
    table.push(keys, torch.ones((cap, dim), device=dev))
    clock.barrier(); torch.cuda.synchronize(); clock.check()
    assert torch.allclose(table.to_dense(), before + world, atol=1e-5)
    if rank == dst:
        assert 3 in table.local_block_ids()
    if rank == src:
        assert 3 not in table.local_block_ids()
    clock.barrier(); torch.cuda.synchronize()
    if rank == 0:
        print("MULTIGPU_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
