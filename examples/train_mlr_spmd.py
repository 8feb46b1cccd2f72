#!/usr/bin/env python
"""SPMD library example: sparse MLR on a `label idx:val ...` file, one process per GPU, no job server.

    python examples/train_mlr_spmd.py --input bin/sample_mlr --features 784 --classes 10 --epochs 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/train_mlr_spmd.py --input data.libsvm --features 1000000 --classes 1000

Every rank loads its byte-range split of the file, the model table is sharded block-wise over the GPUs' HBM, each
step is two kernels (pull + dot + softmax, gradient + push) and the ranks stay within `--clock_slack` mini-batches of
each other through device clocks in peer memory (no host barrier, no NCCL call in the loop).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True)
    ap.add_argument("--features", type=int, required=True)
    ap.add_argument("--classes", type=int, required=True)
    ap.add_argument("--batch_size", type=int, default=256)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--step_size", type=float, default=0.05)
    ap.add_argument("--clock_slack", type=int, default=0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.models.mlr import MLRDeviceModel, load_batches
    from harmony_b200.parallel.device_clock import DeviceClock
    from harmony_b200.parallel.peer import PeerArena, PeerGroup
    world, rank, lr = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))      # bootstrap only
    group = PeerGroup.from_env()
    arena = PeerArena(lr, group)
    table = DeviceTable(TableConfiguration(table_id="model", dense=True, capacity=a.features, row_dim=a.classes,
                                           num_total_blocks=max(world, min(1024, a.features)), backend="device",
                                           update_function={"name": "add", "params": {"init": "gaussian",
                                                                                      "std": 0.001, "seed": 0}}),
                        arena, group)
    model = MLRDeviceModel(table, a.classes, step_size=a.step_size)
    clock = DeviceClock(arena, group, slack=a.clock_slack)
    batches = [b.to("cuda") for b in load_batches(a.input, a.batch_size, rank, world, a.classes, a.features)]
    steps = max(len(batches), 1)
    if world > 1:                                    # every rank runs the same number of steps per epoch
        t = torch.tensor([steps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN); steps = int(t.item())
    for epoch in range(a.epochs):
        model.stats.zero_()
        n = 0
        for b in batches[:steps]:
            clock.wait()
            model.step(b)
            clock.tick()
            n += b.batch_size
        loss = float(model.stats[0]) / max(n, 1)
        print(f"[rank {rank}] epoch {epoch}: loss/instance {loss:.4f}", flush=True)
    clock.barrier(); clock.check()
    torch.cuda.synchronize()
    table.close(); arena.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
