#!/usr/bin/env python
"""NMF rank-256 on a synthetic sparse matrix (BASELINE config 3: 100M x 10k, 8xB200).

Rows are data-parallel (each GPU owns rows_total / N rows + their L factors in HBM), the global
factor R (rank x cols) is a sharded device table.  One step of one worker = pull the batch's
columns (K1), fused SDDMM/SpMM + local-model update (K7+K4), push with SGD+clamp at the owners
(K2+K4).  Metric: mini-batches/sec over the whole box, device-timed, max over ranks.
  python benchmarks/bench_nmf.py --rows_per_gpu 2000000 --steps 20
"""
import argparse

from common import barrier_sync, emit, init_dist, max_over_ranks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--rank", type=int, default=256)
    ap.add_argument("--cols", type=int, default=10000)
    ap.add_argument("--rows_per_gpu", type=int, default=12_500_000)     # 100M rows / 8 GPUs
    ap.add_argument("--nnz_per_row", type=int, default=16)
    ap.add_argument("--batch_rows", type=int, default=131072)
    ap.add_argument("--distinct_batches", type=int, default=8)
    ap.add_argument("--impl", default="harmony", choices=["harmony", "nccl"],
                    help="nccl: replicated R, dense gradient all-reduce, local SGD+clamp (baseline)")
    a = ap.parse_args()
    import torch
    world, rank, lr, dev = init_dist()
    from harmony_b200 import ops
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.models.nmf import NMFBatch, NMFDeviceModel
    from harmony_b200.parallel.device_clock import DeviceClock
    from harmony_b200.parallel.peer import PeerArena, PeerGroup
    group = PeerGroup.from_env() if a.impl == "harmony" else PeerGroup(0, 1, None)
    arena = PeerArena(lr, group, segment_bytes=256 << 20)
    conf = TableConfiguration(table_id="model_table0", dense=True, capacity=a.cols, row_dim=a.rank, num_total_blocks=64,
                              backend="device", update_function={"name": "sgd_clamp", "params": {
                                  "step": 1e-4, "lo": 0.0, "hi": 1e6, "init": "uniform", "init_lo": 0.0, "init_hi": 1.0}})
    table = DeviceTable(conf, arena, group)
    model = NMFDeviceModel(table, a.rank, a.rows_per_gpu, step=1e-4, max_val=1e6, seed=rank)
    clock = DeviceClock(arena, group)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    batches = []
    for i in range(a.distinct_batches):          # synthetic blocks generated on the device
        n, k = a.batch_rows, a.nnz_per_row
        cols = torch.randint(0, a.cols, (n * k,), device=dev, generator=g)
        ucols, slot = torch.unique(cols, return_inverse=True)
        row0 = (i * n) % max(1, a.rows_per_gpu - n)
        batches.append(NMFBatch(torch.arange(0, n * k + 1, k, dtype=torch.int32, device=dev), slot.int(),
                                torch.rand(n * k, device=dev, generator=g),
                                torch.arange(row0, row0 + n, dtype=torch.int32, device=dev), ucols))

    if a.impl == "nccl":
        import torch.distributed as dist
        from harmony_b200.ops import native
        nat = native()
        R = table.to_dense().contiguous()                      # replicated model [cols, rank]
        Rp = torch.zeros((a.cols, table.row_stride), device=dev); Rp[:, : a.rank] = R
        G = torch.zeros_like(Rp)
        full_cols = [b.ucols[b.col_slot.long()].int() for b in batches]   # slot == column id on the replica

        def step_nccl(i):
            b = batches[i]
            G.zero_()
            nat.nmf_batch(b.csr_ptr, full_cols[i], b.val, model._Lbuf, b.l_row, Rp, G, a.rank, 1e-4, 0.0, 0.0, 1e6,
                          model.loss)
            if world > 1:
                dist.all_reduce(G)
            Rp.sub_(G, alpha=1e-4).clamp_(0.0, 1e6)

    def step(b, i=0):
        if a.impl == "nccl":
            return step_nccl(i % len(batches))
        clock.wait()
        model.step_batch(b)
        clock.tick()
    for i in range(a.warmup):
        step(batches[i % len(batches)], i)
    barrier_sync(dev, world)
    ops.reset_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(batches[i % len(batches)], i)
    e1.record()
    barrier_sync(dev, world)
    ms = max_over_ranks(e0.elapsed_time(e1), dev, world)
    clock.check()
    nnz = a.batch_rows * a.nnz_per_row
    emit(rank, {"metric": "mini-batches/sec (whole box, device-timed, max over ranks), NMF", "value": round(world * a.steps / (ms / 1e3), 2),
                "unit": "mini-batches/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True, "scaling": "weak", "dtype": "fp32",
                "data": "synthetic", "gpu_launches": ops.launches(), "impl": a.impl,
                "nnz_per_sec_per_gpu": round(nnz / (ms / a.steps / 1e3), 0),
                "config": {"model": f"NMF rank {a.rank}, {a.rows_per_gpu * world} x {a.cols} sparse, {a.nnz_per_row} nnz/row",
                           "batch_rows": a.batch_rows, "L_bytes_per_gpu": a.rows_per_gpu * a.rank * 4,
                           "parallelism": f"rows dp{world}, R sharded over {world} GPUs (pull/push over peer memory)"}})


if __name__ == "__main__":
    main()
