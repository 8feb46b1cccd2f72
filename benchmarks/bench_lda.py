#!/usr/bin/env python
"""LDA 1M-vocab x 1k-topic, sparse push/pull over peer memory (BASELINE config 4).

Word-topic table int32 [V+1, K+1] sharded over the GPUs (4 GB at 1M x 1k); each GPU owns a slice of
the documents and their topic assignments.  One step = pull the batch vocabulary rows (K1), warp-per-
document collapsed Gibbs sampling (K8), sparse (row, topic, +-1) pushes with clamp/nnz at the owners
(K2+K5).  Metric: mini-batches/sec whole box + tokens/sec.
"""
import argparse

from common import barrier_sync, emit, init_dist, max_over_ranks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--topics", type=int, default=1000)
    ap.add_argument("--docs_per_batch", type=int, default=2048)
    ap.add_argument("--doc_len", type=int, default=200)
    ap.add_argument("--distinct_batches", type=int, default=4)
    ap.add_argument("--zipf", type=float, default=1.1)
    ap.add_argument("--impl", default="harmony", choices=["harmony", "nccl"],
                    help="nccl: full table replica per GPU, sparse deltas all-gathered and applied everywhere (baseline)")
    a = ap.parse_args()
    import numpy as np
    import torch
    world, rank, lr, dev = init_dist()
    from harmony_b200 import ops
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.models.lda import LDADeviceModel, build_lda_batch
    from harmony_b200.parallel.device_clock import DeviceClock
    from harmony_b200.parallel.peer import PeerArena, PeerGroup
    group = PeerGroup.from_env() if a.impl == "harmony" else PeerGroup(0, 1, None)
    arena = PeerArena(lr, group, segment_bytes=1 << 30)
    V, K = a.vocab, a.topics
    conf = TableConfiguration(table_id="model_table0", dense=True, capacity=V + 1, row_dim=K + 1, dtype="int32",
                              num_total_blocks=1024, backend="device",
                              update_function={"name": "lda_sparse", "params": {"num_topics": K}})
    table = DeviceTable(conf, arena, group)
    ntok_batch = a.docs_per_batch * a.doc_len
    model = LDADeviceModel(table, K, V, ntok_batch * a.distinct_batches, alpha=0.1, beta=0.01, seed=rank)
    clock = DeviceClock(arena, group)
    if a.impl == "nccl" and world > 1:
        import torch.distributed as dist

        def exchange(ok, ot, od, sd):
            outs = []
            for t in (ok, ot, od):
                g = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
                dist.all_gather_into_tensor(g, t.contiguous())
                outs.append(g)
            sd = sd.clone()
            dist.all_reduce(sd)
            return outs[0], outs[1], outs[2], sd
        model.exchange = exchange
    rng = np.random.default_rng(7 + rank)
    batches = []
    for i in range(a.distinct_batches):           # word ids follow a Zipf law like real corpora
        words = (rng.zipf(a.zipf, size=ntok_batch) - 1) % V
        docs = list(words.astype(np.int32).reshape(a.docs_per_batch, a.doc_len))
        batches.append(build_lda_batch(docs, i * ntok_batch, V).to(dev))
    for b in batches:                              # initGlobalSettings: random assignment, push counts
        model.step_batch(b, init_mode=True)
    clock.barrier()

    def step(b):
        clock.wait()
        model.step_batch(b)
        clock.tick()
    for i in range(a.warmup):
        step(batches[i % len(batches)])
    barrier_sync(dev, world)
    ops.reset_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(batches[i % len(batches)])
    e1.record()
    barrier_sync(dev, world)
    ms = max_over_ranks(e0.elapsed_time(e1), dev, world)
    clock.check()
    uvoc = int(np.mean([b.vocab.numel() for b in batches]))
    emit(rank, {"metric": "mini-batches/sec (whole box, device-timed, max over ranks), LDA", "value": round(world * a.steps / (ms / 1e3), 2),
                "unit": "mini-batches/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True, "scaling": "weak", "dtype": "int32/fp32",
                "data": "synthetic", "gpu_launches": ops.launches(), "impl": a.impl,
                "tokens_per_sec_whole_box": round(world * ntok_batch / (ms / a.steps / 1e3), 0),
                "pull_bytes_per_step_per_gpu": uvoc * table.row_stride * 4,
                "config": {"model": f"LDA {V}-vocab x {K}-topic (int32 table {(V + 1) * table.row_stride * 4 / 2**30:.2f} GiB sharded)",
                           "docs_per_batch": a.docs_per_batch, "doc_len": a.doc_len, "batch_vocab": uvoc,
                           "parallelism": f"docs dp{world}, word-topic table sharded over {world} GPUs"}})


if __name__ == "__main__":
    main()
