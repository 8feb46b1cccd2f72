"""ONE process, TWO GPUs: the pull / push / ship kernels of GPU 0 working on a model table whose blocks alternate
between GPU 0 and GPU 1 (VMM segments, vmm_grant peer access).  Meant to run under ncu (a single process -- never a
multi-rank launch) to read the NVLink byte counters of the fused compute+collective kernels:
  ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:'mlr_|ship|migrate' --csv --log-file gpurun_out/nvlink.csv python tools/nvlink_probe.py
"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from harmony_b200.et.config import TableConfiguration
from harmony_b200.et.device_table import DeviceTable
from harmony_b200.models.mlr import MLRDeviceModel, synthetic_batch
from harmony_b200.parallel.peer import PeerArena, PeerGroup

assert torch.cuda.device_count() >= 2
F, C, B, NNZ = 200_000, 1000, 2048, 100
arenas = {k: PeerArena(d, PeerGroup(), segment_bytes=1 << 30, self_key=k) for d, k in ((0, "a"), (1, "b"))}
conf = TableConfiguration(table_id="w", dense=True, capacity=F, row_dim=C, num_total_blocks=64, is_ordered=False,
                          backend="device", update_function={"name": "add", "params": {"init": "gaussian", "std": 0.001, "seed": 1}})
owners = ["a" if b % 2 == 0 else "b" for b in range(64)]
tables = {}
for k, dev in (("a", 0), ("b", 1)):
    torch.cuda.set_device(dev)
    tables[k] = DeviceTable(conf, arenas[k], PeerGroup(), block_owners=owners, collective=False)
infos = [tables[k].local_info() for k in ("a", "b")]
for k, dev in (("a", 0), ("b", 1)):
    torch.cuda.set_device(dev)
    tables[k].apply_infos(infos)
    torch.cuda.synchronize(dev)
torch.cuda.set_device(0)
t = tables["a"]
model = MLRDeviceModel(t, C, 1e-4, max_batch=B)
batch = synthetic_batch(B, F, C, NNZ, seed=3).to("cuda:0")
for _ in range(3):
    model.step(batch)              # mlr_fwd_kernel: half of the 4 KB rows are peer loads; bwd: half are peer red.adds
torch.cuda.synchronize(0)
# live migration of 8 blocks from GPU 0 to GPU 1: the bulk-TMA copy writes the new slabs over NVLink
moving = [b for b in range(64) if owners[b] == "a"][:8]
torch.cuda.set_device(1)
locs = tables["b"].alloc_recv(moving)
torch.cuda.set_device(0)
ms = t.send_blocks(locs)
for k, dev in (("a", 0), ("b", 1)):
    torch.cuda.set_device(dev)
    tables[k].flip_blocks(locs, "b")
torch.cuda.set_device(0)
t.drain_blocks(moving)
torch.cuda.synchronize(0)
nbytes = len(moving) * t.block_bytes
print(f"probe ok: remote fraction 0.5, migration {nbytes / 2**20:.0f} MiB in {ms:.3f} ms = {nbytes / ms / 1e6:.1f} GB/s")
