"""Single-shape driver for ncu captures of the tcgen05 dense MLR kernels."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from harmony_b200.parallel.peer import PeerArena, PeerGroup
from harmony_b200.et.config import TableConfiguration
from harmony_b200.et.device_table import DeviceTable
from harmony_b200.models.mlr_dense import DenseMLR
arena = PeerArena(0, PeerGroup(), segment_bytes=1 << 30)
B, F, C = 4096, 8192, 1000
t = DeviceTable(TableConfiguration(table_id="w", dense=True, capacity=F, row_dim=C, num_total_blocks=F // 256, is_ordered=True,
                                   backend="device", update_function={"name": "add", "params": {"init": "gaussian", "std": 0.02}}), arena)
X = torch.randn((B, F), device="cuda")
y = torch.randint(0, C, (B,), device="cuda", dtype=torch.int32)
m = DenseMLR(t, C)
for _ in range(4):
    m.step(X, y, 1e-3)
torch.cuda.synchronize()
print("done")
