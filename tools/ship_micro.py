"""1-GPU micro-benchmark of the ship-mode kernels under the 8-GPU WORKLOAD SHAPE (emulated): this GPU plays one
owner of an 8-way sharded 1M x 1000 model -- 125k local feature rows, 8 x 4096 instances with ~12.5 local non-zeros
each -- with every "peer" buffer mapped to local memory.  Isolates kernel structure from NVLink effects.
  python tools/ship_micro.py [world]"""
import json
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

from harmony_b200.et.config import TableConfiguration
from harmony_b200.et.device_table import DeviceTable
from harmony_b200.models.mlr import build_batch
from harmony_b200.ops import native
from harmony_b200.parallel.peer import PeerArena, PeerGroup

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B, C, NNZ, F = 4096, 1000, 100, 1_000_000
Fl = F // world
n = native()
arena = PeerArena(0, PeerGroup(), segment_bytes=2 << 30)
t = DeviceTable(TableConfiguration(table_id="w", dense=True, capacity=Fl, row_dim=C, num_total_blocks=128,
                                   is_ordered=False, backend="device",
                                   update_function={"name": "add", "params": {"init": "gaussian", "std": 0.001, "seed": 1}}), arena)
stride = t.row_stride
rng = np.random.default_rng(0)
batches = []
for i in range(4):
    cnt = rng.binomial(NNZ, 1.0 / world, size=world * B)
    ptr = np.zeros(world * B + 1, dtype=np.int32); ptr[1:] = np.cumsum(cnt)
    col = rng.integers(0, Fl, size=int(ptr[-1])).astype(np.int32)
    val = rng.random(int(ptr[-1]), dtype=np.float32)
    batches.append(build_batch(ptr, col, val, np.zeros(world * B, dtype=np.int32)).to("cuda"))
rows = world * B
P = torch.zeros((rows, stride), device="cuda")
D = torch.randn((rows, stride), device="cuda") * 1e-6
L = torch.zeros((world, B, stride), device="cuda")            # 8 "workers'" logits buffers, all local
Lptrs = torch.tensor([L[w].data_ptr() for w in range(world)], dtype=torch.int64, device="cuda")
n_tiles = B // 32
tile_done = torch.zeros(world * n_tiles, dtype=torch.int32, device="cuda")
pflag = torch.zeros((world, n_tiles), dtype=torch.int32, device="cuda")
pfptrs = torch.tensor([pflag[w].data_ptr() for w in range(world)], dtype=torch.int64, device="cuda")
err = torch.zeros(1, dtype=torch.int32, device="cuda")
n.ship_preload(0)
bp, rpb, st, _dim, mode = t.view()
epoch = [0]


def timed(fn, reps=10):
    for i in range(3):
        fn(batches[i % 4])
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(batches[i % 4]); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(ts[len(ts) // 2], 4)


def v1_fwd(b):
    n.mlr_partial_forward(bp, rpb, st, C, mode, b.csr_ptr, b.csr_col, b.csr_val, B, Lptrs)


def v2_fwd(sig):
    def f(b):
        epoch[0] += 1
        n.ship_forward(bp, rpb, st, C, mode, b.csr_ptr, b.csr_col, b.csr_val, world, B, 32, P.data_ptr(),
                       tile_done.data_ptr(), pfptrs if sig else None, epoch[0])
    return f


def v3_fwd(sig):
    def f(b):
        epoch[0] += 1
        n.ship3_forward(bp, rpb, st, C, mode, b.csr_ptr, b.csr_col, b.csr_val, world, B, 32, Lptrs,
                        tile_done.data_ptr(), pfptrs if sig else None, epoch[0])
    return f


def k3(b):
    n.mlr_backward_push(bp, rpb, st, C, mode, b.uniq_feat, b.feat_ptr, b.samp, b.xval, D, False)


out = {"world_emulated": world, "nnz_local": int(batches[0].csr_col.numel()), "uniq_local": int(batches[0].uniq_feat.numel())}
out["v1_partial_fwd_ms"] = timed(v1_fwd)
for env in ("cta", "warp"):
    os.environ["HB_SHIP_K1"] = env        # read once per process: only the first value sticks (static), so report it
out["v2_fwd_nosignal_ms"] = timed(v2_fwd(False))
tile_done.zero_(); epoch[0] = 0
out["v2_fwd_signal_ms"] = timed(v2_fwd(True))
tile_done.zero_(); epoch[0] = 0
out["v3_fwd_nosignal_ms"] = timed(v3_fwd(False))
tile_done.zero_(); epoch[0] = 0
out["v3_fwd_signal_ms"] = timed(v3_fwd(True))
out["k3_update_ms"] = timed(k3)
gb = out["nnz_local"] * stride * 4 / 1e9
out["gather_GB"] = round(gb, 3)
out["v3_fwd_TBps"] = round(gb / out["v3_fwd_signal_ms"], 3)
out["k3_rmw_GB"] = round(out["uniq_local"] * stride * 8 / 1e9, 3)
print(json.dumps(out))
