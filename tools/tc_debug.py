import sys, os
sys.path.insert(0, os.getcwd())
import torch
from harmony_b200.ops import native
from harmony_b200.parallel.peer import PeerArena, PeerGroup
from harmony_b200.et.config import TableConfiguration
from harmony_b200.et.device_table import DeviceTable
n = native()
arena = PeerArena(0, PeerGroup(), segment_bytes=64 << 20)
F, C, B = 64, 128, 128
conf = TableConfiguration(table_id="dbg", dense=True, capacity=F, row_dim=C, num_total_blocks=1, is_ordered=True, backend="device", update_function="add")
t = DeviceTable(conf, arena)
blk = t.local_block(0)
W = torch.arange(F * C, device="cuda", dtype=torch.float32).reshape(F, C) % 7 - 3.0
blk[:F, :C] = W
X = (torch.arange(B * F, device="cuda", dtype=torch.float32).reshape(B, F) % 5) - 2.0
ref = X @ W
def run(tag, maps, **kw):
    out = torch.full((B, t.row_stride), -777.0, device="cuda")
    n.dense_logits_tf32(X, maps, t.rows_per_block, t.row_stride, out, **kw)
    torch.cuda.synchronize()
    err = (out[:, :C] - ref).abs().max().item()
    print(tag, "err", err, "ref max", ref.abs().max().item(), "out[0,:4]", out[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist(),
          "zeros", int((out == 0).sum()), flush=True)
    return err
for swz in (1, 0):
    maps = n.make_block_tensormaps(t.block_ptrs.cpu(), t.rows_per_block, t.row_stride, 0, swz)
    for layout in (1, 2):
        for lbo, sbo in ((4096, 512), (4096, 1024), (512, 4096), (4096, 256), (1024, 512)):
            run(f"tma_swz={swz} layout={layout} lbo={lbo} sbo={sbo}", maps, b_lbo=lbo, b_sbo=sbo, b_layout=layout)
