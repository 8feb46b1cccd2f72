"""Times the tcgen05 dense-logits kernel against cuBLAS (torch.matmul tf32 / bf16) on one GPU."""
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from harmony_b200.parallel.peer import PeerArena, PeerGroup
from harmony_b200.et.config import TableConfiguration
from harmony_b200.et.device_table import DeviceTable
from harmony_b200.models.mlr_dense import DenseMLR

arena = PeerArena(0, PeerGroup(), segment_bytes=1 << 30)
res = []
for B, F, C in [(4096, 8192, 1000), (8192, 16384, 1000), (16384, 8192, 4000)]:
    conf = TableConfiguration(table_id=f"w{B}{F}{C}", dense=True, capacity=F, row_dim=C, num_total_blocks=F // 256,
                              is_ordered=True, backend="device",
                              update_function={"name": "add", "params": {"init": "gaussian", "std": 0.02, "seed": 1}})
    t = DeviceTable(conf, arena)
    X = torch.randn((B, F), device="cuda")
    W = t.to_dense()
    Wp = torch.zeros((F, t.row_stride), device="cuda"); Wp[:, :C] = W
    m = DenseMLR(t, C)
    out = torch.empty((B, t.row_stride), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]
    t_tc = timed(lambda: m.logits(X, out))
    torch.backends.cuda.matmul.allow_tf32 = True
    t_cublas_tf32 = timed(lambda: torch.matmul(X, Wp, out=out))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_fp32 = timed(lambda: torch.matmul(X, Wp, out=out))
    ref = X.double() @ W.double()
    err = float((m.logits(X)[:, :C].double() - ref).abs().max() / ref.abs().max())
    flops = 2.0 * B * F * t.row_stride
    r = {"B": B, "F": F, "C": C, "tcgen05_ms": round(t_tc, 4), "tcgen05_tflops": round(flops / t_tc / 1e9, 1),
         "cublas_tf32_ms": round(t_cublas_tf32, 4), "cublas_tf32_tflops": round(flops / t_cublas_tf32 / 1e9, 1),
         "cublas_fp32_ms": round(t_fp32, 4), "rel_err": err}
    print(json.dumps(r), flush=True)
    res.append(r)
    t.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/tc_bench.json", "w"), indent=1)
