"""Correctness + speed of the cta_group::2 kernels against fp64 references / the 1-CTA kernel."""
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from harmony_b200.ops import native
from harmony_b200.parallel.peer import PeerArena, PeerGroup
from harmony_b200.et.config import TableConfiguration
from harmony_b200.et.device_table import DeviceTable
n = native()
arena = PeerArena(0, PeerGroup(), segment_bytes=1 << 30)

def table(F, C, tag):
    return DeviceTable(TableConfiguration(table_id=tag, dense=True, capacity=F, row_dim=C, num_total_blocks=max(1, F // 256),
                                          is_ordered=True, backend="device",
                                          update_function={"name": "add", "params": {"init": "gaussian", "std": 0.05, "seed": 3}}), arena)

def check(B, F, C):
    t = table(F, C, f"t{B}{F}{C}")
    X = torch.randn((B, F), device="cuda")
    W = t.to_dense()
    maps = n.make_block_tensormaps(t.block_ptrs.cpu(), t.rows_per_block, t.row_stride, 0, 1)
    out = torch.full((B, t.row_stride), -7.0, device="cuda")
    n.dense_logits_tf32(X, maps, t.rows_per_block, t.row_stride, out, use_2cta=True)
    torch.cuda.synchronize()
    ref = X.double() @ W.double()
    err = float((out[:, :C].double() - ref).abs().max() / ref.abs().max())
    # backward: W += X^T D
    Bp = (B + 31) // 32 * 32
    Xp = torch.zeros((Bp, F), device="cuda"); Xp[:B] = X
    D = torch.zeros((Bp, t.row_stride), device="cuda"); D[:B, :C] = torch.randn((B, C), device="cuda") * 0.01
    W0 = t.to_dense().clone()
    bp, rpb, st, dim, mode = t.view()
    n.dense_grad_push_tf32(Xp, D, bp, rpb, st, C, mode, use_2cta=True)
    torch.cuda.synchronize()
    Wref = W0.double() + Xp.double().t() @ D[:, :C].double()
    upd = float((Wref - W0.double()).abs().max())
    err2 = float((t.to_dense().double() - Wref).abs().max() / upd)
    print(json.dumps({"B": B, "F": F, "C": C, "fwd_rel_err": err, "bwd_rel_err": err2}), flush=True)
    t.close()
    return err < 3e-3 and err2 < 5e-3

ok = all([check(256, 64, 256), check(300, 1024, 1000), check(1000, 512, 100)])
print("2CTA_OK" if ok else "2CTA_MISMATCH", flush=True)
if ok:
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = []
    for B, F, C in [(4096, 8192, 1000), (8192, 16384, 1000), (16384, 8192, 4000)]:
        t = table(F, C, f"b{B}{F}{C}")
        X = torch.randn((B, F), device="cuda")
        maps = n.make_block_tensormaps(t.block_ptrs.cpu(), t.rows_per_block, t.row_stride, 0, 1)
        out = torch.empty((B, t.row_stride), device="cuda")
        def timed(fn, k=10):
            for _ in range(3): fn()
            ts = []
            for _ in range(k):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            return sorted(ts)[len(ts) // 2]
        t2 = timed(lambda: n.dense_logits_tf32(X, maps, t.rows_per_block, t.row_stride, out, use_2cta=True))
        t1 = timed(lambda: n.dense_logits_tf32(X, maps, t.rows_per_block, t.row_stride, out, use_2cta=False))
        flops = 2.0 * B * F * t.row_stride
        r = {"B": B, "F": F, "C": C, "two_cta_ms": round(t2, 4), "two_cta_tflops": round(flops / t2 / 1e9, 1),
             "one_cta_ms": round(t1, 4), "one_cta_tflops": round(flops / t1 / 1e9, 1)}
        print(json.dumps(r), flush=True); res.append(r)
        t.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/tc_2cta_bench.json", "w"), indent=1)
