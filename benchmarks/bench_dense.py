#!/usr/bin/env python
"""Dense-input MLR on the tensor cores: logits = X[B, F] . W[F, C] as a tcgen05 kind::tf32 GEMM whose B tiles are
TMA-loaded straight from the model table's block slabs (pull fused into the GEMM), timed next to cuBLAS tf32 on the
same tensors, with the clocks sampled during the timed loops (`python bench.py --config dense`).
One GPU; L2 is flushed between timed iterations (256 MB write).  Numerics: max relative error vs an fp64 matmul."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--shapes", default="4096x8192x1000,8192x16384x1000,16384x8192x4000")
    a, _ = ap.parse_known_args()
    import torch
    from bench import ClockSampler
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.models.mlr_dense import DenseMLR
    from harmony_b200.parallel.peer import PeerArena, PeerGroup
    torch.cuda.set_device(0)
    arena = PeerArena(0, PeerGroup(), segment_bytes=1 << 30)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn):
        for _ in range(max(3, a.warmup)):
            fn()
        ts = []
        for _ in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[0]
    res = []
    sampler = ClockSampler(0).start()
    for shape in a.shapes.split(","):
        B, F, C = (int(x) for x in shape.split("x"))
        conf = TableConfiguration(table_id=f"w{B}{F}{C}", dense=True, capacity=F, row_dim=C, num_total_blocks=F // 256,
                                  is_ordered=True, backend="device",
                                  update_function={"name": "add", "params": {"init": "gaussian", "std": 0.02, "seed": 1}})
        t = DeviceTable(conf, arena)
        X = torch.randn((B, F), device="cuda")
        W = t.to_dense()
        Wp = torch.zeros((F, t.row_stride), device="cuda"); Wp[:, :C] = W
        m = DenseMLR(t, C)
        msk = DenseMLR(t, C, stream_k=True)
        out = torch.empty((B, t.row_stride), device="cuda")
        t_tc, t_tc_min = timed(lambda: m.logits(X, out))
        t_sk, _ = timed(lambda: msk.logits(X, out))
        torch.backends.cuda.matmul.allow_tf32 = True
        t_cb, t_cb_min = timed(lambda: torch.matmul(X, Wp, out=out))
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = X.double() @ W.double()
        err = float((m.logits(X)[:, :C].double() - ref).abs().max() / ref.abs().max())
        flops = 2.0 * B * F * t.row_stride
        res.append({"B": B, "F": F, "C": C, "tcgen05_ms": round(t_tc, 4), "tcgen05_tflops": round(flops / t_tc / 1e9, 1),
                    "tcgen05_tflops_best": round(flops / t_tc_min / 1e9, 1),
                    "tcgen05_persistent_stream_k_tflops": round(flops / t_sk / 1e9, 1),
                    "cublas_tf32_ms": round(t_cb, 4), "cublas_tf32_tflops": round(flops / t_cb / 1e9, 1),
                    "frac_of_cublas_tf32": round(t_cb / t_tc, 3), "rel_err_vs_fp64": err})
        t.close()
    clocks = sampler.stop()
    head = res[0]
    print(json.dumps({"metric": "TFLOP/s (tf32, device-timed median), dense-input MLR logits GEMM with the pull fused "
                                "(B operand TMA-loaded from the table's block slabs)",
                      "value": head["tcgen05_tflops"], "unit": "TFLOP/s", "n_gpus": 1, "steps": a.steps,
                      "warmup": max(3, a.warmup), "ms_per_step": head["tcgen05_ms"], "higher_is_better": True,
                      "dtype": "tf32 (fp32 operands, fp32 accumulate)", "data": "synthetic",
                      "vs_baseline": head["frac_of_cublas_tf32"], "vs_baseline_what": "cuBLAS tf32 on the same tensors",
                      "clocks": clocks, "gpu_launches": a.steps, "shapes": res,
                      "config": {"model": f"dense MLR logits {head['B']}x{head['F']}x{head['C']}",
                                 "l2_flush": "256 MB write between timed iterations"}}), flush=True)


if __name__ == "__main__":
    main()
