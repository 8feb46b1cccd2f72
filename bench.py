#!/usr/bin/env python
"""Headline benchmark: MLR 1M-feature x 1000-class, model sharded over N B200s (BASELINE.json).

Metric: mini-batches/sec, whole box (N workers, one mini-batch per worker per step), device
timed, max over ranks.  Every GPU is both a shard owner and a worker (the co-located mode of the
reference job server, DolphinJobEntity.java:80-82).  One step of one worker =
  PULL  rows of the mini-batch's features from the owners' HBM (fused with the sparse dot + softmax)
  PUSH  one reduce-add row per distinct feature into the owners' HBM (fused with `w += delta`)
  SSP   device clock publish/wait (clock_slack, default 0 like the reference)
Weak scaling: per-GPU work (one mini-batch of --batch instances) is fixed as N grows.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N --steps K --warmup W
  --impl reference : the unmodified reference (Java/REEF; not installable here -> "unavailable")
  --impl nccl      : NCCL all-reduce baseline of the same app (replicated model, dense gradient)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="harmony", choices=["harmony", "reference", "nccl"])
    ap.add_argument("--features", type=int, default=1_000_000)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=4096, help="instances per mini-batch per worker")
    ap.add_argument("--nnz", type=int, default=100, help="non-zeros per instance")
    ap.add_argument("--num_server_blocks", type=int, default=1024)
    ap.add_argument("--step_size", type=float, default=1e-4)
    ap.add_argument("--clock_slack", type=int, default=0)
    ap.add_argument("--distinct_batches", type=int, default=4)
    ap.add_argument("--zipf", type=float, default=0.0)
    ap.add_argument("--no_e2e", action="store_true")
    ap.add_argument("--ship_phases", action="store_true",
                    help="ship mode: also time the five parts of a step (partial fwd / barrier / softmax+bcast / "
                         "barrier / update) with CUDA events, outside the headline region")
    ap.add_argument("--mode", default="auto", choices=["auto", "pull", "ship"],
                    help="pull: workers read/reduce model rows over NVLink (PS pull/push); "
                         "ship: owners compute the sparse dot, only [B,C] activations cross NVLink; "
                         "auto: ship when N > 1")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi SM clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    # NVML bit masks of nvmlDeviceGetCurrentClocksEventReasons (nvml.h)
    NVML_REASONS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def _nvml_sampler(self):
        """Fast path: in-process NVML queries (~0.1 ms each) so even a 15 ms timed region gets several samples.
        Returns a zero-argument callable producing one sample in the nvidia-smi column layout, or None."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)

            def sample():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = int(get_reasons(h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:  # noqa: BLE001
                    pw = 0.0
                flag = lambda nm: "Active" if mask & self.NVML_REASONS[nm] else "Not Active"  # noqa: E731
                return [str(sm), str(mx), f"{pw:.1f}", flag("hw_slowdown"), flag("hw_thermal_slowdown"),
                        flag("sw_thermal_slowdown"), flag("sw_power_cap")]
            sample()                      # fail here, not in the loop
            return sample
        except Exception:  # noqa: BLE001
            return None

    def _smi_sample(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        return parts if len(parts) >= 7 else None

    def start(self):
        fast = self._nvml_sampler()
        self.source = "nvml" if fast else "nvidia-smi"

        def run():
            while not self._stop.is_set():
                try:
                    parts = fast() if fast else self._smi_sample()
                    if parts:
                        self.samples.append(parts)
                except Exception:  # noqa: BLE001
                    pass
                self._stop.wait(0.003 if fast else 0.1)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for nm, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples), "source": getattr(self, "source", None)}


def reference_arm(args):
    """The reference is 100% Java 8 / Maven / Apache REEF (no setup.py / pyproject): it cannot be
    pip-installed into baseline/_ref and there is no JDK in the image."""
    ok = os.path.isdir(os.path.join(ROOT, "baseline", "_ref")) and any(
        n.endswith(".jar") for _, _, fs in os.walk(os.path.join(ROOT, "baseline", "_ref")) for n in fs)
    if not ok or subprocess.run(["which", "java"], capture_output=True).returncode != 0:
        print(json.dumps({"impl": "reference",
                          "unavailable": "snuspl/harmony is Java 8/Maven/REEF: no Python package to pip-install, "
                                         "no JDK/Maven/~/.m2 in the image, no network"}))
        return 0
    print(json.dumps({"impl": "reference", "unavailable": "JVM launch of the reference is not wired"}))
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    from harmony_b200 import ops
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.models.mlr import MLRDeviceModel, synthetic_batch
    from harmony_b200.parallel.device_clock import DeviceClock
    from harmony_b200.parallel.peer import PeerArena, PeerGroup

    group = PeerGroup.from_env()
    F, C, B = args.features, args.classes, args.batch

    def barrier_sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # synthetic input blocks (one block == one mini-batch), pinned on the host + resident in HBM
    host_batches = [synthetic_batch(B, F, C, args.nnz, seed=1000 * rank + i, zipf=args.zipf).pin()
                    for i in range(args.distinct_batches)]
    dev_batches = [hb.to(dev) for hb in host_batches]
    h2d_bytes = host_batches[0].nbytes()

    mode = args.mode if args.mode != "auto" else ("ship" if world > 1 else "pull")
    if args.impl != "harmony":
        mode = "n/a"
    if args.impl == "harmony":
        arena = PeerArena(local_rank, group, segment_bytes=2 << 30)
        conf = TableConfiguration(table_id="model_table0", dense=True, capacity=F, row_dim=C,
                                  num_total_blocks=args.num_server_blocks, is_ordered=False, backend="device",
                                  update_function={"name": "add", "params": {"init": "gaussian", "std": 0.001,
                                                                              "seed": 42}})
        table = DeviceTable(conf, arena, group)
        model = MLRDeviceModel(table, C, args.step_size, max_batch=B)
        clock = DeviceClock(arena, group, slack=args.clock_slack)
        if mode == "ship":
            from harmony_b200.models.mlr_ship import MLRShipModel, route_batches
            ship = MLRShipModel(table, arena, group, clock, C, args.step_size, B)
            # 2-D partition of the input: every rank derives its owner-side slices of ALL workers'
            # blocks (same synthetic seeds on every rank: no exchange needed for the benchmark)
            routed_host = []
            for i in range(args.distinct_batches):
                allw = [host_batches[i] if r == rank else
                        synthetic_batch(B, F, C, args.nnz, seed=1000 * r + i, zipf=args.zipf) for r in range(world)]
                routed_host.append(route_batches(table, allw, rank).pin())
            host_batches = routed_host
            dev_batches = [rb.to(dev) for rb in host_batches]
            h2d_bytes = host_batches[0].nbytes()
            model.stats = ship.stats

            def step(batch):
                ship.step(batch)
            parallelism = (f"ps-sharded-model x{world} (co-located), owner-side sparse dot; [B,C] activations "
                           f"reduce-scattered / all-gathered over peer memory; {args.num_server_blocks} blocks")
        else:
            def step(batch):
                clock.wait()
                model.forward(batch, True)
                model.backward_push(batch)
                clock.tick()
            parallelism = (f"ps-sharded-model x{world} (co-located workers), rows pulled / pushed over peer "
                           f"memory; {args.num_server_blocks} blocks")
    else:  # NCCL all-reduce baseline: replicated model, dense gradient all-reduce, dense SGD update
        arena = PeerArena(local_rank, PeerGroup(0, 1, None), segment_bytes=2 << 30)
        one = PeerGroup(0, 1, None)
        conf = TableConfiguration(table_id="model_replica", dense=True, capacity=F, row_dim=C,
                                  num_total_blocks=1, is_ordered=True, backend="device",
                                  update_function={"name": "add", "params": {"init": "gaussian", "std": 0.001,
                                                                              "seed": 42}})
        table = DeviceTable(conf, arena, one)
        gconf = TableConfiguration(table_id="grad_replica", dense=True, capacity=F, row_dim=C, num_total_blocks=1,
                                   is_ordered=True, backend="device", update_function="add")
        gtable = DeviceTable(gconf, arena, one)
        model = MLRDeviceModel(table, C, args.step_size, max_batch=B)
        gmodel = MLRDeviceModel(gtable, C, args.step_size, max_batch=B)
        gmodel._D = model._D
        W, G = table.local_block(0), gtable.local_block(0)

        def step(batch):
            G.zero_()
            model.forward(batch, True)
            gmodel.backward_push(batch)           # dense local gradient (pre-scaled by -step)
            if world > 1:
                dist.all_reduce(G)
            W.add_(G)
        clock = None
        parallelism = f"dp{world} replicated model + NCCL all-reduce of the dense gradient"

    # ---------------- device-timed run ----------------
    for i in range(args.warmup):
        step(dev_batches[i % len(dev_batches)])
    barrier_sync()
    ops.reset_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    model.stats.zero_()
    barrier_sync()
    ev0.record()
    for i in range(args.steps):
        step(dev_batches[i % len(dev_batches)])
    ev1.record()
    barrier_sync()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ops.launches()
    clocks = sampler.stop() if rank == 0 else None
    if clock is not None:
        clock.check()
    loss_dev = float(model.stats[0]) / max(1, args.steps * B)

    # ---------------- per-phase device timing (outside the headline region) ----------------
    phase = None
    if args.impl == "harmony" and mode == "pull":
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_pull = t_push = 0.0
        reps = min(args.steps, 10)
        for i in range(reps):
            b = dev_batches[i % len(dev_batches)]
            barrier_sync()
            evs[0].record(); model.forward(b, True); evs[1].record(); model.backward_push(b); evs[2].record()
            torch.cuda.synchronize(dev)
            t_pull += evs[0].elapsed_time(evs[1]); t_push += evs[1].elapsed_time(evs[2])
        phase = {"pull_fwd_ms": max_over_ranks(t_pull / reps), "bwd_push_ms": max_over_ranks(t_push / reps)}
    elif args.impl == "harmony" and mode == "ship" and args.ship_phases:
        reps = min(args.steps, 10)
        acc = {}
        for i in range(reps):
            barrier_sync()
            for k, v in ship.step_timed(dev_batches[i % len(dev_batches)]).items():
                acc[k] = acc.get(k, 0.0) + v
        phase = {k: max_over_ranks(v / reps) for k, v in sorted(acc.items())}

    # ---------------- end-to-end: pinned host -> H2D every step, loss D2H every step ----------------
    e2e = None
    if not args.no_e2e:
        from harmony_b200.models.mlr import MLRBatch

        def rebuild(views):
            if mode == "ship":
                from harmony_b200.models.mlr_ship import RoutedBatch
                return RoutedBatch(MLRBatch(*views[:8]), views[8])
            return MLRBatch(*views)
        maxn = [max(t.numel() for t in ts) for ts in zip(*[hb.tensors() for hb in host_batches])]
        staging = [[torch.empty(n, dtype=t.dtype, device=dev) for n, t in zip(maxn, host_batches[0].tensors())]
                   for _ in range(2)]                       # double-buffered device staging

        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
        ready = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(i):
            """H2D of step i's inputs from pinned host memory on the copy stream (double buffered: the
            buffer was last read by step i-2, whose loss has already been read back => it is free)."""
            hbt = host_batches[i % len(host_batches)]
            views = []
            with torch.cuda.stream(copy_stream):
                for dst, src in zip(staging[i % 2], hbt.tensors()):
                    v = dst[: src.numel()]
                    v.copy_(src, non_blocking=True)
                    views.append(v)
                ready[i % 2].record(copy_stream)
            return views

        def e2e_run(n):
            last = 0.0
            nxt = prefetch(0)
            for i in range(n):
                views = nxt
                main_stream.wait_event(ready[i % 2])        # this step's inputs have landed
                model.stats.zero_()
                step(rebuild(views))
                if i + 1 < n:
                    nxt = prefetch(i + 1)                   # overlaps with the kernels of step i
                last = float(model.stats.cpu()[0])          # D2H of this step's loss (syncs the step)
            return last
        e2e_run(max(3, args.warmup))
        barrier_sync()
        t0 = time.perf_counter()
        last_loss = e2e_run(args.steps)
        barrier_sync()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {"value": round(world * args.steps / (e2e_ms / 1e3), 3), "unit": "mini-batches/s",
               "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 8,
               "ms_per_step": round(e2e_ms / args.steps, 4), "last_loss_per_instance": round(last_loss / B, 5)}

    if rank == 0:
        value = world * args.steps / (ms / 1e3)
        uniq = int(np.mean([(hb.routed if mode == "ship" else hb).uniq_feat.numel() for hb in host_batches]))
        nnz = B * args.nnz
        row_bytes = table.row_stride * 4
        remote_frac = (world - 1) / world
        out = {
            "metric": "mini-batches/sec (whole box, device-timed, max over ranks), MLR 1M-feature x 1000-class",
            "value": round(value, 3), "unit": "mini-batches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "impl": args.impl,
            "config": {"model": f"MLR {F}-feature x {C}-class (fp32 master weights, {F * row_bytes / 2**30:.2f} GiB sharded)",
                       "global_batch": B * world, "batch_per_worker": B, "nnz_per_instance": args.nnz,
                       "seq_len": None, "parallelism": parallelism, "mode": mode,
                       "clock_slack": 0 if mode == "ship" else args.clock_slack,
                       "l2_flush": "inputs larger than L2: ~%.2f GB of model rows touched per step per GPU vs 126 MB L2"
                                   % ((nnz + uniq) * row_bytes / 1e9),
                       "sgd": "step_size %.1e, decay_rate 0.9 / decay_period 5 epochs (applied per epoch)" % args.step_size},
            "clocks": clocks, "gpu_launches": launches, "loss_per_instance": round(loss_dev, 5),
            "phases": None if phase is None else {k: round(v, 4) for k, v in phase.items()}
            if "pull_fwd_ms" not in phase else {
                "pull_fwd_ms": round(phase["pull_fwd_ms"], 4), "bwd_push_ms": round(phase["bwd_push_ms"], 4),
                "pull_GBps_per_gpu": round(nnz * row_bytes / (phase["pull_fwd_ms"] / 1e3) / 1e9, 1),
                "push_GBps_per_gpu": round(uniq * row_bytes / (phase["bwd_push_ms"] / 1e3) / 1e9, 1),
                "pull_nvlink_GBps_per_gpu": round(nnz * row_bytes * remote_frac / (phase["pull_fwd_ms"] / 1e3) / 1e9, 1),
                "push_nvlink_GBps_per_gpu": round(uniq * row_bytes * remote_frac / (phase["bwd_push_ms"] / 1e3) / 1e9, 1)},
            "nvlink_bytes_per_step_per_gpu": int((2 * B * world * row_bytes if mode == "ship"
                                                   else (nnz + uniq) * row_bytes) * remote_frac),
        }
        if e2e is not None:
            out["e2e"] = e2e
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
