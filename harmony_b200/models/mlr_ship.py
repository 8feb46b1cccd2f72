"""MLR in "ship the activations" mode (see ops/csrc/mlr_ship_ops.cu).

Every GPU is owner AND worker (the reference's co-located mode).  The input table is 2-D
partitioned at load time: worker w keeps its instances' labels, owner g keeps -- for every worker --
the non-zeros whose feature it owns (`RoutedBatch`).  One BSP step then moves only [B, C]
activations over NVLink:
    partial logits  --red.add-->  worker      (compute fused with reduce-scatter over peer memory)
    error rows D    --store--->   every owner (compute fused with all-gather over peer memory)
while all the 4 KB model-row traffic (gather for the dot, read-modify-write for the update) stays in
local HBM.  Same arithmetic as the pull/push mode; clock_slack is 0 by construction.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from ..et.device_table import DeviceTable
from ..et.partitioner import PART_ORDERED
from ..parallel.device_clock import DeviceClock
from ..parallel.peer import PeerArena, PeerGroup
from .mlr import MLRBatch, build_batch


def owner_of_features(table: DeviceTable, feats: np.ndarray, owner_ranks: Optional[Sequence[int]] = None) -> np.ndarray:
    """Rank of the GPU owning each feature row.  `owner_ranks[b]` overrides table.owners when the owners are
    executor ids (job-server mode) instead of SPMD ranks."""
    feats = np.asarray(feats, dtype=np.int64)
    blocks = feats // table.rows_per_block if table.mode == PART_ORDERED else feats % table.num_blocks
    owners = table.owners if owner_ranks is None else owner_ranks
    return np.asarray(owners, dtype=np.int64)[blocks]


@dataclass
class RoutedBatch:
    """Owner-side view of one BSP step: rows = world * B instances (row = worker * B + b)."""
    routed: MLRBatch          # CSR + by-feature grouping restricted to the features this GPU owns
    labels: torch.Tensor      # int32 [B] labels of THIS worker's instances

    def to(self, device, non_blocking=True) -> "RoutedBatch":
        return RoutedBatch(self.routed.to(device, non_blocking), self.labels.to(device, non_blocking=non_blocking))

    def pin(self) -> "RoutedBatch":
        return RoutedBatch(self.routed.pin(), self.labels.pin_memory())

    def tensors(self):
        return self.routed.tensors() + (self.labels,)

    def nbytes(self) -> int:
        return self.routed.nbytes() + self.labels.numel() * 4


def route_batches(table: DeviceTable, worker_batches: Sequence[MLRBatch], rank: int,
                  owner_ranks: Optional[Sequence[int]] = None) -> RoutedBatch:
    """Build rank's owner-side slices from every worker's mini-batch (done once per input block)."""
    B = worker_batches[0].batch_size
    ptrs, cols, vals = [np.zeros(1, dtype=np.int64)], [], []
    for w, mb in enumerate(worker_batches):
        assert mb.batch_size == B, "ship mode needs equally sized mini-batches"
        ptr, col, val = mb.csr_ptr.numpy().astype(np.int64), mb.csr_col.numpy(), mb.csr_val.numpy()
        mine = owner_of_features(table, col, owner_ranks) == rank
        rows = np.repeat(np.arange(B), np.diff(ptr))[mine]
        counts = np.bincount(rows, minlength=B)
        ptrs.append(ptrs[-1][-1] + np.cumsum(counts))
        cols.append(col[mine])
        vals.append(val[mine])
    ptr_all = np.concatenate(ptrs)
    routed = build_batch(ptr_all.astype(np.int32), np.concatenate(cols), np.concatenate(vals),
                         np.zeros(len(ptr_all) - 1, dtype=np.int32))
    return RoutedBatch(routed, worker_batches[rank].labels.clone())


class MLRShipModelV1:
    """(round-1 implementation, kept for A/B runs and for row strides > 1024 floats)
    One BSP step of MLR in ship mode: owner-side partial logits reduce-added into the worker's buffer over peer
    memory, softmax + error rows stored into every owner's buffer, owner-local row update
    (ops/csrc/mlr_ship_ops.cu).  Same arithmetic as MLRTrainer.java:374-398,475-505."""

    def __init__(self, table: DeviceTable, arena: PeerArena, group: PeerGroup, clock: DeviceClock,
                 num_classes: int, step_size: float, batch_size: int) -> None:
        from ..ops import native
        self._n = native()
        self.table, self.arena, self.group, self.clock = table, arena, group, clock
        self.rank, self.world = group.rank, group.world
        self.num_classes, self.step_size, self.B = num_classes, float(step_size), batch_size
        self.device = table.device
        stride = table.row_stride
        lp, lsi, loff = arena.alloc(batch_size * stride * 4)
        dp, dsi, doff = arena.alloc(self.world * batch_size * stride * 4)
        self.L = arena.tensor(lp, [batch_size, stride], torch.float32)
        self.D = arena.tensor(dp, [self.world * batch_size, stride], torch.float32)
        self.L.zero_()
        self.D.zero_()
        torch.cuda.synchronize(self.device)
        infos = group.all_gather_object({"rank": self.rank, "L": (lsi, loff), "D": (dsi, doff),
                                         "handles": {si: arena.segment_info(si) for si in {lsi, dsi}}})
        lptrs, dptrs = [0] * self.world, [0] * self.world
        for info in infos:
            r = info["rank"]
            for si, h in info["handles"].items():
                arena.map_peer(r, si, h)
            lptrs[r] = arena.resolve(r, *info["L"])
            dptrs[r] = arena.resolve(r, *info["D"])
        self.logit_ptrs = torch.tensor(lptrs, dtype=torch.int64, device=self.device)
        self.d_ptrs = torch.tensor(dptrs, dtype=torch.int64, device=self.device)
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.device)
        group.barrier()

    def step(self, rb: RoutedBatch) -> None:
        from ..ops import count_launch
        bp, rpb, stride, _dim, mode = self.table.view()
        r = rb.routed
        self._n.mlr_partial_forward(bp, rpb, stride, self.num_classes, mode, r.csr_ptr, r.csr_col, r.csr_val,
                                    self.B, self.logit_ptrs)
        count_launch()
        self.clock.barrier()                     # every owner's partial logits have landed
        self._n.mlr_softmax_bcast(self.L, rb.labels, self.num_classes, self.step_size, self.d_ptrs, self.rank,
                                  self.stats)
        count_launch()
        self.clock.barrier()                     # every worker's error rows have landed
        self._n.mlr_backward_push(bp, rpb, stride, self.num_classes, mode, r.uniq_feat, r.feat_ptr, r.samp,
                                  r.xval, self.D, False)     # red.add beats a load-add-store RMW here (521 us vs 818 us measured)
        count_launch()

    def step_timed(self, rb: RoutedBatch) -> dict:
        """`step()` with CUDA events around its five parts (bench.py --ship_phases; every rank must call it, the two
        barriers are collective).  Returns milliseconds per part; the caller synchronises before reading them."""
        bp, rpb, stride, _dim, mode = self.table.view()
        r = rb.routed
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        self._n.mlr_partial_forward(bp, rpb, stride, self.num_classes, mode, r.csr_ptr, r.csr_col, r.csr_val,
                                    self.B, self.logit_ptrs)
        ev[1].record()
        self.clock.barrier()
        ev[2].record()
        self._n.mlr_softmax_bcast(self.L, rb.labels, self.num_classes, self.step_size, self.d_ptrs, self.rank,
                                  self.stats)
        ev[3].record()
        self.clock.barrier()
        ev[4].record()
        self._n.mlr_backward_push(bp, rpb, stride, self.num_classes, mode, r.uniq_feat, r.feat_ptr, r.samp,
                                  r.xval, self.D, False)
        ev[5].record()
        torch.cuda.synchronize(self.device)
        names = ("partial_fwd_ms", "barrier1_ms", "softmax_bcast_ms", "barrier2_ms", "bwd_update_ms")
        return {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}

    def decay_step(self, rate: float) -> None:
        self.step_size *= rate


class MLRShipModel:
    """Ship mode, round 2 (ops/csrc/mlr_nvls_ops.cu): three launches per BSP step and no barrier.

    K1 (owner-side partial logits) pushes its rows into the workers' logits buffers with `red.add` over NVLink and
    flags each 8-instance tile; K2 (softmax) runs on a high-priority side stream NEXT TO K1, consumes tiles as all
    owners have flagged them and broadcasts the error rows with `multimem.st` on the multicast mapping of a
    `SymmetricRegion` (NVLS; unicast peer stores when there is no multicast); K3 (owner-side row update) acquires the
    "all error rows have landed" counter inside the kernel.  `reduce_scatter="nvls"` swaps the push for an in-switch
    `multimem.ld_reduce` by K2.  Same arithmetic as MLRTrainer.java:374-398,475-505.  `lam` != 0 uses the lazy L2
    scale (w = s * v, `_scales`): logits are multiplied by s, deltas divided by the next s, no dense pass."""

    TILE_ROWS = 8           # instances per flag: one K1 CTA owns a tile (one system fence per tile), K2 runs it as
                            # 8 warps x 1 instance -> short tail after the last K1 tiles

    def __init__(self, table: DeviceTable, arena: PeerArena, group: PeerGroup, clock: Optional[DeviceClock],
                 num_classes: int, step_size: float, batch_size: int, multicast: bool = True,
                 lam: float = 0.0, reduce_scatter: Optional[str] = None) -> None:
        """reduce_scatter: "push" (default) -- K1's epilogue red.adds the partial logits into the worker's buffer over
        NVLink (fire-and-forget, hidden behind the gathers; measured faster at 8 GPUs); "nvls" -- the owners keep their
        partials and the worker's K2 sums them in the switch with multimem.ld_reduce (pull-based: latency-bound).
        The all-gather of the error rows is multimem.st in both."""
        import os
        from ..ops import native
        from ..parallel.peer import SymmetricRegion
        self.rs = reduce_scatter or os.environ.get("HB_SHIP_RS", "push")
        assert self.rs in ("push", "nvls"), self.rs
        self._n = native()
        self.table, self.arena, self.group, self.clock = table, arena, group, clock
        self.rank, self.world = group.rank, group.world
        self.num_classes, self.step_size, self.B = num_classes, float(step_size), int(batch_size)
        self.lam = float(lam)
        self.device = table.device
        stride = table.row_stride
        if stride > 1024:
            raise ValueError("MLRShipModel needs row_stride <= 1024 floats; use MLRShipModelV1 for wider rows")
        self.stride = stride
        self._n.ship_preload(arena.device)       # lazy kernel loading must not happen while another kernel spins
        self.n_tiles = -(-self.B // self.TILE_ROWS)
        rows = self.world * self.B
        buf = rows * stride * 4
        lbuf = self.B * stride * 4
        self._off_d = 0
        self._off_p = buf                               # "nvls" only: this owner's partial logits of every worker
        self._off_l = buf + (buf if self.rs == "nvls" else 0)      # "push": the logits buffer the owners add into
        self._off_pflag = self._off_l + lbuf
        self._off_dcount = self._off_pflag + ((self.n_tiles * 4 + 255) // 256) * 256
        total = self._off_dcount + 256
        self.region = SymmetricRegion(arena, group, total, multicast=multicast)
        reg = self.region
        reg.tensor(0, [reg.nbytes // 4], torch.int32).zero_()
        torch.cuda.synchronize(self.device)
        group.barrier()
        self.D = reg.tensor(self._off_d, [rows, stride], torch.float32)
        self.L = reg.tensor(self._off_l, [self.B, stride], torch.float32)
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=self.device)  # noqa: E731
        self.src_ptrs = i64([p + self._off_p for p in reg.peer_ptrs])
        self.logit_ptrs = i64([p + self._off_l for p in reg.peer_ptrs])
        self.l_local = i64([reg.local_ptr + self._off_l])
        self.dst_ptrs = i64([p + self._off_d for p in reg.peer_ptrs])
        self.pflag_ptrs = i64([p + self._off_pflag for p in reg.peer_ptrs])
        self.dcount_ptrs = i64([p + self._off_dcount for p in reg.peer_ptrs])
        self.nvls = reg.mc_ptr != 0
        self.mc_p = reg.mc_ptr + self._off_p if (self.nvls and self.rs == "nvls") else 0
        self.mc_d = reg.mc_ptr + self._off_d if self.nvls else 0
        self.tile_done = torch.zeros(max(2, self.world * self.n_tiles), dtype=torch.int32, device=self.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.device)
        self.epoch = 0
        self.side = torch.cuda.Stream(device=self.device, priority=-1)
        self._ev_in = torch.cuda.Event()
        self._ev_done = torch.cuda.Event()
        self.scale = 1.0          # lazy L2 scale s of the model table (w = s * v); identical on every rank (BSP)
        torch.cuda.synchronize(self.device)
        group.barrier()

    # launches of one step (also used one by one by `step_timed`)
    def _k1(self, rb: RoutedBatch) -> None:
        bp, rpb, stride, _dim, mode = self.table.view()
        r = rb.routed
        if self.rs == "push":
            self._n.ship3_forward(bp, rpb, stride, self.num_classes, mode, r.csr_ptr, r.csr_col, r.csr_val, self.world,
                                  self.B, self.TILE_ROWS, self.logit_ptrs, self.tile_done.data_ptr(), self.pflag_ptrs,
                                  self.epoch)
            return
        self._n.ship_forward(bp, rpb, stride, self.num_classes, mode, r.csr_ptr, r.csr_col, r.csr_val, self.world,
                             self.B, self.TILE_ROWS, self.region.local_ptr + self._off_p, self.tile_done.data_ptr(),
                             self.pflag_ptrs, self.epoch)

    def _k2(self, rb: RoutedBatch, logit_scale: float, delta_step: float) -> None:
        if self.rs == "push":       # the reduced logits are local; re-arm (zero) them; broadcast D with multimem.st
            self._n.ship_softmax(0, self.l_local, 0, self.mc_d, self.dst_ptrs, self.rank * self.B, self.stride,
                                 rb.labels, self.num_classes, delta_step, logit_scale, self.TILE_ROWS,
                                 self.region.local_ptr + self._off_pflag, self.epoch * self.world, self.dcount_ptrs,
                                 self.stats, self.err, True)
            return
        self._n.ship_softmax(self.mc_p, self.src_ptrs, self.rank * self.B, self.mc_d, self.dst_ptrs,
                             self.rank * self.B, self.stride, rb.labels, self.num_classes, delta_step, logit_scale,
                             self.TILE_ROWS, self.region.local_ptr + self._off_pflag, self.epoch * self.world,
                             self.dcount_ptrs, self.stats, self.err)

    def _k3(self, rb: RoutedBatch) -> None:
        bp, rpb, stride, _dim, mode = self.table.view()
        r = rb.routed
        self._n.mlr_backward_push(bp, rpb, stride, self.num_classes, mode, r.uniq_feat, r.feat_ptr, r.samp, r.xval,
                                  self.D, False, self.region.local_ptr + self._off_dcount,
                                  self.epoch * self.world * self.n_tiles, self.err)

    def _scales(self) -> tuple:
        """(logit scale, delta step) of this step under the lazy L2 decay: W' = (1 - c) W + delta with
        c = step * lam * B_total  <=>  s' = s (1 - c), v' = v + delta / s'."""
        if self.lam == 0.0:
            return 1.0, self.step_size
        s_old = self.scale
        self.scale = s_old * (1.0 - self.step_size * self.lam * self.B * self.world)
        return s_old, self.step_size / self.scale

    def step(self, rb: RoutedBatch) -> None:
        from ..ops import count_launch
        self.epoch += 1
        ls, ds = self._scales()
        cur = torch.cuda.current_stream(self.device)
        self._ev_in.record(cur)
        self.side.wait_event(self._ev_in)                # this step's inputs are on the device
        with torch.cuda.stream(self.side):
            self._k2(rb, ls, ds)                         # resident first: it spins on tile flags while K1 runs
            self._ev_done.record(self.side)
        self._k1(rb)
        self._k3(rb)
        cur.wait_event(self._ev_done)                    # stats / the next step's inputs are ordered after K2
        count_launch(3)
        if self.lam != 0.0 and self.scale < 1e-3:
            self.renormalize()

    def renormalize(self) -> None:
        """Fold the lazy scale into the rows (v <- s v, s <- 1): one dense pass every few hundred steps."""
        self.table.scale_local(self.scale)
        self.scale = 1.0

    def step_profiled(self, rb: RoutedBatch) -> dict:
        """`step()` with its real schedule (K2 resident on the side stream next to K1) and CUDA events at the kernel
        boundaries: how long K1 runs with K2 next to it, how long K2 outlives K1, what K3 costs after its wait."""
        self.epoch += 1
        ls, ds = self._scales()
        cur = torch.cuda.current_stream(self.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record(cur)
        self.side.wait_event(ev[0])
        with torch.cuda.stream(self.side):
            self._k2(rb, ls, ds)
            ev[3].record(self.side)
        self._k1(rb)
        ev[1].record(cur)
        self._k3(rb)
        ev[2].record(cur)
        cur.wait_event(ev[3])
        ev[4].record(cur)
        torch.cuda.synchronize(self.device)
        return {"k1_with_k2_resident_ms": ev[0].elapsed_time(ev[1]), "k2_end_ms": ev[0].elapsed_time(ev[3]),
                "k3_incl_wait_ms": ev[1].elapsed_time(ev[2]), "step_ms": ev[0].elapsed_time(ev[4])}

    def step_timed(self, rb: RoutedBatch) -> dict:
        """One step with the three kernels serialised on the current stream and CUDA events around each
        (bench.py --ship_phases).  No overlap here by construction: it shows what each part costs alone."""
        self.epoch += 1
        ls, ds = self._scales()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        self._k1(rb)
        ev[1].record()
        self._k2(rb, ls, ds)
        ev[2].record()
        self._k3(rb)
        ev[3].record()
        torch.cuda.synchronize(self.device)
        names = ("k1_partial_fwd_ms", "k2_reduce_softmax_bcast_ms", "k3_update_ms")
        return {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}

    def check(self) -> None:
        e = int(self.err.item())
        if e != 0:
            raise TimeoutError(f"ship-mode flag wait timed out (stage {e}: 2 = logits tile, 3 = error rows)")

    def decay_step(self, rate: float) -> None:
        self.step_size *= rate

    def close(self) -> None:
        self.region.close()
