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


class MLRShipModel:
    """One BSP step of MLR in ship mode: owner-side partial logits reduce-added into the worker's buffer over peer
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
