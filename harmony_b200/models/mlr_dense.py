"""Dense-input MLR on the tensor cores (tcgen05 / TMEM / TMA), see ops/csrc/mlr_dense_tc.cu.

`DenseMLR.logits(X)` = X[B, F] . W[F, C] where W is the (possibly multi-GPU) model table: the B
operand tiles are TMA-loaded from the owners' slabs through one tensor map per block, so the pull
is fused into the GEMM.  The table must be ORDERED (block b = rows [b*R, (b+1)*R)) with R % 32 == 0.
"""
from __future__ import annotations

import torch

from ..et.device_table import DeviceTable
from ..et.partitioner import PART_ORDERED


class DenseMLR:
    """Dense-input MLR on the tensor cores: logits = X . W and W += X^T D as tcgen05 kind::tf32 GEMMs whose B tiles
    are TMA-loaded from the owners' slabs (pull fused) and whose epilogue reduce-adds into the owners' rows
    (push fused); ops/csrc/mlr_dense_tc.cu."""

    def __init__(self, table: DeviceTable, num_classes: int, use_2cta: bool = True, stream_k: bool = False) -> None:
        # use_2cta: cta_group::2 kernels (256 x 256 tile per CTA pair, ~640 TFLOP/s tf32 measured vs ~340 for
        # the single-CTA 128 x 128 kernel, which is L2-bandwidth bound at 32 FLOP per staged byte)
        # stream_k: the persistent variant -- one cluster per TPC, whole waves strided + a stream-K tail, double-buffered
        # TMEM so the epilogue of one work unit overlaps the MMAs of the next.  Measured slower than the tile-parallel
        # kernel at 2 of 3 shapes on B200 (516 / 645 / 513 vs 512 / 610 / 582 TFLOP/s): off by default
        self.use_2cta = use_2cta
        self.stream_k = stream_k and use_2cta
        from ..ops import native
        self._n = native()
        assert table.mode == PART_ORDERED, "dense MLR needs an ordered (range partitioned) model table"
        assert table.rows_per_block % 32 == 0, "rows_per_block must be a multiple of 32"
        self.table, self.num_classes = table, num_classes
        self.device = table.device
        self._bmaps = None
        self._ptr_snapshot = None

    def _maps(self) -> torch.Tensor:
        ptrs = self.table.block_ptrs.cpu()
        if self._bmaps is None or not torch.equal(ptrs, self._ptr_snapshot):   # re-encode after a migration
            self._bmaps = self._n.make_block_tensormaps(ptrs, self.table.rows_per_block, self.table.row_stride,
                                                        self.device.index)
            self._ptr_snapshot = ptrs
        return self._bmaps

    def logits(self, X: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        from ..ops import count_launch
        B = X.shape[0]
        if out is None:
            out = torch.empty((B, self.table.row_stride), dtype=torch.float32, device=self.device)
        self._n.dense_logits_tf32(X, self._maps(), self.table.rows_per_block, self.table.row_stride, out,
                                  use_2cta=self.use_2cta, stream_k=self.stream_k)
        count_launch()
        return out

    # ---- full training step: tensor-core forward -> softmax -> tensor-core backward fused with the push ----
    def step(self, X: torch.Tensor, labels: torch.Tensor, step_size: float, stats: torch.Tensor = None) -> None:
        """X [B, F] fp32 dense instances, labels int32 [B].  W += -step * X^T (softmax(XW) - Y)."""
        from ..ops import count_launch
        B, F = X.shape
        Bp = (B + 31) // 32 * 32
        if Bp != B:                                  # the batch is the K dimension of the backward GEMM
            Xp = torch.zeros((Bp, F), dtype=torch.float32, device=self.device)
            Xp[:B] = X
        else:
            Xp = X.contiguous()
        stride = self.table.row_stride
        if getattr(self, "_L", None) is None or self._L.shape[0] < Bp:
            self._L = torch.empty((Bp, stride), dtype=torch.float32, device=self.device)
            self._D = torch.zeros((Bp, stride), dtype=torch.float32, device=self.device)
            self._dptr = torch.tensor([self._D.data_ptr()], dtype=torch.int64, device=self.device)
        L, D = self._L[:Bp], self._D[:Bp]
        if Bp != B:
            D[B:].zero_()
        self.logits(Xp, L)
        if stats is None:
            stats = getattr(self, "stats", None)
            if stats is None:
                stats = self.stats = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._n.mlr_softmax_bcast(L[:B], labels, self.num_classes, float(step_size), self._dptr, 0, stats)
        count_launch()
        bp, rpb, st, _dim, mode = self.table.view()
        self._n.dense_grad_push_tf32(Xp, D, bp, rpb, st, self.num_classes, mode, use_2cta=self.use_2cta, stream_k=self.stream_k)
        count_launch()
