"""NMF mini-batch on the device: pull (K1) -> fused SDDMM/SpMM + local-model update (K7+K4) ->
push fused with SGD+clamp at the owners (K2+K4).  See ops/csrc/nmf_ops.cu."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch

from ..et.device_table import DeviceTable


@dataclass
class NMFBatch:
    """One NMF mini-batch on the device: CSR of the rows, distinct columns, slot of every non-zero in the staging
    buffer."""

    csr_ptr: torch.Tensor    # int32 [n+1]
    col_slot: torch.Tensor   # int32 [nnz]  index into `ucols`
    val: torch.Tensor        # float32 [nnz]
    l_row: torch.Tensor      # int32 [n]    row of the worker-local L matrix
    ucols: torch.Tensor      # int64 [U]    distinct columns (= keys of the global model to pull/push)

    def tensors(self):
        return (self.csr_ptr, self.col_slot, self.val, self.l_row, self.ucols)

    def to(self, device, non_blocking=True) -> "NMFBatch":
        return NMFBatch(*[t.to(device, non_blocking=non_blocking) for t in self.tensors()])

    def pin(self) -> "NMFBatch":
        return NMFBatch(*[t.pin_memory() for t in self.tensors()])

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())


def build_nmf_batch(ptr: np.ndarray, cols: np.ndarray, vals: np.ndarray, l_rows: np.ndarray) -> NMFBatch:
    ucols, slot = np.unique(np.asarray(cols, dtype=np.int64), return_inverse=True)
    t = torch.from_numpy
    return NMFBatch(t(np.asarray(ptr, dtype=np.int32)), t(slot.astype(np.int32)), t(np.asarray(vals, dtype=np.float32)),
                    t(np.asarray(l_rows, dtype=np.int32)), t(ucols))


def synthetic_nmf_batch(n_rows: int, n_cols: int, nnz_per_row: int, row_base: int, seed: int) -> NMFBatch:
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, n_cols, size=(n_rows, nnz_per_row))
    cols.sort(axis=1)
    vals = rng.random(n_rows * nnz_per_row, dtype=np.float32)
    ptr = np.arange(0, n_rows * nnz_per_row + 1, nnz_per_row)
    return build_nmf_batch(ptr, cols.reshape(-1), vals, np.arange(row_base, row_base + n_rows))


class NMFDeviceModel:
    """Device NMF step: pull distinct columns of R into an L2-resident staging buffer, nmf_batch_kernel (SDDMM +
    both SpMMs + local L update), push with SGD + clamp at the owners.  Reference:
    dolphin/mlapps/nmf/NMFTrainer.java:156-210,328-367."""

    def __init__(self, table: DeviceTable, rank: int, n_local_rows: int, step: float, lam: float = 0.0,
                 max_val: float = 1e6, init_lo: float = 0.0, init_hi: float = 1.0, seed: int = 0) -> None:
        from ..ops import native
        self._n = native()
        self.table, self.rank = table, rank
        self.step, self.lam, self.max_val = float(step), float(lam), float(max_val)
        self.device = table.device
        g = torch.Generator(device=self.device).manual_seed(seed)
        rp = (rank + 3) // 4 * 4                                   # rows padded for 128-bit accesses
        self._Lbuf = torch.zeros((n_local_rows, rp), dtype=torch.float32, device=self.device)
        self.L = self._Lbuf[:, :rank]
        self.L.copy_((init_lo + (init_hi - init_lo) * torch.rand((n_local_rows, rank), device=self.device,
                                                                 generator=g)).clamp_(0.0, max_val))
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._R = self._G = None

    def append_rows(self, rows: torch.Tensor) -> int:
        """Grow the HBM-resident local factor by `rows` ([n, rank]: the L rows of input records that migrated in with
        their blocks); returns the index of the first new row.  Capacity doubles, so appends are amortised O(1)."""
        n0, n = self.L.shape[0], rows.shape[0]
        used = getattr(self, "_used", n0)
        if used + n > self._Lbuf.shape[0]:
            grown = torch.zeros((max(2 * self._Lbuf.shape[0], used + n), self._Lbuf.shape[1]), dtype=torch.float32,
                                device=self.device)
            grown[:used] = self._Lbuf[:used]
            self._Lbuf = grown
        self._Lbuf[used:used + n, : self.rank] = rows.to(self.device, torch.float32)
        self._used = used + n
        self.L = self._Lbuf[: self._used, : self.rank]
        return used

    def _staging(self, U: int):
        if self._R is None or self._R.shape[0] < U:
            self._R = torch.empty((U, self.table.row_stride), dtype=torch.float32, device=self.device)
            self._G = torch.empty_like(self._R)
        return self._R[:U], self._G[:U]

    def step_batch(self, b: NMFBatch) -> None:
        from ..ops import count_launch
        R, G = self._staging(b.ucols.numel())
        self._n.table_pull(*self.table.view(), b.ucols, R)            # PULL the batch's columns once
        count_launch()
        G.zero_()
        self._n.nmf_batch(b.csr_ptr, b.col_slot, b.val, self._Lbuf, b.l_row, R, G, self.rank, self.step, self.lam,
                          0.0, self.max_val, self.loss)
        count_launch(2 if self.lam != 0.0 else 1)
        self.table.push(b.ucols, G)                                    # PUSH + clamp(w - step g) at the owners

    def evaluate(self, b: NMFBatch) -> float:
        R, _ = self._staging(b.ucols.numel())
        self._n.table_pull(*self.table.view(), b.ucols, R)
        rows = torch.repeat_interleave(b.l_row.long(), (b.csr_ptr[1:] - b.csr_ptr[:-1]).long())
        e = (self.L[rows] * R[b.col_slot.long(), : self.rank]).sum(1) - b.val
        return float((e * e).sum())


def nmf_reference_step(L: torch.Tensor, Rfull: torch.Tensor, b: NMFBatch, step: float, lam: float, max_val: float
                       ) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """Plain PyTorch fp32 reference of one mini-batch: returns (L_new, R_new, loss)."""
    rows_local = torch.repeat_interleave(torch.arange(b.l_row.numel(), device=L.device),
                                         (b.csr_ptr[1:] - b.csr_ptr[:-1]).long())
    lrow = b.l_row.long()[rows_local]
    cols = b.ucols[b.col_slot.long()]
    e = (L[lrow] * Rfull[cols]).sum(1) - b.val
    gL = torch.zeros_like(L)
    gL[b.l_row.long()] = 2 * lam * L[b.l_row.long()]
    gL.index_add_(0, lrow, 2 * e[:, None] * Rfull[cols])
    gR = torch.zeros_like(Rfull)
    gR[b.ucols] = 2 * lam * Rfull[b.ucols]
    gR.index_add_(0, cols, 2 * e[:, None] * L[lrow])
    Ln = L.clone()
    Ln[b.l_row.long()] = (L[b.l_row.long()] - step * gL[b.l_row.long()]).clamp(0, max_val)
    Rn = Rfull.clone()
    Rn[b.ucols] = (Rfull[b.ucols] - step * gR[b.ucols]).clamp(0, max_val)
    return Ln, Rn, float((e * e).sum())
