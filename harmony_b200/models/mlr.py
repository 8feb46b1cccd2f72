"""Multinomial logistic regression on a sharded, feature-major model table (device path).

One mini-batch = PULL (rows of the features present in the batch, read from the owners' HBM)
+ COMP (sparse dot, softmax, error) + PUSH (one reduce-add row per distinct feature, applied in
the owners' HBM).  Reference semantics: dolphin/mlapps/mlr/MLRTrainer.java:185-187 (pull),
:190-300 + :374-398 (compute: delta = -step * ((p - y) x + lambda w) summed over the batch),
:303-307 (push), MLRETModelUpdateFunction.java:51-62 (w += delta), :309-316 (step decay).

The L2 term is applied at the owner (`scale_local`): every worker's mini-batch contributes
`-step * lambda * B * w`, i.e. the owner multiplies its shard by (1 - step*lambda*B_total).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from ..et.device_table import DeviceTable


@dataclass
class MLRBatch:
    """A mini-batch in CSR (forward) + grouped-by-feature (backward/push) form."""
    csr_ptr: torch.Tensor     # int32 [B+1]
    csr_col: torch.Tensor     # int32 [nnz]
    csr_val: torch.Tensor     # float32 [nnz]
    labels: torch.Tensor      # int32 [B]
    uniq_feat: torch.Tensor   # int32 [U]
    feat_ptr: torch.Tensor    # int32 [U+1]
    samp: torch.Tensor        # int32 [nnz]
    xval: torch.Tensor        # float32 [nnz]

    @property
    def batch_size(self) -> int:
        return int(self.labels.numel())

    @property
    def nnz(self) -> int:
        return int(self.csr_col.numel())

    def tensors(self):
        return (self.csr_ptr, self.csr_col, self.csr_val, self.labels, self.uniq_feat, self.feat_ptr,
                self.samp, self.xval)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())

    def to(self, device, non_blocking: bool = True) -> "MLRBatch":
        return MLRBatch(*[t.to(device, non_blocking=non_blocking) for t in self.tensors()])

    def pin(self) -> "MLRBatch":
        return MLRBatch(*[t.pin_memory() for t in self.tensors()])


def build_batch(csr_ptr: np.ndarray, csr_col: np.ndarray, csr_val: np.ndarray, labels: np.ndarray) -> MLRBatch:
    """Host-side preprocessing done once per input block at load time (a block IS a mini-batch:
    ETTrainingDataProvider.java:71-80): group the non-zeros by feature for the push phase."""
    csr_ptr = np.asarray(csr_ptr, dtype=np.int32)
    csr_col = np.asarray(csr_col, dtype=np.int32)
    csr_val = np.asarray(csr_val, dtype=np.float32)
    labels = np.asarray(labels, dtype=np.int32)
    t = torch.from_numpy
    from ..ops import native_or_none
    nat = native_or_none()
    if nat is not None:        # counting sort in the native runtime (ops/csrc/loader.cpp); NumPy twin below
        uniq, feat_ptr, samp, xval = nat.group_by_feature(t(np.ascontiguousarray(csr_ptr)), t(np.ascontiguousarray(csr_col)),
                                                          t(np.ascontiguousarray(csr_val)))
        return MLRBatch(t(csr_ptr), t(csr_col), t(csr_val), t(labels), uniq, feat_ptr, samp, xval)
    B = len(labels)
    rows = np.repeat(np.arange(B, dtype=np.int32), np.diff(csr_ptr))
    order = np.argsort(csr_col, kind="stable")
    sc = csr_col[order]
    if len(sc):
        starts = np.flatnonzero(np.r_[True, sc[1:] != sc[:-1]]).astype(np.int32)
        uniq = sc[starts]
    else:
        starts = np.zeros(0, dtype=np.int32)
        uniq = np.zeros(0, dtype=np.int32)
    feat_ptr = np.r_[starts, np.int32(len(sc))].astype(np.int32)
    t = torch.from_numpy
    return MLRBatch(t(csr_ptr), t(csr_col), t(csr_val), t(labels), t(np.ascontiguousarray(uniq)),
                    t(feat_ptr), t(np.ascontiguousarray(rows[order])), t(np.ascontiguousarray(csr_val[order])))


def synthetic_batch(batch_size: int, num_features: int, num_classes: int, nnz_per_row: int, seed: int,
                    zipf: float = 0.0) -> MLRBatch:
    """Synthetic sparse instances: `nnz_per_row` distinct-ish features per instance, uniform over
    the feature space (or Zipf-skewed when `zipf` > 0), values U(0,1), labels from a hidden model."""
    rng = np.random.default_rng(seed)
    nnz = batch_size * nnz_per_row
    if zipf > 0:
        cols = (rng.zipf(1.0 + zipf, size=nnz) - 1) % num_features
    else:
        cols = rng.integers(0, num_features, size=nnz)
    cols = cols.astype(np.int32).reshape(batch_size, nnz_per_row)
    cols.sort(axis=1)
    vals = rng.random(nnz, dtype=np.float32)
    labels = (cols[:, 0].astype(np.int64) * 2654435761 % num_classes).astype(np.int32)
    ptr = np.arange(0, nnz + 1, nnz_per_row, dtype=np.int32)
    return build_batch(ptr, cols.reshape(-1), vals, labels)


class MLRDeviceModel:
    """Step functions over a DeviceTable whose row f = class weights of feature f."""

    def __init__(self, table: DeviceTable, num_classes: int, step_size: float, lam: float = 0.0,
                 max_batch: int = 8192, lazy_l2: bool = True) -> None:
        """`lazy_l2`: keep the L2 decay as a scalar s with w = s * v (the table stores v): the forward multiplies
        the logits by s, pushes divide the deltas by the NEXT s, nothing touches the 3.8 GiB shard.  Exact for
        workers that step in lock-step (BSP: every holder of the table advances s identically); asynchronous
        workers (SSP slack > 0 through the job server) pass lazy_l2=False and get the dense owner-side decay."""
        from ..ops import native
        self._n = native()
        self.table = table
        self.num_classes = num_classes
        self.step_size = float(step_size)
        self.lam = float(lam)
        self.lazy_l2 = bool(lazy_l2)
        self.scale = 1.0           # s: w = s * v
        self._delta_step = None    # step / s' of the mini-batch in flight (lazy mode)
        self.device = table.device
        self._D = torch.empty((max_batch, table.row_stride), dtype=torch.float32, device=self.device)
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.device)

    def _scratch(self, B: int) -> torch.Tensor:
        if B > self._D.shape[0]:
            self._D = torch.empty((B, self.table.row_stride), dtype=torch.float32, device=self.device)
        return self._D

    def forward(self, batch: MLRBatch, write_delta: bool = True) -> None:
        """PULL + dot + softmax + loss (+ pre-scaled error D). Accumulates into self.stats."""
        from ..ops import count_launch
        bp, rpb, stride, _dim, mode = self.table.view()
        D = self._scratch(batch.batch_size) if write_delta else None
        step = self.step_size if self._delta_step is None else self._delta_step
        self._n.mlr_forward(bp, rpb, stride, self.num_classes, mode, batch.csr_ptr, batch.csr_col,
                            batch.csr_val, batch.labels, step, D, self.stats, self.scale)
        count_launch()

    def backward_push(self, batch: MLRBatch) -> None:
        """grad rows + PUSH + `w += delta` at the owners."""
        from ..ops import count_launch
        bp, rpb, stride, _dim, mode = self.table.view()
        self._n.mlr_backward_push(bp, rpb, stride, self.num_classes, mode, batch.uniq_feat, batch.feat_ptr,
                                  batch.samp, batch.xval, self._scratch(batch.batch_size))
        count_launch()

    def apply_l2(self, total_batch_instances: int) -> None:
        """Dense owner-side decay (non-lazy mode): W <- (1 - step * lam * B_total) W on the local shard."""
        if self.lam != 0.0 and not self.lazy_l2:
            self.table.scale_local(1.0 - self.step_size * self.lam * total_batch_instances)

    def step(self, batch: MLRBatch, total_batch_instances: Optional[int] = None) -> None:
        """W' = (1 - c) W + delta, c = step * lam * B_total (MLRTrainer.java:387-397).  Lazy mode: s' = (1 - c) s and
        v' = v + delta / s' -- the forward sees s, the error rows are pre-scaled by step / s'."""
        n_inst = total_batch_instances or batch.batch_size
        if self.lam != 0.0 and self.lazy_l2:
            s_next = self.scale * (1.0 - self.step_size * self.lam * n_inst)
            self._delta_step = self.step_size / s_next
            self.forward(batch, True)
            self.backward_push(batch)
            self._delta_step = None
            self.scale = s_next
            if self.scale < 1e-3:
                self.fold_scale()
            return
        self.forward(batch, True)
        self.backward_push(batch)
        self.apply_l2(n_inst)

    def fold_scale(self) -> None:
        """v <- s v, s <- 1 on the local shard (before checkpoints / dense exports, or when s gets small).  Every
        holder of the table calls it at the same logical point (each folds its own blocks)."""
        if self.scale != 1.0:
            self.table.scale_local(self.scale)
            self.scale = 1.0

    def evaluate(self, batch: MLRBatch) -> Tuple[float, float]:
        """(sum loss, accuracy) of the batch under the current model (no update)."""
        self.stats.zero_()
        self.forward(batch, write_delta=False)
        s = self.stats.cpu()
        return float(s[0]), float(s[1]) / max(1, batch.batch_size)

    def decay_step(self, rate: float) -> None:
        self.step_size *= rate


def mlr_reference_step(W: torch.Tensor, batch: MLRBatch, num_classes: int, step: float, lam: float = 0.0
                       ) -> Tuple[torch.Tensor, float, float]:
    """Plain PyTorch fp32 reference of one mini-batch (used by tests). W: [F, C]."""
    B = batch.batch_size
    ptr = batch.csr_ptr.long().cpu()
    rows = torch.repeat_interleave(torch.arange(B), ptr[1:] - ptr[:-1]).to(W.device)
    cols = batch.csr_col.long().to(W.device)
    vals = batch.csr_val.to(W.device)
    X = torch.sparse_coo_tensor(torch.stack([rows, cols]), vals, (B, W.shape[0])).coalesce()
    logits = torch.sparse.mm(X, W)
    logp = torch.log_softmax(logits, dim=1)
    p = logp.exp().clamp(1e-7, 1 - 1e-7)
    y = batch.labels.long().to(W.device)
    loss = float(-(torch.log(p[torch.arange(B, device=W.device), y])).sum())
    acc = float((logits.argmax(1) == y).float().mean())
    err = p.clone()
    err[torch.arange(B, device=W.device), y] -= 1.0
    delta = -step * torch.sparse.mm(X.t(), err)
    Wn = W + delta
    if lam != 0.0:
        Wn = Wn - step * lam * B * W
    return Wn, loss, acc


def load_batches(path: str, batch_size: int, rank: int = 0, world: int = 1, num_classes: int = 0,
                 num_features: int = 0, drop_last: bool = False):
    """SPMD data loading for library users: rank `rank` of `world` reads its byte-range split of a `label idx:val ...`
    file (exact Hadoop-style line ownership, native multi-threaded parser when the extension is built) and returns
    its mini-batches as `MLRBatch`es (CSR + by-feature grouping) ready for `.pin()` / `.to("cuda")`.
    Labels / feature ids are validated against `num_classes` / `num_features` when given."""
    from ..common.dataloader import get_splits, make_parser, read_split_bytes
    params = {}
    if num_classes:
        params["classes"] = num_classes
    if num_features:
        params["features"] = num_features
    parser = make_parser({"name": "mlr", "params": params})
    records = parser.parse_bytes(read_split_bytes(get_splits(path, world)[rank]))
    out = []
    for lo in range(0, len(records), batch_size):
        chunk = records[lo:lo + batch_size]
        if drop_last and len(chunk) < batch_size:
            break
        ptr = np.zeros(len(chunk) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(r[0]) for r in chunk])
        col = np.concatenate([r[0] for r in chunk]) if chunk else np.zeros(0, dtype=np.int32)
        val = np.concatenate([r[1] for r in chunk]) if chunk else np.zeros(0, dtype=np.float32)
        out.append(build_batch(ptr, col, val, np.asarray([r[2] for r in chunk], dtype=np.int32)))
    return out
