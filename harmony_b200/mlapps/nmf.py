"""Non-negative matrix factorisation (D ~ L R) app.

Reference: dolphin/mlapps/nmf/NMFTrainer.java:144-153,156-210,213-217,328-367,375-406,414-456,
NMFETModelUpdateFunction.java:43-52 (server: R_j = clamp(R_j - step g, 0, max_val); the SERVER step
size is the initial -step_size, Appendix E.3), NMFModelGenerator.java:59-101 (init U(init_min,
init_max)), NMFETDataParser.java:82-106 (`row: col,val col,val ...`, 1-based, vals >= 0),
NMFParameters.java:26-45, NMFJob.java:37-54.

Global model R: key = column -> float[rank] (PS table).  Local model L: key = input key -> float[rank]
(worker-local table updated with update_no_reply).  `-features <num_cols>` makes R a dense slab
table (required by the device backend); without it R is a generic KV table on the host.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np

from ..common.dataloader import DataParser, register_parser
from ..dolphin import params as P
from ..dolphin.conf import ETDolphinConfiguration
from ..dolphin.worker import Trainer
from ..et.config import TableConfiguration
from ..utils.flags import Configuration

LOG = logging.getLogger(__name__)


class NMFDataParser(DataParser):
    """`rowIdx: col,val col,val ...` (1-based) -> (row, cols int32 0-based, vals float32)."""
    index_bound_param = "features"

    def record_indices(self, record):
        return record[1]

    native_mode = 1

    def from_arrays(self, head, ptr, idx, val):
        rows = head.astype(np.int64) - 1
        return [(int(rows[i]), idx[ptr[i]:ptr[i + 1]], val[ptr[i]:ptr[i + 1]]) for i in range(len(rows))]

    def parse_line(self, line: str):
        head, _, rest = line.partition(":")
        row = int(head.strip())
        if row <= 0:
            raise ValueError("row index must be positive")
        cols, vals = [], []
        for tok in rest.split():
            c, _, v = tok.partition(",")
            c, v = int(c), float(v)
            if c <= 0:
                raise ValueError("column index must be positive")
            if v < 0:
                raise ValueError("matrix values must be non-negative")
            cols.append(c - 1)
            vals.append(v)
        return row - 1, np.asarray(cols, dtype=np.int32), np.asarray(vals, dtype=np.float32)


register_parser("nmf", NMFDataParser)


def _update_spec(conf: Configuration, seed: int) -> Dict[str, Any]:
    return {"name": "sgd_clamp", "params": {"step": float(conf["step_size"]), "lo": 0.0,
                                             "hi": float(conf.get("max_val", 1e6)), "init": "uniform",
                                             "init_lo": float(conf.get("init_min", 0.0)),
                                             "init_hi": float(conf.get("init_max", 1.0)), "seed": seed,
                                             "row_dim": int(conf["rank"]), "dtype": "float32"}}


def batch_arrays(batch: Sequence[Tuple[Any, Any]]):
    keys = [k for k, _ in batch]
    ptr = np.zeros(len(batch) + 1, dtype=np.int64)
    for i, (_, (_r, c, _v)) in enumerate(batch):
        ptr[i + 1] = ptr[i] + len(c)
    cols = np.concatenate([b[1][1] for b in batch]).astype(np.int64) if batch else np.zeros(0, dtype=np.int64)
    vals = np.concatenate([b[1][2] for b in batch]).astype(np.float32) if batch else np.zeros(0, dtype=np.float32)
    return keys, ptr, cols, vals


def nmf_gradients(L: np.ndarray, R: np.ndarray, ptr, slot, vals, lam: float):
    """Vectorised K7: returns (gL [n, rank], gR [U, rank], loss)."""
    n = len(ptr) - 1
    rows = np.repeat(np.arange(n), np.diff(ptr))
    e = (L[rows] * R[slot]).sum(axis=1) - vals
    gL = 2.0 * lam * L
    np.add.at(gL, rows, 2.0 * e[:, None] * R[slot])
    gR = 2.0 * lam * R
    np.add.at(gR, slot, 2.0 * e[:, None] * L[rows])
    return gL.astype(np.float32), gR.astype(np.float32), float((e * e).sum())


class NMFTrainer(Trainer):
    """Host NMF trainer: pull the batch's columns of R, gradients for L (local model table) and R, push.
    Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/mlapps/nmf/NMFTrainer.java:144-153,156-210,328-367,375-406."""

    def __init__(self, params: Dict[str, Any], model_accessor, local_model_table=None, context=None,
                 input_table=None, **_: Any) -> None:
        self.acc = model_accessor
        self.local = local_model_table
        self._params, self._context, self._input = params, context, input_table
        self.rank = int(params["rank"])
        self.step = float(params["step_size"])
        self.lam = float(params.get("lambda", 0.0))
        self.decay_rate = float(params.get("decay_rate", 0.9))
        self.decay_period = int(params.get("decay_period", 5))
        self.print_mat = bool(params.get("print_mat", False))
        self.last_loss = 0.0

    def set_mini_batch_data(self, mini_batch) -> None:
        self._keys, self._ptr, self._cols, self._vals = batch_arrays(mini_batch)
        self._ucols, self._slot = np.unique(self._cols, return_inverse=True)

    def pull_model(self) -> None:
        self._R = np.asarray(self.acc.pull_many(self._ucols), dtype=np.float32)

    def local_compute(self) -> None:
        L = np.asarray(self.local.multi_get_or_init(self._keys).get(timeout=600), dtype=np.float32)
        gL, gR, self.last_loss = nmf_gradients(L, self._R, self._ptr, self._slot, self._vals, self.lam)
        self.local.multi_update_no_reply(self._keys, list(gL))      # L_i updated by the local table's fn
        self._gR = gR

    def push_update(self) -> None:
        self.acc.push_many(self._ucols, self._gR)
        self._R = self._gR = None

    def on_epoch_finished(self, epoch_idx: int) -> None:
        if self.decay_rate != 1 and (epoch_idx + 1) % self.decay_period == 0:
            self.step *= self.decay_rate     # (unused by the server-side update, like the reference)

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        data = list(input_data)
        if not data:
            return {"loss": 0.0}
        keys, ptr, cols, vals = batch_arrays(data)
        ucols, slot = np.unique(cols, return_inverse=True)
        R = np.asarray(self.acc.pull_many(ucols), dtype=np.float32)
        L = np.asarray(self.local.multi_get_or_init(keys).get(timeout=600), dtype=np.float32)
        rows = np.repeat(np.arange(len(keys)), np.diff(ptr))
        e = (L[rows] * R[slot]).sum(axis=1) - vals
        return {"loss": float((e * e).sum())}

    def summary(self) -> Dict[str, float]:
        return {"last_batch_loss": self.last_loss}

    def cleanup(self) -> None:
        """-print_mat true: store this worker's rows of L and the columns of R it used (NMFTrainer.java:263-300 prints
        them; here they go through the DataStorer under -base_dir)."""
        if not self.print_mat or self.local is None:
            return
        row_of = {k: rec[0] for k, rec in self._input.get_local_tablet().get_data_iterator()} if self._input else {}
        items = sorted(((row_of.get(k, k), v) for k, v in self.local.get_local_tablet().get_data_iterator()),
                       key=lambda kv: kv[0])
        store_matrices(self._params, self._context, [k for k, _ in items],
                       np.asarray([v for _, v in items], dtype=np.float32))


def store_matrices(params, context, row_keys, L, col_keys=None, R=None) -> str:
    from ..common.datastorer import LocalFSDataStorer
    ex = getattr(getattr(context, "executor", None), "executor_id", "worker")
    ds = LocalFSDataStorer(str(params.get("base_dir", "/tmp/harmony_b200")))
    sub = os.path.join("nmf", str(params.get("job_id", "job")))
    text = "".join(f"{int(k) + 1}: " + " ".join(f"{x:.6g}" for x in row) + "\n" for k, row in zip(row_keys, L))
    path = ds.store_data(os.path.join(sub, f"L-{ex}.txt"), text.encode())
    if R is not None:
        text = "".join(f"{int(k) + 1}: " + " ".join(f"{x:.6g}" for x in row) + "\n" for k, row in zip(col_keys, R))
        ds.store_data(os.path.join(sub, f"R-{ex}.txt"), text.encode())
    return path


def model_table_conf(conf: Configuration, table_id: str) -> TableConfiguration:
    cols = int(conf.get("features", 0) or 0)
    backend = conf.get("backend", "host")
    if cols > 0:
        return TableConfiguration(table_id=table_id, dense=True, capacity=cols, row_dim=int(conf["rank"]),
                                  update_function=_update_spec(conf, seed=11),
                                  backend=backend if backend in ("host", "device") else "host")
    return TableConfiguration(table_id=table_id, update_function=_update_spec(conf, seed=11))


def local_model_table_conf(conf: Configuration, table_id: str) -> TableConfiguration:
    return TableConfiguration(table_id=table_id, update_function=_update_spec(conf, seed=23))


APP = ETDolphinConfiguration(
    app_id="NMF", trainer_class="harmony_b200.mlapps.nmf:NMFTrainer", input_parser="nmf",
    model_table_conf=model_table_conf, local_model_table_conf=local_model_table_conf, parameters=P.NMF,
    device_trainer_class="harmony_b200.mlapps.nmf_device:NMFDeviceTrainer")
APP_PATH = "harmony_b200.mlapps.nmf:APP"


def main_submit(argv: List[str]) -> int:
    from ..dolphin.launcher import main_submit as go
    return go(APP_PATH, APP, argv)


def main_run(argv: List[str]) -> int:
    from ..dolphin.launcher import main_run as go
    return go(APP, argv)
