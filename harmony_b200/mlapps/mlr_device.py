"""GPU-resident MLR trainer: the Dolphin Trainer SPI mapped onto the fused sm_100a kernels.

pull_model  -> mlr_fwd_kernel      (PULL of the batch's feature rows from the owners' HBM fused with
                                    the sparse dot, softmax, loss and the pre-scaled error)
push_update -> mlr_bwd_push_kernel (gradient rows + PUSH + `w += delta` fused at the owners)
Input blocks are converted once into device-resident CSR/CSC tensors (the input table lives in HBM).
"""
from __future__ import annotations

import logging
from typing import Any, Dict

import numpy as np

from ..dolphin.worker import Trainer
from .mlr import batch_to_csr

LOG = logging.getLogger(__name__)


class MLRDeviceTrainer(Trainer):
    """Dolphin Trainer over the fused MLR kernels: pull mode (rows move), ship mode (-mlr_mode ship, activations
    move, BSP) or dense input (tcgen05).  Reference arithmetic: MLRTrainer.java:374-398,475-505."""

    fused_step = True      # pull / compute / push are back-to-back kernels on one stream
    wants_sorted_batches = False
    REROUTE_CHECK_EPOCHS = 16      # ship mode: compare ownership views (rendezvous through the master) every 16th epoch

    def __init__(self, params: Dict[str, Any], model_table, input_table=None, **_: Any) -> None:
        import torch
        from ..models.mlr import MLRDeviceModel
        self._torch = torch
        self.params = params
        self.table = model_table
        self.dtable = model_table.device_table
        self.comps = model_table.components
        self.num_classes = int(params["classes"])
        self.model = MLRDeviceModel(self.dtable, self.num_classes, float(params["step_size"]),
                                    float(params.get("lambda", 0.0)),
                                    lazy_l2=False)   # job-server workers may run SSP-asynchronously: dense decay
        self.decay_rate = float(params.get("decay_rate", 0.9))
        self.decay_period = int(params.get("decay_period", 5))
        self._cache: Dict[Any, Any] = {}
        self._batch = None
        self.num_workers = int(params.get("num_job_workers", 1))
        self.num_features = int(params["features"])
        self.ship = None
        if params.get("mlr_mode") == "ship":
            self._init_ship(params, input_table, _.get("context"))
        self.dense = None
        if params.get("dense_input") and self.dtable.mode == 0 and self.dtable.rows_per_block % 32 == 0 \
                and self.num_features % 32 == 0:
            from ..models.mlr_dense import DenseMLR
            self.dense = DenseMLR(self.dtable, self.num_classes)      # tcgen05 / TMEM / TMA path
            self.dense.stats = self.model.stats

    # ---- "ship the activations" mode inside the job server (see models/mlr_ship.py) ----
    def _init_ship(self, params, input_table, context) -> None:
        from ..models.mlr import build_batch
        from ..models.mlr_ship import MLRShipModel, route_batches
        from ..parallel.device_clock import DeviceClock
        group = getattr(context, "job_group", None)
        tablet = input_table.get_local_tablet()
        blocks = sorted(tablet.get_block_ids())
        host, first_key = [], {}
        for s_idx, b in enumerate(blocks):
            data = sorted(tablet.get_data_in_block(b), key=lambda kv: kv[0])
            if not data:
                continue
            first_key[min(k for k, _ in data)] = len(host)
            ptr, col, val, labels = batch_to_csr(data)
            host.append((ptr, col, val, labels))
        if group is None:
            return
        sizes = group.all_gather_object((len(host), max((len(h[3]) for h in host), default=0)), tag="ship-sizes")
        n_steps, B = sizes[0][0], max(sz[1] for sz in sizes)
        if any(sz[0] != n_steps for sz in sizes) or n_steps == 0:
            return                                        # uneven blocks: stay in pull mode (same on every worker)
        padded = []
        for ptr, col, val, labels in host:                # pad every mini-batch to B rows (label -1 = no instance)
            n = len(labels)
            ptr = np.concatenate([ptr, np.full(B - n, ptr[-1], dtype=ptr.dtype)])
            labels = np.concatenate([labels, np.full(B - n, -1, dtype=labels.dtype)])
            padded.append(build_batch(ptr.astype(np.int32), col, val, labels))
        everyone = group.all_gather_object([[t.numpy() for t in mb.tensors()[:4]] for mb in padded], tag="ship-data")
        ex_ids = list(params.get("job_executor_ids") or [])
        owner_ranks = [ex_ids.index(o) if o in ex_ids else -1 for o in self.dtable.owners]
        if any(r < 0 for r in owner_ranks):
            return
        ctx = context.executor.device_ctx
        clock = DeviceClock(ctx.arena, group)
        self.ship = MLRShipModel(self.dtable, ctx.arena, group, clock, self.num_classes, self.model.step_size, B)
        self.ship.stats = self.model.stats
        self._ship_group, self._ship_everyone, self._ship_ex_ids = group, everyone, ex_ids
        self._ship_route(owner_ranks)
        self._ship_first_key = first_key
        self.wants_sorted_batches = True

    def _ship_route(self, owner_ranks) -> None:
        """(Re)build this owner's slices of every worker's mini-batches for the given block -> owner map."""
        from ..models.mlr import build_batch
        from ..models.mlr_ship import route_batches
        group = self._ship_group
        self._ship_owner_ranks = list(self.dtable.owners)
        self._ship_steps = []
        for s_idx in range(len(self._ship_everyone[0])):
            allw = [build_batch(*self._ship_everyone[w][s_idx]) for w in range(group.world)]
            self._ship_steps.append(route_batches(self.dtable, allw, group.rank, owner_ranks).to(self.dtable.device))

    def _ship_reroute_if_moved(self) -> None:
        """After a live re-partition the kernels stay CORRECT with the old routing -- every row access resolves
        through the pointer table, a moved row is simply reached over NVLink -- but the dot should run where the rows
        now live.  At an epoch boundary (all workers are in lock-step there) the workers compare their ownership
        views and, once they agree on a new map, re-route their input slices."""
        try:
            views = self._ship_group.all_gather_object(list(self.dtable.owners), tag="ship-owners", timeout=30.0)
        except TimeoutError:
            return                                    # a peer left the lock-step (job is ending): keep the routing
        mine = list(self.dtable.owners)
        if any(v != mine for v in views) or mine == self._ship_owner_ranks:
            return                                    # a flip is still travelling, or nothing moved
        owner_ranks = [self._ship_ex_ids.index(o) if o in self._ship_ex_ids else -1 for o in mine]
        if any(r < 0 for r in owner_ranks):
            return                                    # a block left the job's executors: keep resolving over NVLink
        self._ship_route(owner_ranks)

    def _to_dense_batch(self, mini_batch):
        key = ("dense", mini_batch[0][0], len(mini_batch))
        b = self._cache.get(key)
        if b is None:
            X = np.zeros((len(mini_batch), self.num_features), dtype=np.float32)
            y = np.zeros(len(mini_batch), dtype=np.int32)
            for i, (_, (idx, val, label)) in enumerate(mini_batch):
                X[i, idx] = val
                y[i] = label
            b = (self._torch.from_numpy(X).to(self.dtable.device), self._torch.from_numpy(y).to(self.dtable.device))
            self._cache[key] = b
        return b

    def _to_device_batch(self, mini_batch):
        from ..models.mlr import build_batch
        if self.dense is not None:
            return self._to_dense_batch(mini_batch)
        key = (mini_batch[0][0], len(mini_batch)) if mini_batch else None
        b = self._cache.get(key)
        if b is None:
            ptr, col, val, labels = batch_to_csr(mini_batch)
            b = build_batch(ptr, col, val, labels).to(self.dtable.device)
            self._cache[key] = b
        return b

    def set_mini_batch_data(self, mini_batch) -> None:
        if self.ship is not None:
            mk = getattr(mini_batch, "min_key", None)           # O(1) on a block snapshot (cached)
            self._batch = self._ship_steps[self._ship_first_key[mk() if mk else min(k for k, _ in mini_batch)]]
            return
        self._batch = self._to_device_batch(mini_batch)

    def pull_model(self) -> None:
        self.comps.gate.acquire_read()          # (uncontended: live migration does not stop accessors any more)
        self.model.stats.zero_()
        if self.dense is None and self.ship is None:
            self.model.forward(self._batch, True)

    def local_compute(self) -> None:
        pass                                    # fused into the pull / push kernels

    def push_update(self) -> None:
        try:
            if self.ship is not None:
                self.ship.step_size = self.model.step_size
                self.ship.step(self._batch)               # BSP lockstep over the job's workers (device clocks)
            elif self.dense is not None:
                X, y = self._batch
                self.dense.step(X, y, self.model.step_size, self.model.stats)   # GEMM -> softmax -> GEMM+push
            else:
                self.model.backward_push(self._batch)
            bsz = self.ship.B if self.ship is not None else (
                self._batch[0].shape[0] if self.dense is not None else self._batch.batch_size)
            self.model.apply_l2(bsz * self.num_workers)
            # no stream sync here: the next mini-batches of the lease queue behind this one on the stream; the
            # worker loop calls flush() before it reports progress (WorkerTasklet.java:122-149 blocks 4x per batch)
        finally:
            self.comps.gate.release_read()

    def flush(self) -> None:
        """Wait for the kernels of the mini-batches issued so far (end of a lease / epoch)."""
        self._torch.cuda.current_stream(self.dtable.device).synchronize()
        if self.ship is not None:
            self.ship.check()

    def summary(self) -> Dict[str, float]:
        return {"last_batch_loss": float(self.model.stats[0])}

    def on_epoch_finished(self, epoch_idx: int) -> None:
        if self.decay_rate != 1 and (epoch_idx + 1) % self.decay_period == 0:
            self.model.decay_step(self.decay_rate)
        if self.ship is not None and (epoch_idx + 1) % self.REROUTE_CHECK_EPOCHS == 0:
            self._ship_reroute_if_moved()       # a control-plane rendezvous: every worker, at the same epochs

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        out = {}
        for name, data in (("", list(input_data)), ("test", [(None, t) for t in test_data])):
            if not data:
                continue
            if self.dense is not None:
                continue
            b = self._to_device_batch(data)
            loss, acc = self.model.evaluate(b)
            out["loss" if not name else "testLoss"] = loss
            out["avgAccuracy" if not name else "testAvgAccuracy"] = acc
        return out
