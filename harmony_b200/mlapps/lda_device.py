"""GPU-resident LDA trainer (Dolphin Trainer SPI over ops/csrc/lda_ops.cu + the table kernels)."""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

from ..dolphin.worker import Trainer


class LDADeviceTrainer(Trainer):
    """Dolphin Trainer over LDADeviceModel; initial random assignment pushed as counts (LDATrainer.java:113-194),
    likelihoods via the lgamma-sum kernel."""

    fused_step = True      # pull / compute / push are back-to-back kernels on one stream

    def __init__(self, params: Dict[str, Any], model_table, local_model_table=None, input_table=None,
                 **_: Any) -> None:
        import torch
        from ..models.lda import LDADeviceModel
        self._torch = torch
        self.table, self.comps = model_table, model_table.components
        self.dtable = model_table.device_table
        self.local_table, self.input = local_model_table, input_table
        # per-epoch write-back of the HBM-resident assignments is only needed when plans can move worker blocks
        self._elastic = "EmptyPlanOptimizer" not in str(params.get("optimizer", "EmptyPlanOptimizer"))
        self.K, self.V = int(params["num_topics"]), int(params["num_vocabs"])
        self.alpha, self.beta = float(params.get("alpha", 0.001)), float(params.get("beta", 0.01))
        # token layout of the worker-local assignment array: documents in key order
        self._docs = sorted(input_table.get_local_tablet().get_data_iterator(), key=lambda kv: kv[0])
        self._tok_base: Dict[Any, int] = {}
        base = 0
        for k, d in self._docs:
            self._tok_base[k] = base
            base += len(d)
        self.model = LDADeviceModel(self.dtable, self.K, self.V, base, self.alpha, self.beta,
                                    seed=hash(str(params.get("job_id"))) % 100000)
        self.model._used = base
        self._doc_len = {k: len(d) for k, d in self._docs}
        self._cache: Dict[Any, Any] = {}
        self._batch = None
        self._restored = False
        if local_model_table is not None and self._docs and \
                local_model_table.get_local_tablet().get_num_data_items() > 0:
            # started by a plan on blocks that migrated in: the assignments published by the previous owner are in the
            # local-model table -- adopt them instead of drawing (and pushing the counts of) a fresh random assignment
            self._load_assignments([k for k, _ in self._docs])
            self._restored = True

    # ---- worker-side elasticity: assignments travel in the local-model table with the input blocks ----
    def _load_assignments(self, keys) -> list:
        """Copy the (assignments, doc-topic counts) entries of `keys` from the local-model table into HBM.  Returns
        the keys that have no entry yet (new documents: they need an initial assignment)."""
        vals = self.local_table.multi_get(keys).get(timeout=600)
        missing = []
        for k, v in zip(keys, vals):
            if v is None:
                missing.append(k)
                continue
            zz = np.asarray(v[0], dtype=np.int32)
            lo = self._tok_base[k]
            self.model.z[lo: lo + len(zz)] = self._torch.from_numpy(zz).to(self.model.z.device)
        return missing

    def _adopt(self, mini_batch) -> None:
        """Documents without a token range here migrated in with their input block: give them one, fetch their
        assignments from the local-model table (or initialise them if the previous owner never published any)."""
        new = [(k, d) for k, d in mini_batch if k not in self._tok_base]
        if not new:
            return
        new.sort(key=lambda kv: kv[0])
        first = self.model.grow_tokens(sum(len(d) for _, d in new))
        for k, d in new:
            self._tok_base[k] = first
            self._doc_len[k] = len(d)
            first += len(d)
        self._docs.extend(new)
        missing = self._load_assignments([k for k, _ in new]) if self.local_table is not None else [k for k, _ in new]
        if missing:
            miss = set(missing)
            self.model.step_batch(self._to_batch([(k, d) for k, d in new if k in miss]), init_mode=True)

    def _write_back(self) -> None:
        if self.local_table is None or not self._docs:
            return
        here = {k for k, _ in self.input.get_local_tablet().get_data_iterator()}
        docs = [(k, d) for k, d in self._docs if k in here]
        if not docs:
            return
        z = self.model.z.cpu().numpy()
        keys, models = [], []
        for k, d in docs:
            zz = z[self._tok_base[k]: self._tok_base[k] + len(d)].astype(np.int32)
            keys.append(k)
            models.append((zz, np.bincount(zz, minlength=self.K).astype(np.int32)))
        self.local_table.multi_put(keys, models).get(timeout=600)

    def on_epoch_finished(self, epoch_idx: int) -> None:
        super().on_epoch_finished(epoch_idx)
        if self._elastic:
            self._write_back()        # worker blocks can only move under an optimizer's plans

    def flush(self) -> None:
        self._torch.cuda.current_stream(self.dtable.device).synchronize()

    def _to_batch(self, mini_batch):
        from ..models.lda import build_lda_batch
        mb = sorted(mini_batch, key=lambda kv: kv[0])
        key = (mb[0][0], len(mb))
        b = self._cache.get(key)
        if b is None:
            self._adopt(mb)
            # documents of one input block are contiguous in key order => contiguous token range (also for blocks
            # that migrated in: `_adopt` lays a block's documents out back to back)
            b = build_lda_batch([d for _, d in mb], self._tok_base[mb[0][0]], self.V).to(self.dtable.device)
            self._cache[key] = b
        return b

    def init_global_settings(self) -> None:
        """Random initial assignment of every local token; the counts are pushed to the owners."""
        if not self._docs or self._restored:
            return
        self.comps.gate.acquire_read()
        try:
            self.model.step_batch(self._to_batch(self._docs), init_mode=True)
            self._torch.cuda.synchronize(self.dtable.device)
        finally:
            self.comps.gate.release_read()

    def set_mini_batch_data(self, mini_batch) -> None:
        self._batch = self._to_batch(mini_batch)

    def pull_model(self) -> None:
        self.comps.gate.acquire_read()

    def local_compute(self) -> None:
        pass

    def push_update(self) -> None:
        try:
            self.model.step_batch(self._batch)
        finally:
            self.comps.gate.release_read()

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        data = sorted(input_data, key=lambda kv: kv[0])
        if not data:
            return {"docLLH": 0.0, "wordLLH": 0.0}
        b = self._to_batch(data)
        # K9 on the device: the lgamma sums run over the pulled count rows where they are (k12_ops.cu)
        import math
        from ..ops import native
        nat = native()
        rows = self.dtable.pull(b.vocab)                                   # [U + 1, stride] int32, last = summary
        ndk = self.model.doc_topic_counts(b)
        lens = self._torch.tensor([len(d) for _, d in data], dtype=self._torch.float64, device=rows.device)
        K, V, a, be = self.K, self.V, self.alpha, self.beta
        n_words = rows.shape[0] - 1
        word = K * (math.lgamma(V * be) - V * math.lgamma(be)) + nat.lgamma_sum_i32(rows[:-1], K, be) \
            + math.lgamma(be) * K * (V - n_words) - nat.lgamma_sum_i32(rows[-1:], K, V * be)
        doc = len(data) * (math.lgamma(K * a) - K * math.lgamma(a)) + nat.lgamma_sum_i32(ndk, K, a) \
            - float(self._torch.lgamma(lens + K * a).sum())
        return {"docLLH": doc, "wordLLH": word}

    def on_worker_stopped(self) -> None:
        self.cleanup()          # publish the HBM-resident local model before its table blocks migrate

    def cleanup(self) -> None:
        self._write_back()
