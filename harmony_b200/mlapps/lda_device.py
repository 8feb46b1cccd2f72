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
        self._cache: Dict[Any, Any] = {}
        self._batch = None

    def _to_batch(self, mini_batch):
        from ..models.lda import build_lda_batch
        mb = sorted(mini_batch, key=lambda kv: kv[0])
        key = (mb[0][0], len(mb))
        b = self._cache.get(key)
        if b is None:
            # documents of one input block are contiguous in key order => contiguous token range
            b = build_lda_batch([d for _, d in mb], self._tok_base[mb[0][0]], self.V).to(self.dtable.device)
            self._cache[key] = b
        return b

    def init_global_settings(self) -> None:
        """Random initial assignment of every local token; the counts are pushed to the owners."""
        if not self._docs:
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
            self._torch.cuda.current_stream(self.dtable.device).synchronize()
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
        if self.local_table is not None and self._docs:
            z = self.model.z.cpu().numpy()
            keys, models = [], []
            for k, d in self._docs:
                zz = z[self._tok_base[k]: self._tok_base[k] + len(d)].astype(np.int32)
                keys.append(k)
                models.append((zz, np.bincount(zz, minlength=self.K).astype(np.int32)))
            self.local_table.multi_put(keys, models).get(timeout=600)
