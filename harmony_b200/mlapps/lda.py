"""Latent Dirichlet allocation (collapsed Gibbs sampling) app.

Reference: dolphin/mlapps/lda/LDATrainer.java:113-194 (initGlobalSettings: random assignment, push
initial counts), :210-262 (per batch: pull the batch vocabulary rows + the summary row, sample,
push deltas), :317-357, :363-376; SparseLDASampler.java:94-274; LDAETModelUpdateFunction.java:38-64
(sparse (topic, delta) pairs, clamp >= 0, last slot = #non-zero topics); LDAStatCalculator.java:103-184
(doc / word log-likelihood); LDAETDataParser.java:33-53 (`w w w ...`); LDAParameters.java:26-43.

Global model: key word in [0, V) plus key V (topic totals) -> int[K+1].  Local model: doc key ->
(assignments, topic counts).  On the device backend the global model is a dense int32 slab table
and deltas are pushed as (row, topic, +-1) triples (ops/csrc/table_ops.cu push_sparse_i32_kernel).
"""
from __future__ import annotations

import logging
import math
from typing import Any, Dict, List, Tuple

import numpy as np

from ..common.dataloader import DataParser, register_parser
from ..dolphin import params as P
from ..dolphin.conf import ETDolphinConfiguration
from ..dolphin.worker import Trainer
from ..et.config import TableConfiguration
from ..utils.flags import Configuration

LOG = logging.getLogger(__name__)


class LDADataParser(DataParser):
    """One document per line: word ids.  Reference: dolphin/mlapps/lda/LDAETDataParser.java."""

    index_bound_param = "num_vocabs"

    def record_indices(self, record):
        return record

    native_mode = 2

    def from_arrays(self, head, ptr, idx, val):
        return [idx[ptr[i]:ptr[i + 1]] for i in range(len(head))]

    def parse_line(self, line: str):
        return np.asarray([int(t) for t in line.split()], dtype=np.int32)


register_parser("lda", LDADataParser)


def lgamma_sum(x: np.ndarray) -> float:
    from scipy.special import gammaln
    return float(gammaln(x).sum())


def word_llh(word_rows: np.ndarray, summary: np.ndarray, K: int, V: int, beta: float) -> float:
    """LDAStatCalculator.computeWordLLH over dense count rows [n_words, K] and the totals [K]."""
    r = K * (math.lgamma(V * beta) - V * math.lgamma(beta))
    r += lgamma_sum(np.maximum(word_rows[:, :K], 0) + beta)
    r += math.lgamma(beta) * K * (V - word_rows.shape[0])      # words never seen: all-zero rows
    r -= lgamma_sum(np.maximum(summary[:K], 0) + V * beta)
    return r


def doc_llh(doc_topic_counts: np.ndarray, doc_lens: np.ndarray, K: int, alpha: float) -> float:
    n = len(doc_lens)
    r = n * (math.lgamma(K * alpha) - K * math.lgamma(alpha))
    r += lgamma_sum(np.maximum(doc_topic_counts, 0) + alpha)
    r -= lgamma_sum(doc_lens + K * alpha)
    return r


def deltas_to_pairs(delta_rows: Dict[int, Dict[int, int]]) -> Tuple[List[int], List[np.ndarray]]:
    keys, vals = [], []
    for w, d in delta_rows.items():
        flat = [x for t, c in d.items() if c != 0 for x in (t, c)]
        if flat:
            keys.append(w)
            vals.append(np.asarray(flat, dtype=np.int32))
    return keys, vals


class LDATrainer(Trainer):
    """Host trainer: dense per-token conditional (same distribution as the SparseLDA buckets)."""

    def __init__(self, params: Dict[str, Any], model_accessor, local_model_table=None, input_table=None,
                 **_: Any) -> None:
        self.acc = model_accessor
        self.local = local_model_table
        self.input = input_table
        self.K = int(params["num_topics"])
        self.V = int(params["num_vocabs"])
        self.alpha = float(params.get("alpha", 0.001))
        self.beta = float(params.get("beta", 0.01))
        self.rng = np.random.default_rng(int(params.get("seed", 0)) + hash(str(params.get("job_id"))) % 1000)
        self._changes: Dict[int, Dict[int, int]] = {}

    def _inc(self, w: int, t: int, c: int) -> None:
        d = self._changes.setdefault(w, {})
        d[t] = d.get(t, 0) + c

    def init_global_settings(self) -> None:
        self._changes = {}
        keys, models = [], []
        for key, doc in self.input.get_local_tablet().get_data_iterator():
            z = self.rng.integers(0, self.K, size=len(doc)).astype(np.int32)
            counts = np.bincount(z, minlength=self.K).astype(np.int32)
            keys.append(key)
            models.append((z, counts))
            for w, t in zip(doc.tolist(), z.tolist()):
                self._inc(w, t, 1)
                self._inc(self.V, t, 1)
        if keys:
            self.local.multi_put(keys, models).get(timeout=600)
            ks, vs = deltas_to_pairs(self._changes)
            self.acc.push_many(ks, vs)
        self._changes = {}

    def set_mini_batch_data(self, mini_batch) -> None:
        self._batch = list(mini_batch)
        self._words = sorted({int(w) for _, doc in self._batch for w in doc.tolist()})

    def pull_model(self) -> None:
        rows = self.acc.pull_many(self._words + [self.V])
        self._rows = {w: np.array(r, dtype=np.int64) for w, r in zip(self._words, rows[:-1])}
        self._summary = np.array(rows[-1], dtype=np.int64)

    def local_compute(self) -> None:
        K, a, b, vb = self.K, self.alpha, self.beta, self.V * self.beta
        self._changes = {}
        keys = [k for k, _ in self._batch]
        models = self.local.multi_get(keys).get(timeout=600)
        new_models = []
        for (key, doc), lm in zip(self._batch, models):
            z, ndk = np.array(lm[0]), np.array(lm[1], dtype=np.int64)
            for i, w in enumerate(doc.tolist()):
                old = int(z[i])
                row = self._rows[w]
                ndk[old] -= 1
                row[old] = max(row[old] - 1, 0)
                self._summary[old] = max(self._summary[old] - 1, 0)
                p = (ndk + a) * (np.maximum(row[:K], 0) + b) / (np.maximum(self._summary[:K], 0) + vb)
                c = np.cumsum(p)
                new = int(np.searchsorted(c, self.rng.random() * c[-1], side="right"))
                new = min(new, K - 1)
                z[i] = new
                ndk[new] += 1
                row[new] += 1
                self._summary[new] += 1
                if new != old:
                    self._inc(w, old, -1); self._inc(w, new, 1)
                    self._inc(self.V, old, -1); self._inc(self.V, new, 1)
            new_models.append((z.astype(np.int32), ndk.astype(np.int32)))
        self.local.multi_put(keys, new_models).get(timeout=600)

    def push_update(self) -> None:
        ks, vs = deltas_to_pairs(self._changes)
        if ks:
            self.acc.push_many(ks, vs)
        self._changes = {}

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        data = list(input_data)
        words = sorted({int(w) for _, doc in data for w in doc.tolist()})
        rows = self.acc.pull_many(words + [self.V])
        wr = np.stack([np.asarray(r) for r in rows[:-1]]) if words else np.zeros((0, self.K + 1))
        keys = [k for k, _ in data]
        lms = self.local.multi_get(keys).get(timeout=600) if keys else []
        ndk = np.stack([np.asarray(m[1]) for m in lms]) if lms else np.zeros((0, self.K))
        lens = np.asarray([len(d) for _, d in data], dtype=np.float64)
        return {"docLLH": doc_llh(ndk, lens, self.K, self.alpha),
                "wordLLH": word_llh(wr, np.asarray(rows[-1]), self.K, self.V, self.beta)}


def model_table_conf(conf: Configuration, table_id: str) -> TableConfiguration:
    K, V = int(conf["num_topics"]), int(conf["num_vocabs"])
    backend = conf.get("backend", "host")
    if backend == "device":
        return TableConfiguration(table_id=table_id, dense=True, capacity=V + 1, row_dim=K + 1, dtype="int32",
                                  update_function={"name": "lda_sparse", "params": {"num_topics": K}},
                                  backend="device")
    return TableConfiguration(table_id=table_id, update_function={"name": "lda_sparse", "params": {"num_topics": K}})


def local_model_table_conf(conf: Configuration, table_id: str) -> TableConfiguration:
    return TableConfiguration(table_id=table_id, update_function="assign")


APP = ETDolphinConfiguration(
    app_id="LDA", trainer_class="harmony_b200.mlapps.lda:LDATrainer", input_parser="lda",
    model_table_conf=model_table_conf, local_model_table_conf=local_model_table_conf, parameters=P.LDA,
    device_trainer_class="harmony_b200.mlapps.lda_device:LDADeviceTrainer")
APP_PATH = "harmony_b200.mlapps.lda:APP"


def main_submit(argv: List[str]) -> int:
    from ..dolphin.launcher import main_submit as go
    return go(APP_PATH, APP, argv)


def main_run(argv: List[str]) -> int:
    from ..dolphin.launcher import main_run as go
    return go(APP, argv)
