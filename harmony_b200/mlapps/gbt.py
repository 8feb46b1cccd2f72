"""Gradient-boosted trees.

Reference: dolphin/mlapps/gbt/GBTTrainer.java:244-760 -- every mini-batch pulls the forest of a
label (keys label*num_keys + k, each a list of trees), computes residuals, grows ONE tree of depth
<= max_depth_of_tree by exact greedy splits (continuous features: sorted scan; categorical
features: category groups) with gain gL^2/(2 nL + lambda) + gR^2/(2 nR + lambda) - parent
(:895-905), leaves no smaller than leaf_min_size, splits with gain <= gamma rejected, and pushes
the tree to a random key of the label; GBTETModelUpdateFunction.java:40-48 appends it.
Metadata (`-metadata_path`, sample_gbt.meta): `idx:val ...`, val == 0 -> numerical feature,
val != 0 -> categorical with that many categories; entry `features:val` describes y (0 =
regression, k = k-class classification).  Input: `y idx:val ...`.
"""
from __future__ import annotations

import random
from typing import Any, Dict, List, Tuple

import numpy as np

from ..common.dataloader import DataParser, register_parser, strip_scheme
from ..dolphin import params as P
from ..dolphin.conf import ETDolphinConfiguration
from ..dolphin.worker import Trainer
from ..et.config import TableConfiguration
from ..utils.flags import Configuration


class GBTDataParser(DataParser):
    """`y idx:val ...` -> (idx, val, y).  Reference: dolphin/mlapps/gbt/GBTETDataParser.java."""

    index_bound_param = "features"

    def record_indices(self, record):
        return record[0]

    native_mode = 0

    def from_arrays(self, head, ptr, idx, val):
        return [(idx[ptr[i]:ptr[i + 1]], val[ptr[i]:ptr[i + 1]], float(head[i])) for i in range(len(head))]

    def parse_line(self, line: str):
        toks = line.replace(":", " ").split()
        y = float(toks[0])
        idx = np.asarray([int(t) for t in toks[1::2]], dtype=np.int32)
        val = np.asarray([float(t) for t in toks[2::2]], dtype=np.float32)
        return idx, val, y


register_parser("gbt", GBTDataParser)


def read_metadata(path: str, num_features: int) -> Tuple[Dict[int, int], int]:
    """Returns ({feature: num_categories or 0}, y_type) ; y_type 0 = regression else #classes."""
    types = {i: 0 for i in range(num_features)}
    ytype = 0
    if path:
        with open(strip_scheme(path)) as f:
            for line in f:
                s = line.strip()
                if not s or s.startswith("#"):
                    continue
                for tok in s.split():
                    a, _, b = tok.partition(":")
                    i, v = int(a), int(float(b))
                    if i == num_features:
                        ytype = v
                    else:
                        types[i] = v
    return types, ytype


class GBTree:
    """Array-encoded binary tree: node i has children 2i+1 / 2i+2."""

    def __init__(self, max_depth: int) -> None:
        n = 2 ** (max_depth + 1) - 1
        self.feature = np.full(n, -1, dtype=np.int32)     # -1 = leaf / absent
        self.threshold = np.zeros(n, dtype=np.float32)
        self.categories: Dict[int, frozenset] = {}         # node -> categories going LEFT
        self.value = np.zeros(n, dtype=np.float32)
        self.is_leaf = np.zeros(n, dtype=bool)

    def predict(self, x: np.ndarray) -> float:
        i = 0
        while not self.is_leaf[i]:
            f = self.feature[i]
            if f < 0:
                return 0.0
            if i in self.categories:
                i = 2 * i + 1 if int(x[f]) in self.categories[i] else 2 * i + 2
            else:
                i = 2 * i + 1 if x[f] <= self.threshold[i] else 2 * i + 2
        return float(self.value[i])

    def predict_many(self, X: np.ndarray) -> np.ndarray:
        return np.asarray([self.predict(x) for x in X], dtype=np.float64)


class GBTTrainer(Trainer):
    """Per mini-batch: pull the label's forest, residuals, grow one tree by exact greedy splits, push it to a
    random key of the label.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/mlapps/gbt/GBTTrainer.java:244-760."""

    def __init__(self, params: Dict[str, Any], model_accessor, **_: Any) -> None:
        self.acc = model_accessor
        self.num_features = int(params["features"])
        self.num_keys = max(1, int(params.get("num_keys", 1) or 1))
        self.step = float(params["step_size"])
        self.lam = float(params.get("lambda", 0.0))
        self.gamma = float(params.get("gamma", 0.0))
        self.max_depth = int(params.get("max_depth_of_tree", 7))
        self.leaf_min = int(params.get("leaf_min_size", 4))
        self.types, self.ytype = read_metadata(params.get("metadata_path", ""), self.num_features)
        self.labels = [0] if self.ytype == 0 else list(range(self.ytype))
        self.rng = random.Random(int(params.get("seed", 0)))
        self._trees: List[Tuple[int, GBTree]] = []

    def _dense(self, batch):
        X = np.zeros((len(batch), self.num_features), dtype=np.float32)
        y = np.zeros(len(batch), dtype=np.float64)
        for i, (_, (idx, val, yy)) in enumerate(batch):
            X[i, idx] = val
            y[i] = yy
        return X, y

    def set_mini_batch_data(self, mini_batch) -> None:
        self._X, self._y = self._dense(mini_batch)

    def _forest(self, label: int) -> List[GBTree]:
        keys = [label * self.num_keys + k for k in range(self.num_keys)]
        return [t for lst in self.acc.pull_many(keys) for t in (lst or [])]

    def pull_model(self) -> None:
        self._forests = {lb: self._forest(lb) for lb in self.labels}

    def _predict(self, forest: List[GBTree], X: np.ndarray) -> np.ndarray:
        out = np.zeros(len(X), dtype=np.float64)
        for t in forest:
            out += t.predict_many(X)
        return out

    def _residual(self, label: int) -> np.ndarray:
        pred = self._predict(self._forests[label], self._X)
        if self.ytype == 0:
            return self._y - pred
        return (np.abs(self._y - label) < 1e-9).astype(np.float64) - pred

    def _gain(self, gl, gr, nl, nr, total):
        return gl * gl / (2 * nl + self.lam) + gr * gr / (2 * nr + self.lam) - total

    def _best_split(self, X, g):
        n = len(g)
        G = g.sum()
        total = G * G / (2 * n + self.lam)
        best = (self.gamma, None, None, None)       # (gain, feature, threshold, left categories)
        for f in range(self.num_features):
            col = X[:, f]
            if self.types.get(f, 0) == 0:
                order = np.argsort(col, kind="stable")
                cs = np.cumsum(g[order])
                sv = col[order]
                for i in range(self.leaf_min - 1, n - self.leaf_min):
                    if sv[i] == sv[i + 1]:
                        continue
                    gain = self._gain(cs[i], G - cs[i], i + 1, n - i - 1, total)
                    if gain > best[0]:
                        best = (gain, f, float((sv[i] + sv[i + 1]) / 2), None)
            else:
                cats = np.unique(col.astype(np.int64))
                stats = sorted(((g[col == c].mean(), c) for c in cats))
                left: List[int] = []
                gl, nl = 0.0, 0
                for _, c in stats[:-1]:                 # categories ordered by mean gradient
                    m = col == c
                    left.append(int(c)); gl += g[m].sum(); nl += int(m.sum())
                    if nl < self.leaf_min or n - nl < self.leaf_min:
                        continue
                    gain = self._gain(gl, G - gl, nl, n - nl, total)
                    if gain > best[0]:
                        best = (gain, f, 0.0, frozenset(left))
        return best

    def _build(self, X, g) -> GBTree:
        tree = GBTree(self.max_depth)
        stack = [(0, np.arange(len(g)), 0)]
        while stack:
            node, idx, depth = stack.pop()
            gi = g[idx]
            split = (None, None, None, None)
            if depth < self.max_depth and len(idx) >= 2 * self.leaf_min:
                split = self._best_split(X[idx], gi)
            if split[1] is None:
                tree.is_leaf[node] = True
                tree.value[node] = self.step * gi.sum() / (len(idx) + self.lam)
                continue
            _, f, thr, cats = split
            tree.feature[node], tree.threshold[node] = f, thr
            if cats is not None:
                tree.categories[node] = cats
                m = np.isin(X[idx, f].astype(np.int64), list(cats))
            else:
                m = X[idx, f] <= thr
            stack.append((2 * node + 1, idx[m], depth + 1))
            stack.append((2 * node + 2, idx[~m], depth + 1))
        return tree

    def local_compute(self) -> None:
        self._trees = [(lb, self._build(self._X, self._residual(lb))) for lb in self.labels]

    def push_update(self) -> None:
        for lb, t in self._trees:
            self.acc.push(lb * self.num_keys + self.rng.randrange(self.num_keys), t)

    def evaluate_model(self, input_data, test_data) -> Dict[str, float]:
        out = {}
        for name, data in (("train", list(input_data)), ("test", [(None, t) for t in test_data])):
            if not data:
                continue
            X, y = self._dense(data)
            preds = np.stack([self._predict(self._forest(lb), X) for lb in self.labels], axis=1)
            if self.ytype == 0:
                out[f"{name}_mse"] = float(((preds[:, 0] - y) ** 2).mean())
            else:
                out[f"{name}_accuracy"] = float((preds.argmax(1) == y.astype(np.int64)).mean())
        return out

    def summary(self) -> Dict[str, float]:
        r = self._residual(self.labels[0])
        return {"last_batch_residual_mse": float((r * r).mean())}


def model_table_conf(conf: Configuration, table_id: str) -> TableConfiguration:
    return TableConfiguration(table_id=table_id, update_function="append")


APP = ETDolphinConfiguration(app_id="GBT", trainer_class="harmony_b200.mlapps.gbt:GBTTrainer", input_parser="gbt",
                             model_table_conf=model_table_conf, parameters=P.GBT,
                             device_trainer_class="harmony_b200.mlapps.gbt_device:GBTDeviceTrainer",
                             device_trainer_host_model=True)
APP_PATH = "harmony_b200.mlapps.gbt:APP"


def main_submit(argv: List[str]) -> int:
    from ..dolphin.launcher import main_submit as go
    return go(APP_PATH, APP, argv)


def main_run(argv: List[str]) -> int:
    from ..dolphin.launcher import main_run as go
    return go(APP, argv)
