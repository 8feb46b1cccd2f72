"""Synthetic sample-data writers used by the tests (the reference ships tiny sample files)."""
import os

import numpy as np


def write_mlr_sample(path, n=200, features=64, classes=4, nnz=8, seed=0):
    rng = np.random.default_rng(seed)
    Wtrue = rng.standard_normal((features, classes))
    with open(path, "w") as f:
        f.write("# synthetic mlr sample\n")
        for _ in range(n):
            idx = np.sort(rng.choice(features, nnz, replace=False))
            val = rng.random(nnz).astype(np.float32) + 0.5
            label = int(np.argmax(val @ Wtrue[idx]))
            f.write(str(label) + " " + " ".join(f"{i}:{v:.4f}" for i, v in zip(idx, val)) + "\n")
    return path


def write_nmf_sample(path, rows=60, cols=40, rank=4, density=0.3, seed=0):
    rng = np.random.default_rng(seed)
    L, R = rng.random((rows, rank)), rng.random((rank, cols))
    D = L @ R
    with open(path, "w") as f:
        f.write("# synthetic nmf sample\n")
        for i in range(rows):
            cs = np.flatnonzero(rng.random(cols) < density)
            if len(cs) == 0:
                cs = np.array([rng.integers(cols)])
            f.write(f"{i + 1}: " + " ".join(f"{c + 1},{D[i, c]:.4f}" for c in cs) + "\n")
    return path


def write_lda_sample(path, docs=60, vocab=50, topics=3, doc_len=30, seed=0):
    rng = np.random.default_rng(seed)
    phi = rng.dirichlet(np.full(vocab, 0.05), topics)
    with open(path, "w") as f:
        f.write("# synthetic lda sample\n")
        for _ in range(docs):
            theta = rng.dirichlet(np.full(topics, 0.3))
            z = rng.choice(topics, doc_len, p=theta)
            ws = [rng.choice(vocab, p=phi[t]) for t in z]
            f.write(" ".join(str(w) for w in ws) + "\n")
    return path


class FailingTrainer:
    """Trainer that raises in local_compute on ONE executor (fail-fast test)."""

    def __init__(self, params, model_accessor, context=None, **_):
        self.ctx = context
        self.victim = params.get("victim")

    def init_global_settings(self): pass
    def set_mini_batch_data(self, b): pass
    def pull_model(self): pass

    def local_compute(self):
        if self.ctx.executor.executor_id == self.victim:
            raise RuntimeError("injected trainer failure")

    def push_update(self): pass
    def on_epoch_finished(self, e): pass
    def cleanup(self): pass
    def summary(self): return {}


class RendezvousTrainer(FailingTrainer):
    """Trainer whose CONSTRUCTOR gathers an object from every worker of the job through the control-plane group
    (what the MLR ship mode does with its buffer addresses)."""

    def __init__(self, params, model_accessor, context=None, **kw):
        super().__init__(params, model_accessor, context=context, **kw)
        g = context.job_group
        self.seen = g.all_gather_object((g.rank, context.executor.executor_id), tag="hello")

    def local_compute(self): pass

    def summary(self):
        return {"group": [list(x) for x in self.seen]}
