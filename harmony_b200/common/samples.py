"""Synthetic sample inputs in the reference's text formats.

The reference ships small sample files next to its launch scripts (jobserver/bin/sample_mlr, sample_nmf,
sample_lda, sample_lasso, sample_gbt(+.meta), *_test, sample_host_to_bandwidth, inputs/adj_list,
inputs/shortest_path).  This repo generates equivalents instead of shipping data:
`python -m harmony_b200.common.samples [dir]` (bin/run_all.sh and the EXAMPLE USAGE lines use them).
Formats: MLR/Lasso/GBT `y idx:val ...`; NMF `row: col,val ...` (1-based); LDA one document of word ids per line;
graphs `vertex target [weight] ...`; `#` lines are comments.
"""
from __future__ import annotations

import os
import sys

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


def write_lasso_sample(path, n=200, features=10, seed=0, noise=0.01):
    rng = np.random.default_rng(seed)
    w = np.zeros(features)
    w[rng.choice(features, max(1, features // 4), replace=False)] = rng.uniform(1.0, 3.0, max(1, features // 4))
    rng2 = np.random.default_rng(seed + 1000)
    with open(path, "w") as f:
        f.write("# synthetic lasso sample\n")
        for _ in range(n):
            x = rng2.standard_normal(features)
            f.write(f"{x @ w + noise * rng2.standard_normal():.5f} " +
                    " ".join(f"{i}:{v:.5f}" for i, v in enumerate(x)) + "\n")
    return path


def write_gbt_sample(path, meta_path=None, n=160, features=784, active=12, seed=0):
    """Regression target on a few of `features` sparse numerical features (+ metadata: all numerical, y real)."""
    rng = np.random.default_rng(seed)
    cols = rng.choice(features, active, replace=False)
    rng2 = np.random.default_rng(seed + 1000)
    with open(path, "w") as f:
        f.write("# synthetic gbt sample\n")
        for _ in range(n):
            v = rng2.random(active)
            y = 3.0 * (v[0] > 0.5) + 2.0 * v[1] - 1.5 * (v[2] < 0.3) + 0.05 * rng2.standard_normal()
            order = np.argsort(cols)
            f.write(f"{y:.4f} " + " ".join(f"{cols[j]}:{v[j]:.4f}" for j in order) + "\n")
    if meta_path:
        with open(meta_path, "w") as f:
            f.write("# feature:type (0 = numerical, k = categorical with k values); the last entry describes y\n")
            f.write(" ".join(f"{i}:0" for i in range(features + 1)) + "\n")
    return path


def write_graph_samples(dir_path, vertices=40, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(dir_path, exist_ok=True)
    adj = {v: sorted(set(int(t) for t in rng.choice(vertices, rng.integers(1, 5)) if t != v)) for v in range(vertices)}
    with open(os.path.join(dir_path, "adj_list"), "w") as f:
        for v, ts in adj.items():
            f.write(f"{v} " + " ".join(str(t) for t in ts) + "\n")
    with open(os.path.join(dir_path, "shortest_path"), "w") as f:
        for v, ts in adj.items():
            ts = ts + ([(v + 1) % vertices] if (v + 1) % vertices not in ts else [])   # keep the graph connected
            f.write(f"{v} " + " ".join(f"{t} {int(rng.integers(1, 10))}" for t in ts) + "\n")
    return dir_path


def write_all(dir_path: str, force: bool = False) -> str:
    """Everything the launch scripts' EXAMPLE USAGE lines reference, sized like the reference's samples."""
    os.makedirs(dir_path, exist_ok=True)
    p = lambda name: os.path.join(dir_path, name)  # noqa: E731

    def need(name):
        return force or not os.path.exists(p(name))
    if need("sample_mlr"):
        write_mlr_sample(p("sample_mlr"), n=1000, features=784, classes=10, nnz=60, seed=0)
    if need("sample_mlr_test"):
        write_mlr_sample(p("sample_mlr_test"), n=200, features=784, classes=10, nnz=60, seed=0)
    if need("sample_nmf"):
        write_nmf_sample(p("sample_nmf"), rows=250, cols=100, rank=8, density=0.2, seed=0)
    if need("sample_lda"):
        write_lda_sample(p("sample_lda"), docs=200, vocab=17935, topics=10, doc_len=60, seed=0)
    if need("sample_lasso"):
        write_lasso_sample(p("sample_lasso"), n=256, features=10, seed=0)
    if need("sample_lasso_test"):
        write_lasso_sample(p("sample_lasso_test"), n=50, features=10, seed=0, noise=0.02)
    if need("sample_gbt") or need("sample_gbt.meta"):
        write_gbt_sample(p("sample_gbt"), p("sample_gbt.meta"), n=160, features=784, seed=0)
    if need("sample_gbt_test"):
        write_gbt_sample(p("sample_gbt_test"), None, n=40, features=784, seed=0)
    if need("sample_host_to_bandwidth"):
        import socket
        with open(p("sample_host_to_bandwidth"), "w") as f:
            f.write("# hostname bits-per-second (consumed by -host_to_bandwidth_file_path)\n")
            f.write(f"{socket.gethostname()} {770e9 * 8:.0f}\nlocalhost {770e9 * 8:.0f}\n")
    if need(os.path.join("inputs", "adj_list")) or need(os.path.join("inputs", "shortest_path")):
        write_graph_samples(p("inputs"))
    return dir_path


if __name__ == "__main__":
    print(write_all(sys.argv[1] if len(sys.argv) > 1 else "."))
