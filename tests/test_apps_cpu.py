"""NMF / LDA on the host backend through the stand-alone launcher (run_<app>.sh path)."""
import os
import tempfile

import numpy as np
import pytest

from harmony_b200.dolphin.launcher import ETDolphinLauncher
from harmony_b200.mlapps import lda, nmf

from helpers import write_lda_sample, write_nmf_sample


def test_nmf_standalone_loss_decreases():
    tmp = tempfile.mkdtemp(prefix="hbnmf")
    data = write_nmf_sample(os.path.join(tmp, "sample_nmf"), rows=80, cols=30, rank=4)
    out = {}
    for epochs in (1, 25):
        res = ETDolphinLauncher.launch(nmf.APP, ["-input", data, "-max_num_epochs", str(epochs),
                                                 "-num_mini_batches", "4", "-num_worker_blocks", "4",
                                                 "-num_server_blocks", "4", "-step_size", "0.01", "-rank", "4",
                                                 "-features", "30", "-num_dolphin_workers", "2", "-num_servers", "1",
                                                 "-decay_rate", "1.0"])
        assert res.ok
        out[epochs] = sum(r["last_batch_loss"] for r in res.results)
    assert out[25] < 0.5 * out[1], out


def test_nmf_kv_model_table_without_num_cols():
    tmp = tempfile.mkdtemp(prefix="hbnmf2")
    data = write_nmf_sample(os.path.join(tmp, "sample_nmf"), rows=40, cols=20, rank=3)
    res = ETDolphinLauncher.launch(nmf.APP, ["-input", data, "-max_num_epochs", "2", "-num_mini_batches", "2",
                                             "-num_worker_blocks", "2", "-num_server_blocks", "4",
                                             "-step_size", "0.01", "-rank", "3", "-num_dolphin_workers", "1",
                                             "-num_servers", "0"])
    assert res.ok


def test_lda_standalone_runs_and_conserves_counts():
    tmp = tempfile.mkdtemp(prefix="hblda")
    data = write_lda_sample(os.path.join(tmp, "sample_lda"), docs=40, vocab=30, topics=3, doc_len=20)
    res = ETDolphinLauncher.launch(lda.APP, ["-input", data, "-max_num_epochs", "3", "-num_mini_batches", "4",
                                             "-num_worker_blocks", "4", "-num_server_blocks", "4",
                                             "-num_topics", "3", "-num_vocabs", "30", "-num_dolphin_workers", "2",
                                             "-num_servers", "1"])
    assert res.ok
    assert sum(r["num_batches"] for r in res.results) >= 12


def test_lda_host_trainer_keeps_global_counts_consistent():
    """After init + a few sweeps the summary row equals the column sums of the word rows and the
    total equals the number of tokens (the invariant the reference's update function maintains)."""
    from harmony_b200.et import TableConfiguration, local_cluster
    from harmony_b200.dolphin.worker import ETModelAccessor
    master, execs, launcher = local_cluster(2)
    try:
        K, V = 4, 25
        conf = dict(num_topics=K, num_vocabs=V, backend="host")
        master.create_table(lda.model_table_conf(conf, "m"), execs).get(10)
        master.create_table(lda.local_model_table_conf(conf, "l"), execs).get(10)
        master.create_table(TableConfiguration(table_id="in", num_total_blocks=4, is_ordered=True,
                                               update_function="void"), execs).get(10)
        ex = launcher.executors[execs[0].id]
        tin, tm, tl = (ex.tables.get_table(t) for t in ("in", "m", "l"))
        rng = np.random.default_rng(0)
        docs = [rng.integers(0, V, 15).astype(np.int32) for _ in range(12)]
        part = ex.tables.get_components("in").partitioner
        local_blocks = ex.tables.get_components("in").ownership.get_current_local_block_ids()
        keys = [part.key_range(local_blocks[i % len(local_blocks)])[0] + i for i in range(len(docs))]
        tin.multi_put(keys, docs).get(10)
        tr = lda.LDATrainer(dict(conf, job_id="j"), ETModelAccessor(tm), local_model_table=tl, input_table=tin)
        tr.init_global_settings()
        batch = list(tin.get_local_tablet().get_data_iterator())
        for _ in range(3):
            tr.set_mini_batch_data(batch); tr.pull_model(); tr.local_compute(); tr.push_update()
        rows = tm.multi_get(list(range(V + 1))).get(10)
        mat = np.stack([r if r is not None else np.zeros(K + 1, dtype=np.int32) for r in rows])
        assert mat[V, :K].sum() == sum(len(d) for d in docs)
        np.testing.assert_array_equal(mat[:V, :K].sum(axis=0), mat[V, :K])
        np.testing.assert_array_equal((mat[:, :K] > 0).sum(axis=1), mat[:, K])      # nnz bookkeeping
        m = tr.evaluate_model(batch, [])
        assert np.isfinite(m["docLLH"]) and np.isfinite(m["wordLLH"])
    finally:
        master.close()


def test_lasso_recovers_sparse_model():
    from harmony_b200.mlapps import lasso
    tmp = tempfile.mkdtemp(prefix="hblasso")
    rng = np.random.default_rng(1)
    F = 12
    w = np.zeros(F); w[[1, 5, 9]] = [2.0, -3.0, 1.5]
    path = os.path.join(tmp, "sample_lasso")
    with open(path, "w") as f:
        for _ in range(256):          # 4 even blocks of 64 (> F): a tiny first batch could zero a support
            x = rng.standard_normal(F)     # weight, which the reference's skip-zero rule then never revisits
            y = x @ w + 0.01 * rng.standard_normal()
            f.write(f"{y:.5f} " + " ".join(f"{i}:{v:.5f}" for i, v in enumerate(x)) + "\n")
    res = ETDolphinLauncher.launch(lasso.APP, ["-input", path, "-max_num_epochs", "10", "-num_mini_batches", "4",
                                               "-num_worker_blocks", "4", "-num_server_blocks", "2", "-features", str(F),
                                               "-features_per_partition", "6", "-step_size", "1.0", "-lambda", "0.5",
                                               "-model_gaussian", "0.1", "-num_dolphin_workers", "1",
                                               "-num_servers", "1", "-decay_rate", "1.0"])
    assert res.ok
    assert res.results[0]["last_batch_loss"] < 0.05


def test_gbt_regression_reduces_residual():
    from harmony_b200.mlapps import gbt
    tmp = tempfile.mkdtemp(prefix="hbgbt")
    rng = np.random.default_rng(2)
    path, meta = os.path.join(tmp, "sample_gbt"), os.path.join(tmp, "sample_gbt.meta")
    with open(path, "w") as f:
        for _ in range(240):
            x = rng.random(4)
            c = int(rng.integers(0, 3))
            y = 3.0 * (x[0] > 0.5) + 2.0 * x[1] + (1.5 if c == 2 else 0.0)
            f.write(f"{y:.4f} " + " ".join(f"{i}:{v:.4f}" for i, v in enumerate(x)) + f" 4:{c}\n")
    with open(meta, "w") as f:
        f.write("0:0 1:0 2:0 3:0 4:3 5:0\n")
    out = {}
    for epochs in (1, 8):
        res = ETDolphinLauncher.launch(gbt.APP, ["-input", path, "-max_num_epochs", str(epochs), "-num_mini_batches", "3",
                                                 "-num_worker_blocks", "3", "-num_server_blocks", "4", "-features", "5",
                                                 "-step_size", "0.5", "-lambda", "0.0", "-gamma", "0.0",
                                                 "-max_depth_of_tree", "3", "-leaf_min_size", "4", "-num_keys", "2",
                                                 "-metadata_path", meta, "-num_dolphin_workers", "1", "-num_servers", "1"])
        assert res.ok
        out[epochs] = res.results[0]["last_batch_residual_mse"]
    assert out[8] < 0.3 * out[1], out


@pytest.mark.parametrize("lam", [0.0, 0.05])
def test_host_mlr_gradient_matches_autograd(lam):
    """The host MLR trainer's pushed delta == -step * d/dW [ sum_i CE_i + lam * B / 2 * |W|^2 ] by torch autograd
    (the same objective the GPU kernels are tested against: models/mlr.py:mlr_reference_step)."""
    import torch
    from harmony_b200.mlapps.mlr import MLRTrainer
    rng = np.random.default_rng(0)
    F, C, B, step = 40, 5, 16, 0.1
    W0 = rng.standard_normal((F, C)).astype(np.float32) * 0.3
    batch = []
    for i in range(B):
        idx = np.sort(rng.choice(F, 6, replace=False)).astype(np.int32)
        batch.append((i, (idx, rng.random(6).astype(np.float32), int(rng.integers(0, C)))))

    class Acc:
        pushed = None

        def pull_many(self, keys):
            return W0[np.asarray(keys)]

        def push_many(self, keys, deltas):
            Acc.pushed = (np.asarray(keys), np.asarray(deltas))
    tr = MLRTrainer({"classes": C, "features": F, "step_size": step, "lambda": lam, "decay_rate": 0.9,
                     "decay_period": 5}, Acc())
    tr.set_mini_batch_data(batch); tr.pull_model(); tr.local_compute(); tr.push_update()
    keys, delta = Acc.pushed
    W = torch.tensor(W0, dtype=torch.float64, requires_grad=True)
    X = torch.zeros((B, F), dtype=torch.float64)
    y = torch.tensor([b[1][2] for b in batch])
    for i, (_, (idx, val, _l)) in enumerate(batch):
        X[i, idx.astype(np.int64)] = torch.tensor(val, dtype=torch.float64)
    loss = torch.nn.functional.cross_entropy(X @ W, y, reduction="sum") + lam * B / 2 * (W * W).sum()
    loss.backward()
    want = (-step * W.grad).numpy()
    got = np.zeros((F, C))
    got[keys] = delta
    touched = np.unique(np.concatenate([b[1][0] for b in batch]))
    rows = np.arange(F) if lam else touched
    np.testing.assert_allclose(got[rows], want[rows], rtol=2e-4, atol=2e-6)
    if not lam:
        assert np.all(got[np.setdiff1d(np.arange(F), touched)] == 0)


def test_host_nmf_gradients_match_autograd():
    """nmf_gradients == d/d(L, R) [ sum_nnz (L_i . R_j - D_ij)^2 + lam (|L|^2 + |R|^2) ] by torch autograd."""
    import torch
    from harmony_b200.mlapps.nmf import nmf_gradients
    rng = np.random.default_rng(1)
    n, U, rank, lam = 7, 5, 4, 0.03
    ptr = np.array([0, 2, 2, 5, 6, 9, 11, 12])
    slot = rng.integers(0, U, 12)
    vals = rng.random(12)
    L0, R0 = rng.random((n, rank)), rng.random((U, rank))
    gL, gR, loss = nmf_gradients(L0.copy(), R0.copy(), ptr, slot, vals, lam)
    L = torch.tensor(L0, requires_grad=True); R = torch.tensor(R0, requires_grad=True)
    rows = torch.tensor(np.repeat(np.arange(n), np.diff(ptr)))
    e = (L[rows] * R[torch.tensor(slot)]).sum(1) - torch.tensor(vals)
    obj = (e * e).sum() + lam * ((L * L).sum() + (R * R).sum())
    obj.backward()
    np.testing.assert_allclose(gL, L.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gR, R.grad.numpy(), rtol=1e-5, atol=1e-6)
    assert abs(loss - float((e * e).sum())) < 1e-9


def test_nmf_print_mat_stores_the_local_model_through_the_datastorer(tmp_path):
    """-print_mat true: every worker writes its rows of L (keyed by the matrix row id) under -base_dir."""
    import glob
    data = write_nmf_sample(str(tmp_path / "n"), rows=40, cols=20, rank=3)
    res = ETDolphinLauncher.launch(nmf.APP, ["-input", data, "-max_num_epochs", "2", "-num_mini_batches", "2",
                                             "-num_worker_blocks", "2", "-num_server_blocks", "4", "-step_size", "0.01",
                                             "-rank", "3", "-features", "20", "-num_dolphin_workers", "2",
                                             "-num_servers", "1", "-print_mat", "true", "-base_dir", str(tmp_path / "o")])
    assert res.ok
    files = sorted(glob.glob(str(tmp_path / "o" / "nmf" / "*" / "L-*.txt")))
    assert len(files) == 2
    rows = sorted(int(line.split(":")[0]) for f in files for line in open(f))
    assert rows == list(range(1, 41))
    assert all(len(line.split(":")[1].split()) == 3 for f in files for line in open(f))


def test_directory_input_with_part_files(tmp_path):
    """-input may be a directory of part files (the HDFS output layout): hidden / underscore files are skipped and the
    records of every part are loaded exactly once."""
    from harmony_b200.common.samples import write_mlr_sample
    d = tmp_path / "parts"
    d.mkdir()
    for i in range(3):
        write_mlr_sample(str(d / f"part-0000{i}"), n=40, features=64, classes=4, seed=i)
    (d / "_SUCCESS").write_text("")
    (d / ".part-00000.crc").write_text("garbage that must not be parsed")
    from harmony_b200.mlapps import mlr
    res = ETDolphinLauncher.launch(mlr.APP, ["-input", str(d), "-max_num_epochs", "1", "-num_mini_batches", "6",
                                             "-num_worker_blocks", "6", "-num_server_blocks", "4", "-step_size", "0.05",
                                             "-classes", "4", "-features", "64", "-num_dolphin_workers", "2",
                                             "-num_servers", "1"])
    assert res.ok, res.messages
    assert sum(r["num_batches"] for r in res.results) == 6
