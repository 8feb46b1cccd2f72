"""NMF / LDA device kernels against PyTorch / numpy references + the apps through the job server."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import write_lda_sample, write_nmf_sample  # noqa: E402


@pytest.fixture(scope="module")
def arena():
    from harmony_b200.parallel.peer import PeerArena, PeerGroup
    a = PeerArena(0, PeerGroup(), segment_bytes=256 << 20)
    yield a
    a.close()


@pytest.mark.parametrize("rank,lam", [(8, 0.0), (256, 0.0), (40, 0.01)])
def test_nmf_step_matches_reference(arena, rank, lam):
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.models.nmf import NMFDeviceModel, nmf_reference_step, synthetic_nmf_batch
    cols, rows = 500, 300
    conf = TableConfiguration(table_id=f"R{rank}{lam}", dense=True, capacity=cols, row_dim=rank, num_total_blocks=16,
                              backend="device", update_function={"name": "sgd_clamp", "params": {
                                  "step": 0.01, "lo": 0.0, "hi": 5.0, "init": "uniform", "init_lo": 0.0,
                                  "init_hi": 1.0, "seed": 5}})
    t = DeviceTable(conf, arena)
    m = NMFDeviceModel(t, rank, rows, step=0.01, lam=lam, max_val=5.0, seed=1)
    b = synthetic_nmf_batch(rows, cols, 12, 0, seed=3).to("cuda")
    L0, R0 = m.L.clone(), t.to_dense().clone()
    m.loss.zero_()
    m.step_batch(b)
    Ln, Rn, loss = nmf_reference_step(L0, R0, b, 0.01, lam, 5.0)
    torch.testing.assert_close(m.L, Ln, rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(t.to_dense(), Rn, rtol=2e-3, atol=2e-4)
    assert abs(float(m.loss[0]) - loss) / loss < 2e-3
    t.close()


def test_lda_sampler_invariants_and_convergence(arena):
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.mlapps.lda import doc_llh, word_llh
    from harmony_b200.models.lda import LDADeviceModel, build_lda_batch
    K, V, nd, dl = 8, 200, 300, 40
    rng = np.random.default_rng(0)
    phi = rng.dirichlet(np.full(V, 0.02), K)
    docs = []
    for _ in range(nd):
        th = rng.dirichlet(np.full(K, 0.2))
        zs = rng.choice(K, dl, p=th)
        docs.append(np.array([rng.choice(V, p=phi[z]) for z in zs], dtype=np.int32))
    conf = TableConfiguration(table_id="lda_m", dense=True, capacity=V + 1, row_dim=K + 1, dtype="int32",
                              num_total_blocks=8, backend="device")
    t = DeviceTable(conf, arena)
    m = LDADeviceModel(t, K, V, nd * dl, alpha=0.1, beta=0.01, seed=3)
    b = build_lda_batch(docs, 0, V).to("cuda")
    m.step_batch(b, init_mode=True)

    def check():
        M = t.to_dense().cpu().numpy()
        assert M[V, :K].sum() == nd * dl
        np.testing.assert_array_equal(M[:V, :K].sum(axis=0), M[V, :K])
        np.testing.assert_array_equal((M[:, :K] > 0).sum(axis=1), M[:, K])
        assert (M >= 0).all()
        ndk = m.doc_topic_counts(b).cpu().numpy()
        assert (ndk.sum(1) == dl).all()
        return word_llh(M[:V], M[V], K, V, 0.01) + doc_llh(ndk, np.full(nd, dl, dtype=np.float64), K, 0.1)
    l0 = check()
    for _ in range(30):
        m.step_batch(b)
    l1 = check()
    assert l1 > l0 + 0.05 * abs(l0), (l0, l1)      # Gibbs sweeps increase the joint log-likelihood
    t.close()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_nmf_and_lda_jobs_on_device_backend():
    from harmony_b200.dolphin.launcher import DolphinJobLauncher
    from harmony_b200.jobserver.client import JobServer
    from harmony_b200.mlapps import lda, nmf
    from harmony_b200.utils.flags import Configuration
    tmp = tempfile.mkdtemp(prefix="hbgpu3")
    d_nmf = write_nmf_sample(os.path.join(tmp, "sample_nmf"), rows=120, cols=40, rank=4)
    d_lda = write_lda_sample(os.path.join(tmp, "sample_lda"), docs=80, vocab=50, topics=3, doc_len=25)
    port = _free_port()
    js = JobServer(Configuration({"num_executors": 2, "port": port, "backend": "device",
                                  "chkp_temp_path": tmp + "/t", "chkp_commit_path": tmp + "/c"})).start()
    try:
        a1 = ["-input", d_nmf, "-max_num_epochs", "20", "-num_mini_batches", "4", "-num_worker_blocks", "4",
              "-num_server_blocks", "8", "-step_size", "0.01", "-rank", "4", "-features", "40", "-decay_rate", "1.0",
              "-port", str(port)]
        a2 = ["-input", d_lda, "-max_num_epochs", "4", "-num_mini_batches", "4", "-num_worker_blocks", "4",
              "-num_server_blocks", "8", "-num_topics", "3", "-num_vocabs", "50", "-port", str(port)]
        # two concurrent jobs sharing the same two GPU executors
        assert DolphinJobLauncher.submit_job(nmf.APP_PATH, nmf.APP, a1, port=port).startswith("true")
        assert DolphinJobLauncher.submit_job(lda.APP_PATH, lda.APP, a2, port=port).startswith("true")
        ok1, r1 = js.driver.wait_job("NMF-0", timeout=120)
        ok2, r2 = js.driver.wait_job("LDA-1", timeout=120)
        assert ok1, r1
        assert ok2, r2
        assert all(np.isfinite(r["last_batch_loss"]) for r in r1)
    finally:
        js.stop(kill=True)


def test_device_pregel_pagerank_and_sssp(arena):
    from harmony_b200.parallel.peer import PeerGroup
    from harmony_b200.pregel.device import DevicePregel, partition_adjacency
    rng = np.random.default_rng(0)
    V = 500
    adj = {v: [(int(t), int(rng.integers(1, 9))) for t in rng.choice(V, size=rng.integers(1, 6), replace=False)]
           for v in range(V)}
    eng = DevicePregel(arena, PeerGroup(), V, partition_adjacency(adj, V, 0, 1))
    pr = eng.pagerank(10, 0.85)
    ref = np.ones(V)
    for _ in range(10):
        nxt = np.full(V, 0.15)
        for v, es in adj.items():
            for t, _ in es:
                nxt[t] += 0.85 * ref[v] / len(es)
        ref = nxt
    np.testing.assert_allclose([pr[v] for v in range(V)], ref, rtol=1e-4)
    sp = eng.shortest_path(source=3)
    import heapq
    dist = {3: 0}; pq = [(0, 3)]
    while pq:
        d, v = heapq.heappop(pq)
        if d > dist.get(v, 1 << 60):
            continue
        for t, w in adj[v]:
            if d + w < dist.get(t, 1 << 60):
                dist[t] = d + w; heapq.heappush(pq, (d + w, t))
    for v in range(V):
        assert sp[v] == dist.get(v, 2 ** 31 - 1), v
    eng.close()


# ---- K12 / K9: Lasso coordinate descent, GBT split search / prediction, lgamma sums (ops/csrc/k12_ops.cu) ----------
def test_lasso_cd_kernel_matches_float64_coordinate_descent(arena):
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    from harmony_b200.mlapps.lasso import sthresh
    from harmony_b200.ops import native
    F, fpp, B, lam, step = 48, 12, 700, 0.3, 0.5
    conf = TableConfiguration(table_id="lasso_w", dense=True, capacity=F // fpp, row_dim=fpp, num_total_blocks=4,
                              is_ordered=False, backend="device",
                              update_function={"name": "add", "params": {"init": "gaussian", "std": 0.1, "seed": 5}})
    t = DeviceTable(conf, arena)
    rng = np.random.default_rng(0)
    wt = np.zeros(F); wt[[3, 17, 40]] = [2.0, -1.0, 0.7]
    X = rng.standard_normal((B, F)).astype(np.float32)
    X[:, 7] = 0.0                                                      # an all-zero column is skipped
    y = (X @ wt + 0.01 * rng.standard_normal(B)).astype(np.float32)
    w0 = t.to_dense().reshape(-1).double().cpu().numpy()
    # float64 reference of the reference algorithm
    w = w0.copy(); pre = X.astype(np.float64) @ w
    for j in range(F):
        if abs(w[j]) < 1e-9:
            continue
        col = X[:, j].astype(np.float64); nrm = float(col @ col)
        if nrm < 1e-9:
            continue
        pre -= col * w[j]
        w[j] = sthresh(float(col @ (y - pre)) / nrm, lam, nrm)
        pre += col * w[j]
    dev = t.device
    Xt = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)
    w_new = torch.zeros(F, device=dev); w_old = torch.zeros(F, device=dev); stats = torch.zeros(4, device=dev)
    native().lasso_cd(*t.view(), Xt, torch.from_numpy(y).to(dev), lam, step, w_new, w_old, stats)
    torch.cuda.synchronize()
    np.testing.assert_allclose(w_new.cpu().numpy(), w, rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(w_old.cpu().numpy(), w0, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(t.to_dense().reshape(-1).cpu().numpy(), w0 + step * (w - w0), rtol=2e-3, atol=2e-4)
    assert abs(float(stats[0]) - float(((pre - y) ** 2).mean())) < 1e-3
    t.close()


def _gbt_data(n, seed):
    rng = np.random.default_rng(seed)
    X = rng.random((n, 5)).astype(np.float32)
    X[:, 2] = np.round(X[:, 2] * 8) / 8                                # repeated values: equal neighbours are no split
    X[:, 4] = rng.integers(0, 5, n)                                    # categorical
    y = 3.0 * (X[:, 0] > 0.5) + 2.0 * X[:, 1] + np.where(X[:, 4] == 2, 1.5, 0.0) + 0.05 * rng.standard_normal(n)
    return X, y.astype(np.float64)


@pytest.mark.parametrize("depth,leaf_min,lam,gamma", [(3, 4, 0.0, 0.0), (5, 2, 1.0, 0.05)])
def test_gbt_device_tree_matches_cpu_builder(depth, leaf_min, lam, gamma):
    from harmony_b200.mlapps.gbt import GBTTrainer
    from harmony_b200.models.gbt import DeviceTreeBuilder
    X, y = _gbt_data(600, 3)
    types = {0: 0, 1: 0, 2: 0, 3: 0, 4: 5}
    cpu = GBTTrainer.__new__(GBTTrainer)
    cpu.num_features, cpu.types, cpu.lam, cpu.gamma = 5, types, lam, gamma
    cpu.max_depth, cpu.leaf_min, cpu.step = depth, leaf_min, 0.5
    ref = cpu._build(X, y)
    b = DeviceTreeBuilder("cuda:0", 5, types, depth, leaf_min, lam, gamma)
    Xd = torch.from_numpy(X).to("cuda:0")
    t = b.build(Xd, b.presort(Xd), torch.from_numpy(y).to("cuda:0"), 0.5)
    got = b.to_host_tree(t)
    np.testing.assert_array_equal(got.is_leaf, ref.is_leaf)
    np.testing.assert_array_equal(got.feature, ref.feature)
    np.testing.assert_allclose(got.threshold, ref.threshold, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(got.value, ref.value, rtol=1e-5, atol=1e-6)
    assert got.categories == ref.categories
    assert (~ref.is_leaf & (ref.feature >= 0)).sum() >= 3               # a real tree was grown
    # forest prediction kernel == per-sample host traversal
    forest = [ref, got]
    pred = b.predict(Xd, b.pack_forest(forest)).cpu().numpy()
    np.testing.assert_allclose(pred, sum(tr.predict_many(X) for tr in forest), rtol=1e-6, atol=1e-6)


def test_lgamma_sum_kernel():
    from scipy.special import gammaln
    from harmony_b200.ops import native
    x = torch.randint(-3, 500, (777, 40), dtype=torch.int32, device="cuda:0")
    got = native().lgamma_sum_i32(x, 33, 0.1)
    ref = float(gammaln(np.maximum(x[:, :33].cpu().numpy(), 0) + 0.1).sum())
    assert abs(got - ref) < 1e-6 * abs(ref)


def test_lasso_and_gbt_jobs_on_device_backend():
    """Same scenarios as tests/test_apps_cpu.py, on the device backend (LassoDeviceTrainer / GBTDeviceTrainer)."""
    from harmony_b200.dolphin.launcher import ETDolphinLauncher
    from harmony_b200.mlapps import gbt, lasso
    tmp = tempfile.mkdtemp(prefix="hbk12")
    rng = np.random.default_rng(1)
    F = 12
    w = np.zeros(F); w[[1, 5, 9]] = [2.0, -3.0, 1.5]
    path = os.path.join(tmp, "sample_lasso")
    with open(path, "w") as f:
        for _ in range(256):          # 4 even blocks of 64 (> F): a tiny first batch could zero a support
            x = rng.standard_normal(F)     # weight, which the reference's skip-zero rule then never revisits
            f.write(f"{x @ w + 0.01 * rng.standard_normal():.5f} " + " ".join(f"{i}:{v:.5f}" for i, v in enumerate(x)) + "\n")
    res = ETDolphinLauncher.launch(lasso.APP, ["-input", path, "-max_num_epochs", "10", "-num_mini_batches", "4",
                                               "-num_worker_blocks", "4", "-num_server_blocks", "2", "-features", str(F),
                                               "-features_per_partition", "6", "-step_size", "1.0", "-lambda", "0.5",
                                               "-model_gaussian", "0.1", "-num_dolphin_workers", "1",
                                               "-num_servers", "0", "-decay_rate", "1.0", "-backend", "device"])
    assert res.ok and res.results[0]["trainer"] == "LassoDeviceTrainer", res.results
    assert res.results[0]["last_batch_loss"] < 0.05
    rng = np.random.default_rng(2)
    path, meta = os.path.join(tmp, "sample_gbt"), os.path.join(tmp, "sample_gbt.meta")
    with open(path, "w") as f:
        for _ in range(240):
            x = rng.random(4); c = int(rng.integers(0, 3))
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
                                                 "-metadata_path", meta, "-num_dolphin_workers", "1", "-num_servers", "0",
                                                 "-backend", "device"])
        assert res.ok and res.results[0]["trainer"] == "GBTDeviceTrainer", res.results
        out[epochs] = res.results[0]["last_batch_residual_mse"]
    assert out[8] < 0.3 * out[1], out
