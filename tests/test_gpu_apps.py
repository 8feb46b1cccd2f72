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
