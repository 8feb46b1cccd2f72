"""Numerics of the sm_100a kernels against plain PyTorch fp32 references (single GPU)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def arena():
    from harmony_b200.parallel.peer import PeerArena, PeerGroup
    a = PeerArena(0, PeerGroup(), segment_bytes=256 << 20)
    yield a
    a.close()


def _table(arena, capacity, row_dim, nblocks, ordered, fn, dtype="float32"):
    from harmony_b200.et.config import TableConfiguration
    from harmony_b200.et.device_table import DeviceTable
    conf = TableConfiguration(table_id=f"t{np.random.randint(1 << 30)}", dense=True, capacity=capacity,
                              row_dim=row_dim, dtype=dtype, num_total_blocks=nblocks, is_ordered=ordered,
                              update_function=fn, backend="device")
    return DeviceTable(conf, arena)


@pytest.mark.parametrize("ordered", [True, False])
@pytest.mark.parametrize("row_dim", [10, 64, 1000])
def test_pull_push_add(arena, ordered, row_dim):
    cap = 5000
    t = _table(arena, cap, row_dim, 16, ordered, {"name": "add", "params": {"init": "gaussian", "std": 0.01, "seed": 3}})
    W0 = t.to_dense().clone()
    assert W0.shape == (cap, row_dim)
    assert abs(float(W0.std()) - 0.01) < 0.002 and abs(float(W0.mean())) < 1e-3
    g = torch.Generator(device="cuda").manual_seed(1)
    keys = torch.randint(0, cap, (3000,), device="cuda", generator=g)
    deltas = torch.randn((3000, row_dim), device="cuda", generator=g)
    t.push(keys, deltas)
    ref = W0.clone().index_add_(0, keys, deltas)
    torch.testing.assert_close(t.to_dense(), ref, rtol=1e-4, atol=1e-4)
    out = t.pull(keys[:100])
    torch.testing.assert_close(out, ref[keys[:100]], rtol=1e-4, atol=1e-4)
    t.close()


def test_push_sgd_clamp(arena):
    cap, d = 1000, 32
    t = _table(arena, cap, d, 8, True, {"name": "sgd_clamp", "params": {"step": 0.5, "lo": 0.0, "hi": 1.0,
                                                                         "init": "uniform", "init_lo": 0.0, "init_hi": 1.0}})
    W0 = t.to_dense().clone()
    keys = torch.randperm(cap, device="cuda")[:500]
    g = torch.randn((500, d), device="cuda")
    t.push(keys, g)
    ref = W0.clone()
    ref[keys] = (W0[keys] - 0.5 * g).clamp(0.0, 1.0)
    torch.testing.assert_close(t.to_dense(), ref, rtol=1e-5, atol=1e-5)
    t.close()


def test_push_adagrad(arena):
    cap, d = 512, 48
    t = _table(arena, cap, d, 4, False, {"name": "adagrad", "params": {"step": 0.1, "eps": 1e-8}})
    keys = torch.randperm(cap, device="cuda")[:200]
    g = torch.randn((200, d), device="cuda")
    t.push(keys, g)
    ref = torch.zeros((cap, d), device="cuda")
    ref[keys] = -0.1 * g / torch.sqrt(g * g + 1e-8)
    torch.testing.assert_close(t.to_dense(), ref, rtol=1e-3, atol=1e-4)
    t.close()


def test_lda_sparse_push(arena):
    V, K = 300, 20
    t = _table(arena, V + 1, K + 1, 8, False, None, dtype="int32")
    t.update_fn  # void
    keys = torch.tensor([5, 5, 7, 7, 300], device="cuda")
    topics = torch.tensor([1, 2, 3, 3, 0], dtype=torch.int32, device="cuda")
    deltas = torch.tensor([2, 1, 4, -1, 9], dtype=torch.int32, device="cuda")
    t.push_sparse_i32(keys, topics, deltas)
    M = t.to_dense()
    assert M[5, 1] == 2 and M[5, 2] == 1 and M[5, K] == 2
    assert M[7, 3] == 3 and M[7, K] == 1 and M[300, 0] == 9 and M[300, K] == 1
    # decrement below zero clamps at zero and fixes the nnz counter
    t.push_sparse_i32(torch.tensor([5], device="cuda"), torch.tensor([2], dtype=torch.int32, device="cuda"),
                      torch.tensor([-5], dtype=torch.int32, device="cuda"))
    M = t.to_dense()
    assert M[5, 2] == 0 and M[5, K] == 1
    t.close()


@pytest.mark.parametrize("K", [20, 1000])
def test_lda_sparse_pull_through_topic_bitmaps(arena, K):
    """Sparse PULL (SparseArrayCodec.java:58-101 on the GPU): bitmap first, then only the non-zero 16 B chunks; the
    staging rows must equal the dense pull, before and after the owner-side bitmap rebuild, and after a restore."""
    V = 500
    t = _table(arena, V + 1, K + 1, 8, False, {"name": "lda_sparse", "params": {"num_topics": K}}, dtype="int32")
    assert t.bitmap_off > 0 and t.row_stride >= K + 1 + 32
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 4000
    keys = torch.randint(0, V + 1, (n,), device="cuda", generator=g)
    topics = torch.randint(0, K, (n,), device="cuda", generator=g).int()
    t.push_sparse_i32(keys, topics, torch.randint(1, 4, (n,), device="cuda", generator=g).int())
    t.push_sparse_i32(keys[: n // 2], topics[: n // 2], torch.full((n // 2,), -9, dtype=torch.int32, device="cuda"))  # zeros
    all_keys = torch.arange(V + 1, device="cuda")
    dense = t.pull(all_keys)                                            # [V + 1, K + 1]
    out = torch.full((V + 1, t.row_stride), -7, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    t.pull_lda_sparse(all_keys, out, cnt)
    assert torch.equal(out[:, : K + 1], dense)
    stale = int(cnt)
    t.rebuild_bitmaps()
    cnt.zero_()
    t.pull_lda_sparse(all_keys, out, cnt)
    assert torch.equal(out[:, : K + 1], dense) and int(cnt) <= stale
    assert int(cnt) * 16 < (V + 1) * t.row_stride * 4                 # fewer bytes than the dense rows
    # ASSIGN (checkpoint restore) maintains the bitmaps too
    t2 = _table(arena, V + 1, K + 1, 8, False, {"name": "lda_sparse", "params": {"num_topics": K}}, dtype="int32")
    t2.put_rows(all_keys, dense)
    t2.pull_lda_sparse(all_keys, out)
    assert torch.equal(out[:, : K + 1], dense)
    t.close(); t2.close()


@pytest.mark.parametrize("classes,features", [(10, 784), (1000, 20000), (1500, 5000)])
def test_mlr_step_matches_reference(arena, classes, features):
    from harmony_b200.models.mlr import MLRDeviceModel, mlr_reference_step, synthetic_batch
    t = _table(arena, features, classes, 32, True, {"name": "add", "params": {"init": "gaussian", "std": 0.05, "seed": 7}})
    model = MLRDeviceModel(t, classes, step_size=0.01)
    batch = synthetic_batch(257, features, classes, 13, seed=5).to("cuda")
    W0 = t.to_dense().clone()
    model.stats.zero_()
    model.step(batch)
    Wref, loss_ref, acc_ref = mlr_reference_step(W0, batch, classes, 0.01)
    torch.testing.assert_close(t.to_dense(), Wref, rtol=2e-3, atol=2e-5)
    s = model.stats.cpu()
    assert abs(float(s[0]) - loss_ref) / max(1.0, abs(loss_ref)) < 2e-3
    assert abs(float(s[1]) / 257 - acc_ref) < 0.02
    t.close()


def test_migrate_block_single_rank_copy(arena):
    """migration copy kernel: bulk slab copy is exact."""
    from harmony_b200.ops import native
    n = native()
    src = torch.randn(1 << 20, device="cuda")
    dst = torch.zeros_like(src)
    n.migrate_copy(src.data_ptr(), dst.data_ptr(), src.numel() * 4)
    torch.cuda.synchronize()
    assert torch.equal(src, dst)


def test_device_clock_single(arena):
    from harmony_b200.parallel.device_clock import DeviceClock
    c = DeviceClock(arena)
    c.tick(); c.wait(); c.barrier()
    torch.cuda.synchronize()
    c.check()
    assert int(c.clocks[0]) == 2


def test_mlr_ship_mode_single_gpu_matches_reference(arena):
    from harmony_b200.models.mlr import mlr_reference_step, synthetic_batch
    from harmony_b200.models.mlr_ship import MLRShipModel, route_batches
    from harmony_b200.parallel.device_clock import DeviceClock
    from harmony_b200.parallel.peer import PeerGroup
    classes, features = 1000, 30000
    t = _table(arena, features, classes, 64, False, {"name": "add", "params": {"init": "gaussian", "std": 0.05, "seed": 7}})
    g = PeerGroup()
    clock = DeviceClock(arena, g)
    b = synthetic_batch(300, features, classes, 17, seed=8)
    rb = route_batches(t, [b], 0).to("cuda")
    from harmony_b200.models.mlr_ship import MLRShipModelV1
    import functools
    for cls in (MLRShipModel, functools.partial(MLRShipModel, reduce_scatter="nvls"), MLRShipModelV1):
        m = cls(t, arena, g, clock, classes, 0.02, 300)     # push RS / pull RS (3 launches + flags) / v1 (barriers)
        for _ in range(2):
            W0 = t.to_dense().clone()
            m.stats.zero_()
            m.step(rb)
            torch.cuda.synchronize()
            Wref, loss_ref, _ = mlr_reference_step(W0, b.to("cuda"), classes, 0.02)
            torch.testing.assert_close(t.to_dense(), Wref, rtol=2e-3, atol=2e-5)
            assert abs(float(m.stats[0]) - loss_ref) / loss_ref < 2e-3
        if hasattr(m, "check"):
            m.check()
    t.close()


@pytest.mark.parametrize("mode", ["pull", "ship"])
def test_mlr_lazy_l2_matches_reference_with_lambda(arena, mode):
    """W' = (1 - step lam B) W + delta (MLRTrainer.java:387-397) through the lazy scale: no dense pass per step."""
    from harmony_b200.models.mlr import MLRDeviceModel, mlr_reference_step, synthetic_batch
    from harmony_b200.models.mlr_ship import MLRShipModel, route_batches
    from harmony_b200.parallel.peer import PeerGroup
    classes, features, lam, step = 100, 5000, 0.05, 0.02
    t = _table(arena, features, classes, 16, False, {"name": "add", "params": {"init": "gaussian", "std": 0.05, "seed": 7}})
    hb = synthetic_batch(128, features, classes, 9, seed=3)
    W = t.to_dense().clone()
    if mode == "pull":
        m = MLRDeviceModel(t, classes, step, lam=lam)
        stepf = lambda: m.step(hb.to("cuda"))              # noqa: E731
        fold = m.fold_scale
    else:
        m = MLRShipModel(t, arena, PeerGroup(), None, classes, step, 128, lam=lam)
        rb = route_batches(t, [hb], 0).to("cuda")
        stepf = lambda: m.step(rb)                         # noqa: E731
        fold = m.renormalize
    for _ in range(5):
        stepf()
        W, _, _ = mlr_reference_step(W, hb.to("cuda"), classes, step, lam)
    torch.cuda.synchronize()
    assert m.scale < 1.0                                   # nothing was folded yet: the table holds v = w / s
    fold()
    torch.testing.assert_close(t.to_dense(), W, rtol=2e-3, atol=2e-5)
    t.close()


def test_bulk_migration_copy_snapshot_and_drain(arena):
    """K10: one bulk-TMA launch copies a LIST of slabs to their new place + a snapshot; the drain forwards what
    landed in the old slabs afterwards (new += old - snapshot)."""
    from harmony_b200.ops import native
    n = native()
    nb, elems = 5, (3 << 20) // 4 + 4096                   # not a multiple of the 48 KB chunk: exercises the tail
    old = [torch.randn(elems, device="cuda") for _ in range(nb)]
    new = [torch.zeros(elems, device="cuda") for _ in range(nb)]
    snap = [torch.zeros(elems, device="cuda") for _ in range(nb)]
    i64 = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device="cuda")   # noqa: E731
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    n.migrate_bulk(i64(old), i64(new), i64(snap), elems * 4, err)
    torch.cuda.synchronize()
    assert int(err) == 0
    for o, a, b in zip(old, new, snap):
        assert torch.equal(o, a) and torch.equal(o, b)
    ref = [a.clone() for a in new]
    for k, o in enumerate(old):                            # late updates on the old slabs + fresh ones on the new
        o[k::97] += 1.0
        new[k][5::101] += 2.0
        ref[k][k::97] += 1.0
        ref[k][5::101] += 2.0
    fw = torch.zeros(1, dtype=torch.int64, device="cuda")
    n.migrate_drain(i64(old), i64(snap), i64(new), elems * 4, False, fw)
    torch.cuda.synchronize()
    for a, r in zip(new, ref):
        torch.testing.assert_close(a, r, rtol=0, atol=1e-6)
    assert int(fw) == sum(len(range(k, elems, 97)) for k in range(nb))
    # bounded update functions (NMF's SGD + clamp): the forwarded delta is projected back into [lo, hi]
    o, sn, nw = torch.full((1024,), 0.9, device="cuda"), torch.full((1024,), 0.2, device="cuda"), \
        torch.full((1024,), 0.6, device="cuda")
    fw.zero_()
    n.migrate_drain(i64([o]), i64([sn]), i64([nw]), 4096, False, fw, True, 0.0, 1.0)      # 0.6 + 0.7 -> clamp at 1.0
    torch.cuda.synchronize()
    assert int(fw) == 1024 and torch.allclose(nw, torch.ones_like(nw))


def test_key_guard_clamps_and_reports(arena):
    t = _table(arena, 1000, 32, 8, True, {"name": "add", "params": {}})
    t.push(torch.tensor([5, 10 ** 9, -3], device="cuda"), torch.ones((3, 32), device="cuda"))
    torch.cuda.synchronize()
    with pytest.raises(IndexError, match="out-of-range"):
        t.check()
    t.check()                                              # the guard words were reset
    d = t.to_dense()
    assert float(d[5].sum()) == 32.0 and float(d[0].sum()) == 64.0     # bad keys were clamped to key 0
    t.close()


def test_int32_table_put_rows_is_bit_exact(arena):
    """Checkpoint restore of the LDA count table: ASSIGN must not convert counts to float bits."""
    t = _table(arena, 500, 17, 8, False, {"name": "add", "params": {}}, dtype="int32")
    keys = torch.arange(0, 500, 3, device="cuda")
    vals = torch.randint(0, 1000, (keys.numel(), 17), device="cuda", dtype=torch.int32)
    t.put_rows(keys, vals)
    got = t.pull(keys)
    assert got.dtype == torch.int32 and torch.equal(got, vals)
    t.put_rows(keys, vals.float())                         # float input is rounded, not reinterpreted
    assert torch.equal(t.pull(keys), vals)
    t.close()


def test_scale_local_is_one_launch(arena):
    from harmony_b200 import ops
    t = _table(arena, 4000, 64, 16, True, {"name": "add", "params": {"init": "gaussian", "std": 1.0, "seed": 1}})
    W0 = t.to_dense().clone()
    ops.reset_launches()
    t.scale_local(0.5)
    assert ops.launches() == 1
    torch.testing.assert_close(t.to_dense(), W0 * 0.5)
    t.close()


@pytest.mark.parametrize("two_cta", [False, True, "stream_k"])
@pytest.mark.parametrize("B,F,C", [(128, 256, 128), (300, 1024, 1000), (4096, 4096, 1000)])
def test_dense_logits_tcgen05_matches_fp32_matmul(arena, B, F, C, two_cta):
    """tcgen05 (kind::tf32) GEMM whose B operand is TMA-loaded from the table's block slabs."""
    from harmony_b200.models.mlr_dense import DenseMLR
    t = _table(arena, F, C, max(1, F // 256), True, {"name": "add", "params": {"init": "gaussian", "std": 0.05, "seed": 3}})
    assert t.rows_per_block % 32 == 0
    X = torch.randn((B, F), device="cuda")
    W = t.to_dense()
    out = DenseMLR(t, C, use_2cta=bool(two_cta), stream_k=two_cta == "stream_k").logits(X)
    torch.cuda.synchronize()
    ref = X.double() @ W.double()
    err = (out[:, :C].double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err < 2e-3 * scale + 1e-3, (err, scale)       # tf32 inputs (10-bit mantissa), fp32 accumulate
    assert float(out[:, C:].abs().max()) == 0.0 if out.shape[1] > C else True
    t.close()


@pytest.mark.parametrize("two_cta", [False, True, "stream_k"])
@pytest.mark.parametrize("B,F,C", [(64, 256, 100), (500, 2048, 1000)])
def test_dense_mlr_step_tcgen05_matches_reference(arena, B, F, C, two_cta):
    """Forward GEMM + softmax + backward GEMM whose epilogue reduce-adds into the table rows."""
    from harmony_b200.models.mlr_dense import DenseMLR
    t = _table(arena, F, C, max(1, F // 256), True, {"name": "add", "params": {"init": "gaussian", "std": 0.05, "seed": 3}})
    X = torch.randn((B, F), device="cuda") * 0.5
    y = torch.randint(0, C, (B,), device="cuda", dtype=torch.int32)
    W0 = t.to_dense().clone()
    m = DenseMLR(t, C, use_2cta=bool(two_cta), stream_k=two_cta == "stream_k")
    m.step(X, y, 0.01)
    torch.cuda.synchronize()
    logits = X.double() @ W0.double()
    p = torch.softmax(logits, dim=1)
    loss_ref = float(-torch.log(p[torch.arange(B), y.long()]).sum())
    p[torch.arange(B), y.long()] -= 1.0
    Wref = W0.double() - 0.01 * (X.double().t() @ p)
    err = float((t.to_dense().double() - Wref).abs().max())
    upd = float((Wref - W0.double()).abs().max())
    assert err < 5e-3 * upd + 1e-6, (err, upd)
    assert abs(float(m.stats[0]) - loss_ref) / loss_ref < 5e-3
    t.close()


def test_sm_share_budget_is_per_thread_and_kernels_stay_correct(arena):
    """GPU twin of the resource-typed task-unit slots: a tasklet thread launches for its job's SHARE of the SMs."""
    import threading
    from harmony_b200.ops import native
    n = native()
    full = n.sm_count()
    assert full >= 100 and n.get_sm_share() == 1.0
    seen = {}

    def other():
        n.set_sm_share(0.25)
        seen["other"] = n.sm_count()
    th = threading.Thread(target=other); th.start(); th.join()
    assert seen["other"] == round(full * 0.25) and n.sm_count() == full        # thread-local
    try:
        n.set_sm_share(0.25)
        t = _table(arena, 3000, 64, 16, False, {"name": "add", "params": {"init": "gaussian", "std": 0.1, "seed": 2}})
        W0 = t.to_dense().clone()
        keys = torch.randint(0, 3000, (5000,), device="cuda")
        d = torch.randn((5000, 64), device="cuda")
        t.push(keys, d)
        torch.testing.assert_close(t.to_dense(), W0.clone().index_add_(0, keys, d), rtol=1e-4, atol=1e-4)
        t.close()
    finally:
        n.set_sm_share(1.0)
