"""Host-side logic of the SPMD paths, testable without a GPU: ship-mode routing (which owner sees which non-zero),
the PeerGroup bootstrap over gloo with 2 processes, and the control-plane rendezvous group."""
import os
import subprocess
import sys
import textwrap
from types import SimpleNamespace

import numpy as np
import pytest

from harmony_b200.et.partitioner import PART_HASHED, PART_ORDERED
from harmony_b200.models.mlr import synthetic_batch
from harmony_b200.models.mlr_ship import owner_of_features, route_batches

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_table(mode, F, world, num_blocks):
    rpb = -(-F // num_blocks)
    owners = [b % world for b in range(num_blocks)]
    return SimpleNamespace(mode=mode, rows_per_block=rpb, num_blocks=num_blocks, owners=owners)


@pytest.mark.parametrize("mode", [PART_ORDERED, PART_HASHED])
@pytest.mark.parametrize("world", [2, 3])
def test_ship_routing_covers_every_nonzero_exactly_once(mode, world):
    """Sum over owners of the partial logits of the routed slices == X . W of the original batches: every non-zero
    goes to exactly one owner, rows keep their (worker, instance) identity, values are untouched."""
    F, C, B, nnz = 500, 7, 33, 9
    table = _stub_table(mode, F, world, 16)
    rng = np.random.default_rng(0)
    W = rng.standard_normal((F, C))
    batches = [synthetic_batch(B, F, C, nnz, seed=10 + w) for w in range(world)]
    full = np.zeros((world * B, C))
    for w, mb in enumerate(batches):
        ptr, col, val = mb.csr_ptr.numpy(), mb.csr_col.numpy(), mb.csr_val.numpy()
        rows = np.repeat(np.arange(B), np.diff(ptr))
        np.add.at(full, w * B + rows, val[:, None] * W[col])
    partial_sum = np.zeros_like(full)
    total_nnz = 0
    for rank in range(world):
        rb = route_batches(table, batches, rank)
        r = rb.routed
        ptr, col, val = r.csr_ptr.numpy(), r.csr_col.numpy(), r.csr_val.numpy()
        assert len(ptr) == world * B + 1 and ptr[0] == 0 and ptr[-1] == len(col)
        assert np.all(owner_of_features(table, col) == rank)                 # only features this rank owns
        rows = np.repeat(np.arange(world * B), np.diff(ptr))
        np.add.at(partial_sum, rows, val[:, None] * W[col])
        total_nnz += len(col)
        np.testing.assert_array_equal(rb.labels.numpy(), batches[rank].labels.numpy())
        # feature-major grouping of the routed slice (what the owner-side update kernel walks)
        fp, uf = r.feat_ptr.numpy(), r.uniq_feat.numpy()
        assert fp[-1] == len(col) and np.all(np.diff(uf) > 0)
        for j in (0, len(uf) // 2, len(uf) - 1):
            members = np.sort(r.samp.numpy()[fp[j]:fp[j + 1]])
            np.testing.assert_array_equal(members, np.sort(rows[col == uf[j]]))
    assert total_nnz == sum(len(mb.csr_col) for mb in batches)
    np.testing.assert_allclose(partial_sum, full, rtol=1e-6, atol=1e-6)
    # executor-id owners (job-server mode) are mapped through owner_ranks
    table.owners = [f"executor-{b % world}" for b in range(16)]
    ranks = [b % world for b in range(16)]
    rb = route_batches(table, batches, 0, ranks)
    assert np.all(owner_of_features(table, rb.routed.csr_col.numpy(), ranks) == 0)


def test_peer_group_bootstrap_over_gloo_two_processes(tmp_path):
    """`PeerGroup.from_env()` is the only use of torch.distributed in the SPMD path (handle exchange, barriers):
    run it with the gloo backend on 2 CPU processes (127.0.0.1 rendezvous)."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        dist.init_process_group("gloo")
        from harmony_b200.parallel.peer import PeerGroup
        g = PeerGroup.from_env()
        assert (g.rank, g.world) == (int(os.environ["RANK"]), 2)
        got = g.all_gather_object({{"rank": g.rank, "handle": bytes([g.rank]) * 4}})
        assert [x["rank"] for x in got] == [0, 1] and got[1]["handle"] == b"\\x01" * 4
        g.barrier()
        dist.destroy_process_group()
        print("ok", g.rank)
    """))
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert sorted(o.strip().splitlines()[-1] for o in outs) == ["ok 0", "ok 1"]


def test_device_batch_builders_index_the_staging_buffers_consistently():
    """Host preprocessing of the NMF / LDA / MLR device steps: slots index the pulled staging rows, the last
    vocabulary entry is the LDA summary key, the MLR by-feature grouping is a permutation of the CSR non-zeros."""
    from harmony_b200.models.lda import build_lda_batch
    from harmony_b200.models.nmf import build_nmf_batch
    rng = np.random.default_rng(3)
    # NMF: ucols[col_slot] reproduces the original columns; ucols sorted & distinct
    ptr = np.array([0, 3, 3, 7]); cols = rng.integers(0, 50, 7); vals = rng.random(7).astype(np.float32)
    b = build_nmf_batch(ptr, cols, vals, np.array([5, 6, 7]))
    np.testing.assert_array_equal(b.ucols.numpy()[b.col_slot.numpy()], cols)
    assert np.all(np.diff(b.ucols.numpy()) > 0) and b.csr_ptr.dtype.is_floating_point is False
    # LDA: vocab = distinct words + [V]; tok_slot maps every token to its word; token window addresses z
    docs = [rng.integers(0, 30, n).astype(np.int32) for n in (4, 0, 9)]
    lb = build_lda_batch(docs, tok_base=100, V=30)
    words = np.concatenate(docs)
    v = lb.vocab.numpy()
    assert v[-1] == 30 and np.all(np.diff(v[:-1]) > 0)
    np.testing.assert_array_equal(v[lb.tok_slot.numpy()], words)
    assert (lb.tok_lo, lb.tok_hi) == (100, 113) and lb.doc_ptr.tolist() == [0, 4, 4, 13]
    # MLR: (samp, uniq_feat[group], xval) is a permutation of (row, col, val)
    mb = synthetic_batch(17, 200, 5, 6, seed=1)
    rows = np.repeat(np.arange(17), np.diff(mb.csr_ptr.numpy()))
    fp = mb.feat_ptr.numpy()
    feat_of = np.repeat(mb.uniq_feat.numpy(), np.diff(fp))
    a = sorted(zip(rows.tolist(), mb.csr_col.numpy().tolist(), mb.csr_val.numpy().tolist()))
    c = sorted(zip(mb.samp.numpy().tolist(), feat_of.tolist(), mb.xval.numpy().tolist()))
    assert a == c


def test_load_batches_splits_a_file_exactly_over_ranks(tmp_path):
    from harmony_b200.common.samples import write_mlr_sample
    from harmony_b200.models.mlr import load_batches
    path = write_mlr_sample(str(tmp_path / "d"), n=333, features=64, classes=4, nnz=6, seed=5)
    one = load_batches(path, 50)
    assert [b.batch_size for b in one] == [50] * 6 + [33]
    per_rank = [load_batches(path, 50, r, 3, num_classes=4, num_features=64) for r in range(3)]
    assert sum(b.batch_size for bs in per_rank for b in bs) == 333            # every line owned by exactly one rank
    all_vals = np.sort(np.concatenate([b.csr_val.numpy() for bs in per_rank for b in bs]))
    np.testing.assert_array_equal(all_vals, np.sort(np.concatenate([b.csr_val.numpy() for b in one])))
    assert all(b.batch_size == 50 for b in load_batches(path, 50, drop_last=True))
    with pytest.raises(ValueError):
        load_batches(path, 50, num_features=32)                                 # feature ids reach 63


class _ThreadGroup:
    """PeerGroup stand-in for N threads of one process (all_gather_object through a barrier)."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self.pg, self._s = rank, world, None, shared

    def all_gather_object(self, obj):
        import threading
        s = self._s
        with s["lock"]:
            s["slots"][self.rank] = obj
        s["bar"].wait(timeout=20)
        out = list(s["slots"])
        s["bar"].wait(timeout=20)          # nobody overwrites a slot before everyone has read
        return out

    def barrier(self):
        self.all_gather_object(None)


@pytest.mark.parametrize("fail_stage", [None, "add_device", "bind", "map"])
def test_symmetric_region_downgrades_all_ranks_together_when_multicast_setup_fails(fail_stage):
    """The NVLS set-up (cuMulticastCreate -> fd -> add device -> bind -> map) votes after every stage: a failure on ONE
    rank must leave EVERY rank on the unicast path -- nobody may be left waiting inside a bind.  (Driver calls are
    faked; the fd really travels through the FdServer.)"""
    import threading
    from harmony_b200.parallel.peer import SymmetricRegion
    world = 2
    shared = {"lock": threading.Lock(), "slots": [None] * world, "bar": threading.Barrier(world)}

    class FakeNative:
        def __init__(self, rank):
            self.rank = rank

        def vmm_supported(self, dev): return [True, True]
        def mc_granularity(self, w, b, rec): return 1 << 21

        def mc_create(self, dev, w, nbytes):
            return 7, os.open(os.devnull, os.O_RDONLY)

        def mc_import(self, fd, dev, w, nbytes):
            os.fstat(fd)                       # the fd that arrived over the unix socket is a live descriptor
            return 8

        def _maybe_fail(self, stage):
            if fail_stage == stage and self.rank == 1:
                raise RuntimeError(f"injected {stage} failure")

        def mc_add_device(self, mc, dev): self._maybe_fail("add_device")
        def mc_bind(self, mc, off, seg, moff, nbytes): self._maybe_fail("bind")

        def mc_map(self, mc, dev):
            self._maybe_fail("map")
            return 0xABC00000 + self.rank

        def close_fd(self, fd): os.close(fd)

    class FakeArena:
        kind = "vmm"

        def __init__(self, rank):
            self.device, self._n, self.rank = rank, FakeNative(rank), rank

        def new_segment(self, nbytes, align): return 0
        def segment_base(self, seg): return 0x10000000 * (self.rank + 1)
        def alloc_in(self, seg, nbytes): return self.segment_base(seg), seg, 0
        def segment_info(self, seg): return {"kind": "vmm", "base": self.segment_base(seg)}
        def map_peer(self, r, si, h): pass
        def resolve(self, r, si, off): return 0x10000000 * (r + 1) + off

    regions, errors = [None] * world, []

    def run(rank):
        try:
            regions[rank] = SymmetricRegion(FakeArena(rank), _ThreadGroup(rank, world, shared), 1000)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ths]
    [t.join(30) for t in ths]
    assert not errors and all(r is not None for r in regions), errors
    assert [r.peer_ptrs for r in regions] == [[0x10000000, 0x20000000]] * 2
    if fail_stage is None:
        assert [r.mc_ptr for r in regions] == [0xABC00000, 0xABC00001]
    else:
        assert [r.mc_ptr for r in regions] == [0, 0]             # both downgraded, nobody hung
        assert "injected" in regions[1].mc_error


def test_lazy_l2_scale_algebra_matches_the_dense_decay():
    """The lazy L2 decay (models/mlr.py, models/mlr_ship.py): the table stores v with w = s v; one step is
    s' = s (1 - c), v' = v + delta / s' and must equal the reference's W' = (1 - c) W + delta
    (MLRTrainer.java:387-397) -- also across a fold (v <- s v, s <- 1)."""
    rng = np.random.default_rng(0)
    W = rng.standard_normal((50, 7))
    v, s = W.copy(), 1.0
    c = 1e-2 * 0.5 * 64
    for step in range(30):
        delta = 1e-2 * rng.standard_normal(W.shape)            # what the push of this step adds (pre-scaled by -step)
        W = (1 - c) * W + delta
        s_next = s * (1 - c)
        v = v + delta / s_next
        s = s_next
        if s < 1e-3 or step == 11:                             # fold_scale / renormalize
            v, s = s * v, 1.0
        np.testing.assert_allclose(s * v, W, rtol=1e-9, atol=1e-12)
