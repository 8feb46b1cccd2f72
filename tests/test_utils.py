"""Utility layer (reference: utils/src/test DAGImplTest, LongRangeUtilsTest; StateMachine; Tang flags)."""
import json
import os

import numpy as np
import pytest

from harmony_b200.utils import DAG, Configuration, FlagError, IllegalStateError, Param, ParamRegistry, StateMachine
from harmony_b200.utils import trace
from harmony_b200.utils.codecs import (DenseVectorCodec, IntegerCodec, LongCodec, SerializableCodec, SparseArrayCodec,
                                       SparseVectorCodec)
from harmony_b200.utils.misc import generate_dense_long_ranges, get_num_units, murmur3_32


def test_state_machine_transitions():
    sm = (StateMachine.new_builder().add_state("A").add_state("B").add_state("C").add_transition("A", "B")
          .add_transition("B", "C").set_initial_state("A").build())
    sm.check_state("A")
    with pytest.raises(IllegalStateError):
        sm.set_state("C")
    assert sm.compare_and_set_state("A", "B") and not sm.compare_and_set_state("A", "B")
    sm.set_state("C")
    with pytest.raises(IllegalStateError):
        sm.check_state("A")
    with pytest.raises(ValueError):
        StateMachine.new_builder().add_state("A").add_state("A")


def test_dag_roots_degrees_cycles():
    d = DAG()
    for v in "abcd":
        assert d.add_vertex(v)
    assert not d.add_vertex("a")
    d.add_edge("a", "b"); d.add_edge("a", "c"); d.add_edge("b", "d"); d.add_edge("c", "d")
    assert d.get_root_vertices() == {"a"} and d.get_in_degree("d") == 2 and d.is_adjacent("a", "b")
    with pytest.raises(ValueError):
        d.add_edge("d", "a")
    d.remove_vertex("a")
    assert d.get_root_vertices() == {"b", "c"} and d.get_num_vertices() == 3
    d.remove_edge("b", "d")
    assert d.get_in_degree("d") == 1
    assert set(json.loads(d.to_json())["vertices"]) == {"b", "c", "d"}


def test_flags_parse_merge_serialize():
    reg = ParamRegistry([Param("num_executors", int, 1, group="driver"), Param("local", bool, True, group="client"),
                         Param("step_size", float, None, group="user", required=False), Param("input", str, None)])
    conf, unknown = reg.parse(["-num_executors", "5", "-local", "false", "-step_size", "0.1", "-foo", "bar"],
                              allow_unknown=True)
    assert conf["num_executors"] == 5 and conf["local"] is False and conf["step_size"] == 0.1 and unknown == ["-foo", "bar"]
    with pytest.raises(FlagError):
        reg.parse(["-foo", "bar"])
    with pytest.raises(FlagError):
        reg.parse(["-num_executors"])
    back = Configuration.deserialize(conf.serialize())
    assert dict(back) == dict(conf)
    assert dict(conf.extract(reg.by_group("driver"))) == {"num_executors": 5}
    assert conf.merge({"x": 1}).with_values(y=2)["y"] == 2


def test_long_ranges_and_murmur():
    r = generate_dense_long_ranges([5, 1, 2, 3, 9, 10])
    assert r == [(1, 3), (5, 5), (9, 10)] and get_num_units(r) == 6
    assert murmur3_32(b"") == 0 and murmur3_32(b"hello") == 0x248bfa47 and murmur3_32(b"hello", 1) != murmur3_32(b"hello")


def test_codecs_roundtrip():
    assert IntegerCodec().decode(IntegerCodec().encode(-7)) == -7
    assert LongCodec().decode(LongCodec().encode(2 ** 40)) == 2 ** 40
    v = np.arange(5, dtype=np.float32) * 0.5
    np.testing.assert_array_equal(DenseVectorCodec().decode(DenseVectorCodec().encode(v)), v)
    n, i, x = SparseVectorCodec().decode(SparseVectorCodec().encode((100, [3, 50], [1.5, -2.0])))
    assert n == 100 and list(i) == [3, 50] and list(x) == [1.5, -2.0]
    row = np.array([0, 4, 0, 2, 2], dtype=np.int32)
    np.testing.assert_array_equal(SparseArrayCodec(4).decode(SparseArrayCodec(4).encode(row)), row)
    assert SerializableCodec().decode(SerializableCodec().encode({"a": [1, 2]})) == {"a": [1, 2]}


def test_trace_spans_dump(tmp_path):
    trace.enable(True); trace.clear()
    with trace.span("PULL", job="j"):
        with trace.span("COMP"):
            pass
    p = tmp_path / "t.json"
    assert trace.dump(str(p)) == 2
    ev = json.loads(p.read_text())["traceEvents"]
    assert {e["name"] for e in ev} == {"PULL", "COMP"} and all(e["ph"] == "X" for e in ev)
    trace.enable(False)
    t = trace.Tracer(); t.start_timer(); t.record_time(2)
    assert t.total_elapsed_time() >= 0 and t.avg_time_per_elem() >= 0


def test_linalg_vector_matrix_api():
    import torch
    from harmony_b200.common.linalg import MatrixFactory, VectorFactory
    vf, mf = VectorFactory(), MatrixFactory()
    a, b = vf.create_dense([1, 2, 3]), vf.create_dense_ones(3)
    assert a.dot(b) == 6 and a.add(b).t.tolist() == [2, 3, 4] and a.sub(1).t.tolist() == [0, 1, 2]
    s = vf.create_sparse([0, 2], [5.0, -1.0], 3)
    assert not s.is_dense() and s.active_size() == 2 and a.dot(s) == 2 and s.dot(a) == 2
    a.axpy(2.0, s)
    assert a.t.tolist() == [11, 2, 1] and a.slice(1, 3).t.tolist() == [2, 1]
    assert vf.concat_dense([b, b]).length() == 6 and a.scale(2).div(2).t.tolist() == [11, 2, 1]
    m = mf.create_dense([1, 2, 3, 4, 5, 6], 2, 3)            # column-major
    assert m.get(1, 0) == 2 and m.rows() == 2 and m.columns() == 3
    assert m.mmul(vf.create_dense([1, 1, 1])).t.tolist() == [9, 12]
    assert m.transpose().mmul(m).t.shape == (3, 3) and m.slice_column(2).t.tolist() == [5, 6]
    sp = mf.horzcat_vec_sparse([s, s])
    assert not sp.is_dense() and sp.to_dense().t.tolist() == [[5, 5], [0, 0], [-1, -1]]
    assert sp.transpose().mmul(vf.create_dense([1, 1, 1])).t.tolist() == [4, 4]
    m.put_row(0, vf.create_dense([0, 0, 0])); m.addi(1).muli(2)
    assert m.t.tolist() == [[2, 2, 2], [6, 10, 14]]


def test_sample_generator_writes_every_launch_script_input(tmp_path):
    """bin/make_samples.sh equivalents of the reference's sample files, parseable by the registered parsers."""
    from harmony_b200.common.dataloader import make_parser
    from harmony_b200.common.samples import write_all
    import harmony_b200.mlapps.gbt, harmony_b200.mlapps.lasso, harmony_b200.mlapps.lda  # noqa: F401,E401
    import harmony_b200.mlapps.mlr, harmony_b200.mlapps.nmf, harmony_b200.pregel.graph  # noqa: F401,E401
    d = write_all(str(tmp_path))
    for name, parser, n in [("sample_mlr", "mlr", 1000), ("sample_mlr_test", "mlr", 200), ("sample_nmf", "nmf", 250),
                            ("sample_lda", "lda", 200), ("sample_lasso", "lasso", 256), ("sample_gbt", "gbt", 160),
                            ("inputs/adj_list", "graph_none_edge_value", 40), ("inputs/shortest_path", "graph", 40)]:
        p = make_parser(parser)
        with open(os.path.join(d, name)) as f:
            recs = [p.parse_line(ln) for ln in f if ln.strip() and not ln.startswith("#")]
        assert len(recs) == n, name
    from harmony_b200.mlapps.gbt import read_metadata
    types, ytype = read_metadata(os.path.join(d, "sample_gbt.meta"), 784)
    assert ytype == 0 and len(types) == 784
    from harmony_b200.dolphin.optimizer import parse_host_to_value_file
    assert parse_host_to_value_file(os.path.join(d, "sample_host_to_bandwidth"))["localhost"] > 1e12


def test_native_loader_matches_python_parsers(tmp_path):
    """ops/csrc/loader.cpp (multi-threaded text parsers, counting-sort batch grouping) against the Python / NumPy
    twins; skipped when the extension has not been built."""
    from harmony_b200.ops import native_or_none
    nat = native_or_none()
    if nat is None:
        pytest.skip("native extension not built")
    from harmony_b200.common import dataloader as dl
    from harmony_b200.common.samples import write_all
    d = write_all(str(tmp_path))
    for name, parser in [("sample_mlr", "mlr"), ("sample_nmf", "nmf"), ("sample_lda", "lda"),
                         ("sample_lasso", "lasso"), ("sample_gbt", "gbt")]:
        p = dl.make_parser(parser)
        data = open(os.path.join(d, name), "rb").read()
        assert len(data) >= dl.NATIVE_PARSE_MIN_BYTES
        fast = p.parse_bytes(data)
        slow = p.parse(data.decode().splitlines())
        assert len(fast) == len(slow) > 0, name
        for a, b in zip(fast, slow):
            a = a if isinstance(a, tuple) else (a,)
            b = b if isinstance(b, tuple) else (b,)
            for x, y in zip(a, b):
                np.testing.assert_array_equal(np.asarray(x), np.asarray(y), err_msg=name)
    # byte-range splits: every line lands in exactly one split, same as the line reader
    path = os.path.join(d, "sample_mlr")
    for k in (1, 3, 7):
        splits = dl.get_splits(path, k)
        by_bytes = b"".join(dl.read_split_bytes(s) for s in splits).decode().splitlines()
        by_lines = [ln for s in splits for ln in dl.read_split(s)]
        assert by_bytes == by_lines and len(by_lines) == 1001          # 1000 records + the comment line
    with pytest.raises(ValueError):
        dl.make_parser("nmf").parse_bytes(b"0: 1,2\n" * 1000)           # row index must be positive
    # feature grouping == stable argsort
    import torch
    from harmony_b200.models.mlr import synthetic_batch
    b = synthetic_batch(257, 5000, 10, 13, seed=4)
    col, val, ptr = b.csr_col.numpy(), b.csr_val.numpy(), b.csr_ptr.numpy()
    order = np.argsort(col, kind="stable")
    rows = np.repeat(np.arange(257, dtype=np.int32), np.diff(ptr))
    uniq, fptr, samp, xval = nat.group_by_feature(torch.from_numpy(ptr), torch.from_numpy(col), torch.from_numpy(val))
    np.testing.assert_array_equal(samp.numpy(), rows[order])
    np.testing.assert_array_equal(xval.numpy(), val[order])
    np.testing.assert_array_equal(uniq.numpy(), np.unique(col))
    assert fptr[0] == 0 and fptr[-1] == len(col) and np.all(np.diff(fptr.numpy()) > 0)


def test_unified_cli_and_flag_reference(tmp_path, capsys):
    """`python -m harmony_b200 ...`: sample generation, flag listing (docs/FLAGS.md is generated from it) and an
    ET example through the same dispatcher the bin/ scripts use."""
    from harmony_b200.__main__ import all_flags, flags_markdown, main
    assert main(["samples", str(tmp_path)]) == 0 and os.path.exists(tmp_path / "sample_mlr")
    names = [p.name for _, p in all_flags()]
    for must in ("num_executors", "scheduler", "max_num_epochs", "num_mini_batches", "clock_slack", "optimizer",
                 "features_per_partition", "num_topics", "max_depth_of_tree", "num_pregel_workers", "mlr_mode"):
        assert must in names
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert open(os.path.join(root, "docs", "FLAGS.md")).read() == flags_markdown()     # regenerate when flags change
    assert main(["et-example", "simple"]) == 0
    assert main(["run", "nosuchapp"]) == 2 and main(["frobnicate"]) == 2


def test_bench_reference_arm_reports_unavailable():
    """bench.py --impl reference: the reference is Java/REEF and cannot be installed here; the arm must say so in
    one JSON line and exit 0 (driver contract)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "3", "--warmup", "3"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and isinstance(line["unavailable"], str) and line["unavailable"]


def test_native_parser_fuzz_against_python():
    """Random well-formed inputs with awkward spacing / number formats (tabs, CRLF, `+`, exponents, blank and comment
    lines, very long lines) through the multi-threaded native parser and the Python parsers."""
    import random
    from harmony_b200.ops import native_or_none
    if native_or_none() is None:
        pytest.skip("native extension not built")
    from harmony_b200.common import dataloader as dl
    rng = random.Random(0)

    def num(x):
        return rng.choice([f"{x}", f"{x:.3f}", f"{x:.2e}", f"+{x:.4f}" if x >= 0 else f"{x:.4f}"])
    sep = lambda: rng.choice([" ", "  ", "\t", " \t "])  # noqa: E731
    eol = lambda: rng.choice(["\n", "\r\n", " \n", "\n\n", "\n# c\n"])  # noqa: E731
    for name, gen in (
        ("mlr", lambda: str(rng.randint(0, 9)) + "".join(sep() + f"{rng.randint(0, 999)}:{num(rng.uniform(-5, 5))}"
                                                          for _ in range(rng.randint(0, 40)))),
        ("lasso", lambda: num(rng.uniform(-9, 9)) + "".join(sep() + f"{rng.randint(0, 99)}:{num(rng.uniform(-5, 5))}"
                                                             for _ in range(rng.randint(1, 300)))),
        ("nmf", lambda: f"{rng.randint(1, 500)}:" + "".join(sep() + f"{rng.randint(1, 99)},{num(rng.uniform(0, 5))}"
                                                           for _ in range(rng.randint(1, 30)))),
        ("lda", lambda: sep().join(str(rng.randint(0, 9999)) for _ in range(rng.randint(1, 120)))),
    ):
        text = "".join(gen() + eol() for _ in range(600))
        data = text.encode()
        p = dl.make_parser(name)
        fast, slow = p.parse_bytes(data), p.parse(text.splitlines())
        assert len(fast) == len(slow) == 600, name
        for a, b in zip(fast, slow):
            a = a if isinstance(a, tuple) else (a,)
            b = b if isinstance(b, tuple) else (b,)
            for x, y in zip(a, b):
                np.testing.assert_allclose(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64),
                                           rtol=1e-6, atol=0, err_msg=name)


def test_split_reader_fuzz_every_line_in_exactly_one_split(tmp_path):
    """Random files (with / without trailing newline, empty lines, one huge line) cut into 1..40 byte-range splits:
    the concatenation of the splits' lines is the file's lines, for the line reader and the byte reader alike."""
    import random
    from harmony_b200.common import dataloader as dl
    rng = random.Random(1)
    for case in range(40):
        lines = ["x" * rng.choice([0, 1, 5, 40, 400]) + str(i) for i in range(rng.randint(1, 120))]
        if case % 7 == 0:
            lines[rng.randrange(len(lines))] = "y" * 5000
        body = "\n".join(lines) + ("\n" if case % 2 else "")
        path = tmp_path / f"f{case}"
        path.write_text(body)
        want = body.split("\n")
        if want and want[-1] == "":
            want = want[:-1]
        k = rng.randint(1, 40)
        splits = dl.get_splits(str(path), k)
        assert len(splits) == k
        by_lines = [ln for s in splits for ln in dl.read_split(s)]
        by_bytes = b"".join(dl.read_split_bytes(s) for s in splits).decode()
        by_bytes = by_bytes.split("\n")
        if by_bytes and by_bytes[-1] == "":
            by_bytes = by_bytes[:-1]
        assert by_lines == want, (case, k)
        assert by_bytes == want, (case, k)


def test_memory_metric_tracker_windows_share_one_sampler_thread():
    """Short windows (shorter than the period) still report a sample; longer ones collect several; the sampler thread is
    reused across windows and retires by itself."""
    import threading
    import time
    from harmony_b200.common.metric import MemoryMetricTracker
    t = MemoryMetricTracker(0.01)
    t.start()
    out = t.stop()                                   # empty window: falls back to one sample
    assert out["METRIC_MEMORY_MAX_MB"] > 0 and out["METRIC_MEMORY_AVG_MB"] <= out["METRIC_MEMORY_MAX_MB"]
    first = t._thread
    for _ in range(3):
        t.start()
        time.sleep(0.05)
        assert t._thread is first                    # no thread churn per window
        out = t.stop()
        assert out["METRIC_MEMORY_MAX_MB"] > 0
    deadline = time.time() + 2
    while t._thread is not None and time.time() < deadline:
        time.sleep(0.01)
    assert t._thread is None                         # idle for a few periods: retired
    t.start()
    assert t._thread is not None and t._thread is not first
    t.stop()
    t.close()
    assert not any(th.name == "memory-metric" and th.is_alive() and th is first for th in threading.enumerate())


def test_parsers_resolve_in_a_fresh_executor_process():
    """An executor PROCESS has imported no app module when its first TableLoad arrives: every built-in parser name and
    a 'module:Class' user parser must resolve there; an unknown name says what is registered."""
    import subprocess
    import sys
    code = ("from harmony_b200.common.dataloader import make_parser\n"
            "for n in ('example', 'mlr', 'nmf', 'lda', 'lasso', 'gbt', 'graph'):\n"
            "    make_parser(n)\n"
            "p = make_parser('harmony_b200.mlapps.examples:ExampleDataParser')\n"
            "assert type(p).__name__ == 'ExampleDataParser'\n"
            "try:\n"
            "    make_parser('nope')\n"
            "except KeyError as e:\n"
            "    assert 'registered' in str(e)\n"
            "else:\n"
            "    raise SystemExit('unknown parser accepted')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=root), capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stderr


def test_fd_server_hands_out_duplicates_of_registered_fds():
    """VMM segments are shared between executor processes as POSIX fds over a unix socket (SCM_RIGHTS)."""
    import os
    from harmony_b200.parallel.peer import FdServer
    r, w = os.pipe()
    srv = FdServer.get()
    srv.register("pipe-w", lambda: os.dup(w))
    fd = FdServer.fetch(os.getpid(), "pipe-w")
    try:
        assert fd != w
        os.write(fd, b"hb")
        assert os.read(r, 2) == b"hb"
    finally:
        os.close(fd)
    with pytest.raises(KeyError):
        FdServer.fetch(os.getpid(), "no-such-allocation")
    srv.unregister("pipe-w")
    with pytest.raises(KeyError):
        FdServer.fetch(os.getpid(), "pipe-w")
    os.close(r); os.close(w)


def test_native_sm_share_is_thread_local():
    from harmony_b200.ops import native_or_none
    n = native_or_none()
    if n is None:
        pytest.skip("extension not built")
    import threading
    base = n.sm_count()
    out = {}
    th = threading.Thread(target=lambda: (n.set_sm_share(0.5), out.update(v=n.sm_count())))
    th.start(); th.join()
    assert out["v"] == round(base * 0.5) and n.sm_count() == base and n.get_sm_share() == 1.0
