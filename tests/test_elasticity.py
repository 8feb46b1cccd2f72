"""Elastic re-partitioning on a LIVE job (reference OwnershipFirstMigrationTest, PlanCompilerTest,
SampleOptimizersTest, WorkerStateManagerTest): AddVector self-validates that no push was lost
while servers / workers are added, deleted or switched mid-run."""
import os
import tempfile
import threading
import time

import pytest

from harmony_b200.dolphin.launcher import ETDolphinLauncher
from harmony_b200.dolphin.master import MasterSideMsgSender, WorkerStateManager
from harmony_b200.dolphin.optimizer import (EvaluatorParameters, HomogeneousOptimizer, generate_transfer_steps,
                                            make_optimizer)
from harmony_b200.dolphin.plan import NAMESPACE_SERVER, NAMESPACE_WORKER, Plan, PlanCompiler, TransferStep
from harmony_b200.et.config import TaskletConfiguration
from harmony_b200.mlapps import examples


def _args(optimizer, workers=2, servers=2, epochs=6):
    path = examples.write_dummy_input(48)
    return ["-input", path, "-num_dolphin_workers", str(workers), "-num_servers", str(servers),
            "-max_num_epochs", str(epochs), "-num_mini_batches", "8", "-num_worker_blocks", "8",
            "-num_server_blocks", "8", "-num_keys", "20", "-vector_size", "4", "-delta", "2",
            "-compute_time_ms", "40", "-optimizer", optimizer, "-optimization_interval_ms", "300",
            "-delay_after_optimization_ms", "100", "-min_num_required_batch_metrics", "1",
            "-num_extra_resources", "1"]


@pytest.mark.parametrize("optimizer", ["EmptyPlanOptimizer", "DeleteOneServerOptimizer", "AddOneServerOptimizer",
                                       "MoveModelBlocksOptimizer", "DeleteOneWorkerOptimizer", "AddOneWorkerOptimizer",
                                       "ExchangeOneOptimizer"])
def test_addvector_validates_under_server_reconfiguration(optimizer):
    """Servers AND workers are added / deleted / switched on the live job; AddVector checks in its cleanup that every
    granted mini-batch was applied exactly once (a stopped worker finishes the batch in flight and hands back a
    batch it was granted but did not start; a worker started by a plan gets the master's messages from its first
    sync on)."""
    res = ETDolphinLauncher.launch(examples.ADD_VECTOR, _args(optimizer))
    assert res.ok, res.messages
    if optimizer == "DeleteOneWorkerOptimizer":
        assert any("(S: 2, W: 2) -> (S: 2, W: 1)" in m for m in res.messages), res.messages
    if optimizer == "AddOneWorkerOptimizer":
        assert any("(S: 2, W: 2) -> (S: 2, W: 3)" in m for m in res.messages), res.messages
    if optimizer == "ExchangeOneOptimizer":
        assert any("(S: 2, W: 2) -> (S: 1, W: 3)" in m for m in res.messages), res.messages
    if optimizer == "DeleteOneServerOptimizer":
        assert any("(S: 2, W: 2) -> (S: 1, W: 2)" in m for m in res.messages), res.messages
    if optimizer == "AddOneServerOptimizer":
        assert any("(S: 2, W: 2) -> (S: 3, W: 2)" in m for m in res.messages), res.messages


@pytest.mark.parametrize("optimizer", ["DeleteOneWorkerOptimizer", "AddOneWorkerOptimizer", "ExchangeOneOptimizer"])
def test_addvector_with_minibatch_leases_stays_exact_under_worker_reconfiguration(optimizer):
    """-mini_batch_grant_size 3: a worker removed in the middle of a lease hands the unused part back, a worker whose
    blocks leave idles without holding credits, the tail of the budget is granted batch by batch -- AddVector's
    validation (every granted mini-batch applied exactly once, budget used up exactly) still holds."""
    res = ETDolphinLauncher.launch(examples.ADD_VECTOR, _args(optimizer, epochs=8) + ["-mini_batch_grant_size", "3"])
    assert res.ok, res.messages
    assert any("(S: 2, W: 2) -> " in m for m in res.messages), res.messages


@pytest.mark.parametrize("optimizer", ["DeleteOneWorkerOptimizer", "AddOneWorkerOptimizer", "ExchangeOneOptimizer",
                                       "AddOneServerOptimizer"])
def test_addvector_reconfigures_over_executor_processes(optimizer):
    """The same live plans with one OS process per executor and the TCP control plane (`-executor_mode process`, the
    one-process-per-GPU layout): executors are allocated / stopped as processes, tasklets start and stop remotely,
    deferred master messages and hand-backs cross sockets -- the sums stay exact."""
    res = ETDolphinLauncher.launch(examples.ADD_VECTOR, _args(optimizer) + ["-executor_mode", "process",
                                                                           "-backend", "host"])
    assert res.ok, res.messages
    assert any("(S: 2, W: 2) -> " in m for m in res.messages), res.messages


def test_addinteger_standalone():
    res = ETDolphinLauncher.launch(examples.ADD_INTEGER, _args("EmptyPlanOptimizer", epochs=2))
    assert res.ok


def _compiler():
    tc = lambda e: TaskletConfiguration(f"t-{e}", "x:Y")  # noqa: E731
    return PlanCompiler("model", "input", tc, tc)


def _types(plan):
    out = {}
    for op in plan.ops():
        out[op.op_type] = out.get(op.op_type, 0) + 1
    return out


def test_plan_compiler_shapes():
    c = _compiler()
    # move only
    p = Plan(); p.transfer_steps[NAMESPACE_SERVER].append(TransferStep("s0", "s1", 3))
    assert _types(c.compile(p)) == {"MOVE": 1}
    # add a worker: allocate -> (associate input, subscribe model) -> start ; move feeds the new worker first
    p = Plan(); p.evaluators_to_add[NAMESPACE_WORKER].append("NEW"); p.transfer_steps[NAMESPACE_WORKER].append(TransferStep("w0", "NEW", 2))
    et = c.compile(p, 1)
    assert _types(et) == {"ALLOCATE": 1, "ASSOCIATE": 1, "SUBSCRIBE": 1, "START": 1, "MOVE": 1}
    ops = {o.op_type: o for o in et.ops()}
    assert et.dag.is_adjacent(ops["ALLOCATE"], ops["ASSOCIATE"]) and et.dag.is_adjacent(ops["ASSOCIATE"], ops["MOVE"])
    assert et.dag.is_adjacent(ops["MOVE"], ops["START"]) and et.dag.get_root_vertices() == {ops["ALLOCATE"]}
    # delete a server: stop -> move -> unassociate -> deallocate
    p = Plan(); p.evaluators_to_delete[NAMESPACE_SERVER].append("s1"); p.transfer_steps[NAMESPACE_SERVER].append(TransferStep("s1", "s0", 4))
    et = c.compile(p)
    assert _types(et) == {"STOP": 1, "MOVE": 1, "UNASSOCIATE": 1, "DEALLOCATE": 1}
    ops = {o.op_type: o for o in et.ops()}
    assert et.dag.is_adjacent(ops["STOP"], ops["MOVE"]) and et.dag.is_adjacent(ops["MOVE"], ops["UNASSOCIATE"])
    assert et.dag.is_adjacent(ops["UNASSOCIATE"], ops["DEALLOCATE"])
    # switch: delete a server + add a worker => no allocate / deallocate at all
    p = Plan(); p.evaluators_to_delete[NAMESPACE_SERVER].append("s1"); p.evaluators_to_add[NAMESPACE_WORKER].append("NEW")
    p.transfer_steps[NAMESPACE_SERVER].append(TransferStep("s1", "s0", 4)); p.transfer_steps[NAMESPACE_WORKER].append(TransferStep("w0", "NEW", 2))
    et = c.compile(p)
    t = _types(et)
    assert "ALLOCATE" not in t and "DEALLOCATE" not in t and t["MOVE"] == 2 and t["START"] == 1 and t["STOP"] == 1
    assert {m.dst_executor_id for m in et.ops() if m.op_type == "MOVE"} == {"s0", "s1"}
    # infeasible: two adds, no extra resources
    p = Plan(); p.evaluators_to_add[NAMESPACE_WORKER] += ["A", "B"]
    with pytest.raises(RuntimeError):
        c.compile(p, 1)


def test_homogeneous_optimizer_and_transfer_steps():
    opt = HomogeneousOptimizer(num_total_mini_batches=100, default_bandwidth=8e9, opt_benefit_threshold=0.0)
    workers = [EvaluatorParameters(f"w{i}", 10, {"batch_comp_time_sec": 1.0}) for i in range(2)]
    servers = [EvaluatorParameters(f"s{i}", 8, {}) for i in range(4)]
    # compute-bound job with tiny pulls: more workers (fewer servers) is better
    plan = opt.optimize({NAMESPACE_SERVER: servers, NAMESPACE_WORKER: workers}, 6, {"AVG_PULL_SIZE_PER_MINI_BATCH": 1e3})
    assert not plan.is_empty()
    assert len(plan.get_evaluators_to_delete(NAMESPACE_SERVER)) >= 1 and len(plan.get_evaluators_to_add(NAMESPACE_WORKER)) >= 1
    moved = sum(t.num_blocks for t in plan.get_transfer_steps(NAMESPACE_SERVER))
    assert moved >= 8        # the deleted servers' blocks go somewhere
    # sample optimizers produce valid plans
    for name in ("AddOneServerOptimizer", "DeleteOneWorkerOptimizer", "ExchangeOneOptimizer"):
        make_optimizer(name).optimize({NAMESPACE_SERVER: servers, NAMESPACE_WORKER: workers}, 7, {}).validate()


def test_worker_state_manager_barriers_with_optimization():
    released = []

    class Sender:
        def send_release_msg(self, w):
            released.append(w)
    sm = WorkerStateManager(3, Sender())
    for w in ("a", "b"):
        sm.on_sync_msg(w, "INIT")
    assert released == [] and sm.sm.current_state == "INIT"
    sm.on_sync_msg("c", "INIT")
    assert sorted(released) == ["a", "b", "c"] and sm.sm.current_state == "RUN"
    released.clear()
    assert sm.try_enter_optimization() and not sm.try_enter_optimization()
    sm.on_sync_msg("a", "RUN")                 # finishes RUN during optimisation: blocked
    sm.on_sync_msg("d", "INIT")                # late-added worker skips the INIT barrier
    assert released == ["d"]
    sm.on_optimization_finished({"d"}, {"b"})  # b deleted, d added
    assert sm.sm.current_state == "RUN_FINISHING"
    sm.on_sync_msg("c", "RUN"); sm.on_sync_msg("d", "RUN")
    assert sm.sm.current_state == "CLEANUP" and sorted(released[1:]) == ["a", "c", "d"]


def test_heterogeneous_optimizer_swaps_roles_and_compiles_to_switches():
    """The fat-pipe executor should serve and the fast one should work: the optimizer proposes the role swap under the
    same ids (-> SWITCH ops, no re-allocation) and block moves proportional to bandwidth / throughput."""
    from harmony_b200.dolphin.optimizer import HeterogeneousOptimizer
    bw = {"fat": 8e11, "thin": 1e10, "h2": 1e11, "h3": 1e11}
    opt = HeterogeneousOptimizer(num_total_mini_batches=100, default_bandwidth=1e11, host_to_bandwidth=bw)
    servers = [EvaluatorParameters("s_thin", 16, {}, hostname="thin")]
    workers = [EvaluatorParameters("w_fat", 8, {"batch_comp_time_sec": 0.5}, hostname="fat"),
               EvaluatorParameters("w2", 8, {"batch_comp_time_sec": 0.5}, hostname="h2"),
               EvaluatorParameters("w3", 8, {"batch_comp_time_sec": 0.5}, hostname="h3")]
    plan = opt.optimize({NAMESPACE_SERVER: servers, NAMESPACE_WORKER: workers}, 4,
                        {"AVG_PULL_SIZE_PER_MINI_BATCH": 4e9})
    assert not plan.is_empty() and plan.optimizer_type == "HETEROGENEOUS"
    assert plan.get_evaluators_to_delete(NAMESPACE_SERVER) == ["s_thin"] and "w_fat" in plan.get_evaluators_to_add(NAMESPACE_SERVER)
    assert "w_fat" in plan.get_evaluators_to_delete(NAMESPACE_WORKER) and "s_thin" in plan.get_evaluators_to_add(NAMESPACE_WORKER)
    # every model block leaves the old server, every data block of the new server goes to a worker
    assert sum(t.num_blocks for t in plan.get_transfer_steps(NAMESPACE_SERVER) if t.src_id == "s_thin") == 16
    assert sum(t.num_blocks for t in plan.get_transfer_steps(NAMESPACE_WORKER) if t.src_id == "w_fat") == 8
    et = _compiler().compile(plan, 0)
    types = _types(et)
    assert types.get("ALLOCATE", 0) == 0 and types.get("DEALLOCATE", 0) == 0      # pure switches
    assert types["MOVE"] >= 2 and types["START"] >= 2 and types["STOP"] >= 1
    # already optimal: a homogeneous cluster stays as it is
    same = HeterogeneousOptimizer(num_total_mini_batches=100, default_bandwidth=1e11)
    s2 = [EvaluatorParameters("s0", 16, {})]
    w2 = [EvaluatorParameters(f"w{i}", 8, {"batch_comp_time_sec": 0.5}) for i in range(3)]
    assert same.optimize({NAMESPACE_SERVER: s2, NAMESPACE_WORKER: w2}, 4, {"AVG_PULL_SIZE_PER_MINI_BATCH": 1e3}).is_empty()


def test_heterogeneous_optimizer_switches_roles_on_a_live_job(tmp_path):
    """AddVector keeps its sums exact while the optimizer turns the thin-pipe server into a worker and a worker into
    the server (SWITCH ops: stop / move / unassociate / associate / subscribe / start on live executors)."""
    bwf = tmp_path / "bw"
    bwf.write_text("# executor ids on one host\nexecutor-2 1e6\nexecutor-3 1e6\nexecutor-0 8e12\nexecutor-1 8e9\n")
    args = _args("HeterogeneousOptimizer") + ["-host_to_bandwidth_file_path", str(bwf)]
    res = ETDolphinLauncher.launch(examples.ADD_VECTOR, args)
    assert res.ok, res.messages
    changed = [m for m in res.messages if "(S: " in m]
    assert changed, res.messages          # at least one reconfiguration happened and the job still validated


def test_plan_compiler_adds_twin_ops_for_the_worker_local_model_table():
    c = _compiler()
    c.local_model_table_id = "local_model"
    p = Plan(); p.evaluators_to_add[NAMESPACE_WORKER].append("NEW"); p.transfer_steps[NAMESPACE_WORKER].append(TransferStep("w0", "NEW", 2))
    et = c.compile(p, 1)
    assert _types(et) == {"ALLOCATE": 1, "ASSOCIATE": 2, "SUBSCRIBE": 1, "START": 1, "MOVE": 2}
    moves = [o for o in et.ops() if o.op_type == "MOVE"]
    inp, loc = (moves[0], moves[1]) if moves[0].table_id == c.input_table_id else (moves[1], moves[0])
    assert loc.table_id == "local_model" and et.dag.is_adjacent(inp, loc)           # same blocks, input first
    start = next(o for o in et.ops() if o.op_type == "START")
    assert et.dag.is_adjacent(loc, start)
    p = Plan(); p.evaluators_to_delete[NAMESPACE_WORKER].append("w1"); p.transfer_steps[NAMESPACE_WORKER].append(TransferStep("w1", "w0", 3))
    assert _types(c.compile(p)) == {"STOP": 1, "MOVE": 2, "UNASSOCIATE": 2, "UNSUBSCRIBE": 1, "DEALLOCATE": 1}


@pytest.mark.parametrize("app_name,optimizer", [("nmf", "AddOneWorkerOptimizer"), ("nmf", "DeleteOneWorkerOptimizer"),
                                                ("lda", "ExchangeOneOptimizer")])
def test_apps_with_local_model_tables_reconfigure_workers_live(app_name, optimizer, tmp_path):
    """NMF's L rows / LDA's topic assignments live in a worker-local table whose blocks must travel with the input
    blocks when workers are added, deleted or switched."""
    from harmony_b200.common.samples import write_lda_sample, write_nmf_sample
    from harmony_b200.mlapps import lda, nmf
    opt = ["-optimizer", optimizer, "-optimization_interval_ms", "300", "-delay_after_optimization_ms", "100",
           "-min_num_required_batch_metrics", "1", "-num_extra_resources", "1", "-num_dolphin_workers", "2",
           "-num_servers", "2", "-num_mini_batches", "8", "-num_worker_blocks", "8", "-num_server_blocks", "4"]
    if app_name == "nmf":
        d = write_nmf_sample(str(tmp_path / "n"), rows=160, cols=30, rank=4)
        res = ETDolphinLauncher.launch(nmf.APP, ["-input", d, "-max_num_epochs", "30", "-step_size", "0.01", "-rank", "4",
                                                 "-features", "30", "-decay_rate", "1.0"] + opt)
        assert res.ok, res.messages
        assert max(r["last_batch_loss"] for r in res.results) < 8.0, res.results       # ~4 without reconfiguration
    else:
        d = write_lda_sample(str(tmp_path / "l"), docs=80, vocab=50, topics=3, doc_len=30)
        res = ETDolphinLauncher.launch(lda.APP, ["-input", d, "-max_num_epochs", "10", "-num_topics", "3",
                                                 "-num_vocabs", "50"] + opt)
        assert res.ok, res.messages
    assert any("(S: 2, W: 2) -> " in m for m in res.messages), res.messages


def test_plan_compiler_random_plans_compile_to_executable_dags():
    """Random valid plans (adds, deletes, switches in either direction, transfers; with and without a local-model
    table): the compiled DAG must be executable to completion in dependency order and respect the safety orderings
    (a block moves to an executor only after it is associated, leaves one only after its tasklet stopped and before it
    is unassociated; a tasklet starts only after its blocks arrived)."""
    import random
    for seed in range(60):
        rng = random.Random(seed)
        c = _compiler()
        if seed % 2:
            c.local_model_table_id = "local_model"
        servers = [f"s{i}" for i in range(rng.randint(1, 3))]
        workers = [f"w{i}" for i in range(rng.randint(1, 3))]
        p = Plan()
        ns_nodes = {NAMESPACE_SERVER: list(servers), NAMESPACE_WORKER: list(workers)}
        extra = 0
        for ns, prefix in ((NAMESPACE_SERVER, "NS"), (NAMESPACE_WORKER, "NW")):
            if rng.random() < 0.5:
                p.evaluators_to_add[ns].append(prefix + "0"); extra += 1
            if len(ns_nodes[ns]) > 1 and rng.random() < 0.5:
                p.evaluators_to_delete[ns].append(ns_nodes[ns][-1])
        for ns in (NAMESPACE_SERVER, NAMESPACE_WORKER):
            dels, adds = set(p.evaluators_to_delete[ns]), list(p.evaluators_to_add[ns])
            keep = [n for n in ns_nodes[ns] if n not in dels]
            for d in dels:                                   # a deleted node sends everything away
                p.transfer_steps[ns].append(TransferStep(d, rng.choice(keep + adds), rng.randint(1, 4)))
            for a in adds:                                   # a new node receives something
                p.transfer_steps[ns].append(TransferStep(rng.choice(keep), a, rng.randint(1, 4)))
        p.validate()
        et = c.compile(p, extra)
        ops = list(et.ops())
        order, guard = [], 0
        while et.get_ready_ops():
            op = sorted(et.get_ready_ops(), key=repr)[rng.randrange(len(et.get_ready_ops()))]
            order.append(op)
            et.on_complete(op)
            guard += 1
            assert guard <= len(ops)
        assert len(order) == len(ops), f"seed {seed}: {len(ops) - len(order)} ops can never run (cycle)"
        pos = {id(o): i for i, o in enumerate(order)}
        first = lambda pred: min([pos[id(o)] for o in order if pred(o)], default=None)  # noqa: E731
        for m in [o for o in order if o.op_type == "MOVE"]:
            dst, src = m.dst_executor_id, m.src_executor_id
            a = first(lambda o: o.op_type == "ASSOCIATE" and o.executor_id == dst and o.table_id == m.table_id)
            if a is not None:
                assert a < pos[id(m)], (seed, "move before associate", m)
            st = first(lambda o: o.op_type == "START" and o.executor_id == dst)
            if st is not None:
                assert pos[id(m)] < st, (seed, "start before its blocks arrived", m)
            sp = first(lambda o: o.op_type == "STOP" and o.executor_id == src)
            if sp is not None:
                assert sp < pos[id(m)], (seed, "move before the source tasklet stopped", m)
            un = first(lambda o: o.op_type == "UNASSOCIATE" and o.executor_id == src and o.table_id == m.table_id)
            if un is not None:
                assert pos[id(m)] < un, (seed, "unassociate before its blocks left", m)


@pytest.mark.parametrize("kind", ["HomogeneousOptimizer", "HeterogeneousOptimizer"])
def test_optimizers_random_inputs_yield_valid_conserving_plans(kind):
    """Random clusters (block counts, compute times, bandwidths, extra resources): whatever the optimizer decides must
    validate, conserve the blocks of each namespace, empty every deleted executor, never over-draw a sender, and
    compile."""
    import random
    from harmony_b200.dolphin.optimizer import HeterogeneousOptimizer
    for seed in range(80):
        rng = random.Random(seed)
        ns_ = rng.randint(1, 4); nw = rng.randint(1, 4)
        hosts = {f"h{i}": rng.choice([1e9, 1e10, 1e11, 8e11]) for i in range(ns_ + nw)}
        servers = [EvaluatorParameters(f"s{i}", rng.randint(1, 12), {}, hostname=f"h{i}") for i in range(ns_)]
        workers = [EvaluatorParameters(f"w{i}", rng.randint(1, 12), {"batch_comp_time_sec": rng.uniform(0.01, 1.0)},
                                       hostname=f"h{ns_ + i}") for i in range(nw)]
        cls = HeterogeneousOptimizer if kind == "HeterogeneousOptimizer" else HomogeneousOptimizer
        opt = cls(num_total_mini_batches=100, default_bandwidth=1e10, opt_benefit_threshold=rng.choice([0.0, 0.05]),
                  host_to_bandwidth=hosts)
        extra = rng.randint(0, 2)
        plan = opt.optimize({NAMESPACE_SERVER: servers, NAMESPACE_WORKER: workers}, ns_ + nw + extra,
                            {"AVG_PULL_SIZE_PER_MINI_BATCH": rng.choice([1e2, 1e6, 1e9])})
        plan.validate()
        for ns, nodes in ((NAMESPACE_SERVER, servers), (NAMESPACE_WORKER, workers)):
            blocks = {p.id: p.num_blocks for p in nodes}
            for a in plan.get_evaluators_to_add(ns):
                blocks.setdefault(a, 0)
            total = sum(blocks.values())
            for t in plan.get_transfer_steps(ns):
                assert t.num_blocks > 0 and blocks[t.src_id] >= t.num_blocks, (seed, ns, t, blocks)
                blocks[t.src_id] -= t.num_blocks
                blocks[t.dst_id] = blocks.get(t.dst_id, 0) + t.num_blocks
            assert sum(blocks.values()) == total
            for d in plan.get_evaluators_to_delete(ns):
                assert blocks[d] == 0, (seed, ns, d, blocks)
        if not plan.is_empty():
            _compiler().compile(plan, extra)
