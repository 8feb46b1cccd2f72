"""Elastic optimisation: metric processing, cost-model optimizers, orchestration loop.

Reference (jobserver/.../dolphin/optimizer/):
  api/Optimizer.java (optimize(evalParamsMap, availableEvaluators, modelParams) -> Plan),
  impl/ETOptimizationOrchestrator.java:149-195 (loop every optimization_interval_ms while RUN),
  :208-352 (gather + validate metrics, EMA, optimize, compile, execute, settle delay),
  impl/HomogeneousOptimizer.java:127-210,255-289,311-423,461-528 -- cost model
      T(w) = B / sum_i 1 / (comp_i + pullBytes * max(1/bw_i, w / sum bw_servers))
  over w workers / (N - w) servers, bandwidth-proportional server blocks, throughput-proportional
  worker blocks, greedy transfer steps, benefit threshold;
  impl/SampleOptimizers.java:100-363 (Add/Delete one server/worker, Exchange), EmptyPlanOptimizer,
  impl/MetricProcessor.java:49-173 (EMA / windowed averages), Bandwidth/CoreInfoParser.
  The ILP `hetero/` optimizer needs Gurobi and is excluded from the reference build (pom.xml:196-200).
On a B200 box the bandwidth term is the NVLink bandwidth of each GPU (uniform through NVSwitch).
"""
from __future__ import annotations

import heapq
import logging
import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Set, Tuple

from ..et.config import TaskletConfiguration
from ..et.driver import AllocatedExecutor, ETMaster
from ..et.plan import LoggingPlanExecutor, PlanExecutor
from .master import SERVER_TASKLET, WORKER_TASKLET, DolphinMaster
from .plan import (HETEROGENEOUS, HOMOGENEOUS, NAMESPACE_SERVER, NAMESPACE_WORKER, Plan, PlanCompiler, TransferStep, empty_plan)

LOG = logging.getLogger(__name__)
NEW_WORKER_ID_PREFIX, NEW_SERVER_ID_PREFIX = "NEW_WORKER-", "NEW_SERVER-"
AVG_PULL_SIZE_PER_MINI_BATCH, TOTAL_PULLS_PER_MINI_BATCH = "AVG_PULL_SIZE_PER_MINI_BATCH", "TOTAL_PULLS_PER_MINI_BATCH"
NVLINK_BPS = 770e9 * 8       # measured peer copy bandwidth per direction per GPU, in bit/s


@dataclass
class EvaluatorParameters:
    """What the optimizer knows about one executor: id, number of blocks, averaged metrics.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/optimizer/api/EvaluatorParameters.java."""

    id: str
    num_blocks: int
    metrics: Dict[str, float] = field(default_factory=dict)    # comp_time, pull_time, bytes, ...
    hostname: str = "localhost"


def parse_host_to_value_file(path: str, cast=float) -> Dict[str, Any]:
    """`hostname value` per line (sample_host_to_bandwidth); `#` comments."""
    out: Dict[str, Any] = {}
    if not path:
        return out
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            h, v = s.split()[:2]
            out[h] = cast(v)
    return out


class MetricProcessor:
    """Collapses each worker's batch metrics into one averaged record (plain mean, windowed mean
    or exponentially weighted mean)."""

    @staticmethod
    def process_worker_metrics(raw: Dict[str, List[Dict[str, Any]]], window: int = 0, weight: float = 0.0
                               ) -> Dict[str, Dict[str, float]]:
        out: Dict[str, Dict[str, float]] = {}
        for w, ms in raw.items():
            ms = [m for m in ms if m.get("kind", "batch") == "batch"]
            if window > 0:
                ms = ms[-window:]
            if not ms:
                continue
            keys = ("batch_time_sec", "batch_comp_time_sec", "batch_pull_time_sec", "batch_push_time_sec",
                    "num_batch_data_instances", "bytes_received_get_resp", "total_pull_count")
            agg = {}
            for k in keys:
                vals = [float(m.get(k, 0.0)) for m in ms]
                if weight > 0.0:
                    acc = vals[0]
                    for v in vals[1:]:
                        acc = weight * v + (1.0 - weight) * acc
                    agg[k] = acc
                else:
                    agg[k] = sum(vals) / len(vals)
            out[w] = agg
        return out


class Optimizer:
    """SPI: (evaluator parameters, available executors, model parameters) -> Plan.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/optimizer/api/Optimizer.java."""

    def optimize(self, eval_params: Dict[str, List[EvaluatorParameters]], available_evaluators: int,
                 model_params: Dict[str, float]) -> Plan:
        raise NotImplementedError


class EmptyPlanOptimizer(Optimizer):
    """Never changes anything (the default -optimizer).  Reference: optimizer/impl/EmptyPlanOptimizer.java."""

    def optimize(self, eval_params, available_evaluators, model_params) -> Plan:
        return empty_plan()


@dataclass
class _Summary:
    id: str
    num_blocks: int
    throughput: float
    bandwidth: float
    num_optimal_blocks: int = 0


class HomogeneousOptimizer(Optimizer):
    """Cost-model optimizer: picks the worker / server split minimising estimated epoch time = max(compute,
    communication) with bandwidth-proportional server blocks, then emits add / delete / move steps.  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/optimizer/impl/HomogeneousOptimizer.java:127-210,255-289,311-423,461-528."""

    def __init__(self, num_total_mini_batches: int, default_bandwidth: float = NVLINK_BPS,
                 opt_benefit_threshold: float = 0.0, host_to_bandwidth: Optional[Dict[str, float]] = None) -> None:
        self.B = num_total_mini_batches
        self.default_bw = default_bandwidth
        self.threshold = opt_benefit_threshold
        self.host_bw = host_to_bandwidth or {}

    def _bw(self, p: EvaluatorParameters) -> float:
        # entries of -host_to_bandwidth_file_path may name a host or one executor id (the executors of one box share
        # the hostname; an executor whose GPU is shared or throttled is rated individually)
        return self.host_bw.get(p.id, self.host_bw.get(p.hostname, self.default_bw)) / 8.0    # bit/s -> byte/s

    def _sorted(self, params: List[EvaluatorParameters], available: int, unit_cost, prefix: str) -> List[_Summary]:
        nodes = [_Summary(p.id, p.num_blocks, 1.0 / max(unit_cost(p), 1e-12), self._bw(p)) for p in params]
        nodes.sort(key=lambda s: -s.throughput)
        if params:
            avg_cost = sum(unit_cost(p) for p in params) / len(params)
            avg_bw = sum(self._bw(p) for p in params) / len(params)
            for i in range(max(0, available - len(params) - 1)):
                nodes.append(_Summary(f"{prefix}{i}", 0, 1.0 / max(avg_cost, 1e-12), avg_bw))
        return nodes

    def total_cost(self, num_workers: int, avg_pull: float, available: int, workers: List[_Summary],
                   servers: List[_Summary]) -> float:
        num_servers = available - num_workers
        sbw = sum(s.bandwidth for s in servers[:num_servers])
        inv = 0.0
        for w in workers[:num_workers]:
            term = 1.0 / w.throughput + avg_pull * max(1.0 / w.bandwidth, num_workers / sbw)
            inv += 1.0 / term
        return self.B / inv

    def _cost_map(self, servers, workers, n_use: int, avg_pull: float) -> Dict[int, float]:
        model_blocks = sum(s.num_blocks for s in servers)
        data_blocks = sum(w.num_blocks for w in workers)
        out = {}
        for nw in range(1, n_use):
            if nw <= data_blocks and (n_use - nw) <= model_blocks and nw <= len(workers) and (n_use - nw) <= len(servers):
                out[nw] = self.total_cost(nw, avg_pull, n_use, workers, servers)
        return out

    def optimize(self, eval_params, available_evaluators, model_params) -> Plan:
        sp, wp = eval_params[NAMESPACE_SERVER], eval_params[NAMESPACE_WORKER]
        curr = len(sp) + len(wp)
        extra = available_evaluators - curr
        avg_pull = float(model_params.get(AVG_PULL_SIZE_PER_MINI_BATCH, 0.0))
        servers = self._sorted(sp, available_evaluators, lambda p: 1.0 / self._bw(p), NEW_SERVER_ID_PREFIX)
        workers = self._sorted(wp, available_evaluators,
                               lambda p: max(p.metrics.get("batch_comp_time_sec", 0.0), 1e-9), NEW_WORKER_ID_PREFIX)
        if extra > 0:
            cm_extra = self._cost_map(servers, workers, available_evaluators, avg_pull)
            cm_curr = self._cost_map(servers, workers, curr, avg_pull)
            curr_cost = cm_curr.get(len(wp), float("inf"))
            be = min(cm_extra.items(), key=lambda kv: kv[1]) if cm_extra else (len(wp), float("inf"))
            bc = min(cm_curr.items(), key=lambda kv: kv[1]) if cm_curr else (len(wp), float("inf"))
            if be[1] < bc[1]:
                opt_w, opt_cost, n_use = be[0], be[1], available_evaluators
            else:
                opt_w, opt_cost, n_use = bc[0], bc[1], curr
        else:
            cm = self._cost_map(servers, workers, available_evaluators, avg_pull)
            if not cm:
                return empty_plan()
            curr_cost = float("inf") if extra < 0 else cm.get(len(wp), float("inf"))
            opt_w, opt_cost = min(cm.items(), key=lambda kv: kv[1])
            n_use = available_evaluators
        opt_s = n_use - opt_w
        if curr_cost != float("inf") and (curr_cost - opt_cost) / curr_cost < self.threshold:
            return empty_plan()
        if curr_cost != float("inf") and opt_w == len(wp) and opt_s == len(sp) and opt_cost >= curr_cost:
            return empty_plan()
        plan = Plan(optimizer_type=HOMOGENEOUS, num_available_extra_evals=max(extra, 0))
        self._assign(plan, NAMESPACE_SERVER, servers, opt_s, len(sp), [s.bandwidth for s in servers[:opt_s]])
        sbw = sum(s.bandwidth for s in servers[:opt_s])
        inv = [1.0 / (1.0 / w.throughput + avg_pull * max(1.0 / w.bandwidth, opt_w / sbw)) for w in workers[:opt_w]]
        self._assign(plan, NAMESPACE_WORKER, workers, opt_w, len(wp), inv)
        return plan.validate()

    @staticmethod
    def _assign(plan: Plan, ns: str, nodes: List[_Summary], n_opt: int, n_active: int, weights: List[float]) -> None:
        total = sum(n.num_blocks for n in nodes)
        wsum = sum(weights) or 1.0
        assigned = 0
        for n, w in zip(nodes[:n_opt], weights):
            n.num_optimal_blocks = int(total * w / wsum)
            assigned += n.num_optimal_blocks
        i = 0
        for _ in range(total - assigned):
            nodes[i].num_optimal_blocks += 1
            i = (i + 1) % max(1, n_opt)
        if n_active > n_opt:
            for n in nodes[n_opt:n_active]:
                plan.evaluators_to_delete[ns].append(n.id)
                n.num_optimal_blocks = 0
        elif n_active < n_opt:
            for n in nodes[n_active:n_opt]:
                plan.evaluators_to_add[ns].append(n.id)
        plan.transfer_steps[ns].extend(generate_transfer_steps(nodes[:max(n_opt, n_active)]))


class HeterogeneousOptimizer(HomogeneousOptimizer):
    """Chooses WHICH executors work and which serve, not only how many: executors differ in compute speed (measured
    batch compute time, or `-host_to_core_file_path` for executors that have not worked yet) and in bandwidth
    (`-host_to_bandwidth_file_path`).  The epoch-time model is the homogeneous one evaluated per role assignment,

        T(W, S) = B / sum_{i in W} 1 / (c_i + pull * max(1 / b_i, |W| / sum_{j in S} b_j)),

    searched exhaustively up to 14 executors and by pairwise role swaps beyond; data blocks go to the workers in
    proportion to their throughput, model blocks to the servers in proportion to their bandwidth; an executor that
    changes role is deleted in one namespace and added in the other under the SAME id, which the plan compiler turns
    into a switch (no re-allocation).  The reference keeps this optimizer out of its build because it needs a Gurobi
    licence (optimizer/impl/hetero/{HeterogeneousOptimizer,ILPSolver,ILPPlanGenerator}.java, pom.xml:196-200); the
    search space of one box (<= 16 executors) does not need an ILP solver."""

    EXHAUSTIVE_LIMIT = 14

    def __init__(self, num_total_mini_batches: int, default_bandwidth: float = NVLINK_BPS,
                 opt_benefit_threshold: float = 0.0, host_to_bandwidth: Optional[Dict[str, float]] = None,
                 host_to_cores: Optional[Dict[str, float]] = None) -> None:
        super().__init__(num_total_mini_batches, default_bandwidth, opt_benefit_threshold, host_to_bandwidth)
        self.host_cores = host_to_cores or {}

    def _comp_costs(self, sp: List[EvaluatorParameters], wp: List[EvaluatorParameters]) -> Dict[str, float]:
        """Seconds of compute per mini-batch for every executor; servers are rated by cores relative to the workers."""
        measured = {p.id: max(p.metrics.get("batch_comp_time_sec", 0.0), 1e-9) for p in wp}
        mean_cost = sum(measured.values()) / max(len(measured), 1) if measured else 1.0
        worker_cores = [self.host_cores.get(p.hostname, 1.0) for p in wp] or [1.0]
        mean_cores = sum(worker_cores) / len(worker_cores)
        out = dict(measured)
        for p in sp:
            cores = self.host_cores.get(p.id, self.host_cores.get(p.hostname, mean_cores))
            out[p.id] = mean_cost * mean_cores / max(cores, 1e-9)
        return out

    @staticmethod
    def _epoch_time(B: int, workers: Sequence[str], servers: Sequence[str], cost: Dict[str, float],
                    bw: Dict[str, float], pull: float) -> float:
        sbw = sum(bw[s] for s in servers)
        if not workers or sbw <= 0:
            return float("inf")
        inv = sum(1.0 / (cost[w] + pull * max(1.0 / bw[w], len(workers) / sbw)) for w in workers)
        return B / inv if inv > 0 else float("inf")

    def optimize(self, eval_params, available_evaluators, model_params) -> Plan:
        sp, wp = eval_params[NAMESPACE_SERVER], eval_params[NAMESPACE_WORKER]
        every = {p.id: p for p in list(sp) + list(wp)}
        ids = sorted(every)
        n = len(ids)
        if n < 2 or not wp or not sp:
            return empty_plan(HETEROGENEOUS)
        pull = float(model_params.get(AVG_PULL_SIZE_PER_MINI_BATCH, 0.0))
        cost = self._comp_costs(sp, wp)
        bw = {i: self._bw(every[i]) for i in ids}
        data_blocks = sum(p.num_blocks for p in wp)
        model_blocks = sum(p.num_blocks for p in sp)

        def feasible(k: int) -> bool:          # k workers, n - k servers
            return 1 <= k <= n - 1 and k <= max(data_blocks, 1) and (n - k) <= max(model_blocks, 1)

        def t_of(wset) -> float:
            return self._epoch_time(self.B, [i for i in ids if i in wset], [i for i in ids if i not in wset],
                                    cost, bw, pull)
        cur_w = frozenset(p.id for p in wp)
        cur_t = t_of(cur_w)
        best_w, best_t = cur_w, cur_t
        if n <= self.EXHAUSTIVE_LIMIT:
            for mask in range(1, (1 << n) - 1):
                if not feasible(bin(mask).count("1")):
                    continue
                wset = frozenset(ids[i] for i in range(n) if mask >> i & 1)
                t = t_of(wset)
                if t < best_t - 1e-12:
                    best_w, best_t = wset, t
        else:
            improved = True
            while improved:                     # steepest descent over single moves and pairwise swaps
                improved = False
                cands = [best_w ^ {i} for i in ids] + [best_w ^ {i, j} for i in best_w for j in ids if j not in best_w]
                for c in cands:
                    if feasible(len(c)):
                        t = t_of(c)
                        if t < best_t - 1e-12:
                            best_w, best_t, improved = frozenset(c), t, True
        if best_w == cur_w or (cur_t != float("inf") and (cur_t - best_t) / cur_t < max(self.threshold, 1e-9)):
            return self._rebalance_only(sp, wp, cost, bw, pull)
        plan = Plan(optimizer_type=HETEROGENEOUS)
        new_s = [i for i in ids if i not in best_w]
        self._assign_roles(plan, NAMESPACE_SERVER, sp, new_s, {i: bw[i] for i in new_s})
        sbw = sum(bw[i] for i in new_s)
        wts = {i: 1.0 / (cost[i] + pull * max(1.0 / bw[i], len(best_w) / sbw)) for i in best_w}
        self._assign_roles(plan, NAMESPACE_WORKER, wp, sorted(best_w), wts)
        return plan.validate()

    def _rebalance_only(self, sp, wp, cost, bw, pull) -> Plan:
        """Roles stay; still move blocks when the current distribution is far from proportional."""
        plan = Plan(optimizer_type=HETEROGENEOUS)
        s_ids = [p.id for p in sp]
        sbw = sum(bw[i] for i in s_ids)
        self._assign_roles(plan, NAMESPACE_SERVER, sp, s_ids, {i: bw[i] for i in s_ids})
        self._assign_roles(plan, NAMESPACE_WORKER, wp, [p.id for p in wp],
                           {p.id: 1.0 / (cost[p.id] + pull * max(1.0 / bw[p.id], len(wp) / sbw)) for p in wp})
        moved = sum(t.num_blocks for ns in (NAMESPACE_SERVER, NAMESPACE_WORKER) for t in plan.transfer_steps[ns])
        total = sum(p.num_blocks for p in list(sp) + list(wp))
        return plan if total and moved / total > max(self.threshold, 0.1) else empty_plan(HETEROGENEOUS)

    @staticmethod
    def _assign_roles(plan: Plan, ns: str, current: List[EvaluatorParameters], target_ids: List[str],
                      weights: Dict[str, float]) -> None:
        cur = {p.id: p.num_blocks for p in current}
        total = sum(cur.values())
        wsum = sum(weights.values()) or 1.0
        nodes: Dict[str, _Summary] = {}
        for i in target_ids:
            nodes[i] = _Summary(i, cur.get(i, 0), 0.0, 0.0, int(total * weights[i] / wsum))
        order = sorted(target_ids, key=lambda i: -weights[i])
        for k in range(total - sum(nd.num_optimal_blocks for nd in nodes.values())):
            nodes[order[k % len(order)]].num_optimal_blocks += 1
        for i in target_ids:
            if i not in cur:
                plan.evaluators_to_add[ns].append(i)
        for i in cur:
            if i not in nodes:
                plan.evaluators_to_delete[ns].append(i)
                nodes[i] = _Summary(i, cur[i], 0.0, 0.0, 0)
        plan.transfer_steps[ns].extend(generate_transfer_steps(list(nodes.values())))


def generate_transfer_steps(nodes: Sequence[_Summary]) -> List[TransferStep]:
    """Greedy matching of senders (blocks > optimal) to receivers (blocks < optimal)."""
    senders = [[-(n.num_blocks - n.num_optimal_blocks), i, n] for i, n in enumerate(nodes)
               if n.num_blocks > n.num_optimal_blocks]
    receivers = [[-(n.num_optimal_blocks - n.num_blocks), i, n] for i, n in enumerate(nodes)
                 if n.num_blocks < n.num_optimal_blocks]
    heapq.heapify(senders); heapq.heapify(receivers)
    blocks = {n.id: n.num_blocks for n in nodes}
    steps: List[TransferStep] = []
    while senders and receivers:
        _, si, s = heapq.heappop(senders)
        _, ri, r = heapq.heappop(receivers)
        to_send = blocks[s.id] - s.num_optimal_blocks
        to_recv = r.num_optimal_blocks - blocks[r.id]
        mv = min(to_send, to_recv)
        if mv <= 0:
            continue
        steps.append(TransferStep(s.id, r.id, mv))
        blocks[s.id] -= mv
        blocks[r.id] += mv
        if blocks[s.id] > s.num_optimal_blocks:
            heapq.heappush(senders, [-(blocks[s.id] - s.num_optimal_blocks), si, s])
        if blocks[r.id] < r.num_optimal_blocks:
            heapq.heappush(receivers, [-(r.num_optimal_blocks - blocks[r.id]), ri, r])
    return steps


# ---- sample optimizers (deterministic plans for tests / demos) --------------------------------
class _OneShot(Optimizer):
    MAX_CALLS = 1

    def __init__(self, **_: Any) -> None:
        self.calls = 0

    def optimize(self, eval_params, available_evaluators, model_params) -> Plan:
        self.calls += 1
        if self.calls > self.MAX_CALLS:
            return empty_plan()
        return self._plan(eval_params, available_evaluators).validate()


def _most_blocks(params: List[EvaluatorParameters]) -> EvaluatorParameters:
    return max(params, key=lambda p: p.num_blocks)


class AddOneServerOptimizer(_OneShot):
    """Sample optimizer: add one server and give it half of the most loaded server's blocks.  Reference:
    optimizer/impl/SampleOptimizers.java:100-363."""

    def _plan(self, ep, avail):
        p = Plan()
        src = _most_blocks(ep[NAMESPACE_SERVER])
        p.evaluators_to_add[NAMESPACE_SERVER].append(NEW_SERVER_ID_PREFIX + "0")
        p.transfer_steps[NAMESPACE_SERVER].append(TransferStep(src.id, NEW_SERVER_ID_PREFIX + "0", max(1, src.num_blocks // 2)))
        p.num_available_extra_evals = 1
        return p


class AddOneWorkerOptimizer(_OneShot):
    """Sample optimizer: add one worker with half of the most loaded worker's input blocks.  Reference:
    SampleOptimizers.java:100-363."""

    def _plan(self, ep, avail):
        p = Plan()
        src = _most_blocks(ep[NAMESPACE_WORKER])
        p.evaluators_to_add[NAMESPACE_WORKER].append(NEW_WORKER_ID_PREFIX + "0")
        p.transfer_steps[NAMESPACE_WORKER].append(TransferStep(src.id, NEW_WORKER_ID_PREFIX + "0", max(1, src.num_blocks // 2)))
        p.num_available_extra_evals = 1
        return p


class DeleteOneServerOptimizer(_OneShot):
    """Sample optimizer: delete the last server after moving its blocks to the first.  Reference:
    SampleOptimizers.java:100-363 (used by OwnershipFirstMigrationTest)."""

    def _plan(self, ep, avail):
        p = Plan()
        servers = ep[NAMESPACE_SERVER]
        if len(servers) < 2:
            return p
        victim, dst = servers[-1], servers[0]
        p.evaluators_to_delete[NAMESPACE_SERVER].append(victim.id)
        if victim.num_blocks:
            p.transfer_steps[NAMESPACE_SERVER].append(TransferStep(victim.id, dst.id, victim.num_blocks))
        return p


class DeleteOneWorkerOptimizer(_OneShot):
    """Sample optimizer: delete the last worker after moving its input blocks away.  Reference:
    SampleOptimizers.java:100-363."""

    def _plan(self, ep, avail):
        p = Plan()
        workers = ep[NAMESPACE_WORKER]
        if len(workers) < 2:
            return p
        victim, dst = workers[-1], workers[0]
        p.evaluators_to_delete[NAMESPACE_WORKER].append(victim.id)
        if victim.num_blocks:
            p.transfer_steps[NAMESPACE_WORKER].append(TransferStep(victim.id, dst.id, victim.num_blocks))
        return p


class ExchangeOneOptimizer(_OneShot):
    """Turn one server into a worker (switch): delete in SERVER + add in WORKER."""

    def _plan(self, ep, avail):
        p = Plan()
        servers = ep[NAMESPACE_SERVER]
        if len(servers) < 2:
            return p
        victim, dst = servers[-1], servers[0]
        p.evaluators_to_delete[NAMESPACE_SERVER].append(victim.id)
        if victim.num_blocks:
            p.transfer_steps[NAMESPACE_SERVER].append(TransferStep(victim.id, dst.id, victim.num_blocks))
        src = _most_blocks(ep[NAMESPACE_WORKER])
        p.evaluators_to_add[NAMESPACE_WORKER].append(NEW_WORKER_ID_PREFIX + "0")
        p.transfer_steps[NAMESPACE_WORKER].append(TransferStep(src.id, NEW_WORKER_ID_PREFIX + "0", max(1, src.num_blocks // 2)))
        return p


class MoveModelBlocksOptimizer(_OneShot):
    """Re-partition the model table: move half of the most loaded server's blocks to the least
    loaded one (used by the 4-concurrent-jobs config for its mid-run elastic re-partition)."""

    def _plan(self, ep, avail):
        p = Plan()
        servers = ep[NAMESPACE_SERVER]
        if len(servers) < 2:
            return p
        src = _most_blocks(servers)
        dst = min(servers, key=lambda s: s.num_blocks)
        if src.id != dst.id and src.num_blocks > 0:
            p.transfer_steps[NAMESPACE_SERVER].append(TransferStep(src.id, dst.id, max(1, src.num_blocks // 2)))
        return p


_OPTIMIZERS: Dict[str, Callable[..., Optimizer]] = {
    "EmptyPlanOptimizer": EmptyPlanOptimizer, "HomogeneousOptimizer": HomogeneousOptimizer,
    "HeterogeneousOptimizer": HeterogeneousOptimizer,
    "AddOneServerOptimizer": AddOneServerOptimizer, "AddOneWorkerOptimizer": AddOneWorkerOptimizer,
    "DeleteOneServerOptimizer": DeleteOneServerOptimizer, "DeleteOneWorkerOptimizer": DeleteOneWorkerOptimizer,
    "ExchangeOneOptimizer": ExchangeOneOptimizer, "MoveModelBlocksOptimizer": MoveModelBlocksOptimizer,
}


def make_optimizer(name: str, **kw: Any) -> Optimizer:
    """`-optimizer <class name>` (the reference's fully qualified Java names are accepted: the last component
    selects the implementation); constructor arguments are matched by name."""
    import inspect
    short = name.split(".")[-1].split("$")[-1]
    cls = _OPTIMIZERS.get(short)
    if cls is None:
        raise ValueError(f"unknown optimizer {name}")
    params = inspect.signature(cls.__init__).parameters
    if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in params.values()):
        return cls(**kw)
    return cls(**{k: v for k, v in kw.items() if k in params})


class ETOptimizationOrchestrator:
    """Periodically: gather metrics -> optimize -> compile -> execute the plan on the live job."""

    def __init__(self, master: ETMaster, dolphin_master: DolphinMaster, conf, model_table_id: str,
                 input_table_id: str, servers: Sequence[AllocatedExecutor], workers: Sequence[AllocatedExecutor],
                 local_model_table_id: Optional[str] = None) -> None:
        self._m, self._dm, self.conf = master, dolphin_master, conf
        self.model_table_id, self.input_table_id = model_table_id, input_table_id
        self.servers: Set[str] = {e.id for e in servers}
        self.workers: Set[str] = {e.id for e in workers}
        epochs, nmb = int(conf.get("max_num_epochs", 1)), int(conf.get("num_mini_batches", 10))
        self.optimizer = make_optimizer(conf.get("optimizer", "EmptyPlanOptimizer"),
                                        num_total_mini_batches=epochs * nmb,
                                        default_bandwidth=float(conf.get("default_bandwidth", NVLINK_BPS)),
                                        opt_benefit_threshold=float(conf.get("opt_benefit_threshold", 0.0)),
                                        host_to_bandwidth=parse_host_to_value_file(
                                            conf.get("host_to_bandwidth_file_path", "")),
                                        host_to_cores=parse_host_to_value_file(
                                            conf.get("host_to_core_file_path", "")))
        ex_name = conf.get("et_plan_executor", "PlanExecutorImpl")
        self.plan_executor = LoggingPlanExecutor(master) if ex_name.endswith("LoggingPlanExecutor") \
            else PlanExecutor(master)
        self.interval = int(conf.get("optimization_interval_ms", 10000)) / 1000.0
        self.delay_after = int(conf.get("delay_after_optimization_ms", 10000)) / 1000.0
        self.min_metrics = int(conf.get("min_num_required_batch_metrics", 3))
        self.window = int(conf.get("moving_avg_window_size", 0))
        self.weight = float(conf.get("metric_weight_factor", 0.0))
        self.num_available = len(self.servers | self.workers) + int(conf.get("num_extra_resources", 0))
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.executed_plans: List[Plan] = []
        p = dolphin_master.worker_params()
        self.compiler = PlanCompiler(
            model_table_id, input_table_id,
            lambda e: TaskletConfiguration(f"{dolphin_master.job_id}-worker-{e}", WORKER_TASKLET, None, p),
            lambda e: TaskletConfiguration(f"{dolphin_master.job_id}-server-{e}", SERVER_TASKLET, None,
                                           {"job_id": dolphin_master.job_id}),
            msg_handler=dolphin_master.on_worker_msg, local_model_table_id=local_model_table_id)

    def start(self) -> None:
        self._thread = threading.Thread(target=self._loop, name="optimizer", daemon=True)
        self._thread.start()

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=120)

    def _loop(self) -> None:
        sm = None
        while not self._stop.wait(self.interval):
            sm = self._dm.state_manager
            if sm is None or sm.sm.current_state != "RUN":
                continue
            try:
                self.optimize_once()
            except Exception:  # noqa: BLE001
                LOG.exception("optimization round failed")

    def _eval_params(self) -> Optional[Tuple[Dict[str, List[EvaluatorParameters]], Dict[str, float]]]:
        mm = self._dm.metric_manager
        raw = mm.get_worker_batch_metrics()
        processed = MetricProcessor.process_worker_metrics(raw, self.window, self.weight)
        model_info = self._m.get_table(self.model_table_id).get_partition_info()
        input_info = self._m.get_table(self.input_table_id).get_partition_info()
        wps = []
        for w in sorted(self.workers):
            if self.min_metrics > 0 and len(raw.get(w, [])) < self.min_metrics:
                return None
            wps.append(EvaluatorParameters(w, len(input_info.get(w, ())), processed.get(w, {})))
        sps = [EvaluatorParameters(s, len(model_info.get(s, ())), {}) for s in sorted(self.servers)]
        pulls = [m.get("bytes_received_get_resp", 0.0) for m in processed.values()]
        model_params = {AVG_PULL_SIZE_PER_MINI_BATCH: (sum(pulls) / len(pulls)) if pulls else 0.0}
        return {NAMESPACE_SERVER: sps, NAMESPACE_WORKER: wps}, model_params

    def optimize_once(self) -> bool:
        ep = self._eval_params()
        if ep is None:
            return False
        plan = self.optimizer.optimize(ep[0], self.num_available, ep[1])
        if plan.is_empty():
            return False
        sm = self._dm.state_manager
        if not sm.try_enter_optimization():
            return False
        mm = self._dm.metric_manager
        mm.stop_metric_collection()
        added_w: Set[str] = set()
        deleted_w: Set[str] = set()
        try:
            et_plan = self.compiler.compile(plan, self.num_available - len(self.servers | self.workers))
            before = (len(self.servers), len(self.workers))
            listener_id = f"opt-{id(plan)}"
            new_worker_tasklets, new_server_tasklets = {}, {}

            def on_op(res):
                if res.op_type == "START":
                    rt = res.payload
                    if "-worker-" in rt.tasklet_id:
                        new_worker_tasklets[rt.executor_id] = rt
                        self._dm.task_runner.register_worker(rt)      # it may sync with the master right away
                    else:
                        new_server_tasklets[rt.executor_id] = rt
            self.plan_executor.register_listener(listener_id, on_op)
            vmap = self.plan_executor.execute(et_plan).get(timeout=600) or {}
            self.plan_executor.unregister_listener(listener_id)
            # ground truth = where tasklets were actually started: a virtual id of the plan is either mapped to a
            # freshly allocated executor (vmap) or was translated into a SWITCH of an existing one
            del vmap
            switched_s2w = set(new_worker_tasklets) & self.servers
            switched_w2s = set(new_server_tasklets) & self.workers
            added_w = set(new_worker_tasklets)
            deleted_w = (set(plan.get_evaluators_to_delete(NAMESPACE_WORKER)) & self.workers) | switched_w2s
            added_s = set(new_server_tasklets)
            deleted_s = (set(plan.get_evaluators_to_delete(NAMESPACE_SERVER)) & self.servers) | switched_s2w
            self.workers = (self.workers | added_w) - deleted_w
            self.servers = (self.servers | added_s) - deleted_s
            self._dm.task_runner.update_executor_entry(new_worker_tasklets, deleted_w, new_server_tasklets, deleted_s)
            for w in deleted_w:
                self._dm.mini_batch_controller.on_worker_deleted(w)
            self._dm.report(f"(S: {before[0]}, W: {before[1]}) -> (S: {len(self.servers)}, W: {len(self.workers)})")
            self.executed_plans.append(plan)
        finally:
            model_info = self._m.get_table(self.model_table_id).get_partition_info()
            input_info = self._m.get_table(self.input_table_id).get_partition_info()
            mm.load_metric_validation_info({w: len(b) for w, b in input_info.items()},
                                           {s: len(b) for s, b in model_info.items()})
            mm.clear_worker_metrics()
            mm.start_metric_collection()
            sm.on_optimization_finished(added_w & self.workers, deleted_w)
        if self.delay_after > 0:
            self._stop.wait(self.delay_after)
        return True
