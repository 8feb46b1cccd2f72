"""Rendezvous group over the control plane: gives the workers of ONE job the small collective
surface (`rank`, `world`, `all_gather_object`, `barrier`) that the SPMD device classes need
(DeviceClock, MLRShipModel, route_batches) when they run inside the job server instead of under
torchrun.  The job's DolphinMaster is the rendezvous point: every worker sends
(tag, seq, payload); once all `world` payloads of a (tag, seq) arrived the master answers each
worker with the rank-ordered list.  Only handles / descriptors travel here -- the data plane stays
in peer memory.
"""
from __future__ import annotations

import threading
from typing import Any, Callable, Dict, List, Tuple


class MasterSideRendezvous:
    """Master half of the job-level gather used by SPMD-style trainers (ship mode): collects one object per worker
    and replies to all.  No reference twin (the reference has no peer-to-peer data plane)."""

    def __init__(self, send_to_worker: Callable[[str, Dict[str, Any]], None]) -> None:
        self._send = send_to_worker
        self._pending: Dict[Tuple[str, int], Dict[str, Any]] = {}
        self._lock = threading.Lock()

    def on_msg(self, body: Dict[str, Any]) -> None:
        key = (body["tag"], body["seq"])
        with self._lock:
            slot = self._pending.setdefault(key, {})
            slot[body["worker_id"]] = (body["rank"], body["payload"])
            done = len(slot) >= body["world"]
            if done:
                self._pending.pop(key)
        if done:
            ordered = [p for _, p in sorted(slot.values(), key=lambda rp: rp[0])]
            for w in slot:
                self._send(w, {"type": "GroupGatherRes", "tag": body["tag"], "seq": body["seq"], "payloads": ordered})


class WorkerSideGroup:
    """Duck-types parallel.peer.PeerGroup."""

    def __init__(self, rank: int, world: int, worker_id: str, send_to_master: Callable[[Dict[str, Any]], None]) -> None:
        self.rank, self.world, self.pg = rank, world, None
        self._worker_id = worker_id
        self._send = send_to_master
        self._seq = 0
        self._waiting: Dict[Tuple[str, int], Any] = {}
        self._results: Dict[Tuple[str, int], List[Any]] = {}
        self._cond = threading.Condition()

    def on_msg(self, body: Dict[str, Any]) -> None:
        with self._cond:
            self._results[(body["tag"], body["seq"])] = body["payloads"]
            self._cond.notify_all()

    def all_gather_object(self, obj: Any, tag: str = "g", timeout: float = 300.0) -> List[Any]:
        if self.world == 1:
            return [obj]
        self._seq += 1
        key = (tag, self._seq)
        self._send({"type": "GroupGather", "tag": tag, "seq": self._seq, "rank": self.rank, "world": self.world,
                    "worker_id": self._worker_id, "payload": obj})
        with self._cond:
            if not self._cond.wait_for(lambda: key in self._results, timeout):
                raise TimeoutError(f"group gather {key} timed out")
            return self._results.pop(key)

    def barrier(self) -> None:
        self.all_gather_object(None, tag="barrier")
