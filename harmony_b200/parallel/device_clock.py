"""Device-side SSP clocks / barriers over peer memory (no host round trip, no NCCL).

The reference synchronises every mini-batch through the driver: MiniBatchBarrier ->
MiniBatchSyncMsg -> MiniBatchController (dolphin/core/master/MiniBatchController.java:51-117)
blocks a worker whose progress exceeds min progress + clock_slack.  Here each GPU keeps an
int32[world] clock array in its arena; after its PUSH a worker writes its clock into slot `rank`
of EVERY GPU's array (st.release.sys over NVLink, after a system fence so the pushes are visible
first) and before its next PULL a tiny kernel spins (bounded) until
min(clocks) >= my_clock - slack.  Everything stays on the stream, so the loop is graph-capturable.
"""
from __future__ import annotations

from typing import Optional

import torch

from .peer import PeerArena, PeerGroup


class DeviceClock:
    """Device-side SSP clocks / barriers over peer memory: st.release.sys publishes a worker's clock into every
    GPU's array, a bounded spin kernel waits for min(clocks) >= mine - slack.  Replaces MiniBatchBarrier ->
    MiniBatchController round trips (dolphin/core/master/MiniBatchController.java:51-117) on the SPMD path."""

    def __init__(self, arena: PeerArena, group: Optional[PeerGroup] = None, slack: int = 0,
                 timeout_cycles: int = 20_000_000_000) -> None:
        from ..ops import native
        self._n = native()
        self.arena = arena
        self.group = group or arena.group
        self.rank, self.world = self.group.rank, self.group.world
        self.slack = int(slack)
        self.timeout_cycles = int(timeout_cycles)
        self.device = torch.device("cuda", arena.device)
        self._ptr, si, off = arena.alloc(max(self.world, 32) * 4)
        self.clocks = arena.tensor(self._ptr, [self.world], torch.int32)
        self.clocks.zero_()
        torch.cuda.synchronize(self.device)
        infos = self.group.all_gather_object((self.rank, si, off, arena.segment_info(si)))
        ptrs = [0] * self.world
        for r, s, o, h in infos:
            arena.map_peer(r, s, h)
            ptrs[r] = arena.resolve(r, s, o)
        self.clock_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.value = 0
        self.group.barrier()

    def tick(self) -> None:
        """Publish that this worker finished one more mini-batch (after its push)."""
        from ..ops import count_launch
        self.value += 1
        self._n.clock_publish(self.clock_ptrs, self.rank, self.value)
        count_launch()

    def wait(self) -> None:
        """Block the STREAM until every worker's clock >= my clock - slack."""
        from ..ops import count_launch
        if self.world == 1:
            return
        need = self.value - self.slack
        if need <= 0:
            return
        self._n.clock_wait(self.clocks, need, self.timeout_cycles, self.err)
        count_launch()

    def barrier(self) -> None:
        """Full device-side barrier: tick, then wait for everyone to reach my clock."""
        from ..ops import count_launch
        self.tick()
        if self.world > 1:
            self._n.clock_wait(self.clocks, self.value, self.timeout_cycles, self.err)
            count_launch()

    def check(self) -> None:
        e = int(self.err.item())
        if e != 0:
            raise TimeoutError(f"device clock wait timed out waiting for rank {e - 1}")
