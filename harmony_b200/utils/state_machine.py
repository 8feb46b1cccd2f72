"""Named-state finite state machine with a legal-transition table.

Used by the table lifecycle, job server, tasklet representers, worker-state barriers and the
op-tracker of the remote-access handler.  Reference behaviour: utils/.../StateMachine.java:27-282
(check_state / set_state / compare_and_set_state; illegal transitions raise).
"""
from __future__ import annotations

import threading
from typing import Dict, Hashable, List, Set, Tuple


class IllegalStateError(RuntimeError):
    """Raised by StateMachine.check_state / set_state on an illegal state or transition."""

    pass


class StateMachine:
    """Named states + legal-transition table with check_state / set_state / compare_and_set_state.  Reference:
    utils/src/main/java/edu/snu/cay/utils/StateMachine.java:27,44,65,96,199-282."""

    class Builder:
        def __init__(self) -> None:
            self._states: Dict[Hashable, str] = {}
            self._transitions: Dict[Hashable, List[Tuple[Hashable, str]]] = {}
            self._initial = None

        def add_state(self, state: Hashable, description: str = "") -> "StateMachine.Builder":
            if state in self._states:
                raise ValueError(f"A state {state} was already added")
            self._states[state] = description
            self._transitions[state] = []
            return self

        def set_initial_state(self, state: Hashable) -> "StateMachine.Builder":
            if state not in self._states:
                raise ValueError(f"Unknown state {state}")
            self._initial = state
            return self

        def add_transition(self, src: Hashable, dst: Hashable, description: str = "") -> "StateMachine.Builder":
            if src not in self._states or dst not in self._states:
                raise ValueError(f"Unknown state in transition {src} -> {dst}")
            if any(d == dst for d, _ in self._transitions[src]):
                raise ValueError(f"A transition from {src} to {dst} was already added")
            self._transitions[src].append((dst, description))
            return self

        def build(self) -> "StateMachine":
            if self._initial is None:
                raise ValueError("An initial state should be set")
            return StateMachine(dict(self._states), {k: list(v) for k, v in self._transitions.items()},
                                self._initial)

    @staticmethod
    def new_builder() -> "StateMachine.Builder":
        return StateMachine.Builder()

    def __init__(self, states, transitions, initial) -> None:
        self._states = states
        self._transitions = transitions
        self._current = initial
        self._lock = threading.RLock()
        self._cond = threading.Condition(self._lock)

    def _possible(self) -> str:
        lines = [f"  {self._current} -> {d} ({desc})" for d, desc in self._transitions[self._current]]
        return "Possible transitions from the current state are\n" + "\n".join(lines)

    def check_state(self, expected: Hashable) -> None:
        with self._lock:
            if self._current != expected:
                raise IllegalStateError(
                    f"The expected current state is {expected} but the actual state is {self._current}\n"
                    + self._possible())

    def set_state(self, state: Hashable) -> None:
        with self._lock:
            if state not in self._states:
                raise IllegalStateError(f"Unknown state {state}")
            if not any(d == state for d, _ in self._transitions[self._current]):
                raise IllegalStateError(f"Illegal transition from {self._current} to {state}\n" + self._possible())
            self._current = state
            self._cond.notify_all()

    def compare_and_set_state(self, expected: Hashable, state: Hashable) -> bool:
        with self._lock:
            if self._current != expected:
                return False
            self.set_state(state)
            return True

    @property
    def current_state(self) -> Hashable:
        with self._lock:
            return self._current

    def get_current_state(self) -> Hashable:
        return self.current_state

    def wait_for(self, states: Set[Hashable], timeout: float | None = None) -> bool:
        with self._cond:
            return self._cond.wait_for(lambda: self._current in states, timeout)
