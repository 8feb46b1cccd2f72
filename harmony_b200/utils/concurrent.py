"""Futures / callback registries / exception-propagating executors.

Reference: services/et/.../common/util/concurrent/{ResultFuture,AggregateFuture,CompletedFuture,
ListenableFuture}.java, common/impl/CallbackRegistry.java:46-85, utils/CatchableExecutors.java:36-78.
"""
from __future__ import annotations

import logging
import threading
import traceback
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, Dict, Generic, List, Optional, TypeVar

T = TypeVar("T")
LOG = logging.getLogger(__name__)


class ResultFuture(Generic[T]):
    """A future completed exactly once by `on_completed`; listeners fire on completion."""

    def __init__(self) -> None:
        self._ev = threading.Event()
        self._result: Optional[T] = None
        self._exc: Optional[BaseException] = None
        self._listeners: List[Callable[[T], None]] = []
        self._lock = threading.Lock()

    def on_completed(self, result: T = None) -> None:
        with self._lock:
            if self._ev.is_set():
                return
            self._result = result
            self._ev.set()
            ls, self._listeners = self._listeners, []
        for cb in ls:
            cb(result)

    def on_failed(self, exc: BaseException) -> None:
        with self._lock:
            if self._ev.is_set():
                return
            self._exc = exc
            self._ev.set()
            ls, self._listeners = self._listeners, []
        for cb in ls:
            try:
                cb(None)
            except Exception:  # noqa: BLE001
                LOG.exception("listener failed")

    def add_listener(self, cb: Callable[[T], None]) -> "ResultFuture[T]":
        with self._lock:
            if not self._ev.is_set():
                self._listeners.append(cb)
                return self
        cb(self._result)
        return self

    def exception(self) -> Optional[BaseException]:
        return self._exc

    def chain(self, out: "ResultFuture", cb: Callable[[T], None],
              on_error: Optional[Callable[[BaseException], None]] = None) -> "ResultFuture[T]":
        """Continue an asynchronous sequence: on success run `cb(result)` (an exception in it fails `out`), on
        failure fail `out` with the same exception -- so an error anywhere in a multi-step control operation
        (init -> peer map -> load ...) reaches whoever waits on the operation's final future instead of leaving it
        pending forever.  `on_error` (optional) runs instead of `out.on_failed` to undo partial work first."""
        fail = on_error or out.on_failed

        def step(result):
            if self._exc is not None:
                fail(self._exc)
                return
            try:
                cb(result)
            except BaseException as e:  # noqa: BLE001
                fail(e)
        return self.add_listener(step)

    def done(self) -> bool:
        return self._ev.is_set()

    is_done = done

    def get(self, timeout: Optional[float] = None) -> T:
        if not self._ev.wait(timeout):
            raise TimeoutError("future timed out")
        if self._exc is not None:
            raise self._exc
        return self._result

    result = get


class CompletedFuture(ResultFuture[T]):
    """A ResultFuture that is already complete."""

    def __init__(self, result: T = None) -> None:
        super().__init__()
        self.on_completed(result)


class AggregateFuture(ResultFuture[List[T]]):
    """Completes when `n` partial results arrived (fan-out / fan-in of control messages)."""

    def __init__(self, n: int) -> None:
        super().__init__()
        self._n = n
        self._parts: List[T] = []
        self._plock = threading.Lock()
        if n == 0:
            self.on_completed([])

    def on_partial(self, result: T = None) -> None:
        with self._plock:
            self._parts.append(result)
            done = len(self._parts) >= self._n
            parts = list(self._parts)
        if done:
            self.on_completed(parts)


class CallbackRegistry:
    """identifier -> one-shot callback."""

    def __init__(self) -> None:
        self._cbs: Dict[Any, Callable[[Any], None]] = {}
        self._lock = threading.Lock()

    def register(self, ident: Any, cb: Callable[[Any], None]) -> None:
        with self._lock:
            if ident in self._cbs:
                raise KeyError(f"callback for {ident} already registered")
            self._cbs[ident] = cb

    def on_completed(self, ident: Any, result: Any = None) -> None:
        with self._lock:
            cb = self._cbs.pop(ident, None)
        if cb is None:
            raise KeyError(f"no callback registered for {ident}")
        cb(result)


class CatchableExecutor(ThreadPoolExecutor):
    """Thread pool whose task exceptions are logged (and re-raised from the future)."""

    def submit(self, fn, *args, **kwargs):  # type: ignore[override]
        def wrapped():
            try:
                return fn(*args, **kwargs)
            except BaseException:  # noqa: BLE001
                LOG.error("task failed:\n%s", traceback.format_exc())
                raise
        return super().submit(wrapped)


def new_fixed_thread_pool(n: int, name: str = "hb-pool") -> CatchableExecutor:
    return CatchableExecutor(max_workers=max(1, n), thread_name_prefix=name)


class CountDownLatch:
    """java.util.concurrent.CountDownLatch equivalent."""

    def __init__(self, count: int) -> None:
        self._count = count
        self._cond = threading.Condition()

    def count_down(self) -> None:
        with self._cond:
            self._count -= 1
            if self._count <= 0:
                self._cond.notify_all()

    def get_count(self) -> int:
        with self._cond:
            return self._count

    def await_(self, timeout: Optional[float] = None) -> bool:
        with self._cond:
            return self._cond.wait_for(lambda: self._count <= 0, timeout)


class ReadWriteLock:
    """Fair-ish reader/writer lock (writers are not starved): the per-block ownership lock."""

    def __init__(self) -> None:
        self._cond = threading.Condition()
        self._readers = 0
        self._writer = False
        self._waiting_writers = 0

    def acquire_read(self) -> None:
        with self._cond:
            while self._writer or self._waiting_writers > 0:
                self._cond.wait()
            self._readers += 1

    def release_read(self) -> None:
        with self._cond:
            self._readers -= 1
            if self._readers == 0:
                self._cond.notify_all()

    def acquire_write(self) -> None:
        with self._cond:
            self._waiting_writers += 1
            while self._writer or self._readers > 0:
                self._cond.wait()
            self._waiting_writers -= 1
            self._writer = True

    def release_write(self) -> None:
        with self._cond:
            self._writer = False
            self._cond.notify_all()
