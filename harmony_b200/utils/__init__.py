from .state_machine import StateMachine, IllegalStateError  # noqa: F401
from .dag import DAG  # noqa: F401
from .flags import Param, ParamRegistry, Configuration, FlagError, config_to_argv  # noqa: F401
from .concurrent import (ResultFuture, CompletedFuture, AggregateFuture, CallbackRegistry,  # noqa: F401
                         CatchableExecutor, new_fixed_thread_pool, CountDownLatch, ReadWriteLock)
