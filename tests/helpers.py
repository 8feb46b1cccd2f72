"""Synthetic sample-data writers used by the tests (the reference ships tiny sample files)."""
import os

import numpy as np


from harmony_b200.common.samples import (write_gbt_sample, write_lasso_sample, write_lda_sample,  # noqa: F401,E402
                                         write_mlr_sample, write_nmf_sample)


class FailingTrainer:
    """Trainer that raises in local_compute on ONE executor (fail-fast test)."""

    def __init__(self, params, model_accessor, context=None, **_):
        self.ctx = context
        self.victim = params.get("victim")

    def init_global_settings(self): pass
    def set_mini_batch_data(self, b): pass
    def pull_model(self): pass

    def local_compute(self):
        if self.ctx.executor.executor_id == self.victim:
            raise RuntimeError("injected trainer failure")

    def push_update(self): pass
    def on_epoch_finished(self, e): pass
    def cleanup(self): pass
    def summary(self): return {}


class RendezvousTrainer(FailingTrainer):
    """Trainer whose CONSTRUCTOR gathers an object from every worker of the job through the control-plane group
    (what the MLR ship mode does with its buffer addresses)."""

    def __init__(self, params, model_accessor, context=None, **kw):
        super().__init__(params, model_accessor, context=context, **kw)
        g = context.job_group
        self.seen = g.all_gather_object((g.rank, context.executor.executor_id), tag="hello")

    def local_compute(self): pass

    def summary(self):
        return {"group": [list(x) for x in self.seen]}


from harmony_b200.et.tasklet import Tasklet as _Tasklet  # noqa: E402


class EchoTasklet(_Tasklet):
    """Returns its `value` parameter (a tasklet that needs an execution slot, for the runtime tests)."""

    def run(self):
        return self.context.params.get("value")


class FlushCountingTrainer(FailingTrainer):
    """Stands in for a fused device trainer on the CPU: records when the worker loop asks it to `flush()` (wait for
    the kernels of the mini-batches issued so far) relative to the mini-batches it was handed."""

    fused_step = True
    events = []            # class-level log: ("batch",) / ("flush",) in call order, per process

    def __init__(self, params, model_accessor, context=None, **kw):
        super().__init__(params, model_accessor, context=context, **kw)
        FlushCountingTrainer.events = []

    def local_compute(self):
        pass

    def push_update(self):
        FlushCountingTrainer.events.append("batch")

    def flush(self):
        FlushCountingTrainer.events.append("flush")

    def summary(self):
        return {"events": list(FlushCountingTrainer.events)}
