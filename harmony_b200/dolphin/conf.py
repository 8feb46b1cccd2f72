"""App descriptor for Dolphin jobs (the ETDolphinConfiguration builder equivalent).

Reference: dolphin/core/client/ETDolphinConfiguration.java:45-64,175-278 -- Trainer, input parser,
input key/value codecs, model update function + codecs, optional local-model table (fn + codecs),
hasInputDataKey, extra parameter classes.  Codecs disappear (tables hold typed tensors / python
values); the table *shape* of a model is described by callables of the parsed flags.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional

from ..et.config import TableConfiguration
from ..utils.flags import Configuration, Param

TableConfFn = Callable[[Configuration, str], TableConfiguration]


@dataclass
class ETDolphinConfiguration:
    """App descriptor: trainer, input parser, model / local-model table factories, extra parameters (+ the GPU
    trainer).  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/core/client/ETDolphinConfiguration.java:45-64,175-278."""

    app_id: str                                   # "MLR", "NMF", ...
    trainer_class: str                            # "module:Class"
    input_parser: Any                             # parser name / spec (common/dataloader.py)
    model_table_conf: TableConfFn                 # (conf, table_id) -> TableConfiguration
    local_model_table_conf: Optional[TableConfFn] = None
    has_input_data_key: bool = False
    parameters: List[Param] = field(default_factory=list)
    device_trainer_class: Optional[str] = None    # GPU-resident trainer (fused kernels)
    device_trainer_host_model: bool = False       # the GPU trainer keeps its model in a host table (GBT forest)

    @property
    def has_local_model_table(self) -> bool:
        return self.local_model_table_conf is not None
