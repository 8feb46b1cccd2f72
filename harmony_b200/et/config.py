"""Configurations of the Elastic Table service.

Reference: services/et/.../configuration/{TableConfiguration.java:36-82,179-202,
ExecutorConfiguration, ResourceConfiguration, RemoteAccessConfiguration, TaskletConfiguration}.
All configurations are plain dataclasses that serialise to dicts (they travel inside
TableInit / TaskletStart control messages and are stored with checkpoints).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

NUM_TOTAL_BLOCKS_DEFAULT = 1024   # configuration/parameters/NumTotalBlocks.java:21-23
CHUNK_SIZE_DEFAULT = 2048         # configuration/parameters/ChunkSize.java:23
COMM_QUEUE_SIZE_DEFAULT = 1024
NUM_COMM_THREADS_DEFAULT = 2
NUM_TASKLETS_DEFAULT = 4


@dataclass
class TableConfiguration:
    """Everything that defines a table: id, blocks, mutability, ordering, update function, input / checkpoint path,
    parser, dense row layout, backend.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/configuration/TableConfiguration.java."""

    table_id: str
    # --- layout ---
    dense: bool = False               # dense slab table (bounded int key space) vs generic KV table
    capacity: int = 0                 # dense: key space is [0, capacity)
    row_dim: int = 0                  # dense: elements per row
    dtype: str = "float32"            # dense: float32 | int32
    num_total_blocks: int = NUM_TOTAL_BLOCKS_DEFAULT
    is_mutable: bool = True
    is_ordered: bool = False
    # --- behaviour ---
    update_function: Any = None       # name | {"name", "params"}
    chunk_size: int = CHUNK_SIZE_DEFAULT
    # --- data sources ---
    chkp_path: Optional[str] = None
    input_path: Optional[str] = None
    data_parser: Any = None           # name | {"name", "params"} (common/dataloader.py)
    bulk_loader: str = "none_key"     # none_key | exist_key
    # --- backend ---
    backend: str = "host"             # host | device
    user_params: Dict[str, Any] = field(default_factory=dict)

    def to_dict(self) -> Dict[str, Any]:
        d = dataclasses.asdict(self)
        uf = self.update_function
        if uf is not None and not isinstance(uf, (str, dict)):
            d["update_function"] = uf.spec()
        return d

    @staticmethod
    def from_dict(d: Dict[str, Any]) -> "TableConfiguration":
        names = {f.name for f in dataclasses.fields(TableConfiguration)}
        return TableConfiguration(**{k: v for k, v in d.items() if k in names})


@dataclass
class ResourceConfiguration:
    """Cores / memory / node (here: GPU ordinal) of an executor.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/configuration/ResourceConfiguration.java."""

    num_cores: int = 1
    mem_size_mb: int = 128
    node_names: List[str] = field(default_factory=list)
    device: Optional[int] = None      # GPU ordinal this executor is pinned to (None => CPU)


@dataclass
class RemoteAccessConfiguration:
    """Queue sizes and thread counts of the remote-access handler / sender.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/configuration/RemoteAccessConfiguration.java."""

    handler_queue_size: int = COMM_QUEUE_SIZE_DEFAULT
    sender_queue_size: int = COMM_QUEUE_SIZE_DEFAULT
    num_handler_threads: int = NUM_COMM_THREADS_DEFAULT
    num_sender_threads: int = NUM_COMM_THREADS_DEFAULT


@dataclass
class ExecutorConfiguration:
    """Tasklet slots, resources, remote-access configuration, checkpoint paths of an executor.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/configuration/ExecutorConfiguration.java."""

    resource: ResourceConfiguration = field(default_factory=ResourceConfiguration)
    remote_access: RemoteAccessConfiguration = field(default_factory=RemoteAccessConfiguration)
    num_tasklets: int = NUM_TASKLETS_DEFAULT
    chkp_temp_path: str = "/tmp/chkp"
    chkp_commit_path: str = "/tmp/chkp/commit"
    user_params: Dict[str, Any] = field(default_factory=dict)

    def to_dict(self) -> Dict[str, Any]:
        return dataclasses.asdict(self)

    @staticmethod
    def from_dict(d: Dict[str, Any]) -> "ExecutorConfiguration":
        return ExecutorConfiguration(
            resource=ResourceConfiguration(**d.get("resource", {})),
            remote_access=RemoteAccessConfiguration(**d.get("remote_access", {})),
            num_tasklets=d.get("num_tasklets", NUM_TASKLETS_DEFAULT),
            chkp_temp_path=d.get("chkp_temp_path", "/tmp/chkp"),
            chkp_commit_path=d.get("chkp_commit_path", "/tmp/chkp/commit"),
            user_params=d.get("user_params", {}))


@dataclass
class TaskletConfiguration:
    """Tasklet id, class, custom message handler, user parameters.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/configuration/TaskletConfiguration.java."""

    tasklet_id: str
    tasklet_class: str                # "module:ClassName"
    msg_handler_class: Optional[str] = None
    user_params: Dict[str, Any] = field(default_factory=dict)

    def to_dict(self) -> Dict[str, Any]:
        return dataclasses.asdict(self)

    @staticmethod
    def from_dict(d: Dict[str, Any]) -> "TaskletConfiguration":
        return TaskletConfiguration(**d)
