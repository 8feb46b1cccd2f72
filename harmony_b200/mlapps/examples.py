"""Self-validating example apps used as integration fixtures.

Reference: dolphin/examples/addvector/AddVectorTrainer.java:88,104-181 and addinteger/*: every
mini-batch pulls all keys, sleeps `compute_time_ms` (fake compute) and pushes `+delta` to every
key; cleanup() validates that every element equals delta * epochs * num_mini_batches -- which
only holds if no push is lost or duplicated, even while blocks migrate (OwnershipFirstMigrationTest).
"""
from __future__ import annotations

import os
import tempfile
import time
from typing import Any, Dict, List

import numpy as np

from ..common.dataloader import DataParser, register_parser
from ..dolphin import params as P
from ..dolphin.conf import ETDolphinConfiguration
from ..dolphin.worker import Trainer
from ..et.config import TableConfiguration
from ..utils.flags import Configuration

NUM_VALIDATE_RETRIES, VALIDATE_SLEEP_S = 30, 0.2


class ExampleDataParser(DataParser):
    """Dummy-input parser of the AddInteger / AddVector examples (a record is its line)."""

    def parse_line(self, line: str):
        return line


register_parser("example", ExampleDataParser)


def write_dummy_input(num_records: int, path: str = None) -> str:
    path = path or os.path.join(tempfile.mkdtemp(prefix="hbex"), "empty_file")
    with open(path, "w") as f:
        for i in range(max(1, num_records)):
            f.write(f"{i}\n")
    return path


class AddVectorTrainer(Trainer):
    """Sleep-based fake compute pushing a constant delta vector; validates the final sums in cleanup (self-checking
    under migration).  Reference:
    jobserver/src/main/java/edu/snu/cay/dolphin/examples/addvector/AddVectorTrainer.java:104-181."""

    def __init__(self, params: Dict[str, Any], model_accessor, **_: Any) -> None:
        self.acc = model_accessor
        self.delta = float(params.get("delta", 1))
        self.num_keys = int(params["num_keys"])
        self.vector_size = int(params.get("vector_size", 1))
        self.compute_time = int(params.get("compute_time_ms", 300)) / 1000.0
        self.keys = np.arange(self.num_keys, dtype=np.int64)
        self.expected = self.delta * int(params["max_num_epochs"]) * int(params["num_mini_batches"])

    def set_mini_batch_data(self, mini_batch) -> None:
        pass

    def pull_model(self) -> None:
        self.acc.pull_many(self.keys)

    def local_compute(self) -> None:
        time.sleep(self.compute_time)

    def push_update(self) -> None:
        self.acc.push_many(self.keys, np.full((self.num_keys, self.vector_size), self.delta, dtype=np.float32))

    def _validate(self) -> bool:
        vals = np.asarray(self.acc.pull_many(self.keys))
        return bool(np.all(vals == self.expected))

    def cleanup(self) -> None:
        for _ in range(NUM_VALIDATE_RETRIES):
            if self._validate():
                return
            time.sleep(VALIDATE_SLEEP_S)
        vals = np.asarray(self.acc.pull_many(self.keys))
        raise RuntimeError(f"Validation failed: expected {self.expected}, got min {vals.min()} max {vals.max()}")


def _vec_table(conf: Configuration, table_id: str) -> TableConfiguration:
    backend = conf.get("backend", "host")
    return TableConfiguration(table_id=table_id, dense=True, capacity=int(conf["num_keys"]),
                              row_dim=int(conf.get("vector_size", 1)), update_function="add",
                              backend=backend if backend in ("host", "device") else "host")


ADD_VECTOR = ETDolphinConfiguration(app_id="AddVector", trainer_class="harmony_b200.mlapps.examples:AddVectorTrainer",
                                    input_parser="example", model_table_conf=_vec_table, parameters=P.EXAMPLES)
ADD_VECTOR_PATH = "harmony_b200.mlapps.examples:ADD_VECTOR"


class AddIntegerTrainer(AddVectorTrainer):
    """Scalar flavour: values are python ints in a generic KV table (exercises DictBlock updates)."""

    def __init__(self, params: Dict[str, Any], model_accessor, **kw: Any) -> None:
        super().__init__(params, model_accessor, **kw)
        self.keys = list(range(self.num_keys))
        self.delta = int(params.get("delta", 1))
        self.expected = self.delta * int(params["max_num_epochs"]) * int(params["num_mini_batches"])

    def push_update(self) -> None:
        for k in self.keys:
            self.acc.push(k, self.delta)

    def _validate(self) -> bool:
        return all(v == self.expected for v in self.acc.pull_many(self.keys))


def _int_table(conf: Configuration, table_id: str) -> TableConfiguration:
    return TableConfiguration(table_id=table_id, update_function="add")


ADD_INTEGER = ETDolphinConfiguration(app_id="AddInteger",
                                     trainer_class="harmony_b200.mlapps.examples:AddIntegerTrainer",
                                     input_parser="example", model_table_conf=_int_table, parameters=P.EXAMPLES)
ADD_INTEGER_PATH = "harmony_b200.mlapps.examples:ADD_INTEGER"


def _with_input(argv: List[str]) -> List[str]:
    argv = list(argv)
    if "-input" not in argv:
        n = int(argv[argv.index("-num_training_data") + 1]) if "-num_training_data" in argv else 64
        argv += ["-input", write_dummy_input(n)]
    return argv


def main_run_addvector(argv: List[str]) -> int:
    from ..dolphin.launcher import main_run
    return main_run(ADD_VECTOR, _with_input(argv))


def main_run_addinteger(argv: List[str]) -> int:
    from ..dolphin.launcher import main_run
    return main_run(ADD_INTEGER, _with_input(argv))
