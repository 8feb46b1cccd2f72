"""Checkpoint / restore of tables (executor slave + driver master).

Reference: evaluator/impl/ChkpManagerSlave.java:146-201 (write conf + one file per local block
under the table write-lock, optional sampling ratio), :203-220 (block file = int n, n x (key,value)),
:226-293 (load from temp / commit / arbitrary path by multiPut => re-blocking works), :467-484
(commit temp -> commit path on context stop); driver/impl/ChkpManagerMaster.java:142-168
(chkpId = <tableId>-<ms>), :177-245 (tracking temp vs committed blocks), :309-418 (restore with a
different number of blocks / executors).

File format (one directory per checkpoint id):
    conf.json                 -- serialised TableConfiguration
    <blockId>.blk             -- b"HBCK" + pickle protocol 5 of {"keys": ..., "values": ...}
"""
from __future__ import annotations

import json
import logging
import os
import pickle
import random
import shutil
import threading
from typing import Dict, List, Optional, Sequence

import numpy as np

from .config import TableConfiguration

LOG = logging.getLogger(__name__)
MAGIC = b"HBCK"


def write_block_file(path: str, keys, values) -> None:
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        pickle.dump({"keys": keys, "values": values}, f, protocol=5)
    os.replace(tmp, path)


def read_block_file(path: str):
    with open(path, "rb") as f:
        if f.read(4) != MAGIC:
            raise IOError(f"{path} is not a harmony_b200 block file")
        d = pickle.load(f)
    return d["keys"], d["values"]


def write_conf(dir_path: str, conf: TableConfiguration) -> None:
    with open(os.path.join(dir_path, "conf.json"), "w") as f:
        json.dump(conf.to_dict(), f, default=str)


def read_conf(dir_path: str) -> TableConfiguration:
    with open(os.path.join(dir_path, "conf.json")) as f:
        return TableConfiguration.from_dict(json.load(f))


def _sample(keys, values, ratio: float):
    if ratio >= 1.0 or len(keys) == 0:
        return keys, values
    idx = [i for i in range(len(keys)) if random.random() < ratio]
    if isinstance(keys, np.ndarray):
        return keys[idx], values[idx]
    return [keys[i] for i in idx], [values[i] for i in idx]


class ChkpManagerSlave:
    """Executor side of checkpointing: conf + one file per local block (optionally sampled) into the temp dir under
    the write lock, commit on stop, load with re-blocking.  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/impl/ChkpManagerSlave.java:146-201,203-220,226-293,346-455,467-484."""

    def __init__(self, executor, temp_path: str, commit_path: str) -> None:
        self._ex = executor
        self.temp_root = os.path.join(temp_path, executor.executor_id)
        self.commit_root = commit_path
        self._temp_chkps: Dict[str, List[int]] = {}
        self._lock = threading.Lock()

    def _temp_dir(self, chkp_id: str) -> str:
        return os.path.join(self.temp_root, chkp_id)

    def _commit_dir(self, chkp_id: str) -> str:
        return os.path.join(self.commit_root, chkp_id)

    def checkpoint(self, chkp_id: str, table_id: str, sampling_ratio: float = 1.0) -> List[int]:
        comps = self._ex.tables.get_components(table_id)
        d = self._temp_dir(chkp_id)
        os.makedirs(d, exist_ok=True)
        write_conf(d, comps.conf)
        done: List[int] = []
        for block_id in comps.block_store.block_ids():
            unlock = comps.ownership.hold_write_lock(block_id)   # no concurrent updates while snapshotting
            try:
                block = comps.block_store.try_get(block_id)
                if block is None:
                    continue
                keys, values = block.export_items()
            finally:
                unlock()
            keys, values = _sample(keys, values, sampling_ratio)
            write_block_file(os.path.join(d, f"{block_id}.blk"), keys, values)
            done.append(block_id)
        with self._lock:
            self._temp_chkps[chkp_id] = done
        return done

    def commit(self, chkp_id: str) -> None:
        src, dst = self._temp_dir(chkp_id), self._commit_dir(chkp_id)
        if not os.path.isdir(src):
            return
        os.makedirs(dst, exist_ok=True)
        for name in os.listdir(src):
            shutil.move(os.path.join(src, name), os.path.join(dst, name))
        shutil.rmtree(src, ignore_errors=True)
        with self._lock:
            self._temp_chkps.pop(chkp_id, None)

    def commit_all_local_chkps(self) -> None:
        with self._lock:
            ids = list(self._temp_chkps)
        for cid in ids:
            self.commit(cid)

    def load(self, chkp_id_or_path: str, table_id: str, block_ids: Sequence[int], committed: bool) -> int:
        """Load the given block files into `table_id` through multi_put (re-blocking safe)."""
        if os.path.isdir(chkp_id_or_path):
            d = chkp_id_or_path
        else:
            d = self._commit_dir(chkp_id_or_path) if committed else self._temp_dir(chkp_id_or_path)
        table = self._ex.tables.get_table(table_id)
        n = 0
        for b in block_ids:
            path = os.path.join(d, f"{b}.blk")
            if not os.path.exists(path):
                # another executor's temp dir (same host): search siblings
                path = _find_block_file(os.path.dirname(self.temp_root), chkp_id_or_path, b) or path
            keys, values = read_block_file(path)
            if len(keys):
                table.multi_put(keys, values).get(timeout=120)
                n += len(keys)
        return n


def _find_block_file(temp_parent: str, chkp_id: str, block_id: int) -> Optional[str]:
    if not os.path.isdir(temp_parent):
        return None
    for ex in os.listdir(temp_parent):
        p = os.path.join(temp_parent, ex, chkp_id, f"{block_id}.blk")
        if os.path.exists(p):
            return p
    return None


def list_chkp_blocks(dir_path: str) -> List[int]:
    return sorted(int(n[:-4]) for n in os.listdir(dir_path) if n.endswith(".blk"))
