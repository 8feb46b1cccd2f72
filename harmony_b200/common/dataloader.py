"""File-split data loading + bulk loaders + parsers.

Reference: common/.../dataloader/HdfsSplitManager.java:46-80 (N exact splits of an input path),
ExactNumSplitFileInputFormat.java, HdfsSplitFetcher.java:44-95 / HdfsDataSet.java:46-65 (record
iterator of one split), services/et/.../evaluator/impl/NoneKeyBulkDataLoader.java:57-87 +
LocalKeyGenerator.java:55-113 (assign Long keys out of the key ranges of the LOCAL blocks of an
ordered table, then multiPut), ExistKeyBulkDataLoader.java:53-73, DefaultDataParser.

Local files only (`file://` prefixes accepted); HDFS is replaced by the local FS of the box.
"""
from __future__ import annotations

import os
from dataclasses import asdict, dataclass
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence


import numpy as np

from ..et.exceptions import KeyGenerationError


def strip_scheme(path: str) -> str:
    for pre in ("file://", "hdfs://"):
        if path.startswith(pre):
            return path[len(pre):]
    return path


PATH_FLAGS = ("input", "test_data_path", "metadata_path", "model_path", "local_model_path", "input_chkp_path",
              "host_to_bandwidth_file_path", "host_to_core_file_path")


def resolve_data_paths(conf):
    """Relative data paths are resolved against the current directory first and then against the launch scripts'
    directory (`HB_BIN_DIR`, exported by bin/*.sh): the EXAMPLE USAGE lines name `sample_mlr`, `inputs/adj_list`
    ... the way the reference's scripts do when run from jobserver/bin.  The job server may run in another working
    directory than the client, so the client sends absolute paths."""
    bin_dir = os.environ.get("HB_BIN_DIR")
    changed = {}
    for k in PATH_FLAGS:
        v = conf.get(k)
        if not v or not isinstance(v, str):
            continue
        p = strip_scheme(v)
        if os.path.isabs(p):
            continue
        if os.path.exists(p):
            changed[k] = os.path.abspath(p)
        elif bin_dir and os.path.exists(os.path.join(bin_dir, p)):
            changed[k] = os.path.abspath(os.path.join(bin_dir, p))
    return conf.with_values(**changed) if changed else conf


@dataclass
class SplitInfo:
    """Serialisable byte-range split of an input file.  Reference:
    common/src/main/java/edu/snu/cay/common/dataloader/HdfsSplitInfo.java, HdfsSplitManager.java:46-80."""

    path: str
    start: int
    length: int

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)

    @staticmethod
    def from_dict(d: Dict[str, Any]) -> "SplitInfo":
        return SplitInfo(d["path"], d["start"], d["length"])


def _files_of(path: str) -> List[str]:
    path = strip_scheme(path)
    if os.path.isdir(path):
        return sorted(os.path.join(path, n) for n in os.listdir(path) if not n.startswith((".", "_")))
    return [path]


def get_splits(path: str, num_splits: int) -> List[SplitInfo]:
    """Exactly `num_splits` byte-range splits over the file(s) (some may be empty)."""
    files = _files_of(path)
    sizes = [os.path.getsize(f) for f in files]
    total = sum(sizes)
    num_splits = max(1, num_splits)
    target = max(1, -(-total // num_splits))
    splits: List[SplitInfo] = []
    for f, sz in zip(files, sizes):
        off = 0
        while off < sz:
            ln = min(target, sz - off)
            splits.append(SplitInfo(f, off, ln))
            off += ln
    # merge / pad to the exact number
    while len(splits) > num_splits:
        # merge the two smallest adjacent splits of the same file
        best = None
        for i in range(len(splits) - 1):
            a, b = splits[i], splits[i + 1]
            if a.path == b.path and a.start + a.length == b.start:
                c = a.length + b.length
                if best is None or c < best[0]:
                    best = (c, i)
        if best is None:
            break
        i = best[1]
        a, b = splits[i], splits[i + 1]
        splits[i:i + 2] = [SplitInfo(a.path, a.start, a.length + b.length)]
    while len(splits) < num_splits:
        splits.append(SplitInfo(files[0], sizes[0], 0))
    return splits


NATIVE_PARSE_MIN_BYTES = 1 << 12


def read_split_bytes(split: SplitInfo) -> bytes:
    """The bytes of the lines whose FIRST byte lies inside [start, start+length)."""
    if split.length == 0:
        return b""
    with open(strip_scheme(split.path), "rb") as f:
        end = split.start + split.length
        if split.start > 0:
            f.seek(split.start - 1)
            f.readline()
        pos = f.tell()
        if pos >= end:
            return b""
        data = f.read(end - pos)
        if not data.endswith(b"\n"):
            data += f.readline()          # the last line starts inside the split: it belongs to it entirely
        return data


def read_split(split: SplitInfo) -> Iterator[str]:
    """Lines whose FIRST byte lies inside [start, start+length) (Hadoop line-record semantics)."""
    if split.length == 0:
        return
    with open(strip_scheme(split.path), "rb") as f:
        end = split.start + split.length
        if split.start > 0:
            f.seek(split.start - 1)
            f.readline()          # skip the partial (or exactly-ended) previous line
        pos = f.tell()
        while pos < end:
            line = f.readline()
            if not line:
                break
            pos = f.tell()
            yield line.decode("utf-8", errors="replace").rstrip("\r\n")


# ------------------------------------------------------------------------------------------
# parsers  (value parsers return `value`; key-value parsers return `(key, value)`)
# ------------------------------------------------------------------------------------------
class DataParser:
    """lines -> records (value parsers) or (key, value) pairs (key-value parsers); sparse text formats have a
    native multi-threaded twin (ops/csrc/loader.cpp).  Reference:
    services/et/src/main/java/edu/snu/cay/services/et/evaluator/api/DataParser.java,
    impl/DefaultDataParser.java."""

    has_key = False
    native_mode: Optional[int] = None      # ops/csrc/loader.cpp format id (0 `y i:v`, 1 `r: c,v`, 2 `w w w`)

    def __init__(self, **params: Any) -> None:
        self.params = params

    def parse_bytes(self, data: bytes) -> List[Any]:
        """Parse a whole byte range.  Sparse text formats go through the multi-threaded native parser
        (ops/csrc/loader.cpp) when the extension is built; otherwise line by line in Python."""
        if self.native_mode is not None and len(data) >= NATIVE_PARSE_MIN_BYTES:
            from ..ops import native_or_none
            nat = native_or_none()
            if nat is not None:
                try:
                    head, ptr, idx, val = nat.parse_sparse_text(data, self.native_mode, os.cpu_count() or 4)
                except ValueError as e:
                    raise ValueError(f"{type(self).__name__}: {e}") from None
                idx_np = idx.numpy()
                return self.check_index_bounds(self.from_arrays(head.numpy(), ptr.numpy(), idx_np, val.numpy()), idx_np)
        return self.parse(data.decode("utf-8", errors="replace").splitlines())

    def from_arrays(self, head, ptr, idx, val) -> List[Any]:
        """Records from the native parser's (head[n], ptr[n+1], idx[nnz], val[nnz]) arrays."""
        raise NotImplementedError

    index_bound_param: Optional[str] = None     # name of the param bounding the record indices ("features", ...)

    def record_indices(self, record: Any):
        """The index array of one record (checked against `index_bound_param`); None = nothing to check."""
        return None

    def check_index_bounds(self, records: List[Any], flat_idx=None) -> List[Any]:
        """Indices address rows of a dense model table (and raw HBM on the device backend): reject out-of-range
        input at load time instead of corrupting memory at training time."""
        bound = self.params.get(self.index_bound_param) if self.index_bound_param else None
        if not bound:
            return records
        bound = int(bound)
        if flat_idx is not None:
            bad = flat_idx.size and (int(flat_idx.min()) < 0 or int(flat_idx.max()) >= bound)
        else:
            bad = False
            for r in records:
                idx = self.record_indices(r)
                if idx is not None and len(idx) and (int(np.min(idx)) < 0 or int(np.max(idx)) >= bound):
                    bad = True
                    break
        if bad:
            raise ValueError(f"{type(self).__name__}: index out of range [0, {bound}) (-{self.index_bound_param})")
        return records

    def parse(self, lines: Iterable[str]) -> List[Any]:
        out = []
        for line in lines:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            v = self.parse_line(s)
            if v is not None:
                out.append(v)
        return self.check_index_bounds(out)

    def parse_line(self, line: str) -> Any:
        return line


class DefaultDataParser(DataParser):
    """Every non-empty, non-comment line is one value.  Reference:
    services/et/.../evaluator/impl/DefaultDataParser.java."""

    pass


_PARSERS: Dict[str, Callable[..., DataParser]] = {"default": DefaultDataParser}


def register_parser(name: str, cls: Callable[..., DataParser]) -> None:
    _PARSERS[name] = cls


def make_parser(spec: Any) -> DataParser:
    if isinstance(spec, DataParser):
        return spec
    if spec is None:
        return DefaultDataParser()
    name, params = (spec, {}) if isinstance(spec, str) else (spec["name"], spec.get("params", {}))
    if name not in _PARSERS:
        _ensure_app_parsers()
    if name not in _PARSERS and ":" in name:
        # "package.module:ParserClass" -- a user parser that an executor PROCESS has not imported yet
        import importlib
        mod, _, cls = name.partition(":")
        return getattr(importlib.import_module(mod), cls)(**params)
    if name not in _PARSERS:
        raise KeyError(f"unknown data parser {name!r} (registered: {sorted(_PARSERS)}); user parsers are named "
                       f"'module:Class' or registered with register_parser() in a module the executors import")
    return _PARSERS[name](**params)


def _ensure_app_parsers() -> None:
    # app parsers register themselves on import
    import importlib
    for mod in ("harmony_b200.mlapps.mlr", "harmony_b200.mlapps.nmf", "harmony_b200.mlapps.lda",
                "harmony_b200.mlapps.lasso", "harmony_b200.mlapps.gbt", "harmony_b200.mlapps.examples",
                "harmony_b200.pregel.graph"):
        try:
            importlib.import_module(mod)
        except ImportError:
            pass


# ------------------------------------------------------------------------------------------
# key generation + bulk loaders
# ------------------------------------------------------------------------------------------
class LocalKeyGenerator:
    """Hands out unused Long keys from the key ranges of the blocks owned by this executor."""

    def __init__(self, comps) -> None:
        self._comps = comps
        self._next: Dict[int, int] = {}

    def get_keys(self, n: int) -> List[int]:
        part = self._comps.partitioner
        blocks = self._comps.ownership.get_current_local_block_ids()
        if not blocks:
            raise KeyGenerationError("this executor owns no block of the table")
        if not hasattr(part, "key_range"):
            raise KeyGenerationError("key generation needs an ordered table")
        keys: List[int] = []
        # water-fill: after this call every local block holds (as nearly as possible) the same number of records, no
        # matter how many small splits the records arrive in -- one block is one mini-batch, so uneven blocks mean
        # uneven (or empty) mini-batches and workers that idle while others train
        done = {b: self._next.get(b, part.key_range(b)[0]) - part.key_range(b)[0] for b in blocks}
        level = -(-(sum(done.values()) + n) // len(blocks))
        for b in sorted(blocks, key=lambda x: (done[x], x)):
            lo, hi = part.key_range(b)
            cur = lo + done[b]
            take = min(max(level - done[b], 0), n - len(keys), hi - cur + 1)
            if take > 0:
                keys.extend(range(cur, cur + take))
                self._next[b] = cur + take
            if len(keys) >= n:
                break
        if len(keys) < n:                   # some block's key range is full: overflow into the others
            for b in blocks:
                lo, hi = part.key_range(b)
                cur = self._next.get(b, lo)
                take = min(n - len(keys), hi - cur + 1)
                if take > 0:
                    keys.extend(range(cur, cur + take))
                    self._next[b] = cur + take
                if len(keys) >= n:
                    break
        if len(keys) < n:
            raise KeyGenerationError(f"local key space exhausted: need {n}, got {len(keys)}")
        return keys


def run_bulk_load(executor, table_id: str, file_splits: Sequence[Dict[str, Any]]) -> int:
    comps = executor.tables.get_components(table_id)
    table = comps.table
    parser = make_parser(comps.conf.data_parser)
    total = 0
    keygen = LocalKeyGenerator(comps)
    for sd in file_splits:
        records = parser.parse_bytes(read_split_bytes(SplitInfo.from_dict(sd)))
        if not records:
            continue
        if comps.conf.bulk_loader == "exist_key" or parser.has_key:
            keys = [r[0] for r in records]
            vals = [r[1] for r in records]
        else:
            keys = keygen.get_keys(len(records))
            vals = records
        table.multi_put(keys, vals).get(timeout=300)
        total += len(records)
    return total
