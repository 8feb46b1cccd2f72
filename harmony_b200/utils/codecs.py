"""Codecs: byte (de)serialisers used by checkpoints and the TCP transport.

Reference: utils/{IntegerCodec,LongCodec,NullCodec,StreamingSerializableCodec}.java and
dolphin/mlapps/serialization/{DenseVectorCodec.java:45-95 (len + floats),
SparseVectorCodec.java:41-100 (len, nnz, (idx, val)...)}, lda/SparseArrayCodec.java:58-101 (only
the non-zero (index, value) pairs of a count row travel).  Tables hold typed tensors here, so codecs
are only needed at the process boundary (checkpoint files, wire messages).
"""
from __future__ import annotations

import pickle
import struct
from typing import Any, Tuple

import numpy as np


class Codec:
    """encode / decode SPI (checkpoint files, TCP frames).  Reference: REEF Codec users in utils/."""

    def encode(self, obj: Any) -> bytes:
        raise NotImplementedError

    def decode(self, data: bytes) -> Any:
        raise NotImplementedError


class IntegerCodec(Codec):
    """4-byte big-endian int.  Reference: utils/src/main/java/edu/snu/cay/utils/IntegerCodec.java."""

    def encode(self, obj: int) -> bytes:
        return struct.pack(">i", int(obj))

    def decode(self, data: bytes) -> int:
        return struct.unpack(">i", data[:4])[0]


class LongCodec(Codec):
    """8-byte big-endian long.  Reference: utils/src/main/java/edu/snu/cay/utils/LongCodec.java."""

    def encode(self, obj: int) -> bytes:
        return struct.pack(">q", int(obj))

    def decode(self, data: bytes) -> int:
        return struct.unpack(">q", data[:8])[0]


class NullCodec(Codec):
    """Encodes nothing.  Reference: utils/src/main/java/edu/snu/cay/utils/NullCodec.java."""

    def encode(self, obj: Any) -> bytes:
        return b""

    def decode(self, data: bytes) -> None:
        return None


class SerializableCodec(Codec):
    """Generic object codec.  Reference: utils/src/main/java/edu/snu/cay/utils/StreamingSerializableCodec.java."""

    def encode(self, obj: Any) -> bytes:
        return pickle.dumps(obj, protocol=5)

    def decode(self, data: bytes) -> Any:
        return pickle.loads(data)


class DenseVectorCodec(Codec):
    """int32 length + float32 values (big endian like the reference's DataOutputStream)."""

    def encode(self, obj) -> bytes:
        v = np.asarray(obj, dtype=">f4")
        return struct.pack(">i", v.size) + v.tobytes()

    def decode(self, data: bytes) -> np.ndarray:
        n = struct.unpack(">i", data[:4])[0]
        return np.frombuffer(data, dtype=">f4", count=n, offset=4).astype(np.float32)


class SparseVectorCodec(Codec):
    """int32 length, int32 nnz, then nnz x (int32 index, float32 value)."""

    def encode(self, obj: Tuple[int, np.ndarray, np.ndarray]) -> bytes:
        length, idx, val = obj
        idx = np.asarray(idx, dtype=">i4")
        val = np.asarray(val, dtype=">f4")
        body = np.empty(len(idx), dtype=[("i", ">i4"), ("v", ">f4")])
        body["i"], body["v"] = idx, val
        return struct.pack(">ii", int(length), len(idx)) + body.tobytes()

    def decode(self, data: bytes) -> Tuple[int, np.ndarray, np.ndarray]:
        length, nnz = struct.unpack(">ii", data[:8])
        body = np.frombuffer(data, dtype=[("i", ">i4"), ("v", ">f4")], count=nnz, offset=8)
        return length, body["i"].astype(np.int32), body["v"].astype(np.float32)


class SparseArrayCodec(Codec):
    """LDA count rows: int32 number of non-zeros then (index, value) pairs of the NON-ZERO entries;
    the trailing nnz slot of the row is rebuilt on decode."""

    def __init__(self, num_topics: int) -> None:
        self.k = num_topics

    def encode(self, row) -> bytes:
        row = np.asarray(row, dtype=np.int32)
        nz = np.flatnonzero(row[: self.k])
        pairs = np.empty(2 * len(nz), dtype=">i4")
        pairs[0::2], pairs[1::2] = nz, row[nz]
        return struct.pack(">i", len(nz)) + pairs.tobytes()

    def decode(self, data: bytes) -> np.ndarray:
        n = struct.unpack(">i", data[:4])[0]
        pairs = np.frombuffer(data, dtype=">i4", count=2 * n, offset=4)
        row = np.zeros(self.k + 1, dtype=np.int32)
        row[pairs[0::2]] = pairs[1::2]
        row[self.k] = n
        return row
