"""Write bytes under a base directory (reference: common/.../datastorer/{DataStorer,LocalFSDataStorer}.java)."""
from __future__ import annotations

import os


class DataStorer:
    """Write bytes under a base directory.  Reference:
    common/src/main/java/edu/snu/cay/common/datastorer/DataStorer.java:24-33."""

    def store_data(self, sub_path: str, data: bytes) -> str:
        raise NotImplementedError


class LocalFSDataStorer(DataStorer):
    """DataStorer on the local file system (-base_dir).  Reference:
    common/src/main/java/edu/snu/cay/common/datastorer/LocalFSDataStorer.java:34-55."""

    def __init__(self, base_dir: str) -> None:
        self.base_dir = base_dir

    def store_data(self, sub_path: str, data: bytes) -> str:
        path = os.path.join(self.base_dir, sub_path)
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "wb") as f:
            f.write(data)
        return path
