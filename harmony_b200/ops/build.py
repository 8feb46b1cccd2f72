"""In-tree build + loader of the native extension (`_hb_native*.so`).

The extension is compiled for sm_100a ONLY (`-gencode arch=compute_100a,code=sm_100a`) with
`-lineinfo` so ncu's source page maps to our code.  The built `.so` lives next to this file
(git-ignored, but it travels to the GPU box with the gpurun snapshot), and is imported
directly by path -- no JIT cache under ~/.cache is involved at run time.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_BUILD_DIR = os.path.join(_HERE, "_build")
_NAME = "_hb_native"
_lock = threading.Lock()
_module = None

CUDA_FLAGS = [
    "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--use_fast_math",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
    "-std=c++17",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")) + glob.glob(os.path.join(_CSRC, "*.cpp")))


def _so_path() -> str | None:
    cands = glob.glob(os.path.join(_BUILD_DIR, _NAME + "*.so"))
    return cands[0] if cands else None


def _stale(so: str) -> bool:
    t = os.path.getmtime(so)
    deps = sources() + glob.glob(os.path.join(_CSRC, "*.cuh")) + glob.glob(os.path.join(_CSRC, "*.h"))
    return any(os.path.getmtime(s) > t for s in deps)


def build(verbose: bool = False, force: bool = False):
    """Compile every CUDA/C++ source for sm_100a (cross-compiles without a GPU)."""
    global _module
    from torch.utils import cpp_extension

    os.makedirs(_BUILD_DIR, exist_ok=True)
    so = _so_path()
    if so and not force and not _stale(so):
        return load()
    # torch would otherwise add its own arch list; restrict it to the one we target
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", str(min(8, os.cpu_count() or 4)))
    mod = cpp_extension.load(
        name=_NAME,
        sources=sources(),
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=CUDA_FLAGS,
        extra_include_paths=[_CSRC],
        build_directory=_BUILD_DIR,
        with_cuda=True,
        verbose=verbose,
    )
    _module = mod
    return mod


def load():
    """Import the prebuilt extension by path; raises if it has not been built."""
    global _module
    with _lock:
        if _module is not None:
            return _module
        so = _so_path()
        if so is None:
            raise ImportError(
                "harmony_b200 native extension is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        import torch  # noqa: F401  (libtorch symbols must be loaded first)

        spec = importlib.util.spec_from_file_location(_NAME, so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[_NAME] = mod
        _module = mod
        return mod


def available() -> bool:
    return _so_path() is not None


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    print("built:", _so_path())
