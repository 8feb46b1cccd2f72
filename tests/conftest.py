import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs")
    config.addinivalue_line("markers", "integration: whole-system test (multi-process / multi-thread)")
    config.addinivalue_line("markers", "intensive: long-running stress test")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
        n_gpu = torch.cuda.device_count() if has_gpu else 0
    except Exception:  # noqa: BLE001
        has_gpu, n_gpu = False, 0
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_multi = pytest.mark.skip(reason="needs >= 2 GPUs")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)
        if "multigpu" in item.keywords and n_gpu < 2:
            item.add_marker(skip_multi)


@pytest.fixture(autouse=True)
def _gpu_test_starts_on_device0(request):
    """GPU tests build tensors with `device="cuda"`: start every one of them on device 0 whatever a previous test
    (multi-executor jobs drive several GPUs from this process) left as the thread's current device."""
    if "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(0)
    yield
    if "gpu" in request.keywords:
        # collect the test's cycles (executors, trainers, streams, events) HERE, on the main thread with device 0
        # current, instead of inside whichever worker thread of the NEXT test happens to trigger the collector
        import gc
        gc.collect()
