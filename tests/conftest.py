import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    ngpu = torch.cuda.device_count() if has_gpu else 0
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_multi = pytest.mark.skip(reason="needs >= 2 CUDA devices")
    # a deadlock in a multi-process / networked test must fail THAT test, not hang the whole run (pytest-timeout, when installed;
    # CPU tests only: the longest one takes ~15 s, GPU tests keep their own budgets)
    has_timeout = config.pluginmanager.hasplugin("timeout")
    for item in items:
        if has_timeout and "gpu" not in item.keywords and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(int(os.environ.get("AFB_TEST_TIMEOUT", "300"))))
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(skip_multi)


@pytest.fixture
def scheduler():
    from asyncframework_b200.core import Scheduler
    s = Scheduler(4)
    yield s
    s.stop()
