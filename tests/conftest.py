import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_gpu() -> bool:
    try:
        import ctypes
        from voxel_cone_tracing_b200 import capi
        L = capi.load()
        h = ctypes.c_void_p()
        rc = L.vct_device_create(0, ctypes.byref(h))
        if rc == 0:
            L.vct_device_destroy(h)
        return rc == 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_available():
    return _has_gpu()


def pytest_collection_modifyitems(config, items):
    # gpu tests must FAIL (not skip) on a GPU box if the library is broken; on a CPU-only box the
    # driver deselects them with -m "not gpu".  When someone runs the whole suite without a GPU we skip.
    if config.getoption("-m") and "gpu" in config.getoption("-m") and "not gpu" not in config.getoption("-m"):
        return
    if not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)
