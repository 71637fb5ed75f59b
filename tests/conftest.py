import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the suite needs libfiber_b200.so: build it in-tree if this checkout has not been built yet
    # (nvcc cross-compiles without a GPU; on a box without nvcc the prebuilt .so must have travelled)
    try:
        from fiber_b200 import build
        if build.stale():
            build.build()
    except Exception as e:      # noqa: BLE001
        import warnings
        warnings.warn("could not (re)build libfiber_b200.so: %s" % e)


def _have_gpu():
    try:
        import ctypes
        from fiber_b200 import _abi
        n = ctypes.c_int(0)
        return _abi.load().fbr_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (GPU tests run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json

    def load(name):
        with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as fh:
            return json.load(fh)
    return load
