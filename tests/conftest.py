import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# the flattened-index sidecar (index.load_glb) would be written next to tests/golden/*.glb
os.environ.setdefault("TEC_INDEX_CACHE", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    try:
        import ctypes
        from te_counter_b200 import _lib
        lib = _lib.load_library()
        h = ctypes.c_void_p()
        if lib.tec_create(0, ctypes.byref(h)) != 0:
            return False
        lib.tec_destroy(h)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: skip them where tec_create fails (the CPU-only build box), so that a plain
    `pytest tests` works there too.  On a GPU box nothing is skipped."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_device_present():
        return
    import pytest
    skip = pytest.mark.skip(reason="no CUDA device (tec_create failed)")
    for it in gpu_items:
        it.add_marker(skip)
