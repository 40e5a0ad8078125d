import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# the flattened-index sidecar (index.load_glb) would be written next to tests/golden/*.glb
os.environ.setdefault("TEC_INDEX_CACHE", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
