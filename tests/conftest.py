import sys
from pathlib import Path

import pytest

import os

# tests/test_sharded_gpu.py runs several "virtual ranks" on ONE device, each with its own stream, and
# their device-side waits must not share a hardware queue with the kernels they wait for: ask for 32
# queues before the CUDA context exists (the default of 8 lets unrelated streams alias)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """libemme_b200.so, built in-tree with nvcc if it is stale (cross-compiles without a GPU)."""
    from emme_b200 import build, capi
    build.build_library()
    return capi.load()


@pytest.fixture(scope="session")
def golden():
    import json
    return json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())


@pytest.fixture(scope="session")
def floor():
    import json
    return json.loads((ROOT / "tests" / "golden" / "rounding_floor.json").read_text())
