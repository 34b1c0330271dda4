import sys
from pathlib import Path

import pytest

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
