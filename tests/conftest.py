import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The tests bind the in-tree libb200cd.so: (re)build it when it is missing or older than its sources (a no-op
    when the stamp matches — `python -m multimodal_siamese_cd_b200.build`). Building needs nvcc, not a GPU."""
    from multimodal_siamese_cd_b200 import build
    try:
        build.build()
    except Exception as e:  # noqa: BLE001  (no nvcc on this machine: the tests that need the library fail loudly)
        sys.stderr.write(f"libb200cd.so could not be built here: {e}\n")
    yield

