import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The tests drive the product through libpedoni_cuda.so; build it (nvcc cross-compiles without a GPU)
    # if a fresh checkout has not been through __graft_entry__.build() yet. No-op when up to date.
    from pedoni_b200 import build as _build
    _build.build()


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    return oracle.lib()
