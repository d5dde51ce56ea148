import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "validated_r2: components written at the end of round 1, first run on a B200 in round 2 (no longer gated)")


@pytest.fixture(scope="session")
def amgx():
    """The loaded C-ABI engine, initialised once per session (GPU tests only)."""
    from amgx_b200 import capi
    capi.load_library()
    capi.initialize()
    capi.register_print_callback(None)
    yield capi
    capi.finalize()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    o.set_num_threads(1)   # small inputs: OpenMP fork/join on a 64-core box costs more than the loops
    return o
