"""pytest configuration.  `-m "not gpu"` runs here (no GPU); `-m gpu` runs on a B200 and goes through the C ABI."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import _pkg  # noqa: E402

_pkg.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); everything else must pass on CPU")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, built if needed (nvcc cross-compiles without a GPU)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("b200_build", ROOT / "the-algorithm_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    oracle.build()
    return oracle
