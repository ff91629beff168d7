import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "last: run after every other test (long oracle comparisons)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `last` go to the end, so that under `-x` the short parity tests have all run before them."""
    tail = [it for it in items if it.get_closest_marker("last")]
    if tail:
        items[:] = [it for it in items if not it.get_closest_marker("last")] + tail


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests/golden/reference_vectors.json").read_text())


@pytest.fixture(scope="session")
def port():
    from oracle_libs import load_port
    return load_port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build; None where /root/reference never existed."""
    from oracle_libs import load_ref
    return load_ref()


@pytest.fixture(scope="session")
def cb():
    """The product package; importing it loads the CUDA C-ABI library or raises."""
    import cimba_b200
    return cimba_b200
