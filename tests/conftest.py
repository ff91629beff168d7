import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


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
