import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pf_oracle
    pf_oracle.build()
    return pf_oracle


@pytest.fixture(scope="session")
def golden():
    import json
    return json.loads((ROOT / "tests" / "golden" / "kat_v1.json").read_text())


@pytest.fixture(scope="session")
def toy_ctx(oracle):
    """N=1024 toy BFV context: three 36-bit data primes + special prime, 20-bit batching t."""
    from tests.util import toy_params
    n, primes, t = toy_params()
    return oracle.Context(n, primes, t)
