import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import kc_oracle
    kc_oracle.build()
    return kc_oracle


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads the product library; never falls back to anything else."""
    import __graft_entry__ as ge
    ge.build()
    from katacoffee_b200 import capi
    return capi.lib()


@pytest.fixture(scope="session")
def ctx(built_lib):
    from katacoffee_b200 import backend
    c = backend.createComputeContext(0)
    yield c
    c.close()


def golden(name):
    import json
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.load(f)
