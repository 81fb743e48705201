import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def ctx():
    """A bxw context on cuda:0 with the standard client registry bound. GPU tests only."""
    import ballxworld_b200 as bxw

    c = bxw.Context(0)
    reg = bxw.standard_registry()
    reg.bind(c)
    c.set_gen_ids(*bxw.StdGenerator.resolve_ids(reg))
    yield c
    c.close()
