import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.get()


@pytest.fixture(scope="session")
def golden():
    import gzip
    import json
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

    def load(name):
        path = os.path.join(d, name)
        if name.endswith(".gz"):
            return json.loads(gzip.open(path).read())
        return json.load(open(path))
    return load
