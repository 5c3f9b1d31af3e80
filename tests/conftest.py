import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.basename(p)[:-len(".dump.gz")] for p in glob.glob(os.path.join(GOLDEN, "*.dump.gz")))


_cache = {}


def load_golden(name):
    from oracle import dumpio
    if name not in _cache:
        _cache[name] = dumpio.read_dump(os.path.join(GOLDEN, name + ".dump.gz"))
    return _cache[name]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package
    return load_package
