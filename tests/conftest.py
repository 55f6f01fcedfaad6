import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle.cpu import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The compiled reference (oracle/_ref/libref.so); present here and shipped to the GPU box."""
    from oracle.cpu import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def native():
    """Builds (if needed) and loads libataxxzero.so."""
    from ataxxzero_b200 import build as az_build
    az_build.build()
    import ataxxzero_b200
    return ataxxzero_b200.lib()


@pytest.fixture(scope="session")
def ctx(native):
    import ataxxzero_b200
    c = ataxxzero_b200.Context(device=0, seed=1234)
    yield c
    c.close()


def golden_position(cls, d):
    p = cls()
    p.turn, p.ply, p.blockers = d["turn"], d["ply"], d["blockers"]
    p.pieces[0], p.pieces[1] = d["x"], d["o"]
    return p
