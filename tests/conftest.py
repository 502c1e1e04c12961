import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """native pieces present (the driver runs __graft_entry__.build() before the tests; do it here too)"""
    import __graft_entry__ as g

    from tiktak_b200 import lib

    if not os.path.exists(lib.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "libtiktak_oracle.so")):
        g.build()
    return True
