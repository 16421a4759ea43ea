import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    oracle_lib.build_oracle()
    return oracle_lib.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference. Present here and (prebuilt) on the GPU box; skip if never built."""
    import oracle_lib
    r = oracle_lib.load_ref()
    if r is None:
        pytest.skip("oracle/_ref/libako_ref.so not built (no /root/reference)")
    return r
