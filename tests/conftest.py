import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def small_index():
    import fx
    return fx.SynthIndex("small", n_long=40, n_short=160, n_x=5, n_y=5, with_rollhash=True)


@pytest.fixture(scope="session")
def ref_required():
    import fx
    if not fx.have_ref():
        pytest.skip("oracle/_ref/libfqref.so not built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)")
    return True
