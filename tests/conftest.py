import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_built():
    import kswtest
    kswtest.build_oracle()
    return True


@pytest.fixture(scope="session")
def gpu_ctx():
    import bwa_mem_quickassist_b200 as B
    ctx = B.KswB200(0)
    yield ctx
    ctx.close()
