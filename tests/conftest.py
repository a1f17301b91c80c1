import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def gpu_api():
    """The CUDA path through the C ABI. Fails loudly (no skip, no fallback) when the library or the GPU is missing."""
    from iv_slam_b200 import api
    rc, name, sm, sms = api.device_info(0)
    assert rc == 0, "no usable sm_100 device: rc=%d name=%r sm=%d" % (rc, name, sm)
    return api
