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
def ref():
    """oracle/_ref: the unmodified reference hot path (built by oracle/refbuild/build.sh where /root/reference exists; the
    libraries travel to the GPU box as files)."""
    from oracle import ref_lib
    if not (ref_lib.available("asbuilt") and ref_lib.available("nofma")):
        if not os.path.exists("/root/reference/introspective_ORB_SLAM/src/ORBextractor.cc"):
            pytest.skip("oracle/_ref is not built and the reference sources are not on this machine")
        ref_lib.build()
    ref_lib.lib("asbuilt"), ref_lib.lib("nofma")
    return ref_lib


@pytest.fixture(scope="session")
def gpu_api():
    """The CUDA path through the C ABI. Fails loudly (no skip, no fallback) when the library or the GPU is missing."""
    from iv_slam_b200 import api
    rc, name, sm, sms = api.device_info(0)
    assert rc == 0, "no usable sm_100 device: rc=%d name=%r sm=%d" % (rc, name, sm)
    return api
