"""iv_slam_b200 — B200-native (sm_100a) stereo front-end for IV-SLAM: ORB extraction + stereo matching.

The product is the C-ABI shared library built from iv_slam_b200/csrc (declared in include/ivslam_gpu.h);
this package is the thin Python binding (ctypes) that tests and bench.py use, mirroring the reference's
ORBextractor / Frame::ComputeStereoMatches interface.
"""
from .api import KP_DTYPE  # noqa: F401
