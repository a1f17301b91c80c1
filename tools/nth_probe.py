"""Cycles of the warp nth_element replay for a few sizes (developer tool; needs a library built with -DIVG_SEL_CLOCK:
make -C iv_slam_b200/csrc OUT=../lib/var/selclk.so EXTRA=-DIVG_SEL_CLOCK; IVSLAM_GPU_LIB=.../selclk.so python tools/nth_probe.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api
rng = np.random.default_rng(1)
for n, nth in ((4, 1), (8, 3), (16, 5), (31, 10), (33, 10), (64, 12), (64, 50), (128, 20), (256, 100), (500, 433)):
    keys = rng.integers(20, 255, n).astype(np.uint32) << 8
    for _ in range(2):
        api.debug_nth_element(keys, nth)
