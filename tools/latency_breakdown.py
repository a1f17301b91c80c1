"""Where one stereo frame's latency goes: per-kernel device time at batch 1 (developer tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S

left, right = S.make_stereo_pair(1241, 376, 0)
a = (2000, 1.2, 8, 20, 7)
gL, gR = api.ORBextractor(*a), api.ORBextractor(*a)
def frame():
    kL, dL = gL(left); kR, dR = gR(right)
    return api.compute_stereo_matches(gL, gR, 386.1448, 718.856)
for _ in range(20): frame()
t = time.perf_counter()
for _ in range(200): frame()
print("python, serial L then R then stereo: %.3f ms per frame" % ((time.perf_counter() - t) / 200 * 1e3))
gL.profile_enable(True); gR.profile_enable(True)
for _ in range(50): frame()
pl, pr = gL.profile_read(), gR.profile_read()
tot = 0
for k in pl:
    ms = (pl[k][0] + pr[k][0]) / 50
    if ms > 0:
        tot += ms
        print("%-20s %7.1f us per stereo frame (%d launches)" % (k, ms * 1e3, (pl[k][1] + pr[k][1]) // 50))
print("kernel sum %.1f us per stereo frame (both eyes serial)" % (tot * 1e3))
