"""One stereo frame at a time through two handles (serial L, R, stereo): workload for an ncu launch list at batch 1 (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iv_slam_b200 import api, synthetic as S
left, right = S.make_stereo_pair(1241, 376, 0)
a = (2000, 1.2, 8, 20, 7)
gL, gR = api.ORBextractor(*a), api.ORBextractor(*a)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    gL(left); gR(right)
    api.compute_stereo_matches(gL, gR, 386.1448, 718.856)
