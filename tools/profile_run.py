"""Small fixed workload for ncu: N stereo pairs through one left/right extractor pair (developer tool; run under gpurun + ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L, R = S.make_stereo_batch(1241, 376, n, 100, distinct=8)
a = (2000, 1.2, 8, 20, 7)
gL, gR = api.ORBextractor(*a), api.ORBextractor(*a)
gL.upload(L); gR.upload(R)
for _ in range(steps):
    gL.run(); gR.run(); gL.sync(); gR.sync()
    rc = api.lib().ivg_stereo_match_batch(gL._h, gR._h, 386.1448, 718.856, None, None, gL.cap, 1)
    assert rc == 0
print("launches", gL.launch_count() + gR.launch_count())
