"""Wall time per stereo pair: left/right handles serial vs concurrent streams, for several batch sizes (developer tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S
a = (2000, 1.2, 8, 20, 7)
for n in (128, 512, 1024):
    L, R = S.make_stereo_batch(1241, 376, n, 100, distinct=16)
    gL, gR = api.ORBextractor(*a), api.ORBextractor(*a)
    gL.upload(L); gR.upload(R); gL.sync(); gR.sync()
    def stereo(): assert api.lib().ivg_stereo_match_batch(gL._h, gR._h, 386.1448, 718.856, None, None, gL.cap, 0) == 0
    for mode in ('serial', 'concurrent'):
        def step():
            gL.run()
            if mode == 'serial': gL.sync()
            gR.run(); stereo()
        for _ in range(2): step(); gL.sync(); gR.sync()
        t = time.perf_counter()
        for _ in range(5): step()
        gL.sync(); gR.sync()
        dt = (time.perf_counter() - t) / 5
        print('batch %4d %-10s %.2f ms/step  %.2f us/pair  %.0f pairs/s' % (n, mode, dt * 1e3, dt * 1e6 / n, n / dt))
    gL.close(); gR.close()
