"""Where does the end-to-end arm lose time vs the resident arm? Times the chunked pipeline with copies switched on one by one."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S
from iv_slam_b200.frontend import StereoFrontend
B, chunk, slots = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 2
params = dict(nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)
L, R = S.make_stereo_batch(1241, 376, B, 100, distinct=16)
pL, pR = api.PinnedArray(L.shape, np.uint8), api.PinnedArray(R.shape, np.uint8)
pL.array[...] = L; pR.array[...] = R
fe = StereoFrontend(params, 1241, 376, chunk, slots)
out = fe.alloc_outputs(B)
def step(up, down):
    for ci, s in enumerate(range(0, B, chunk)):
        e = s + chunk
        l, r = fe.slots[ci % slots]
        if up: l.upload(pL.array[s:e]); r.upload(pR.array[s:e])
        l.run(); r.run()
        if down:
            l.download(out["kL"][s:e], out["dL"][s:e], out["nL"][s:e]); r.download(out["kR"][s:e], out["dR"][s:e], out["nR"][s:e])
        api.compute_stereo_matches_batch(l, r, 386.1448, 718.856, out["uRight"][s:e] if down else None, out["depth"][s:e] if down else None, sync=False) if down else \
            api.lib().ivg_stereo_match_batch(l._h, r._h, 386.1448, 718.856, None, None, l.cap, 0)
step(True, True); fe.finish()
for up, down in ((False, False), (True, False), (False, True), (True, True)):
    for _ in range(2): step(up, down)
    fe.finish()
    t = time.perf_counter()
    for _ in range(5): step(up, down)
    fe.finish()
    dt = (time.perf_counter() - t) / 5
    th = time.perf_counter()
    step(up, down)
    host = time.perf_counter() - th
    fe.finish()
    print('upload %-5s download %-5s: %.2f ms/step  %.0f pairs/s   (host enqueue %.2f ms)' % (up, down, dt * 1e3, B / dt, host * 1e3))
