"""Where one eye's one-frame latency goes (developer tool): each phase of ivg_extract alone, wall clock with a sync after it,
then the whole call.  The phases overlap nothing when run like this, so their sum is an upper bound of the call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S

left, _ = S.make_stereo_pair(1241, 376, 0)
g = api.ORBextractor(2000, 1.2, 8, 20, 7)
g.set_graph_mode(True)
pin = api.PinnedArray(left.shape, np.uint8); pin.array[...] = left
cap = g.cap
kps = api.PinnedArray((1, cap), api.KP_DTYPE); desc = api.PinnedArray((1, cap, 32), np.uint8); cnt = api.PinnedArray((1,), np.int32)

def timed(fn, iters=300):
    for _ in range(20): fn()
    ts = []
    for _ in range(iters):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return 1e6 * float(np.median(ts))

for name, img in (("pageable", left), ("pinned", pin.array)):
    up = timed(lambda: (g.upload(img[None]), g.sync()))
    g.upload(img[None]); g.sync()
    run = timed(lambda: (g.run(), g.sync()))
    dl = timed(lambda: (g.download(kps.array, desc.array, cnt.array), g.sync()))
    sy = timed(lambda: g.sync())
    whole = timed(lambda: g(img))
    whole_p = timed(lambda: (g.upload(img[None]), g.run(), g.download(kps.array, desc.array, cnt.array), g.sync()))
    print("%-8s image: upload+sync %.1f  run+sync %.1f  download+sync %.1f  (bare sync %.1f)  operator() into pageable results %.1f  "
          "upload/run/download/sync into pinned results %.1f us" % (name, up, run, dl, sy, whole, whole_p))
