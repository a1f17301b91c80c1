"""N4 measurement: device time of k_prologue (remap + cvtColor fused ingest) on KITTI-shape BGR frames (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H = 1241, 376
gray = S.make_image(W, H, 5)
frames = np.repeat(np.stack([gray, np.roll(gray, 2, 1), 255 - gray], -1)[None], n, 0)
yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
mx = (xx + 3 * np.sin(yy / 50)).astype(np.float32)
my = (yy + 2 * np.cos(xx / 70)).astype(np.float32)
g = api.ORBextractor(2000, 1.2, 8, 20, 7)
for label, maps, fr in (("cvtColor only (BGR)", None, frames), ("remap + cvtColor (BGR)", (mx, my), frames), ("remap only (gray)", (mx, my), np.ascontiguousarray(frames[..., 0]))):
    g.set_rectify_maps(*(maps if maps else (None, None)))
    g.upload_raw(fr); g.sync()
    g.profile_enable(True)
    for _ in range(3):
        g.upload_raw(fr); g.sync()
    ms, cnt = g.profile_read()["k_prologue"]
    g.profile_enable(False)
    cn = 1 if fr.ndim == 3 else 3
    byts = W * H * (cn + 1 + (8 if maps else 0))
    print("%-26s %7.3f us/frame  %6.1f GB/s algorithmic (%d B/px)" % (label, ms * 1e3 / (cnt * n), byts * cnt * n / (ms * 1e-3) / 1e9, byts // (W * H)))
