"""GPU-side time of one eye's kernel chain at batch 1 (CUDA events on the handle's stream around run()), graph mode on/off (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S
left, right = S.make_stereo_pair(1241, 376, 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1          # frames per run (default: one)
for graph in (False, True):
    g = api.ORBextractor(2000, 1.2, 8, 20, 7)
    g.set_graph_mode(graph)
    g.upload(np.stack([left] * B)); g.sync()
    for _ in range(10): g.run()
    g.sync()
    ts = []
    for _ in range(50):
        g.timer_start(); g.run(); g.timer_stop(); ts.append(g.timer_ms())
    print("batch %d graph=%d  run() device time: median %.1f us  min %.1f us" % (B, graph, 1e3 * float(np.median(ts)), 1e3 * min(ts)))
