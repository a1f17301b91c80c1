"""Single-process multi-GPU e2e (one host thread per GPU, ONE pinned result array): pairs/s and per-GPU H2D GB/s (developer tool).
usage: tools/multi_gpu_e2e.py [n_gpus] [pairs_per_gpu] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S
from iv_slam_b200.multi import MultiGpuStereoFrontend

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
W, H = 1241, 376
params = dict(nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)
L, R = S.make_stereo_batch(W, H, B, 100, distinct=16)
pinL, pinR = api.PinnedArray((G * B, H, W), np.uint8), api.PinnedArray((G * B, H, W), np.uint8)
for g in range(G):
    pinL.array[g * B:(g + 1) * B] = L
    pinR.array[g * B:(g + 1) * B] = R
fe = MultiGpuStereoFrontend(params, W, H, list(range(G)), 256, 2)
out = fe.alloc_outputs(G * B)
for _ in range(2):
    fe.process(pinL.array, pinR.array, out, 386.1448, 718.856)
t0 = time.perf_counter()
for _ in range(steps):
    fe.process(pinL.array, pinR.array, out, 386.1448, 718.856)
dt = time.perf_counter() - t0
per = 2 * B * W * H
print("single process, %d GPUs x %d pairs: %.0f pairs/s end to end; per-GPU H2D GB/s: %s; keypoints/frame %.1f"
      % (G, B, G * B * steps / dt, " ".join("%.1f" % (per / s / 1e9) for s in fe.seconds), out["nL"].mean()))
