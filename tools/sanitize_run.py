"""Small workload touching every kernel (extraction with and without cost-map, octree mode, stereo, N1, N2, N4) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from iv_slam_b200 import api, synthetic as S

w, h = 640, 360
L, R = S.make_stereo_batch(w, h, 2, 5, distinct=2)
cost = np.stack([S.make_cost_map(w, h, 6)] * 2)
a = (800, 1.2, 8, 20, 7)
gL, gR = api.ORBextractor(*a, True), api.ORBextractor(*a)
gL.upload(L, cost); gR.upload(R)
gL.run(); gR.run(); gL.sync(); gR.sync()
u, d = api.compute_stereo_matches_batch(gL, gR, 100.0, 400.0)
q, gs, gi = gL.frame_postprocess(0, w, 0, h)
k = np.zeros((2, gL.cap), api.KP_DTYPE); de = np.zeros((2, gL.cap, 32), np.uint8); n = np.zeros(2, np.int32)
gL.download(k, de, n); gL.sync()
m = n[0]
rng = np.random.default_rng(0)
world = np.stack([(k["x"][0, :m] - w / 2) * 5 / 500, (k["y"][0, :m] - h / 2) * 5 / 500, np.full(m, 5.0)], 1).astype(np.float32)
flags = rng.integers(0, 4, m).astype(np.uint8) | 1
cam = (500.0, 500.0, w / 2, h / 2, 100.0)
match, nm = gL.search_by_projection_last(world, de[0, :m], k["octave"][0, :m], k["angle"][0, :m], flags, np.eye(3, dtype=np.float32), np.array([0.01, 0, 0], np.float32), cam, (0, w, 0, h), 0, 7.0, True)
proj = np.stack([k["x"][0, :m] + 1, k["y"][0, :m], k["x"][0, :m] - 3], 1).astype(np.float32)
match2, nm2 = gL.search_by_projection_map(proj, np.full(m, 0.999, np.float32), k["octave"][0, :m], de[0, :m], flags, (0, w, 0, h), None, 3.0, 0.8)
from helpers import bow_scenario
bs = bow_scenario(de[0, :m], k["angle"][0, :m], de[0, :m], 1)
match3, nm3 = gL.search_by_bow(bs["desc"], bs["angle"], bs["flags"], bs["node_slot"], bs["node_start"], bs["node_idx"], 0.7, True)
print("stereo matches", int((u >= 0).sum()), "proj", nm, nm2, "bow", nm3)
gL.set_keypoint_mode(1); gL.upload(L, cost); gL.run(); gL.sync(); gL.set_keypoint_mode(0)
yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
gL.set_rectify_maps(xx * 1.01 - 3, yy * 0.99 + 2)
bgr = np.stack([L[0], L[0], L[0]], -1).copy()
kk, dd = gL.extract_raw(bgr, False, cost[0])
print("raw keypoints", kk.size)
# one frame at a time on a linked pair (graph mode): fused pyramid, 16-slot describe, forked blur, CTA-wide level trim,
# eager stereo index and the speculative matcher, from two host threads
import threading
sL, sR = api.ORBextractor(*a), api.ORBextractor(*a)
sL.set_graph_mode(True); sR.set_graph_mode(True)
for i in range(3):
    tl = threading.Thread(target=lambda: sL(L[i % 2])); tr = threading.Thread(target=lambda: sR(R[i % 2]))
    tl.start(); tr.start(); tl.join(); tr.join()
    us, ds = api.compute_stereo_matches(sL, sR, 100.0, 400.0)
print("single-frame stereo matches", int((us >= 0).sum()))
# the one-call front-end (both eyes + matcher queued from one thread), then both kernel configurations by force
for i in range(3):
    kL1, dL1, kR1, dR1, u1, d1 = api.extract_stereo(sL, sR, L[i % 2], R[i % 2], 100.0, 400.0)
print("one-call stereo matches", int((u1 >= 0).sum()))
for mode in (1, 2, 0):
    sL.debug_force_config(mode); sR.debug_force_config(mode)
    sL(L[0]); sR(R[0])
    us, ds = api.compute_stereo_matches(sL, sR, 100.0, 400.0)
    print("forced configuration", mode, "stereo matches", int((us >= 0).sum()))
