"""N2 measurement: latency of the two SearchByProjection variants on a KITTI-size frame, device vs the CPU oracle (developer tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from iv_slam_b200 import api, synthetic as S
from oracle import oracle_lib as O
from helpers import projection_scenario

w, h, nf = 1241, 376, 2000
left, right = S.make_stereo_pair(w, h, 81)
oL, oR = O.OracleExtractor(nf, 1.2, 8, 20, 7), O.OracleExtractor(nf, 1.2, 8, 20, 7)
last = O.stereo_frame(oL, oR, left, right, None, 386.1448, 718.856)
gL, gR = api.ORBextractor(nf, 1.2, 8, 20, 7), api.ORBextractor(nf, 1.2, 8, 20, 7)
kps, dcur = gL(np.roll(left, 3, axis=1)); gR(np.roll(right, 3, axis=1))
uR, _ = api.compute_stereo_matches(gL, gR, 386.1448, 718.856)
scale = oL.scale_factors()

def timeit(f, n):
    f(); t = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t) / n * 1e3, r

for label, ndup in (("tracking-like (no duplicated points)", 0), ("adversarial (150 duplicated points compete for keypoints)", 150)):
    print(label)
    sc = projection_scenario(last["kL"], last["dL"], last["depth"], w, h, 82, n_dup=ndup)
    gL.frame_postprocess(*sc["bounds"])
    _, gs, gi = O.frame_post(kps, None, *sc["bounds"])
    g_ms, (gm, gn) = timeit(lambda: gL.search_by_projection_last(sc["world"], sc["desc"], sc["octave"], sc["angle"], sc["flags"], sc["Rcw"], sc["tcw"], sc["cam"], sc["bounds"], 0, 7.0, True), 200)
    c_ms, (cm, cn) = timeit(lambda: O.search_by_projection_last(kps, dcur, uR[:kps.size], gs, gi, scale, sc["bounds"], sc["world"], sc["desc"], sc["octave"], sc["angle"], sc["flags"], sc["Rcw"], sc["tcw"], sc["cam"], 0, 7.0, True), 200)
    print("last-frame variant: %d points, %d matches | device %.3f ms per call (host arrays in, match[] out) | CPU oracle %.3f ms | equal %s" % (sc["flags"].size, gn, g_ms, c_ms, bool(gn == cn and np.array_equal(gm[:kps.size], cm))))
    g_ms, (gm, gn) = timeit(lambda: gL.search_by_projection_map(sc["proj"], sc["view_cos"], sc["level"], sc["desc"], sc["mflags"], sc["bounds"], None, 3.0, 0.8), 200)
    c_ms, (cm, cn) = timeit(lambda: O.search_by_projection_map(kps, dcur, uR[:kps.size], gs, gi, scale, sc["bounds"], sc["proj"], sc["view_cos"], sc["level"], sc["desc"], sc["mflags"], None, 3.0, 0.8), 200)
    print("local-map variant:  %d points, %d matches | device %.3f ms per call | CPU oracle %.3f ms | equal %s" % (sc["mflags"].size, gn, g_ms, c_ms, bool(gn == cn and np.array_equal(gm[:kps.size], cm))))
    gL.profile_enable(True)
    for _ in range(20):
        gL.search_by_projection_last(sc["world"], sc["desc"], sc["octave"], sc["angle"], sc["flags"], sc["Rcw"], sc["tcw"], sc["cam"], sc["bounds"], 0, 7.0, True)
    p = gL.profile_read()
    print("kernel time per call: candidates %.1f us, resolve %.1f us" % (p["k_proj_candidates"][0] * 1e3 / 20, p["k_proj_resolve"][0] * 1e3 / 20))
    gL.profile_enable(False)

# SearchByBoW (key frame = the last frame's keypoints, vocabulary nodes simulated by the first descriptor byte)
from helpers import bow_scenario
sc = bow_scenario(last["dL"], last["kL"]["angle"], dcur, 3, 7)      # 128 nodes: DBoW2's level-4 nodes hold tens of keypoints
g_ms, (gm, gn) = timeit(lambda: gL.search_by_bow(sc["desc"], sc["angle"], sc["flags"], sc["node_slot"], sc["node_start"], sc["node_idx"], 0.7, True), 200)
c_ms, (cm, cn) = timeit(lambda: O.search_by_bow(kps, dcur, sc["desc"], sc["angle"], sc["flags"], sc["node_slot"], sc["node_start"], sc["node_idx"], 0.7, True), 200)
print("SearchByBoW:        %d points, %d matches | device %.3f ms per call | CPU oracle %.3f ms | equal %s" % (sc["flags"].size, gn, g_ms, c_ms, bool(gn == cn and np.array_equal(gm[:kps.size], cm))))
gL.profile_enable(True)
for _ in range(20):
    gL.search_by_bow(sc["desc"], sc["angle"], sc["flags"], sc["node_slot"], sc["node_start"], sc["node_idx"], 0.7, True)
p = gL.profile_read()
print("SearchByBoW kernel time per call: match %.1f us, finish %.1f us; node sizes: max %d candidates" % (p["k_proj_candidates"][0] * 1e3 / 20, p["k_proj_resolve"][0] * 1e3 / 20, int(np.diff(sc["node_start"]).max())))
gL.profile_enable(False)
import time as _t
t = _t.perf_counter()
for _ in range(200):
    api.lib().ivg_sync(gL._h)
print("ivg_sync round trip %.1f us" % ((_t.perf_counter() - t) / 200 * 1e6))
