"""CPU tests (no GPU): pin the oracle (oracle/ivslam_oracle.cpp) to the real OpenCV 4.13 primitives and to the
cv2-composed restatement of the reference pipeline.  The reference has no tests or golden vectors for this path
(SURVEY §4), so this is what "parity pinned" rests on."""
import cv2
import numpy as np
import pytest

from iv_slam_b200 import synthetic as S
from oracle.cv2_pipeline import Cv2Extractor, compute_stereo_matches

from helpers import load_golden

cv2.setNumThreads(1)


@pytest.mark.parametrize("sw,sh,dw,dh", [(1241, 376, 1034, 313), (1034, 313, 862, 261), (960, 600, 800, 500),
                                          (346, 105, 288, 88), (97, 61, 81, 51), (640, 480, 640, 480), (300, 200, 251, 167)])
def test_resize_matches_cv2(oracle, sw, sh, dw, dh):
    img = np.random.default_rng(sw * 7 + dh).integers(0, 256, (sh, sw), dtype=np.uint8)
    assert np.array_equal(oracle.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("w,h", [(1241, 376), (346, 105), (64, 48), (9, 7), (37, 5)])
def test_gauss_matches_cv2(oracle, w, h):
    img = np.random.default_rng(w + h).integers(0, 256, (h, w), dtype=np.uint8)
    ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(oracle.gauss7(img), ref)


@pytest.mark.parametrize("th", [7, 12, 20, 50])
@pytest.mark.parametrize("kind", ["texture", "noise", "window"])
def test_fast_matches_cv2(oracle, th, kind):
    if kind == "texture":
        img = S.make_image(320, 200, 5)
    elif kind == "noise":
        img = np.random.default_rng(3).integers(0, 256, (120, 160), dtype=np.uint8)
    else:   # a strided ROI like the reference's cell windows
        img = S.make_image(400, 300, 6)[40:40 + 58, 100:100 + 247]
    det = cv2.FastFeatureDetector_create(th, True)
    ref = np.array([(k.pt[0], k.pt[1], k.response) for k in det.detect(np.ascontiguousarray(img))]).reshape(-1, 3)
    xs, ys, sc = oracle.fast9(img, th)
    assert np.array_equal(ref, np.stack([xs, ys, sc], 1).astype(np.float64))


def test_fast_atan2_matches_cv2(oracle):
    rng = np.random.default_rng(0)
    y = rng.integers(-1500000, 1500000, 20000).astype(np.float32)
    x = rng.integers(-1500000, 1500000, 20000).astype(np.float32)
    y[:50] = 0
    x[25:75] = 0
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(ref, oracle.fast_atan2(y, x))


def test_retain_best_is_nth_element_prefix(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = int(rng.integers(1, 80))
        resp = rng.integers(7, 40, n).astype(np.float32)
        k = int(rng.integers(0, n + 3))
        keep = oracle.retain_best(resp, k)
        if k >= n:
            assert np.array_equal(keep, np.arange(n))
        else:
            assert keep.size == k
            if k:
                cut = np.sort(resp)[::-1][k - 1]
                assert (resp[keep] >= cut).all() and (np.sort(resp[keep])[::-1] == np.sort(resp)[::-1][:k]).all()


@pytest.mark.parametrize("w,h,nf,ini,intro", [(480, 200, 600, 20, False), (400, 300, 500, 12, True), (752, 480, 1200, 20, True)])
def test_pipeline_matches_cv2_composition(oracle, w, h, nf, ini, intro):
    img = S.make_image(w, h, 21)
    cost = S.make_cost_map(w, h, 22) if intro else None
    eo = oracle.OracleExtractor(nf, 1.2, 8, ini, 7, intro)
    ec = Cv2Extractor(nf, 1.2, 8, ini, 7, intro)
    ko, do = eo(img, cost)
    kc, dc = ec(img, cost)
    assert list(eo.features_per_level()) == ec.nper
    assert list(eo.umax()) == ec.umax == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    for l in range(8):
        assert np.array_equal(eo.level(l, 0), ec.last_pyramid[l])
        if ec.last_blur[l] is not None:
            assert np.array_equal(eo.level(l, 1), ec.last_blur[l])
        if intro:
            assert np.array_equal(eo.level(l, 2), ec.last_qpyr[l])
    assert ko.size == kc.size
    for f in ko.dtype.names:
        assert np.array_equal(ko[f], kc[f]), f
    assert np.array_equal(do, dc)


@pytest.mark.parametrize("w,h,nf", [(640, 480, 1000), (1241, 376, 2000)])
def test_octree_mode_matches_independent_restatement(oracle, w, h, nf):
    """Mode 1 (ComputeKeyPointsOctTree + DistributeOctTree, dead code in the reference): the C++ oracle (std::list, like the
    reference) against the independent Python restatement built on cv2's FAST.  Both break the reference's pointer
    tie-break by creation order."""
    img = S.make_image(w, h, 33)
    eo = oracle.OracleExtractor(nf, 1.2, 8, 20, 7)
    eo.set_keypoint_mode(1)
    ec = Cv2Extractor(nf, 1.2, 8, 20, 7)
    ec.kp_mode = 1
    ko, do = eo(img)
    kc, dc = ec(img)
    assert ko.size == kc.size and nf - 50 <= ko.size <= nf + 24
    for f in ko.dtype.names:
        assert np.array_equal(ko[f], kc[f]), f
    assert np.array_equal(do, dc)
    # one keypoint per quadtree node => no duplicates inside a level
    for l in range(8):
        k = eo.level_keypoints(l)
        assert len(set(zip(k["x"].tolist(), k["y"].tolist()))) == k.size


@pytest.mark.parametrize("name", ["small_plain", "small_cost", "kitti_c1", "jackal_c2"])
def test_oracle_reproduces_golden(oracle, name):
    """tests/golden/*.npz were produced by make_golden.py from cv2 4.13; the C++ oracle must reproduce them bit-for-bit."""
    g = load_golden(name)
    nf, ini, mn, intro = (int(v) for v in g["params"])
    mbf, maxD = (float(v) for v in g["calib"])
    eL, eR = oracle.OracleExtractor(nf, 1.2, 8, ini, mn, bool(intro)), oracle.OracleExtractor(nf, 1.2, 8, ini, mn, False)
    r = oracle.stereo_frame(eL, eR, g["left"], g["right"], g.get("cost"), mbf, maxD, threads=2)
    for side, k, d in (("L", "kL", "dL"), ("R", "kR", "dR")):
        assert r[k].size == g[k].size
        for f in r[k].dtype.names:
            assert np.array_equal(r[k][f], g[k][f]), (side, f)
        assert np.array_equal(r[d], g[d])
    assert np.array_equal(r["uRight"], g["uRight"]) and np.array_equal(r["depth"], g["depth"])


def test_stereo_oracle_matches_python_restatement(oracle):
    img, right = S.make_stereo_pair(640, 240, 31)
    eL, eR = oracle.OracleExtractor(800, 1.2, 8, 20, 7), oracle.OracleExtractor(800, 1.2, 8, 20, 7)
    r = oracle.stereo_frame(eL, eR, img, right, None, 386.1448, 718.856, threads=1)
    sc = eL.scale_factors()
    inv = (np.float32(1) / sc).astype(np.float32)
    pyrL, pyrR = [eL.level(l) for l in range(8)], [eR.level(l) for l in range(8)]
    u, d = compute_stereo_matches(r["kL"], r["dL"], r["kR"], r["dR"], pyrL, pyrR, sc, inv, 386.1448, 718.856)
    assert np.array_equal(u, r["uRight"]) and np.array_equal(d, r["depth"])
    assert (u >= 0).sum() > 100


def test_oracle_edge_cases(oracle):
    e = oracle.OracleExtractor(500, 1.2, 8, 20, 7)
    k, d = e(np.full((240, 320), 128, np.uint8))          # flat image: no corners anywhere
    assert k.size == 0 and d.shape == (0, 32)
    with pytest.raises(RuntimeError):                       # grid would be 0 columns: the reference divides by zero
        e(np.zeros((60, 80), np.uint8))
    # non-contiguous rows (a ROI of a larger buffer) give the same result as a compact copy
    big = S.make_image(500, 300, 9)
    roi = big[10:250, 20:420]
    k1, d1 = e(roi)
    k2, d2 = e(np.ascontiguousarray(roi))
    assert np.array_equal(d1, d2) and all(np.array_equal(k1[f], k2[f]) for f in k1.dtype.names)


# ----------------------------------------------------------------------------- N4: input prologue (remap + cvtColor)
@pytest.mark.parametrize("cn,rgb", [(1, False), (3, False), (3, True), (4, False), (4, True)])
def test_prologue_matches_cv2(oracle, cn, rgb):
    import cv2
    rng = np.random.default_rng(40 + cn + int(rgb))
    sh, sw, h, w = 131, 203, 97, 160
    src = rng.integers(0, 256, (sh, sw, cn), dtype=np.uint8)
    if cn == 1:
        src = src[..., 0]
    code = {(3, False): cv2.COLOR_BGR2GRAY, (3, True): cv2.COLOR_RGB2GRAY, (4, False): cv2.COLOR_BGRA2GRAY, (4, True): cv2.COLOR_RGBA2GRAY}.get((cn, rgb))
    gray = (lambda a: a) if code is None else (lambda a: cv2.cvtColor(a, code))
    assert np.array_equal(oracle.prologue(src, rgb), gray(src)), "cvtColor only"
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    cases = {
        "identity": (xx[:, :sw].copy(), yy[:sh].copy()),
        "warp": ((xx * (sw / w) + rng.normal(0, 2.5, xx.shape)).astype(np.float32), (yy * (sh / h) + rng.normal(0, 2.5, xx.shape)).astype(np.float32)),
        "border": ((xx * 1.6 - 30).astype(np.float32), (yy * 1.7 - 25).astype(np.float32)),          # large parts fall outside the source
        "halves": ((xx * (sw / w) + 0.5).astype(np.float32), (yy * (sh / h) + 0.015625).astype(np.float32)),   # ties of cvRound(32*m)
    }
    bad = cases["warp"][0].copy()
    bad[::7, ::5] = np.nan
    bad[1::9, 2::6] = 1e12
    bad[2::11, ::4] = -np.inf
    cases["nonfinite"] = (bad, cases["warp"][1])
    for name, (mx, my) in cases.items():
        ref = gray(cv2.remap(src, mx, my, cv2.INTER_LINEAR))
        assert np.array_equal(oracle.prologue(src, rgb, mx, my), ref), name


def test_prologue_golden_and_real_rectification_maps(oracle):
    g = load_golden("prologue_small")
    assert np.array_equal(oracle.prologue(g["frame"], False, g["mapx"], g["mapy"]), g["gray_bgr"])
    assert np.array_equal(oracle.prologue(g["frame"], True, g["mapx"], g["mapy"]), g["gray_rgb"])
    assert np.array_equal(oracle.prologue(g["frame"], False), g["gray_noremap"])
    assert np.array_equal(oracle.prologue(g["cost"], False, g["mapx"], g["mapy"]), g["cost_remapped"])


# ----------------------------------------------------------------------------- N2: SearchByProjection restatement
def test_transform_point_matches_cv2_gemm(oracle):
    """`Rcw*x3Dw+tcw` (src/ORBmatcher.cc:1408) is one cv::gemm call on CV_32F 3x3 / 3x1 matrices."""
    rng = np.random.default_rng(9)
    for _ in range(3000):
        R = rng.normal(0, 1, (3, 3)).astype(np.float32)
        X = rng.normal(0, 20, (3, 1)).astype(np.float32)
        t = rng.normal(0, 5, (3, 1)).astype(np.float32)
        assert np.array_equal(oracle.transform_point(R, t, X), cv2.gemm(R, X, 1, t, 1)[:, 0])


def _py_features_in_area(kps, gs, gi, bounds, x, y, r, min_level, max_level):
    f = np.float32
    minX, maxX, minY, maxY = (f(b) for b in bounds)
    invW, invH = f(64) / (maxX - minX), f(48) / (maxY - minY)
    x, y, r = f(x), f(y), f(r)
    c0 = max(0, int(np.floor((x - minX - r) * invW)))
    c1 = min(63, int(np.ceil((x - minX + r) * invW)))
    r0 = max(0, int(np.floor((y - minY - r) * invH)))
    r1 = min(47, int(np.ceil((y - minY + r) * invH)))
    if c0 >= 64 or c1 < 0 or r0 >= 48 or r1 < 0:
        return []
    check = min_level > 0 or max_level >= 0
    out = []
    for ix in range(c0, c1 + 1):
        for iy in range(r0, r1 + 1):
            for j in range(gs[ix * 48 + iy], gs[ix * 48 + iy + 1]):
                k = kps[gi[j]]
                if check and (k["octave"] < min_level or (max_level >= 0 and k["octave"] > max_level)):
                    continue
                if abs(f(k["x"]) - x) < r and abs(f(k["y"]) - y) < r:
                    out.append(int(gi[j]))
    return out


def _popcount_rows(a, b):
    return int(np.unpackbits(a ^ b).sum())


def test_search_by_projection_oracle_matches_python_restatement(oracle):
    """Independent Python restatement of both SearchByProjection variants (src/ORBmatcher.cc:45-133, 1372-1519) on a small frame."""
    from helpers import projection_scenario
    f = np.float32
    w, h = 620, 300
    left, right = S.make_stereo_pair(w, h, 71)
    cur_l = np.roll(left, 2, axis=1)
    cur_r = np.roll(right, 2, axis=1)
    oL, oR = oracle.OracleExtractor(500, 1.2, 8, 20, 7), oracle.OracleExtractor(500, 1.2, 8, 20, 7)
    last = oracle.stereo_frame(oL, oR, left, right, None, 386.1448, 718.856)
    cur = oracle.stereo_frame(oL, oR, cur_l, cur_r, None, 386.1448, 718.856)
    sc = projection_scenario(last["kL"], last["dL"], last["depth"], w, h, 5, n_dup=60)
    kps, dcur, uR = cur["kL"], cur["dL"], cur["uRight"]
    _, gs, gi = oracle.frame_post(kps, None, *sc["bounds"])
    scale = oL.scale_factors()
    fx, fy, cx, cy, mbf = (f(v) for v in sc["cam"])
    for mode, th, ori in ((0, 7.0, True), (1, 15.0, True), (2, 7.0, False)):
        match = np.full(kps.size, -1, np.int32)
        blocked = np.zeros(kps.size, bool)
        hist = [[] for _ in range(30)]
        nm = 0
        for i in range(sc["flags"].size):
            if not sc["flags"][i] & 1:
                continue
            pc = cv2.gemm(sc["Rcw"], sc["world"][i].reshape(3, 1), 1, sc["tcw"].reshape(3, 1), 1)[:, 0]
            invz = f(1.0 / np.float64(pc[2]))
            if invz < 0:
                continue
            u = f(f(f(fx * pc[0]) * invz) + cx)
            v = f(f(f(fy * pc[1]) * invz) + cy)
            if u < 0 or u > w or v < 0 or v > h:
                continue
            o = int(sc["octave"][i])
            radius = f(f(th) * scale[o])
            lv = {0: (o - 1, o + 1), 1: (o, -1), 2: (0, o)}[mode]
            best, bi = 256, -1
            for i2 in _py_features_in_area(kps, gs, gi, sc["bounds"], u, v, radius, *lv):
                if blocked[i2]:
                    continue
                if uR[i2] > 0 and abs(f(f(u - f(mbf * invz)) - uR[i2])) > radius:
                    continue
                d = _popcount_rows(sc["desc"][i], dcur[i2])
                if d < best:
                    best, bi = d, i2
            if best <= 100:
                match[bi] = i
                blocked[bi] = bool(sc["flags"][i] & 2)
                nm += 1
                if ori:
                    rot = f(sc["angle"][i] - kps["angle"][bi])
                    if rot < 0:
                        rot = f(rot + f(360))
                    b = int(np.floor(f(rot * f(1.0 / 30)) + 0.5))
                    hist[0 if b == 30 else b].append(bi)
        if ori:
            sizes = [len(x) for x in hist]
            order = sorted(range(30), key=lambda b: (-sizes[b], b))          # ties keep the lower bin first, like the scan
            i1, i2_, i3 = order[:3]
            keep = {i1}
            if not sizes[i2_] < 0.1 * sizes[i1]:
                keep.add(i2_)
                if not sizes[i3] < 0.1 * sizes[i1]:
                    keep.add(i3)
            keep = {b for b in keep if sizes[b] > 0}
            for b in range(30):
                if b not in keep:
                    for idx in hist[b]:
                        match[idx] = -1
                        nm -= 1
        m, n = oracle.search_by_projection_last(kps, dcur, uR, gs, gi, scale, sc["bounds"], sc["world"], sc["desc"], sc["octave"], sc["angle"],
                                                sc["flags"], sc["Rcw"], sc["tcw"], sc["cam"], mode, th, ori)
        assert n == nm and np.array_equal(m, match), "last-frame variant mode %d" % mode
        assert nm > 50
    # local-map variant
    rng = np.random.default_rng(2)
    cur_blocked = (rng.random(kps.size) < 0.1).astype(np.uint8)
    for th in (1.0, 3.0):
        match = np.full(kps.size, -1, np.int32)
        blocked = cur_blocked.astype(bool).copy()
        nm = 0
        for i in range(sc["mflags"].size):
            if not sc["mflags"][i] & 1:
                continue
            lev = int(sc["level"][i])
            r = f(2.5) if np.float64(sc["view_cos"][i]) > 0.998 else f(4.0)
            if th != 1.0:
                r = f(r * f(th))
            rad = f(r * scale[lev])
            cand = []
            for pos, idx in enumerate(_py_features_in_area(kps, gs, gi, sc["bounds"], sc["proj"][i, 0], sc["proj"][i, 1], rad, lev - 1, lev)):
                if blocked[idx]:
                    continue
                if uR[idx] > 0 and abs(f(sc["proj"][i, 2] - uR[idx])) > rad:
                    continue
                cand.append((_popcount_rows(sc["desc"][i], dcur[idx]), pos, idx))
            cand.sort()
            if not cand or cand[0][0] > 100:
                continue
            if len(cand) > 1 and kps["octave"][cand[0][2]] == kps["octave"][cand[1][2]] and f(cand[0][0]) > f(f(0.8) * f(cand[1][0])):
                continue
            match[cand[0][2]] = i
            blocked[cand[0][2]] = bool(sc["mflags"][i] & 2)
            nm += 1
        m, n = oracle.search_by_projection_map(kps, dcur, uR, gs, gi, scale, sc["bounds"], sc["proj"], sc["view_cos"], sc["level"], sc["desc"],
                                               sc["mflags"], cur_blocked, th, 0.8)
        assert n == nm and np.array_equal(m, match), "local-map variant th %.0f" % th
        assert nm > 20


def test_search_by_bow_oracle_matches_python_restatement(oracle):
    """Independent Python restatement of ORBmatcher::SearchByBoW(pKF, F, matches) (src/ORBmatcher.cc:165-294)."""
    from helpers import bow_scenario
    f = np.float32
    w, h = 620, 300
    left, _ = S.make_stereo_pair(w, h, 73)
    e = oracle.OracleExtractor(600, 1.2, 8, 20, 7)
    k_kf, d_kf = e(left)
    k_f, d_f = e(np.roll(left, 2, axis=1))
    sc = bow_scenario(d_kf, k_kf["angle"], d_f, 7)
    for ratio, ori in ((0.7, True), (0.9, False)):
        match = np.full(k_f.size, -1, np.int32)
        hist = [[] for _ in range(30)]
        nm = 0
        for i in range(sc["flags"].size):
            s = sc["node_slot"][i]
            if not sc["flags"][i] & 1 or s < 0:
                continue
            b1, b2, bi = 256, 256, -1
            for j in range(sc["node_start"][s], sc["node_start"][s + 1]):
                idx = int(sc["node_idx"][j])
                if match[idx] >= 0:
                    continue
                d = _popcount_rows(sc["desc"][i], d_f[idx])
                if d < b1:
                    b2, b1, bi = b1, d, idx
                elif d < b2:
                    b2 = d
            if b1 <= 50 and f(b1) < f(f(ratio) * f(b2)):
                match[bi] = i
                nm += 1
                if ori:
                    rot = f(sc["angle"][i] - k_f["angle"][bi])
                    if rot < 0:
                        rot = f(rot + f(360))
                    b = int(np.floor(f(rot * f(1.0 / 30)) + 0.5))
                    hist[0 if b == 30 else b].append(bi)
        if ori:
            sizes = [len(x) for x in hist]
            order = sorted(range(30), key=lambda b: (-sizes[b], b))
            keep = {order[0]}
            if not sizes[order[1]] < 0.1 * sizes[order[0]]:
                keep.add(order[1])
                if not sizes[order[2]] < 0.1 * sizes[order[0]]:
                    keep.add(order[2])
            keep = {b for b in keep if sizes[b] > 0}
            for b in range(30):
                if b not in keep:
                    for idx in hist[b]:
                        match[idx] = -1
                        nm -= 1
        m, n = oracle.search_by_bow(k_f, d_f, sc["desc"], sc["angle"], sc["flags"], sc["node_slot"], sc["node_start"], sc["node_idx"], ratio, ori)
        assert n == nm and np.array_equal(m, match), "ratio %.1f" % ratio
        assert nm > 100
