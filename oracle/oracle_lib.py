"""ctypes loader for the CPU oracle (oracle/ivslam_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (iv_slam_b200/) never imports it.
(The unmodified reference sources built over an OpenCV-compat layer are oracle/_ref, loaded by oracle/ref_lib.py;
tests/test_ref_pin.py holds the two to bit-for-bit equality.)

The library is compiled with -march=native, so it is rebuilt per host CPU: the .so lives in
oracle/_build/<hash of the CPU flags>/ and is (re)built on first use on a new machine.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def _cpu_tag():
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:12]


def build(force=False):
    out_dir = os.path.join(_HERE, "_build", _cpu_tag())
    so = os.path.join(out_dir, "libivslam_oracle.so")
    src = os.path.join(_HERE, "ivslam_oracle.cpp")
    inc = os.path.join(_HERE, "..", "include", "ivslam_brief_pattern.inc")
    stale = (not os.path.exists(so)) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in (src, inc))
    if force or stale:
        os.makedirs(out_dir, exist_ok=True)
        cmd = ["g++", "-O3", "-march=native", "-ffp-contract=off", "-std=c++17", "-fPIC", "-Wall",
               "-pthread", "-shared", "-o", so + ".tmp%d" % os.getpid(), src]
        subprocess.check_call(cmd, cwd=_HERE)
        os.replace(so + ".tmp%d" % os.getpid(), so)
    return so


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    u8p, i32p, f32p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_void_p
    L.orc_resize_linear_u8.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, C.c_int, C.c_size_t]
    L.orc_resize_linear_u8.restype = None
    L.orc_gauss7_u8.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t]
    L.orc_gauss7_u8.restype = None
    L.orc_fast9.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, vp, vp, vp, C.c_int]
    L.orc_fast9.restype = C.c_int
    L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.orc_fast_atan2.restype = C.c_float
    L.orc_fast_atan2_array.argtypes = [vp, vp, vp, C.c_int]
    L.orc_fast_atan2_array.restype = None
    L.orc_retain_best.argtypes = [vp, vp, C.c_int, C.c_int]
    L.orc_retain_best.restype = C.c_int
    L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_extractor_create.restype = vp
    L.orc_extractor_destroy.argtypes = [vp]
    L.orc_extractor_destroy.restype = None
    L.orc_set_trig_mode.argtypes = [vp, C.c_int]
    L.orc_set_trig_mode.restype = None
    L.orc_set_keypoint_mode.argtypes = [vp, C.c_int]
    L.orc_set_keypoint_mode.restype = None
    for name in ("orc_features_per_level", "orc_scale_factors", "orc_umax"):
        getattr(L, name).argtypes = [vp, vp]
        getattr(L, name).restype = C.c_int
    L.orc_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t, vp, vp, C.c_int, i32p]
    L.orc_extract.restype = C.c_int
    L.orc_compute_pyramid.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t]
    L.orc_compute_pyramid.restype = C.c_int
    L.orc_level_size.argtypes = [vp, C.c_int, i32p, i32p]
    L.orc_level_size.restype = C.c_int
    L.orc_get_level.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t]
    L.orc_get_level.restype = C.c_int
    L.orc_get_level_keypoints.argtypes = [vp, C.c_int, vp, C.c_int]
    L.orc_get_level_keypoints.restype = C.c_int
    L.orc_level_grid.argtypes = [vp, C.c_int, vp]
    L.orc_level_grid.restype = C.c_int
    L.orc_stats.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.orc_stats.restype = None
    L.orc_stage_seconds.argtypes = [vp, vp]
    L.orc_stage_seconds.restype = None
    L.orc_frame_post.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, vp]
    L.orc_frame_post.restype = None
    L.orc_search_by_projection_last.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int] + [C.c_float] * 4 + [C.c_int, vp, vp, vp, vp, vp, vp, vp] + [C.c_float] * 5 + [C.c_int, C.c_float, C.c_int, vp]
    L.orc_search_by_projection_last.restype = C.c_int
    L.orc_search_by_projection_map.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int] + [C.c_float] * 4 + [C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp]
    L.orc_search_by_projection_map.restype = C.c_int
    L.orc_search_by_bow.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, C.c_float, C.c_int, vp]
    L.orc_search_by_bow.restype = C.c_int
    L.orc_transform_point.argtypes = [vp, vp, vp, vp]
    L.orc_transform_point.restype = None
    L.orc_prologue.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, vp, vp, C.c_size_t, vp, C.c_int, C.c_int, C.c_size_t]
    L.orc_prologue.restype = None
    L.orc_stereo_match.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_float, C.c_float, vp, vp, vp, vp]
    L.orc_stereo_match.restype = C.c_int
    L.orc_stereo_frame.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t, C.c_float, C.c_float,
                                   C.c_int, vp, vp, i32p, vp, vp, i32p, vp, vp, C.c_int]
    L.orc_stereo_frame.restype = C.c_int
    L.orc_stereo_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int,
                                   C.c_size_t, C.c_float, C.c_float, C.c_int, vp, vp]
    L.orc_stereo_batch.restype = C.c_int
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


# ---------------------------------------------------------------- primitives
def resize_linear(src, dw, dh):
    src = _u8(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
    return dst


def search_by_projection_last(kps, desc_cur, uRight, grid_start, grid_idx, scale, bounds, world_pos, desc, octave, angle, flags, Rcw, tcw,
                              cam, mode=0, th=7.0, check_orientation=True):
    """N2 restatement of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono): (match[N], nmatches)."""
    f32, i32, u8 = np.float32, np.int32, np.uint8
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    a = [np.ascontiguousarray(x, t) for x, t in ((desc_cur, u8), (uRight, f32), (grid_start, i32), (grid_idx, i32), (scale, f32), (world_pos, f32),
                                                   (desc, u8), (octave, i32), (angle, f32), (flags, u8), (Rcw, f32), (tcw, f32))]
    match = np.zeros(max(kps.size, 1), i32)
    nm = lib().orc_search_by_projection_last(_p(kps), kps.size, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), a[4].size, *[float(v) for v in bounds],
                                             a[9].size, _p(a[5]), _p(a[6]), _p(a[7]), _p(a[8]), _p(a[9]), _p(a[10]), _p(a[11]),
                                             *[float(v) for v in cam], int(mode), float(th), int(bool(check_orientation)), _p(match))
    return match[:kps.size], nm


def search_by_projection_map(kps, desc_cur, uRight, grid_start, grid_idx, scale, bounds, proj, view_cos, level, desc, flags, cur_blocked=None,
                             th=1.0, nnratio=0.8):
    """N2 restatement of ORBmatcher::SearchByProjection(F, vpMapPoints, th): (match[N], nmatches)."""
    f32, i32, u8 = np.float32, np.int32, np.uint8
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    a = [np.ascontiguousarray(x, t) for x, t in ((desc_cur, u8), (uRight, f32), (grid_start, i32), (grid_idx, i32), (scale, f32), (proj, f32),
                                                   (view_cos, f32), (level, i32), (desc, u8), (flags, u8))]
    cb = np.ascontiguousarray(cur_blocked, u8) if cur_blocked is not None else None
    match = np.zeros(max(kps.size, 1), i32)
    nm = lib().orc_search_by_projection_map(_p(kps), kps.size, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), a[4].size, *[float(v) for v in bounds],
                                            a[9].size, _p(a[5]), _p(a[6]), _p(a[7]), _p(a[8]), _p(a[9]), _p(cb) if cb is not None else None,
                                            float(th), float(nnratio), _p(match))
    return match[:kps.size], nm


def search_by_bow(kps, desc_cur, desc, angle, flags, node_slot, node_start, node_idx, nnratio=0.7, check_orientation=True):
    """N2 restatement of ORBmatcher::SearchByBoW(pKF, F, matches) on flattened inputs: (match[N], nmatches)."""
    f32, i32, u8 = np.float32, np.int32, np.uint8
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    a = [np.ascontiguousarray(x, t) for x, t in ((desc_cur, u8), (desc, u8), (angle, f32), (flags, u8), (node_slot, i32), (node_start, i32), (node_idx, i32))]
    match = np.zeros(max(kps.size, 1), i32)
    nm = lib().orc_search_by_bow(_p(kps), kps.size, _p(a[0]), a[3].size, _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), a[5].size - 1, _p(a[5]), _p(a[6]),
                                 float(nnratio), int(bool(check_orientation)), _p(match))
    return match[:kps.size], nm


def transform_point(R, t, X):
    """x3Dc = Rcw*x3Dw+tcw with cv::gemm's small-matrix float semantics."""
    R, t, X = (np.ascontiguousarray(a, np.float32) for a in (R, t, X))
    out = np.zeros(3, np.float32)
    lib().orc_transform_point(_p(R), _p(t), _p(X), _p(out))
    return out


def prologue(frame, rgb=False, mapx=None, mapy=None):
    """N4: cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) with CV_32FC1 maps (optional) then cvtColor(*2GRAY) (if 3/4 channels)."""
    frame = np.ascontiguousarray(frame, np.uint8)
    sh, sw = frame.shape[:2]
    cn = 1 if frame.ndim == 2 else frame.shape[2]
    if mapx is not None:
        mapx, mapy = np.ascontiguousarray(mapx, np.float32), np.ascontiguousarray(mapy, np.float32)
        H, W = mapx.shape
    else:
        H, W = sh, sw
    out = np.zeros((H, W), np.uint8)
    lib().orc_prologue(_p(frame), sw, sh, sw * cn, cn, int(bool(rgb)), _p(mapx) if mapx is not None else None,
                       _p(mapy) if mapy is not None else None, W, _p(out), W, H, W)
    return out


def gauss7(src):
    src = _u8(src)
    dst = np.empty_like(src)
    lib().orc_gauss7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast9(img, th, nms=True):
    """Returns (x, y, score) int32 arrays in OpenCV's emission order (row-major)."""
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    cap = max(16, img.shape[0] * img.shape[1])
    xs, ys, sc = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orc_fast9(_p(img), img.shape[1], img.shape[0], img.strides[0], int(th), int(nms), _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().orc_fast_atan2_array(_p(y), _p(x), _p(out), y.size)
    return out


def retain_best(resp, n):
    """cv::KeyPointsFilter::retainBest(v, n) + v.resize(n): returns the surviving original indices, in order."""
    r = np.ascontiguousarray(resp, np.float32).copy()
    ids = np.arange(r.size, dtype=np.int32)
    m = lib().orc_retain_best(_p(r), _p(ids), r.size, int(n))
    return ids[:m].copy()


# ---------------------------------------------------------------- extractor
class OracleExtractor:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:54-128)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, enableIntrospection=False):
        self.h = lib().orc_extractor_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, int(enableIntrospection))
        if not self.h:
            raise ValueError("bad extractor parameters")
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.cap = nfeatures + 64

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_extractor_destroy(self.h)
            self.h = None

    def set_trig_mode(self, mode):
        lib().orc_set_trig_mode(self.h, mode)

    def set_keypoint_mode(self, mode):
        """0: ComputeKeyPointsOld (the reference's live path), 1: ComputeKeyPointsOctTree (dead there, optional here)."""
        lib().orc_set_keypoint_mode(self.h, mode)
        self.cap = self.nfeatures + 64 if mode == 0 else self.nfeatures + 64 + 4 * self.nlevels

    def features_per_level(self):
        out = np.zeros(self.nlevels, np.int32)
        lib().orc_features_per_level(self.h, _p(out))
        return out

    def scale_factors(self):
        out = np.zeros(self.nlevels, np.float32)
        lib().orc_scale_factors(self.h, _p(out))
        return out

    def umax(self):
        out = np.zeros(16, np.int32)
        lib().orc_umax(self.h, _p(out))
        return out

    def __call__(self, image, mask=None):
        """operator()(image, mask, keypoints, descriptors) -> (keypoints[KP_DTYPE], descriptors[N,32])."""
        image = _u8(image)
        if mask is not None:
            mask = _u8(mask)
            assert mask.shape == image.shape
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int(0)
        rc = lib().orc_extract(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0],
                               _p(mask), mask.strides[0] if mask is not None else 0, _p(kps), _p(desc), self.cap, C.byref(n))
        if rc:
            raise RuntimeError("oracle extract failed rc=%d" % rc)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def compute_pyramid(self, image):
        image = _u8(image)
        lib().orc_compute_pyramid(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0])

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        lib().orc_level_size(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, level, which=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        rc = lib().orc_get_level(self.h, level, which, _p(out), out.strides[0])
        return out if rc == 0 else None

    def level_keypoints(self, level):
        k = np.zeros(self.cap * 2, KP_DTYPE)
        n = lib().orc_get_level_keypoints(self.h, level, _p(k), k.size)
        return k[:n].copy()

    def level_grid(self, level):
        out = np.zeros(4, np.int32)
        rc = lib().orc_level_grid(self.h, level, _p(out))
        return None if rc else tuple(int(v) for v in out)

    def stage_seconds(self):
        """Accumulated seconds per stage: pyramid, FAST, selection, orientation, blur, descriptors."""
        out = np.zeros(6, np.float64)
        lib().orc_stage_seconds(self.h, _p(out))
        return dict(zip(("pyramid", "fast", "select", "orient", "blur", "describe"), out.tolist()))

    def stats(self):
        a, b = C.c_long(), C.c_long()
        lib().orc_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value


def stereo_match(left, right, kL, dL, kR, dR, mbf, maxD, debug=False):
    """Frame::ComputeStereoMatches -> (mvuRight, mvDepth[, bestDist, bestSAD])."""
    kL = np.ascontiguousarray(kL, KP_DTYPE)
    kR = np.ascontiguousarray(kR, KP_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8)
    dR = np.ascontiguousarray(dR, np.uint8)
    N = kL.size
    uR = np.empty(N, np.float32)
    dep = np.empty(N, np.float32)
    bd = np.empty(N, np.int32)
    sad = np.empty(N, np.int32)
    rc = lib().orc_stereo_match(left.h, right.h, _p(kL), N, _p(dL), _p(kR), kR.size, _p(dR), mbf, maxD, _p(uR), _p(dep), _p(bd), _p(sad))
    if rc:
        raise RuntimeError("oracle stereo failed rc=%d" % rc)
    return (uR, dep, bd, sad) if debug else (uR, dep)


def stereo_frame(left, right, imgL, imgR, cost, mbf, maxD, threads=2):
    imgL, imgR = _u8(imgL), _u8(imgR)
    assert imgL.shape == imgR.shape and imgL.strides == imgR.strides
    if cost is not None:
        cost = _u8(cost)
    cap = left.cap
    kL, kR = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dL, dR = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    uR, dep = np.empty(cap, np.float32), np.empty(cap, np.float32)
    nL, nR = C.c_int(), C.c_int()
    rc = lib().orc_stereo_frame(left.h, right.h, _p(imgL), _p(imgR), imgL.shape[1], imgL.shape[0], imgL.strides[0],
                                _p(cost), cost.strides[0] if cost is not None else 0, mbf, maxD, cap,
                                _p(kL), _p(dL), C.byref(nL), _p(kR), _p(dR), C.byref(nR), _p(uR), _p(dep), threads)
    if rc:
        raise RuntimeError("oracle stereo_frame failed rc=%d" % rc)
    a, b = nL.value, nR.value
    return dict(kL=kL[:a], dL=dL[:a], kR=kR[:b], dR=dR[:b], uRight=uR[:a], depth=dep[:a])


def stereo_batch(params, imgsL, imgsR, mbf, maxD, workers):
    """CPU baseline: frame-parallel over `workers` threads. imgs: [n,H,W] u8 contiguous."""
    imgsL = np.ascontiguousarray(imgsL, np.uint8)
    imgsR = np.ascontiguousarray(imgsR, np.uint8)
    n, H, W = imgsL.shape
    nL = np.zeros(n, np.int32)
    nM = np.zeros(n, np.int32)
    rc = lib().orc_stereo_batch(params["nfeatures"], params["scaleFactor"], params["nlevels"], params["iniThFAST"],
                                params["minThFAST"], n, _p(imgsL), _p(imgsR), W, H, W, mbf, maxD, workers, _p(nL), _p(nM))
    if rc:
        raise RuntimeError("oracle batch failed rc=%d" % rc)
    return nL, nM


def frame_post(kps, cost, minX, maxX, minY, maxY):
    """mvKeyQualScore + AssignFeaturesToGrid (N1) -> (qual[N], gridStart[3073], gridIdx[N])."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    if cost is not None:
        cost = _u8(cost)
    qual = np.zeros(kps.size, np.float32)
    gs = np.zeros(64 * 48 + 1, np.int32)
    gi = np.zeros(max(kps.size, 1), np.int32)
    lib().orc_frame_post(_p(kps), kps.size, _p(cost), cost.strides[0] if cost is not None else 0, minX, maxX, minY, maxY, _p(qual), _p(gs), _p(gi))
    return qual, gs, gi[:kps.size]
