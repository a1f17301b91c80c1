"""ctypes loader for oracle/_ref: the UNMODIFIED reference hot path (ORBextractor.cc as a whole + Frame::ComputeStereoMatches
cut out verbatim) compiled by oracle/refbuild/build.sh against the OpenCV-compat layer.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (iv_slam_b200/) never imports it.

Two builds of the same sources:
  variant "asbuilt"  libivslam_ref.so        the reference's flags, FP contraction on (what a user's binary computes)
  variant "nofma"    libivslam_ref_nofma.so  -ffp-contract=off (the canonical float semantics the oracle restates)
The libraries are built where /root/reference exists (this container) and travel to the GPU box as files; there they
are only loaded, never rebuilt.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle_lib import KP_DTYPE, _p, _u8

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_NAMES = {"asbuilt": "libivslam_ref.so", "nofma": "libivslam_ref_nofma.so"}
_libs = {}


def build():
    """Run the recipe if the reference sources are present (no-op on the GPU box)."""
    subprocess.check_call(["bash", os.path.join(_HERE, "refbuild", "build.sh")])


def available(variant="asbuilt"):
    return os.path.exists(os.path.join(_REF_DIR, _NAMES[variant]))


def lib(variant="asbuilt"):
    if variant in _libs:
        return _libs[variant]
    path = os.path.join(_REF_DIR, _NAMES[variant])
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    vp, i32p = C.c_void_p, C.POINTER(C.c_int)
    L.ref_fp_contract.argtypes = []
    L.ref_fp_contract.restype = C.c_int
    L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ref_extractor_create.restype = vp
    L.ref_extractor_destroy.argtypes = [vp]
    L.ref_extractor_destroy.restype = None
    for name in ("ref_features_per_level", "ref_umax", "ref_scale_factors"):
        getattr(L, name).argtypes = [vp, vp]
        getattr(L, name).restype = C.c_int
    L.ref_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t, vp, vp, C.c_int, i32p]
    L.ref_extract.restype = C.c_int
    L.ref_level_size.argtypes = [vp, C.c_int, C.c_int, i32p, i32p]
    L.ref_level_size.restype = C.c_int
    L.ref_get_level.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t]
    L.ref_get_level.restype = C.c_int
    L.ref_stereo_match.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_float, C.c_float, vp, vp]
    L.ref_stereo_match.restype = C.c_int
    L.ref_stereo_frame.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t, C.c_float, C.c_float,
                                   C.c_int, vp, vp, i32p, vp, vp, i32p, vp, vp, C.c_int]
    L.ref_stereo_frame.restype = C.c_int
    L.ref_stereo_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int,
                                   C.c_size_t, C.c_float, C.c_float, C.c_int, vp, vp, vp]
    L.ref_stereo_batch.restype = C.c_int
    assert L.ref_fp_contract() == (1 if variant == "asbuilt" else 0)
    _libs[variant] = L
    return L


def max_disparity(mbf, mb):
    """maxD as Frame::ComputeStereoMatches computes it (Frame.cc:787-789): float mbf / float mb."""
    return float(np.float32(mbf) / np.float32(mb))


class RefExtractor:
    """ORB_SLAM2::ORBextractor itself (include/ORBextractor.h:54-128), same Python surface as OracleExtractor."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, enableIntrospection=False, variant="asbuilt"):
        self.L = lib(variant)
        self.variant = variant
        self.h = self.L.ref_extractor_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, int(enableIntrospection))
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.cap = nfeatures + 64

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_extractor_destroy(self.h)
            self.h = None

    def features_per_level(self):
        out = np.zeros(self.nlevels, np.int32)
        self.L.ref_features_per_level(self.h, _p(out))
        return out

    def scale_factors(self):
        out = np.zeros(self.nlevels, np.float32)
        self.L.ref_scale_factors(self.h, _p(out))
        return out

    def umax(self):
        out = np.zeros(16, np.int32)
        self.L.ref_umax(self.h, _p(out))
        return out

    def __call__(self, image, mask=None):
        image = _u8(image)
        if mask is not None:
            mask = _u8(mask)
            assert mask.shape == image.shape
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int(0)
        rc = self.L.ref_extract(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0],
                                _p(mask), mask.strides[0] if mask is not None else 0, _p(kps), _p(desc), self.cap, C.byref(n))
        if rc:
            raise RuntimeError("reference extract failed rc=%d" % rc)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def level(self, level, which=0):
        """mvImagePyramid[level] (which=0) or mvQualityImagePyramid[level] (which=2); None when empty."""
        w, h = C.c_int(), C.c_int()
        if self.L.ref_level_size(self.h, level, which, C.byref(w), C.byref(h)):
            return None
        out = np.empty((h.value, w.value), np.uint8)
        self.L.ref_get_level(self.h, level, which, _p(out), out.strides[0])
        return out


def stereo_match(left, right, kL, dL, kR, dR, mbf, mb):
    """Frame::ComputeStereoMatches -> (mvuRight, mvDepth).  Takes the reference's `mb`; compare with max_disparity(mbf, mb)."""
    kL = np.ascontiguousarray(kL, KP_DTYPE)
    kR = np.ascontiguousarray(kR, KP_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8)
    dR = np.ascontiguousarray(dR, np.uint8)
    N = kL.size
    uR = np.full(N, -1, np.float32)
    dep = np.full(N, -1, np.float32)
    rc = left.L.ref_stereo_match(left.h, right.h, _p(kL), N, _p(dL), _p(kR), kR.size, _p(dR), mbf, mb, _p(uR), _p(dep))
    if rc:
        raise RuntimeError("reference stereo failed rc=%d" % rc)
    return uR, dep


def stereo_frame(left, right, imgL, imgR, cost, mbf, mb, threads=2):
    imgL, imgR = _u8(imgL), _u8(imgR)
    assert imgL.shape == imgR.shape and imgL.strides == imgR.strides
    if cost is not None:
        cost = _u8(cost)
    cap = left.cap
    kL, kR = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dL, dR = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    uR, dep = np.full(cap, -1, np.float32), np.full(cap, -1, np.float32)
    nL, nR = C.c_int(), C.c_int()
    rc = left.L.ref_stereo_frame(left.h, right.h, _p(imgL), _p(imgR), imgL.shape[1], imgL.shape[0], imgL.strides[0],
                                 _p(cost), cost.strides[0] if cost is not None else 0, mbf, mb, cap,
                                 _p(kL), _p(dL), C.byref(nL), _p(kR), _p(dR), C.byref(nR), _p(uR), _p(dep), threads)
    if rc:
        raise RuntimeError("reference stereo_frame failed rc=%d" % rc)
    a, b = nL.value, nR.value
    return dict(kL=kL[:a], dL=dL[:a], kR=kR[:b], dR=dR[:b], uRight=uR[:a], depth=dep[:a])


def stereo_batch(params, imgsL, imgsR, mbf, mb, workers, variant="asbuilt", costs=None):
    """CPU timing: frame-parallel over `workers` threads. imgs (and optional left-eye cost-maps): [n,H,W] u8 contiguous."""
    imgsL = np.ascontiguousarray(imgsL, np.uint8)
    imgsR = np.ascontiguousarray(imgsR, np.uint8)
    if costs is not None:
        costs = np.ascontiguousarray(costs, np.uint8)
        assert costs.shape == imgsL.shape
    n, H, W = imgsL.shape
    nL = np.zeros(n, np.int32)
    nM = np.zeros(n, np.int32)
    rc = lib(variant).ref_stereo_batch(params["nfeatures"], params["scaleFactor"], params["nlevels"], params["iniThFAST"],
                                       params["minThFAST"], n, _p(imgsL), _p(imgsR), W, H, W, mbf, mb, workers, _p(nL), _p(nM), _p(costs))
    if rc:
        raise RuntimeError("reference batch failed rc=%d" % rc)
    return nL, nM
