"""The source-compatible C++ shim (shim/ORBextractor.{h,cc}, shim/Frame_ComputeStereoMatches.cc): compiles against a
minimal OpenCV stand-in, and on a GPU reproduces the oracle when driven the way Frame::Frame drives the reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def _build(tmp):
    from iv_slam_b200 import api
    api.lib()
    so = os.path.join(str(tmp), "shim_check.so")
    cmd = ["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-o", so,
           os.path.join(HERE, "native", "shim_check.cpp"), os.path.join(ROOT, "shim", "ORBextractor.cc"),
           os.path.join(ROOT, "shim", "Frame_ComputeStereoMatches.cc"),
           "-I", os.path.join(HERE, "fake_opencv"), "-I", os.path.join(ROOT, "shim"), "-I", os.path.join(ROOT, "include"),
           "-L", os.path.dirname(api.LIB_PATH), "-livslam_gpu", "-Wl,-rpath," + os.path.dirname(api.LIB_PATH)]
    subprocess.check_call(cmd)
    return C.CDLL(so)


def test_shim_compiles_and_refuses_to_run_without_gpu(tmp_path):
    import torch
    L = _build(tmp_path)
    rc = L.shim_construct_only()
    assert rc == (0 if torch.cuda.is_available() else -1)


@pytest.mark.gpu
def test_shim_matches_oracle(tmp_path, gpu_api, oracle):
    from iv_slam_b200 import synthetic as S
    from helpers import assert_keypoints_equal, assert_stereo_close
    L = _build(tmp_path)
    w, h, nf = 800, 400, 1000
    left, right = S.make_stereo_pair(w, h, 17)
    cap = nf + 64
    kps = np.zeros(cap, gpu_api.KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    u = np.zeros(cap, np.float32)
    d = np.zeros(cap, np.float32)
    n, levels, pw, ph = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    s1 = C.c_float()
    pyr1 = np.zeros(w * h, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.shim_stereo_frame(p(left), p(right), w, h, nf, 20, 7, C.c_float(386.1448), C.c_float(718.856), p(kps), p(desc),
                             C.byref(n), p(u), p(d), C.byref(levels), C.byref(s1), p(pyr1), C.byref(pw), C.byref(ph))
    assert rc == 0, "shim_stereo_frame rc=%d (-2..-5: ExtractStereoGPU differs from the two extractor calls + ComputeStereoMatchesGPU)" % rc
    oL, oR = oracle.OracleExtractor(nf, 1.2, 8, 20, 7), oracle.OracleExtractor(nf, 1.2, 8, 20, 7)
    r = oracle.stereo_frame(oL, oR, left, right, None, 386.1448, 718.856)
    m = n.value
    assert_keypoints_equal(kps[:m], r["kL"], "shim")
    assert np.array_equal(desc[:m], r["dL"])
    assert_stereo_close(u[:m], d[:m], r["uRight"], r["depth"], "shim")
    assert levels.value == 8 and np.float32(s1.value) == oL.scale_factors()[1]
    assert np.array_equal(pyr1[:pw.value * ph.value].reshape(ph.value, pw.value), oL.level(1))


@pytest.mark.gpu
def test_shim_extract_raw_matches_oracle(tmp_path, gpu_api, oracle):
    """N4 through the C++ shim: SetRectifyMaps + ExtractRaw == oracle remap/cvtColor + oracle extraction."""
    from iv_slam_b200 import synthetic as S
    from helpers import assert_keypoints_equal
    L = _build(tmp_path)
    sw, sh, w, h, nf = 720, 440, 672, 400, 800
    gray = S.make_image(sw, sh, 31)
    bgr = np.ascontiguousarray(np.stack([gray, np.roll(gray, 1, 0), np.roll(gray, 2, 1)], -1))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    mx = np.ascontiguousarray(xx * 1.05 + 4 * np.sin(yy / 41), np.float32)
    my = np.ascontiguousarray(yy * 1.08 + 3 * np.cos(xx / 29), np.float32)
    cap = nf + 64
    kps = np.zeros(cap, gpu_api.KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    lvl0 = np.zeros((h, w), np.uint8)
    n = C.c_int()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.shim_extract_raw(p(bgr), sw, sh, p(mx), p(my), w, h, nf, p(kps), p(desc), C.byref(n), p(lvl0))
    assert rc == 0
    want = oracle.prologue(bgr, False, mx, my)
    assert np.array_equal(lvl0, want)
    ko, do = oracle.OracleExtractor(nf, 1.2, 8, 20, 7)(want)
    assert_keypoints_equal(kps[:n.value], ko, "shim raw")
    assert np.array_equal(desc[:n.value], do)
