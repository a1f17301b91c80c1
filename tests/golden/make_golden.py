"""Generates the golden vectors under tests/golden/ from the REAL OpenCV primitives.

Run here (build container): python tests/golden/make_golden.py
Source of truth = oracle/cv2_pipeline.py: cv2 4.13 resize / FastFeatureDetector / GaussianBlur / fastAtan2 composed in
the reference's order (src/ORBextractor.cc:1224-1296, :880-1213) + real libstdc++ std::nth_element + glibc cosf/sinf,
and its Python restatement of Frame::ComputeStereoMatches (src/Frame.cc:758-932) using cv2.norm.
Each case stores its input image(s) so the fixtures do not depend on the synthetic generator staying bit-stable.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from iv_slam_b200 import synthetic as S  # noqa: E402
from oracle.cv2_pipeline import Cv2Extractor, compute_stereo_matches  # noqa: E402

CASES = {
    # name: (w, h, seed, nfeatures, iniTh, minTh, introspection, mbf, maxD)
    "small_plain": (480, 200, 11, 600, 20, 7, False, 386.1448, 718.856),
    "small_cost": (400, 300, 12, 500, 12, 7, True, 69.690815, 528.955512),
    "kitti_c1": (1241, 376, 0, 2000, 20, 7, False, 386.1448, 718.856),
    "jackal_c2": (960, 600, 1, 2000, 12, 7, True, 69.690815, 528.955512),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def euroc_like_maps(w, h, sw, sh):
    """Undistort/rectify maps from cv2.initUndistortRectifyMap with an EuRoC-like calibration scaled to (sw, sh) -> (w, h)."""
    import cv2
    fx, fy = 0.61 * sw, 0.95 * sh
    K = np.array([[fx, 0, 0.49 * sw], [0, fy, 0.52 * sh], [0, 0, 1]])
    D = np.array([-0.2834, 0.0740, 0.00019, 1.76e-05])
    a = 0.01
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    P = np.array([[0.45 * w, 0, 0.5 * w], [0, 0.7 * h, 0.5 * h], [0, 0, 1]])
    return cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32F)


def prologue_fixture():
    """N4: raw BGR frame + cost-map -> cv2.remap(INTER_LINEAR) -> cv2.cvtColor(BGR2GRAY): inputs, maps and cv2's outputs."""
    import cv2
    rng = np.random.default_rng(21)
    sw, sh, w, h = 200, 150, 176, 128
    gray = S.make_image(sw, sh, 21)
    frame = np.stack([gray, np.roll(gray, 3, 1), 255 - gray], -1) ^ rng.integers(0, 8, (sh, sw, 3), dtype=np.uint8)
    cost = S.make_cost_map(sw, sh, 22)
    m1, m2 = euroc_like_maps(w, h, sw, sh)
    out = dict(frame=frame, cost=cost, mapx=m1, mapy=m2,
               gray_bgr=cv2.cvtColor(cv2.remap(frame, m1, m2, cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY),
               gray_rgb=cv2.cvtColor(cv2.remap(frame, m1, m2, cv2.INTER_LINEAR), cv2.COLOR_RGB2GRAY),
               gray_noremap=cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY),
               cost_remapped=cv2.remap(cost, m1, m2, cv2.INTER_LINEAR))
    np.savez_compressed(os.path.join(HERE, "prologue_small.npz"), **out)
    print("prologue_small", out["gray_bgr"].shape, "border zeros", int((out["gray_bgr"] == 0).sum()))


def main():
    prologue_fixture()
    for name, (w, h, seed, nf, ini, mn, intro, mbf, maxD) in CASES.items():
        left, right = S.make_stereo_pair(w, h, seed)
        cost = S.make_cost_map(w, h, seed + 1000) if intro else None
        eL, eR = Cv2Extractor(nf, 1.2, 8, ini, mn, intro), Cv2Extractor(nf, 1.2, 8, ini, mn, False)
        kL, dL = eL(left, cost)
        pyrL = eL.last_pyramid
        kR, dR = eR(right, None)
        pyrR = eR.last_pyramid
        sc = np.array(eL.scale, np.float32)
        inv = np.array(eL.inv, np.float32)
        u, d = compute_stereo_matches(kL, dL, kR, dR, pyrL, pyrR, sc, inv, mbf, maxD)
        out = dict(left=left, right=right, params=np.array([nf, ini, mn, int(intro)], np.int32),
                   calib=np.array([mbf, maxD], np.float32), kL=kL, dL=dL, kR=kR, dR=dR, uRight=u, depth=d,
                   pyr_sha=np.array([sha(p) for p in pyrL]), blur_sha=np.array([sha(b) if b is not None else "" for b in eL.last_blur]))
        if cost is not None:
            out["cost"] = cost
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "kL", kL.size, "kR", kR.size, "matched", int((u >= 0).sum()))


if __name__ == "__main__":
    main()
