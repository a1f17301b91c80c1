"""Deterministic synthetic stereo inputs for tests and bench.py (SURVEY §8(d) recipe).

There is no dataset in the image and no network, so every workload is synthetic: textured
grayscale images with a realistic corner density (about 3.2 k raw FAST corners / 2 k kept at
1241x376), a right eye made by a per-row integer disparity shift, and smooth-blob cost-maps for
the introspection configuration.
"""
import numpy as np


def _cv2():
    import cv2
    return cv2


def make_image(w, h, seed):
    """u8 HxW textured image: 6 octaves of bicubic-upsampled noise + random rectangles + blur + noise."""
    cv2 = _cv2()
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), np.float32)
    for o in range(6):
        cell = 4 * (2 ** o)
        gh, gw = h // cell + 2, w // cell + 2
        g = rng.random((gh, gw), dtype=np.float32)
        up = cv2.resize(g, (gw * cell, gh * cell), interpolation=cv2.INTER_CUBIC)[:h, :w]
        acc += (2.0 ** (0.7 * o)) * up
    acc -= acc.min()
    acc = 20.0 + acc * (200.0 / max(float(acc.max()), 1e-6))
    n_rect = (w * h) // 3000
    for _ in range(n_rect):
        rw, rh = int(rng.integers(4, 41)), int(rng.integers(4, 41))
        x0, y0 = int(rng.integers(0, max(1, w - rw))), int(rng.integers(0, max(1, h - rh)))
        acc[y0:y0 + rh, x0:x0 + rw] = float(rng.integers(0, 256))
    acc = cv2.GaussianBlur(acc, (3, 3), 0.8)
    acc += rng.normal(0.0, 2.0, acc.shape).astype(np.float32)
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def shift_right_eye(left):
    """Right image = left shifted per row by d(y) = 4 + floor(40*y/H) px (a pixel at uL appears at uL - d)."""
    h, w = left.shape
    right = np.empty_like(left)
    xs = np.arange(w)
    for y in range(h):
        d = 4 + (40 * y) // h
        right[y] = left[y, np.clip(xs + d, 0, w - 1)]
    return right


def make_stereo_pair(w, h, seed):
    left = make_image(w, h, seed)
    return left, shift_right_eye(left)


def make_cost_map(w, h, seed, n_blobs=8):
    """u8 HxW introspection cost-map: Gaussian blobs (sigma 40-150 px, peak 255) on zero background."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    acc = np.zeros((h, w), np.float32)
    for _ in range(n_blobs):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        s = rng.uniform(40, 150)
        acc = np.maximum(acc, 255.0 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s)))
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def make_stereo_batch(w, h, n, seed0, distinct=None):
    """[n,H,W] left and right stacks. `distinct` limits how many images are generated from scratch; the rest are
    cheap deterministic variants (row-rolled + brightness offset) so large batches build in seconds."""
    distinct = n if distinct is None else min(distinct, n)
    base = [make_image(w, h, seed0 + i) for i in range(distinct)]
    L = np.empty((n, h, w), np.uint8)
    for i in range(n):
        b = base[i % distinct]
        k = i // distinct
        if k == 0:
            L[i] = b
        else:
            L[i] = np.clip(np.roll(b, (7 * k) % h, axis=0).astype(np.int16) + ((k * 5) % 17) - 8, 0, 255).astype(np.uint8)
    R = np.stack([shift_right_eye(L[i]) for i in range(n)])
    return L, R


# Named configurations from BASELINE.json (made concrete in SURVEY §8(d)).
CONFIGS = {
    "C1": dict(w=1241, h=376, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
               introspection=False, mbf=386.1448, maxD=718.856, seed=0),
    "C2": dict(w=960, h=600, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=12, minThFAST=7,
               introspection=True, mbf=69.690815, maxD=528.955512, seed=1, cost_seed=2),
    "C4": dict(w=3840, h=2160, nfeatures=8000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
               introspection=False, mbf=386.1448, maxD=718.856, seed=3),
}


def make_c5_stereo_stress(ex_scale_factors, features_per_level, w=1241, h=376, n=5000, seed=4):
    """C5: stereo-only stress — n synthetic keypoints per eye with dense epipolar bands.

    Returns (kL, dL, kR, dR) with cv::KeyPoint-layout records. Octaves are drawn proportionally to the extractor's
    features-per-level, level coordinates are integers inside [19, dim-19) like real extractor output, right keypoints
    are the left ones shifted by the ground-plane disparity +/- noise, descriptors are random 256-bit strings with
    k in [0,60] flipped bits on the right.
    """
    from .api import KP_DTYPE
    rng = np.random.default_rng(seed)
    sf = np.asarray(ex_scale_factors, np.float32)
    p = np.asarray(features_per_level, np.float64)
    p = p / p.sum()
    octv = rng.choice(len(sf), size=n, p=p).astype(np.int32)
    kL = np.zeros(n, KP_DTYPE)
    kR = np.zeros(n, KP_DTYPE)
    for i in range(n):
        s = sf[octv[i]]
        inv = np.float32(1.0) / s
        lw, lh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
        # keep both eyes' level-x inside [19+8, lw-19-8) so the right keypoint stays valid after the shift
        y = int(rng.integers(19, lh - 19))
        d0 = 4 + (40 * int(y * s)) // h
        dl = int(np.ceil((d0 + 3) / s))
        x = int(rng.integers(19 + dl, lw - 19))
        xr = x - int(np.rint((d0 + rng.integers(-2, 3)) / s))
        xr = min(max(xr, 19), lw - 20)
        yr = min(max(y + int(rng.integers(-1, 2)), 19), lh - 20)
        for k, (xx, yy) in ((kL, (x, y)), (kR, (xr, yr))):
            k["x"][i] = np.float32(xx) * s if octv[i] else np.float32(xx)
            k["y"][i] = np.float32(yy) * s if octv[i] else np.float32(yy)
            k["size"][i] = np.float32(int(np.float32(31) * s))
            k["angle"][i] = 0.0
            k["response"][i] = 20.0
            k["octave"][i] = octv[i]
            k["class_id"][i] = -1
    dL = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    dR = dL.copy()
    for i in range(n):
        k = int(rng.integers(0, 61))
        bits = rng.choice(256, size=k, replace=False)
        for b in bits:
            dR[i, b >> 3] ^= np.uint8(1 << (b & 7))
    perm = rng.permutation(n)
    return kL, dL, kR[perm].copy(), dR[perm].copy()
