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


def fuzz_case(seed):
    """Seeded fuzz geometry shared by the oracle-vs-reference CPU tests and the GPU parity tests: image size, feature count,
    scale factor, level count, threshold, introspection on/off; every fourth case is a pure-noise image."""
    from iv_slam_b200 import synthetic as S
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(160, 1400)), int(rng.integers(120, 900))
    nf = int(rng.integers(150, 4000))
    sf = float(rng.choice([1.2, 1.2, 1.1, 1.3, 1.5]))
    nl = int(rng.integers(3, 9))
    ini = int(rng.choice([12, 20, 20, 35]))
    intro = bool(rng.integers(0, 2))
    noise = seed % 4 == 3                      # a few pure-noise images: almost every pixel passes the FAST reject test
    if noise:
        left = rng.integers(0, 256, (h, w), dtype=np.uint8)
        right = np.roll(left, -7, axis=1)
    else:
        left, right = S.make_stereo_pair(w, h, 2000 + seed)
    cost = S.make_cost_map(w, h, 3000 + seed) if intro else None
    what = "fuzz%d %dx%d nf%d sf%.1f nl%d ini%d intro%d" % (seed, w, h, nf, sf, nl, ini, intro)
    return dict(w=w, h=h, nf=nf, sf=sf, nl=nl, ini=ini, intro=intro, left=left, right=right, cost=cost, what=what)


def reference_mb(mbf, maxD):
    """The reference's `mb` member for a wanted maxD, and the maxD it then really uses: Frame::ComputeStereoMatches computes
    maxD = mbf/mb in float (Frame.cc:787-789), so a comparison with the unmodified reference must feed every
    implementation that value."""
    mb = np.float32(mbf) / np.float32(maxD)
    return float(mb), float(np.float32(mbf) / mb)


# ----------------------------------------------------------------------------- N2 scenarios (SearchByProjection)
PROJ_CAM = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, mbf=386.1448)


def projection_scenario(kps_last, desc_last, depth_last, w, h, seed, n_dup=150):
    """Flattened LastFrame / local-map inputs for the N2 matchers from a last frame's keypoints and stereo depths:
    world points by back-projection (last pose = identity), a small camera motion, random outlier / observation flags and
    `n_dup` duplicated points (same world position, slightly perturbed descriptor) so that several points compete for one
    current keypoint — that is what exercises the order-dependent blocking of the reference loops."""
    rng = np.random.default_rng(seed)
    c = PROJ_CAM
    n0 = kps_last.size
    z = np.where(depth_last > 0, depth_last, 1.0).astype(np.float32)
    world = np.stack([(kps_last["x"] - c["cx"]) * z / c["fx"], (kps_last["y"] - c["cy"]) * z / c["fy"], z], 1).astype(np.float32)
    flags = np.where(depth_last > 0, 1, 0).astype(np.uint8)
    flags &= (rng.random(n0) > 0.1).astype(np.uint8)                    # outliers / no MapPoint
    flags |= (rng.random(n0) < 0.7).astype(np.uint8) << 1                # Observations() > 0
    dup = rng.choice(n0, n_dup, replace=False)
    ddesc = desc_last[dup].copy()
    for r in range(n_dup):                                               # flip 0..5 random bits
        for b in rng.integers(0, 256, rng.integers(0, 6)):
            ddesc[r, b >> 3] ^= np.uint8(1 << (b & 7))
    order = rng.permutation(n0 + n_dup)
    world = np.concatenate([world, world[dup]])[order]
    desc = np.concatenate([desc_last, ddesc])[order]
    octave = np.concatenate([kps_last["octave"], kps_last["octave"][dup]])[order].astype(np.int32)
    angle = np.concatenate([kps_last["angle"], kps_last["angle"][dup]])[order].astype(np.float32)
    flags = np.concatenate([flags, flags[dup] | 1])[order].astype(np.uint8)
    a = 0.004
    Rcw = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    tcw = np.array([0.03, -0.01, -0.12], np.float32)
    # local-map style inputs: projections with the same pose (double precision here: they are inputs, not part of the function)
    pc = world.astype(np.float64) @ Rcw.astype(np.float64).T + tcw.astype(np.float64)
    invz = 1.0 / pc[:, 2]
    u = c["fx"] * pc[:, 0] * invz + c["cx"]
    v = c["fy"] * pc[:, 1] * invz + c["cy"]
    proj = np.stack([u, v, u - c["mbf"] * invz], 1).astype(np.float32)
    inview = (u > 0) & (u < w) & (v > 0) & (v < h) & (pc[:, 2] > 0)
    mflags = ((flags & 1) & inview.astype(np.uint8)) | (flags & 2)
    view_cos = rng.uniform(0.99, 1.0, world.shape[0]).astype(np.float32)
    level = np.clip(octave + rng.integers(-1, 2, octave.size), 0, 7).astype(np.int32)
    return dict(world=world, desc=desc, octave=octave, angle=angle, flags=flags, Rcw=Rcw, tcw=tcw, proj=proj, mflags=mflags.astype(np.uint8),
                view_cos=view_cos, level=level, cam=(c["fx"], c["fy"], c["cx"], c["cy"], c["mbf"]), bounds=(0.0, float(w), 0.0, float(h)))


def bow_scenario(desc_kf, angle_kf, desc_f, seed, node_bits=6):
    """Flattened SearchByBoW inputs.  A stand-in for the DBoW2 feature vectors: the 'vocabulary node' of a descriptor is
    the top `node_bits` bits of its first byte (64 nodes by default, tens of keypoints each; matching keypoints mostly share
    it; 3 bits give lists of hundreds of keypoints, longer than the kernel's shared-memory cache).  Points = the key-frame
    keypoints in the reference's traversal order (nodes ascending, indices ascending inside a node); node lists of the
    frame as CSR; nodes only the key frame has get slot -1 (the reference never visits them)."""
    rng = np.random.default_rng(seed)
    n_nodes = 1 << node_bits
    node_f = (desc_f[:, 0] >> (8 - node_bits)).astype(np.int64)
    node_kf = (desc_kf[:, 0] >> (8 - node_bits)).astype(np.int64)
    present = np.unique(node_f)
    slot_of = -np.ones(n_nodes, np.int64)
    slot_of[present] = np.arange(present.size)
    node_start = np.zeros(present.size + 1, np.int32)
    node_idx = []
    for s, nd in enumerate(present):
        idx = np.nonzero(node_f == nd)[0]
        node_idx.append(idx)
        node_start[s + 1] = node_start[s] + idx.size
    node_idx = np.concatenate(node_idx).astype(np.int32) if node_idx else np.zeros(0, np.int32)
    order = np.lexsort((np.arange(desc_kf.shape[0]), node_kf))              # nodes ascending, index ascending
    flags = (rng.random(order.size) > 0.15).astype(np.uint8)               # pMP && !isBad
    return dict(desc=desc_kf[order], angle=angle_kf[order].astype(np.float32), flags=flags, node_slot=slot_of[node_kf[order]].astype(np.int32),
                node_start=node_start, node_idx=node_idx, order=order)
