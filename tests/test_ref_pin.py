"""CPU tests (no GPU): pin the oracle to the REFERENCE ITSELF.

oracle/_ref is the unmodified /root/reference/introspective_ORB_SLAM/src/ORBextractor.cc (whole file) plus
Frame::ComputeStereoMatches (src/Frame.cc:758-932) and ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1698-1716) cut out
verbatim, compiled by oracle/refbuild/build.sh against an OpenCV-compat layer whose pixel primitives are the cv2-pinned
ones (tests/test_oracle_vs_cv2.py).  Two builds:

  "nofma"   -ffp-contract=off : must equal the restated oracle (oracle/ivslam_oracle.cpp) BIT FOR BIT, everything.
  "asbuilt" the reference's flags (FP contraction on): keypoints and pyramids must still be bit-identical;
            descriptors within the north star's 99.9 % allowance with the fraction REPORTED; disparities within 1e-3 px.

The libraries are built in the container that has /root/reference and travel to the GPU box as files.
"""
import numpy as np
import pytest

from iv_slam_b200 import synthetic as S

from helpers import (assert_descriptors_close, assert_keypoints_equal, assert_stereo_close, fuzz_case, load_golden, reference_mb)


def _kp_identical(a, b, what):
    assert a.size == b.size, "%s: %d vs %d keypoints" % (what, a.size, b.size)
    assert a.tobytes() == b.tobytes(), "%s: keypoint records differ" % what


def _run_all(ref, oracle, variant, nf, sf, nl, ini, mn, intro, left, right, cost, mbf, maxD_nominal):
    """One stereo frame through the unmodified reference (two extractor threads + matcher) and through the oracle."""
    mb, maxD = reference_mb(mbf, maxD_nominal)
    rL, rR = ref.RefExtractor(nf, sf, nl, ini, mn, intro, variant), ref.RefExtractor(nf, sf, nl, ini, mn, False, variant)
    oL, oR = oracle.OracleExtractor(nf, sf, nl, ini, mn, intro), oracle.OracleExtractor(nf, sf, nl, ini, mn, False)
    r = ref.stereo_frame(rL, rR, left, right, cost, mbf, mb, threads=2)
    o = oracle.stereo_frame(oL, oR, left, right, cost, mbf, maxD, threads=2)
    for l in range(nl):     # the public pyramid members (include/ORBextractor.h:91-92)
        assert np.array_equal(rL.level(l, 0), oL.level(l, 0)), "mvImagePyramid[%d]" % l
        if cost is not None and intro:
            assert np.array_equal(rL.level(l, 2), oL.level(l, 2)), "mvQualityImagePyramid[%d]" % l
    assert np.array_equal(rL.features_per_level(), oL.features_per_level())
    assert np.array_equal(rL.scale_factors(), oL.scale_factors())
    assert np.array_equal(rL.umax(), oL.umax())
    return r, o


def _assert_bit_identical(r, o, what):
    _kp_identical(r["kL"], o["kL"], what + " left")
    _kp_identical(r["kR"], o["kR"], what + " right")
    assert np.array_equal(r["dL"], o["dL"]) and np.array_equal(r["dR"], o["dR"]), what + " descriptors"
    assert np.array_equal(r["uRight"], o["uRight"]) and np.array_equal(r["depth"], o["depth"]), what + " stereo"


def _assert_within_north_star(r, o, what):
    assert_keypoints_equal(r["kL"], o["kL"], what + " left")
    assert_keypoints_equal(r["kR"], o["kR"], what + " right")
    # everything but the descriptor sampling is bit-exact even with contraction on: the keypoint records, angle included
    _kp_identical(r["kL"], o["kL"], what + " left")
    _kp_identical(r["kR"], o["kR"], what + " right")
    fl = assert_descriptors_close(r["dL"], o["dL"], what)
    fr = assert_descriptors_close(r["dR"], o["dR"], what)
    assert_stereo_close(r["uRight"], r["depth"], o["uRight"], o["depth"], what)
    return fl, fr


@pytest.mark.parametrize("variant", ["nofma", "asbuilt"])
@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_reference_equals_oracle_on_baseline_configs(ref, oracle, variant, cfg):
    c = S.CONFIGS[cfg]
    left, right = S.make_stereo_pair(c["w"], c["h"], c["seed"])
    cost = S.make_cost_map(c["w"], c["h"], c["cost_seed"]) if c["introspection"] else None
    r, o = _run_all(ref, oracle, variant, c["nfeatures"], c["scaleFactor"], c["nlevels"], c["iniThFAST"], c["minThFAST"],
                    c["introspection"], left, right, cost, c["mbf"], c["maxD"])
    assert r["kL"].size > 1500 and (r["uRight"] >= 0).sum() > 500
    if variant == "nofma":
        _assert_bit_identical(r, o, cfg)
    else:
        print(cfg, "as-built descriptor identical fraction L/R:", _assert_within_north_star(r, o, cfg))


@pytest.mark.parametrize("variant", ["nofma", "asbuilt"])
@pytest.mark.parametrize("seed", range(20))
def test_reference_equals_oracle_on_fuzz_geometries(ref, oracle, variant, seed):
    c = fuzz_case(seed)
    try:   # geometries the reference cannot run (zero grid => division by zero, cells outside the level) are the oracle's -2
        oracle.OracleExtractor(c["nf"], c["sf"], c["nl"], c["ini"], 7, c["intro"])(c["left"], c["cost"])
    except RuntimeError:
        pytest.skip("geometry outside the reference's domain: " + c["what"])
    r, o = _run_all(ref, oracle, variant, c["nf"], c["sf"], c["nl"], c["ini"], 7, c["intro"], c["left"], c["right"], c["cost"], 120.0, 300.0)
    if variant == "nofma":
        _assert_bit_identical(r, o, c["what"])
    else:
        _assert_within_north_star(r, o, c["what"])


@pytest.mark.parametrize("variant", ["nofma", "asbuilt"])
def test_reference_equals_oracle_on_cost_map_extremes(ref, oracle, variant):
    """cost 255 everywhere => all weights 0 => 0/0 budgets: whatever the reference's float code does there, the
    oracle must do the same."""
    left, right = S.make_stereo_pair(640, 400, 55)
    noise = np.random.default_rng(0).integers(0, 256, left.shape, dtype=np.uint8)
    for name, cost in (("255", np.full(left.shape, 255, np.uint8)), ("0", np.zeros(left.shape, np.uint8)),
                       ("128", np.full(left.shape, 128, np.uint8)), ("noise", noise)):
        r, o = _run_all(ref, oracle, variant, 1000, 1.2, 8, 20, 7, True, left, right, cost, 100.0, 400.0)
        (_assert_bit_identical if variant == "nofma" else _assert_within_north_star)(r, o, "cost=" + name)


@pytest.mark.parametrize("w,h,nf,ini,intro", [(752, 480, 1200, 20, True), (960, 600, 5000, 12, False), (960, 600, 2000, 50, True),
                                               (641, 479, 777, 20, False), (320, 240, 300, 20, False)])
def test_reference_equals_oracle_on_other_reference_configs(ref, oracle, w, h, nf, ini, intro):
    left, right = S.make_stereo_pair(w, h, w + nf)
    cost = S.make_cost_map(w, h, 5) if intro else None
    r, o = _run_all(ref, oracle, "nofma", nf, 1.2, 8, ini, 7, intro, left, right, cost, 100.0, 400.0)
    _assert_bit_identical(r, o, "%dx%d/%d" % (w, h, nf))


def test_reference_introspection_flag_without_cost_map_and_cost_map_without_flag(ref, oracle):
    """operator() only weights when BOTH the flag and a mask are present (ORBextractor.cc:1231-1240)."""
    left, right = S.make_stereo_pair(640, 400, 9)
    cost = S.make_cost_map(640, 400, 10)
    for intro, cm in ((True, None), (False, cost)):
        kr, dr = ref.RefExtractor(800, 1.2, 8, 20, 7, intro, "nofma")(left, cm)
        ko, do = oracle.OracleExtractor(800, 1.2, 8, 20, 7, intro)(left, cm)
        _kp_identical(kr, ko, "intro=%s" % intro)
        assert np.array_equal(dr, do)


def test_reference_stereo_stress_c5(ref, oracle):
    """C5: 5000 synthetic keypoints per eye through the verbatim ComputeStereoMatches vs the oracle's restatement."""
    c = S.CONFIGS["C1"]
    left, right = S.make_stereo_pair(c["w"], c["h"], 4)
    mb, maxD = reference_mb(c["mbf"], c["maxD"])
    oL, oR = oracle.OracleExtractor(2000, 1.2, 8, 20, 7), oracle.OracleExtractor(2000, 1.2, 8, 20, 7)
    oL.compute_pyramid(left), oR.compute_pyramid(right)
    kL, dL, kR, dR = S.make_c5_stereo_stress(oL.scale_factors(), oL.features_per_level(), c["w"], c["h"], 5000, 4)
    uo, do = oracle.stereo_match(oL, oR, kL, dL, kR, dR, c["mbf"], maxD)
    for variant in ("nofma", "asbuilt"):
        rL, rR = ref.RefExtractor(2000, 1.2, 8, 20, 7, False, variant), ref.RefExtractor(2000, 1.2, 8, 20, 7, False, variant)
        rL(left), rR(right)      # fills mvImagePyramid, which is all ComputeStereoMatches reads from the extractors
        u, d = ref.stereo_match(rL, rR, kL, dL, kR, dR, c["mbf"], mb)
        assert (u >= 0).sum() > 2000
        if variant == "nofma":
            assert np.array_equal(u, uo) and np.array_equal(d, do)
        else:
            assert_stereo_close(u, d, uo, do, "C5 as-built")


@pytest.mark.parametrize("name", ["small_plain", "small_cost", "kitti_c1", "jackal_c2"])
def test_reference_reproduces_the_cv2_goldens(ref, name):
    """The committed golden vectors were generated from real cv2 primitives composed in Python (tests/golden/make_golden.py);
    the unmodified reference control flow over the compat layer must land on the same bytes."""
    g = load_golden(name)
    nf, ini, mn, intro = (int(v) for v in g["params"])
    mbf, maxD = (float(v) for v in g["calib"])
    mb, maxD_ref = reference_mb(mbf, maxD)
    rL, rR = ref.RefExtractor(nf, 1.2, 8, ini, mn, bool(intro), "nofma"), ref.RefExtractor(nf, 1.2, 8, ini, mn, False, "nofma")
    r = ref.stereo_frame(rL, rR, g["left"], g["right"], g.get("cost"), mbf, mb)
    assert_keypoints_equal(r["kL"], g["kL"], name)
    assert_keypoints_equal(r["kR"], g["kR"], name)
    assert np.array_equal(r["dL"], g["dL"]) and np.array_equal(r["dR"], g["dR"])
    assert_stereo_close(r["uRight"], r["depth"], g["uRight"], g["depth"], name)


def test_asbuilt_fma_descriptor_flip_fraction_report(ref, oracle):
    """SURVEY Q9 quantified: the reference binary as its CMake builds it contracts `x*b + y*a` (ORBextractor.cc:119-121) into
    an FMA.  Over 16 KITTI-shape images count the descriptor bits that differ between that build and the canonical
    (contraction off) semantics the oracle and the CUDA path implement."""
    tot = bad = 0
    ra, rn = ref.RefExtractor(2000, 1.2, 8, 20, 7, False, "asbuilt"), ref.RefExtractor(2000, 1.2, 8, 20, 7, False, "nofma")
    oe = oracle.OracleExtractor(2000, 1.2, 8, 20, 7)
    for seed in range(40, 56):
        img = S.make_image(1241, 376, seed)
        ka, da = ra(img)
        kn, dn = rn(img)
        ko, do = oe(img)
        _kp_identical(ka, kn, "as-built vs canonical keypoints")
        _kp_identical(kn, ko, "canonical vs oracle keypoints")
        assert np.array_equal(dn, do)
        tot += da.size * 8
        bad += int(np.unpackbits(da ^ dn).sum())
    print("descriptor bits differing, reference as-built (FMA) vs canonical: %d of %d (%.3g)" % (bad, tot, bad / tot))
    assert bad / tot <= 1e-3
