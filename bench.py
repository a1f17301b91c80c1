#!/usr/bin/env python
"""bench.py — stereo frames/s of the IV-SLAM stereo front-end (ORB extract L+R + stereo match) on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  * workload (BASELINE.json configs[2] = C3, the shape the metric is quoted on): a batch of 1241x376 synthetic stereo pairs per
    GPU, ORBextractor nFeatures=2000, 8 levels, scale 1.2, iniTh 20 / minTh 7, introspection off, mbf/maxD of KITTI00-02.yaml.
    One "step" = one pass of the hot path over the rank's whole batch (default 1024 pairs per GPU, weak scaling: frames are
    independent, every GPU owns its own contiguous frame batch, no collective on the data path).
  * `value`     device-resident: level-0 images already in HBM when the timed region starts (inputs of one step are ~0.95 GB
                per GPU, far larger than the 126 MB L2, so nothing is cache-warm between steps).
  * `e2e`       the same metric through the public API with HOST (pinned) buffers: H2D of every image, kernels, D2H of
                keypoints / descriptors / counts / uRight / depth inside the timed region.
  * `roofline`  for the dominant kernel (largest share of device time, measured live with CUDA events bracketing each launch
                on the launching stream during the timed region): algorithmic bytes per launch / mean duration against the
                measured HBM peak (MEASURED_PEAKS.json).
  * `cpu_baseline` oracle/_ref — the UNMODIFIED reference ORBextractor.cc + Frame::ComputeStereoMatches compiled with the
                reference's own flags over an OpenCV-compat layer (kind "reference"; the oracle port only if those libraries
                are missing, kind "port") — on a bounded sample, on the host cores.
  * `configs`   (N=1 only) the other BASELINE.json configurations, each with value / e2e / cpu_baseline / roofline and a
                parity check of the CUDA result against oracle/_ref on that configuration's first pair (outside any timed
                region):  C1 one KITTI pair at a time, C2 Jackal 960x600 with an introspection cost-map on the left eye,
                C4 3840x2160 / 8000 features, C5 stereo matcher alone on 5000 synthetic keypoints per eye.
  * `c1_single_pair` drop-in latency of ONE stereo frame through the C++ shim (shim/ORBextractor.cc driven exactly like
                Frame::Frame: two std::threads + matcher) next to the reference's own 2+1-thread CPU time.
`--impl reference` times the reference's CPU implementation of the path (oracle/_ref on all host threads, frame-parallel),
same metric/config.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KITTI = dict(w=1241, h=376, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, mbf=386.1448, maxD=718.856)
METRIC = "stereo frames/sec (ORB extract L+R + stereo match) at 1241x376, 2000 feats"
UNIT = "stereo_frames/s"
PARAM_KEYS = ("nfeatures", "scaleFactor", "nlevels", "iniThFAST", "minThFAST")


def level_sizes(w, h, nlevels=8, sf=1.2):
    s, out = np.float32(1.0), []
    for l in range(nlevels):
        if l:
            s = np.float32(float(s) * float(np.float32(sf)))
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def algorithmic_bytes(w, h, n=2000, n_raw=None, stereo_pair_bytes=2.4e6):
    """Per-IMAGE algorithmic bytes of each kernel; n = kept keypoints, n_raw = raw FAST corners (1.6 n, the
    ratio probed at KITTI shape: 3200 raw / 2000 kept)."""
    lv = level_sizes(w, h)
    P = sum(a * b for a, b in lv)
    L0, L7 = lv[0][0] * lv[0][1], lv[-1][0] * lv[-1][1]
    n_raw = int(1.6 * n) if n_raw is None else n_raw
    return {
        "k_resize_level": 2 * P - L0 - L7,                 # read level l-1, write level l, l = 1..7 (all 7 launches)
        "k_fast_cells": P + 8 * n_raw,                     # read every level once, write the corner lists (K2/K3 of SURVEY §8d)
        "k_gauss7": 2 * P,                                 # read P, write blurred P
        "k_level_select": 4 * n_raw + 8 * n,
        "k_orient_describe": min(961 * n, P) + min(1369 * n, P) + 32 * n + 28 * n,
        "k_stereo_match": stereo_pair_bytes,               # per PAIR (B_stereo)
        "k_stereo_median": 8 * n,
    }, P


def pair_bytes(w, h, n, intro, stereo_pair_bytes):
    """B_pair of SURVEY §8(d): both eyes' extraction + the matcher (+ the cost pyramid build and one read on the left eye)."""
    alg, P = algorithmic_bytes(w, h, n, None, stereo_pair_bytes)
    per_img = sum(v for k, v in alg.items() if not k.startswith("k_stereo"))
    extra = (alg["k_resize_level"] + P) if intro else 0
    return 2 * per_img + extra + stereo_pair_bytes


def stereo_bytes(n_kp, n_matched):
    """B_stereo = N_L (32 + 8 C_scan + 32 C_eval) + 352 M + 8 N_L with the candidate counts probed in SURVEY §8(d) scaled
    linearly with the keypoint density (58 scanned / 11 evaluated per left keypoint at 2000 keypoints per eye)."""
    c_scan, c_eval = 58.0 * n_kp / 2000.0, 11.0 * n_kp / 2000.0
    return n_kp * (32 + 8 * c_scan + 32 * c_eval) + 352 * n_matched + 8 * n_kp


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML, 50 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ CPU side (checker / baseline)
def reference_mb(mbf, maxD):
    """The reference's `mb` for a wanted maxD and the maxD = mbf/mb it then computes in float (Frame.cc:787-789)."""
    mb = np.float32(mbf) / np.float32(maxD)
    return float(mb), float(np.float32(mbf) / mb)


def cpu_backend():
    """oracle/_ref (the unmodified reference sources, as-built flags) when its libraries are present, else the oracle port."""
    from oracle import ref_lib
    if ref_lib.available("asbuilt"):
        ref_lib.lib("asbuilt")
        return "reference"
    return "port"


def cpu_batch_fps(params, imgsL, imgsR, mbf, maxD, workers, costs=None, reps=1, warm=True):
    """Frame-parallel CPU front-end on `workers` threads: (pairs/s, seconds, kind)."""
    kind = cpu_backend()
    mb, maxD = reference_mb(mbf, maxD)
    if kind == "reference":
        from oracle import ref_lib as R
        run = lambda a, b, c: R.stereo_batch(params, a, b, mbf, mb, workers, "asbuilt", c)
    else:
        from oracle import oracle_lib as O
        assert costs is None
        run = lambda a, b, c: O.stereo_batch(params, a, b, mbf, maxD, workers)
    if warm:
        k = min(len(imgsL), workers)
        run(imgsL[:k], imgsR[:k], None if costs is None else costs[:k])
    t = time.perf_counter()
    for _ in range(reps):
        run(imgsL, imgsR, costs)
    dt = time.perf_counter() - t
    return len(imgsL) * reps / dt, dt, kind


def cpu_frame_ms(params, left, right, mbf, maxD, cost=None, intro=False, reps=10):
    """One frame at a time with the reference's own threading (2 extraction threads + matcher, Frame.cc:115-125,:193)."""
    kind = cpu_backend()
    mb, maxD = reference_mb(mbf, maxD)
    a = tuple(params[k] for k in PARAM_KEYS)
    if kind == "reference":
        from oracle import ref_lib as R
        eL, eR = R.RefExtractor(*a, intro), R.RefExtractor(*a, False)
        run = lambda: R.stereo_frame(eL, eR, left, right, cost, mbf, mb, threads=2)
    else:
        from oracle import oracle_lib as O
        eL, eR = O.OracleExtractor(*a, intro), O.OracleExtractor(*a, False)
        run = lambda: O.stereo_frame(eL, eR, left, right, cost, mbf, maxD, threads=2)
    run()
    t = time.perf_counter()
    for _ in range(reps):
        run()
    return 1e3 * (time.perf_counter() - t) / reps, kind


def parity_check(api, params, intro, left, right, cost, mbf, maxD, gL=None, gR=None):
    """CUDA path vs oracle/_ref (or the oracle port) on one pair, outside every timed region: keypoint records byte-identical,
    >= 99.9 % of descriptor bits, disparities within 1e-3 px."""
    kind = cpu_backend()
    mb, maxD = reference_mb(mbf, maxD)
    a = tuple(params[k] for k in PARAM_KEYS)
    own = gL is None
    if own:
        gL, gR = api.ORBextractor(*a, intro), api.ORBextractor(*a, False)
    kL, dL = gL(left, cost)
    kR, dR = gR(right, None)
    u, d = api.compute_stereo_matches(gL, gR, mbf, maxD)
    if kind == "reference":
        from oracle import ref_lib as R
        r = R.stereo_frame(R.RefExtractor(*a, intro), R.RefExtractor(*a, False), left, right, cost, mbf, mb)
    else:
        from oracle import oracle_lib as O
        r = O.stereo_frame(O.OracleExtractor(*a, intro), O.OracleExtractor(*a, False), left, right, cost, mbf, maxD)
    n = kL.size
    ok = kL.tobytes() == r["kL"].tobytes() and kR.tobytes() == r["kR"].tobytes()
    flips = int(np.unpackbits(dL ^ r["dL"]).sum() + np.unpackbits(dR ^ r["dR"]).sum()) if ok else -1
    ok = ok and flips <= 1e-3 * (dL.size + dR.size) * 8
    ok = ok and bool(np.array_equal(u[:n] >= 0, r["uRight"] >= 0)) and (n == 0 or float(np.max(np.abs(u[:n] - r["uRight"]))) <= 1e-3)
    if own:
        gL.close(), gR.close()
    return {"ok": bool(ok), "against": "oracle/_ref (unmodified reference, as-built flags)" if kind == "reference" else "oracle port",
            "keypoints": int(n), "descriptor_bits_differing": flips, "stereo_matches": int((r["uRight"] >= 0).sum())}


# ------------------------------------------------------------------------------------------------ GPU arms
def resident_arm(api, params, intro, imgsL, imgsR, costs, mbf, maxD, steps, warmup, device, barrier=None, sampler=None):
    """Device-resident arm: inputs uploaded once, `steps` timed passes of (left extract, right extract, stereo match) with
    CUDA events around every launch (the per-kernel shares) and around the whole region (the step time)."""
    a = tuple(params[k] for k in PARAM_KEYS)
    B = imgsL.shape[0]
    resL, resR = api.ORBextractor(*a, intro, device=device), api.ORBextractor(*a, False, device=device)
    resR.share_stream(resL)     # one kernel stream: left and right extraction run back to back, not interleaved
    resL.upload(imgsL, costs)
    resR.upload(imgsR)
    resL.sync(), resR.sync()

    def step():
        resL.run()
        resR.run()
        rc = api.lib().ivg_stereo_match_batch(resL._h, resR._h, mbf, maxD, None, None, resL.cap, 0)
        assert rc == 0, rc

    for _ in range(max(warmup, 3)):
        step()
    resL.sync(), resR.sync()
    # a pass without the per-launch profiling events, to show what they cost inside the timed region below
    resL.timer_start()
    for _ in range(steps):
        step()
    resL.timer_stop()
    ms_plain = resL.timer_ms() / steps
    resR.sync()
    launches0 = resL.launch_count() + resR.launch_count()
    resL.profile_enable(True), resR.profile_enable(True)
    if barrier:
        barrier()
    if sampler:
        sampler.start()
    resL.timer_start()
    for _ in range(steps):
        step()
    resL.timer_stop()          # left stream: its last op (stereo) waits for the right stream
    ms_total = resL.timer_ms()
    resR.sync()
    if barrier:
        barrier()
    launches = resL.launch_count() + resR.launch_count() - launches0
    prof = {}
    for k, (ms, cnt) in resL.profile_read().items():
        prof[k] = [ms, cnt]
    for k, (ms, cnt) in resR.profile_read().items():
        prof[k][0] += ms
        prof[k][1] += cnt
    resL.profile_enable(False), resR.profile_enable(False)
    prof = {k: v for k, v in prof.items() if v[1]}
    kps = np.zeros((B, resL.cap), api.KP_DTYPE)
    desc = np.zeros((B, resL.cap, 32), np.uint8)
    cnt = np.zeros(B, np.int32)
    resL.download(kps, desc, cnt)
    u, d = api.compute_stereo_matches_batch(resL, resR, mbf, maxD)
    resL.sync()
    check = {"keypoints_per_pair_left": float(cnt.mean()), "stereo_matches_per_pair": float((u >= 0).sum() / B)}
    resR.close(), resL.close()
    return dict(ms_step=ms_total / steps, ms_step_no_profile_events=ms_plain, launches=int(launches), prof=prof, check=check, B=B)


def roofline_of(prof, alg, B, steps, peak, matched_per_pair=None):
    """Dominant kernel (largest live CUDA-event share) and its algorithmic bytes per launch / mean launch duration."""
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k][0])
    per_launch_ms = prof[dom][0] / max(prof[dom][1], 1)
    if dom == "k_resize_level":
        bytes_per_launch = alg[dom] * B / 7.0      # 7 launches per eye: per-launch bytes = total / 7
    elif dom == "k_stereo_match":
        bytes_per_launch = alg[dom] * B / 2.0      # index + match launches are accounted under one id
    else:
        bytes_per_launch = alg[dom] * B
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "bytes_per_launch": bytes_per_launch, "ms_per_launch": per_launch_ms, "share_of_device_time": prof[dom][0] / tot_ms}


def e2e_arm(StereoFrontend, api, params, intro, W, H, pinL, pinR, pinC, mbf, maxD, chunk, slots, steps, device, shared, barrier=None):
    """End-to-end arm through the public API: pinned host frames in, host result arrays out, every step."""
    B = pinL.shape[0]
    fe = StereoFrontend(params, W, H, min(chunk, B), slots, device=device, introspection=intro, share_kernel_stream=shared)
    out = fe.alloc_outputs(B, pinned=True)
    for _ in range(2):
        fe.process(pinL, pinR, out, mbf, maxD, pinC)
        fe.finish()
    if barrier:
        barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fe.process(pinL, pinR, out, mbf, maxD, pinC)
    fe.finish()
    dt = time.perf_counter() - t0
    h2d = int(pinL.nbytes + pinR.nbytes + (pinC.nbytes if pinC is not None else 0))
    d2h = int(sum(out[k].nbytes for k in ("kL", "dL", "nL", "kR", "dR", "nR", "uRight", "depth")))
    res = dict(seconds=dt, h2d=h2d, d2h=d2h, n_kp=int(out["nL"].sum()), n_match=int((out["uRight"] >= 0).sum()), out=out)
    fe.close()
    return res


def pinned_copy(api, arr):
    p = api.PinnedArray(arr.shape, arr.dtype)
    p.array[...] = arr
    return p


def run_batch_config(name, cfg, B, distinct, chunk, steps, device, api, StereoFrontend, S, peak, cpu_sample):
    """One BASELINE configuration measured like the headline: resident value, e2e, CPU baseline, roofline, parity."""
    W, H, intro = cfg["w"], cfg["h"], cfg["introspection"]
    params = {k: cfg[k] for k in PARAM_KEYS}
    mbf, maxD = cfg["mbf"], reference_mb(cfg["mbf"], cfg["maxD"])[1]
    Lh, Rh = S.make_stereo_batch(W, H, B, cfg["seed"], distinct=distinct)
    Ch = None
    if intro:
        base = [S.make_cost_map(W, H, cfg["cost_seed"] + i) for i in range(min(distinct, B))]
        Ch = np.stack([base[i % len(base)] for i in range(B)])
    parity = parity_check(api, params, intro, Lh[0], Rh[0], None if Ch is None else Ch[0], mbf, maxD)
    pL, pR = pinned_copy(api, Lh), pinned_copy(api, Rh)
    pC = pinned_copy(api, Ch) if Ch is not None else None
    r = resident_arm(api, params, intro, pL.array, pR.array, None if pC is None else pC.array, mbf, maxD, steps, 3, device)
    value = B / (r["ms_step"] * 1e-3)
    alg, P = algorithmic_bytes(W, H, cfg["nfeatures"])
    roof = roofline_of(r["prof"], alg, B, steps, peak)
    pb = pair_bytes(W, H, cfg["nfeatures"], intro, 2.4e6 * cfg["nfeatures"] / 2000.0)
    roof["whole_pipeline_GBps"] = pb * value / 1e9
    roof["whole_pipeline_frac"] = pb * value / 1e9 / peak
    e = e2e_arm(StereoFrontend, api, params, intro, W, H, pL.array, pR.array, None if pC is None else pC.array, mbf, maxD, chunk, 2, steps, device, True)
    ncpu = os.cpu_count() or 1
    n_s = min(cpu_sample, B)
    fps, dt, kind = cpu_batch_fps(params, Lh[:n_s], Rh[:n_s], mbf, maxD, ncpu, None if Ch is None else Ch[:n_s])
    return {"workload": "%s: %dx%d stereo pairs, nFeatures %d, iniTh %d, introspection %s; %d pairs per step (%d distinct images + variants)"
                        % (name, W, H, cfg["nfeatures"], cfg["iniThFAST"], "cost-map on the left eye" if intro else "off", B, distinct),
            "value": value, "unit": UNIT, "ms_per_step": r["ms_step"],
            "e2e": {"value": B * steps / e["seconds"], "unit": UNIT, "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"]},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": ncpu, "kind": kind,
                             "sample": "first %d pairs, frame-parallel over %d threads, %.1f s wall" % (n_s, ncpu, dt)},
            "roofline": roof, "kernel_ms_per_step": {k: v[0] / steps for k, v in r["prof"].items()},
            "check": r["check"], "parity_checked": parity["ok"], "parity": parity}


def run_c1(cfg, api, S, peak, steps=50):
    """C1: ONE 1241x376 pair at a time (the reference's own CPU-runnable case)."""
    W, H = cfg["w"], cfg["h"]
    params = {k: cfg[k] for k in PARAM_KEYS}
    a = tuple(params[k] for k in PARAM_KEYS)
    mbf, maxD = cfg["mbf"], reference_mb(cfg["mbf"], cfg["maxD"])[1]
    left, right = S.make_stereo_pair(W, H, cfg["seed"])
    parity = parity_check(api, params, False, left, right, None, mbf, maxD)
    pL, pR = pinned_copy(api, left[None]), pinned_copy(api, right[None])
    r = resident_arm(api, params, False, pL.array, pR.array, None, mbf, maxD, steps, 5, 0)
    value = 1.0 / (r["ms_step_no_profile_events"] * 1e-3)
    alg, P = algorithmic_bytes(W, H, cfg["nfeatures"])
    roof = roofline_of(r["prof"], alg, 1, steps, peak)
    # e2e: host images in, host keypoints / descriptors / uRight / depth out, two host threads + matcher per frame
    lL, lR = api.ORBextractor(*a, False), api.ORBextractor(*a, False)
    lL.set_graph_mode(True), lR.set_graph_mode(True)

    def one_pair():
        tl = threading.Thread(target=lambda: lL(left))
        tr = threading.Thread(target=lambda: lR(right))
        tl.start(); tr.start(); tl.join(); tr.join()
        api.compute_stereo_matches(lL, lR, mbf, maxD)
    for _ in range(5):
        one_pair()
    t1 = time.perf_counter()
    for _ in range(steps):
        one_pair()
    e2e_ms = 1e3 * (time.perf_counter() - t1) / steps
    lL.close(), lR.close()
    cpu_ms, kind = cpu_frame_ms(params, left, right, mbf, maxD)
    cap = cfg["nfeatures"] + 64
    return {"workload": "C1: single 1241x376 stereo pair, nFeatures 2000, introspection off, one frame per call",
            "value": value, "unit": UNIT, "ms_per_step": r["ms_step_no_profile_events"],
            "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "latency_ms": e2e_ms, "h2d_bytes_per_step": int(2 * W * H),
                    "d2h_bytes_per_step": int(2 * cap * 60 + 2 * cap * 4), "note": "python host threads; the C++ shim number is c1_single_pair"},
            "cpu_baseline": {"value": 1e3 / cpu_ms, "unit": UNIT, "cores": 3, "kind": kind,
                             "sample": "the same pair, 10 repetitions, the reference's own threading (2 extraction threads + matcher)"},
            "roofline": roof, "kernel_ms_per_step": {k: v[0] / steps for k, v in r["prof"].items()},
            "check": r["check"], "parity_checked": parity["ok"], "parity": parity}, cpu_ms


def run_c5(api, S, peak, n_kp=5000, steps=30):
    """C5: the stereo matcher alone (Frame::ComputeStereoMatches) on 5000 synthetic keypoints per eye with dense epipolar bands."""
    c = S.CONFIGS["C1"]
    W, H = c["w"], c["h"]
    a = (2000, 1.2, 8, 20, 7)
    mbf = c["mbf"]
    mb, maxD = reference_mb(mbf, c["maxD"])
    left, right = S.make_stereo_pair(W, H, 4)
    gL, gR = api.ORBextractor(*a), api.ORBextractor(*a)
    gL.compute_pyramid(left), gR.compute_pyramid(right)
    kL, dL, kR, dR = S.make_c5_stereo_stress(gL.GetScaleFactors(), gL.features_per_level(), W, H, n_kp, 4)
    u, d = api.compute_stereo_matches_keypoints(gL, gR, kL, dL, kR, dR, mbf, maxD)
    kind = cpu_backend()
    if kind == "reference":
        from oracle import ref_lib as R
        rL, rR = R.RefExtractor(*a), R.RefExtractor(*a)
        rL(left), rR(right)
        cpu = lambda: R.stereo_match(rL, rR, kL, dL, kR, dR, mbf, mb)
    else:
        from oracle import oracle_lib as O
        oL, oR = O.OracleExtractor(*a), O.OracleExtractor(*a)
        oL.compute_pyramid(left), oR.compute_pyramid(right)
        cpu = lambda: O.stereo_match(oL, oR, kL, dL, kR, dR, mbf, maxD)
    uo, do = cpu()
    ok = bool(np.array_equal(u >= 0, uo >= 0)) and float(np.max(np.abs(u - uo))) <= 1e-3
    t = time.perf_counter()
    for _ in range(5):
        cpu()
    cpu_ms = 1e3 * (time.perf_counter() - t) / 5
    for _ in range(5):
        api.compute_stereo_matches_keypoints(gL, gR, kL, dL, kR, dR, mbf, maxD)
    gL.profile_enable(True)
    t = time.perf_counter()
    for _ in range(steps):
        api.compute_stereo_matches_keypoints(gL, gR, kL, dL, kR, dR, mbf, maxD)
    e2e_ms = 1e3 * (time.perf_counter() - t) / steps
    prof = {k: v for k, v in gL.profile_read().items() if v[1]}
    gL.profile_enable(False)
    dev_ms = sum(v[0] for v in prof.values()) / steps
    matched = int((uo >= 0).sum())
    bs = stereo_bytes(n_kp, matched)
    ms_match = prof["k_stereo_match"][0] / steps
    gL.close(), gR.close()
    return {"workload": "C5: stereo matcher alone, %d keypoints per eye at 1241x376 (dense epipolar bands), Hamming + SAD sub-pixel + median filter" % n_kp,
            "value": 1e3 / dev_ms, "unit": "stereo_frames/s (matcher only)", "ms_per_step": dev_ms,
            "e2e": {"value": 1e3 / e2e_ms, "unit": "stereo_frames/s (matcher only)", "latency_ms": e2e_ms,
                    "h2d_bytes_per_step": int(2 * n_kp * 60), "d2h_bytes_per_step": int(2 * n_kp * 4)},
            "cpu_baseline": {"value": 1e3 / cpu_ms, "unit": "stereo_frames/s (matcher only)", "cores": 1, "kind": kind,
                             "sample": "the same keypoints, 5 repetitions, single thread (the reference matcher is single-threaded)"},
            "roofline": {"bound": "hbm", "kernel": "k_stereo_match", "achieved": bs / (ms_match * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bs / (ms_match * 1e-3) / 1e9 / peak, "bytes_per_launch": bs, "ms_per_launch": ms_match,
                         "note": "B_stereo of SURVEY 8(d) with candidate counts scaled to the keypoint density; index + match launches together"},
            "kernel_ms_per_step": {k: v[0] / steps for k, v in prof.items()},
            "check": {"stereo_matches": matched}, "parity_checked": ok}


def shim_latency(iters=200):
    """One stereo frame through the C++ shim exactly like Frame::Frame (tools/latency.py): ms per frame, graph mode on."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import latency
    return latency.measure(iters)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world, emit):
    """Reference arm: the reference's CPU path on all host threads, frame-parallel; rank 0 only."""
    if rank != 0:
        return
    from iv_slam_b200 import synthetic as S
    params = {k: KITTI[k] for k in PARAM_KEYS}
    workers = os.cpu_count() or 1
    sample = max(workers * 2, args.ref_sample)
    L, R = S.make_stereo_batch(KITTI["w"], KITTI["h"], sample, 100, distinct=min(16, sample))
    for _ in range(args.warmup):
        cpu_batch_fps(params, L[:workers], R[:workers], KITTI["mbf"], KITTI["maxD"], workers)
    t0 = time.perf_counter()
    tot = 0
    kind = "port"
    for _ in range(args.steps):
        _, _, kind = cpu_batch_fps(params, L, R, KITTI["mbf"], KITTI["maxD"], workers, warm=False)
        tot += sample
    dt = time.perf_counter() - t0
    fps = tot / dt
    how = ("oracle/_ref: unmodified reference ORBextractor.cc + Frame::ComputeStereoMatches, reference build flags, OpenCV-compat layer"
           if kind == "reference" else "oracle port (oracle/_ref libraries not present)")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C3: KITTI-shape 1241x376 stereo pairs, nFeatures 2000, 8 levels, scale 1.2, iniTh 20, minTh 7, introspection off",
                       "pairs_per_step": sample, "note": "CPU reference path: host cores only, all threads, frame-parallel; " + how},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": workers, "kind": kind,
                             "sample": "%d pairs per step, %d steps" % (sample, args.steps)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    # Libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line on stdout, so everything
    # else is sent to stderr and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="stereo pairs per GPU per step")
    ap.add_argument("--chunk", type=int, default=256, help="stereo pairs per launch group")
    ap.add_argument("--slots", type=int, default=2)
    ap.add_argument("--distinct", type=int, default=32, help="distinct synthetic base images per rank")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="pairs in the cpu_baseline sample (1024 pairs = ~17 CPU-seconds)")
    ap.add_argument("--ref-sample", type=int, default=128, help="pairs per step for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C2/C4/C5 legs (N=1 only anyway)")
    ap.add_argument("--e2e-streams", choices=["shared", "per-handle"], default="shared", help="kernel streams of the e2e pipeline")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    from iv_slam_b200 import api, sharding, synthetic as S
    from iv_slam_b200.frontend import StereoFrontend

    torch.cuda.set_device(local)
    numa_cpus = sharding.bind_to_gpu_numa_node(local) if world > 1 else None     # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, W, H = args.batch, KITTI["w"], KITTI["h"]
    params = {k: KITTI[k] for k in PARAM_KEYS}
    mbf, maxD = KITTI["mbf"], KITTI["maxD"]
    # this rank's contiguous frame range [f0, f1) of the global synthetic sequence of world*B pairs (weak scaling)
    f0, f1 = sharding.frame_range(rank, world, world * B)
    assert f1 - f0 == B
    Lh, Rh = S.make_stereo_batch(W, H, B, 100 + f0, distinct=args.distinct)
    pinL, pinR = pinned_copy(api, Lh), pinned_copy(api, Rh)

    # ------------------------------------------------------------------ device-resident arm (`value`)
    sampler = ClockSampler(local)
    r = resident_arm(api, params, False, pinL.array, pinR.array, None, mbf, maxD, args.steps, args.warmup, local, barrier, sampler)
    clocks = sampler.result()
    ms_step = sharding.reduce_max(r["ms_step"] * args.steps, "cuda") / args.steps       # max over ranks
    value = world * B / (ms_step * 1e-3)

    # ------------------------------------------------------------------ end-to-end arm (`e2e`)
    e = e2e_arm(StereoFrontend, api, params, False, W, H, pinL.array, pinR.array, None, mbf, maxD, args.chunk, args.slots, args.steps,
                local, args.e2e_streams == "shared", barrier)
    torch.cuda.synchronize()
    e2e_s = sharding.reduce_max(e["seconds"], "cuda")
    e2e_value = world * B * args.steps / e2e_s
    # per-rank host<->device rates of the e2e arm (names the slow rank / link when the aggregate stops scaling)
    rank_rates = [0.0] * world
    rank_rates[rank] = e["h2d"] * args.steps / e["seconds"] / 1e9
    if world > 1:
        t = torch.tensor(rank_rates, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        rank_rates = [float(v) for v in t.tolist()]
    # every rank's result counts, gathered to all ranks: one host result summary over the whole frame sequence
    kp_all = sharding.gather_counts(e["out"]["nL"], world * B, rank, world, "cuda")

    if rank == 0:
        peak, peak_kind = measured_peak()
        alg, P = algorithmic_bytes(W, H)
        prof = r["prof"]
        tot_ms = sum(v[0] for v in prof.values()) or 1.0
        shares = {k: v[0] / tot_ms for k, v in prof.items()}
        roof = roofline_of(prof, alg, B, args.steps, peak)
        dom = roof["kernel"]
        traffic, traffic_source = None, None
        try:     # DRAM bytes per image of the same kernel from the committed ncu --set full capture (NOT measured in this run)
            with open(os.path.join(ROOT, "profiles", "ncu_dram_bytes_per_image.json")) as f:
                per_img = json.load(f).get(dom)
            if per_img:
                traffic = per_img * B
                traffic_source = "profiles/ncu_dram_bytes_per_image.json: dram__bytes_read+write per image from the committed ncu --set full capture at 64 images per launch, scaled to %d images; not measured in this run" % B
        except Exception:
            pass
        # the same kernel against the ISSUE roofline (it is integer-ALU work, not HBM-bound): warp instructions per image from the
        # committed ncu capture / (SMs x 4 schedulers x the SM clock sampled during the run); not measured in this run either
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_warp_inst_per_image.json")) as f:
                wi = json.load(f).get(dom)
            mhz = (clocks or {}).get("sm_mhz")
            if wi and mhz:
                issue_peak = 148 * 4 * mhz * 1e6                  # warp instructions per second, one per scheduler per cycle
                roof["issue_frac"] = wi * B / (roof["ms_per_launch"] * 1e-3) / issue_peak
                roof["issue_frac_source"] = ("profiles/ncu_warp_inst_per_image.json (smsp__inst_executed per image, committed capture) x %d images / "
                                             "launch time / (148 SMs x 4 schedulers x %.0f MHz)" % (B, mhz))
        except Exception:
            pass
        pb = 21.9e6
        roof.update({"traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_kind,
                     "whole_pipeline_GBps": pb * value / world / 1e9, "whole_pipeline_frac": pb * value / world / 1e9 / peak})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "C3: batch of KITTI-shape 1241x376 stereo pairs, nFeatures 2000, 8 levels, scale 1.2, iniTh 20, minTh 7, introspection off",
                       "pairs_per_gpu_per_step": B, "chunk_pairs": args.chunk, "slots": args.slots,
                       "distinct": "%d distinct synthetic images per rank (seeds %d..), the rest row-rolled / brightness-shifted variants"
                                   % (min(args.distinct, B), 100 + f0),
                       "l2": "inputs of one step (%.0f MB per GPU) exceed the 126 MB L2; no flush needed" % (e["h2d"] / 1e6),
                       "parallelism": "frame-parallel, %d independent rank(s), no collective" % world,
                       "numa": ("rank 0 bound to %d CPUs local to its GPU" % len(numa_cpus)) if numa_cpus else "no binding"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"]},
            "gpu_launches": r["launches"],
            "clocks": clocks,
            "roofline": roof,
            "kernel_shares": {k: round(v, 4) for k, v in shares.items()},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "value_without_profiling_events": world * B / (r["ms_step_no_profile_events"] * 1e-3),
            "check": dict(r["check"], keypoints_left_all_ranks=int(kp_all.sum()), frames_all_ranks=int(kp_all.size)),
            "e2e_scaling": {"h2d_GBps_per_rank": [round(v, 2) for v in rank_rates], "h2d_GBps_total": round(sum(rank_rates), 2),
                            "e2e_pairs_per_s_per_gpu": e2e_value / world},
        }
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N=1 only
            ncpu = os.cpu_count() or 1
            n_s = min(args.cpu_sample, B)
            fps_all, dt_all, kind = cpu_batch_fps(params, Lh[:n_s], Rh[:n_s], mbf, maxD, ncpu)
            ms_frame, _ = cpu_frame_ms(params, Lh[0], Rh[0], mbf, maxD, reps=20)
            line["cpu_baseline"] = {"value": fps_all, "unit": UNIT, "cores": ncpu, "kind": kind,
                                    "sample": "first %d pairs of the workload, frame-parallel over %d threads, %.1f s wall" % (n_s, ncpu, dt_all),
                                    "reference_threading_2plus1": {"value": 1e3 / ms_frame, "cores": 3,
                                                                   "note": "one frame at a time, 2 extraction threads + matching, as src/Frame.cc:115-125,:193"}}
        if world == 1 and not args.no_configs:
            del pinL, pinR
            cfgs = {}
            c1, cpu_ms_c1 = run_c1(S.CONFIGS["C1"], api, S, peak)
            cfgs["C1"] = c1
            cfgs["C2"] = run_batch_config("C2", S.CONFIGS["C2"], 256, 8, 128, args.steps, local, api, StereoFrontend, S, peak, 128)
            cfgs["C4"] = run_batch_config("C4", S.CONFIGS["C4"], 16, 2, 8, args.steps, local, api, StereoFrontend, S, peak, 16)
            cfgs["C5"] = run_c5(api, S, peak)
            line["configs"] = cfgs
            try:
                lat = shim_latency()
                line["c1_single_pair"] = {"latency_ms": lat["graph"], "latency_ms_no_graph": lat["plain"],
                                          "latency_ms_pinned_caller_images": lat.get("graph_pinned_input"),
                                          "latency_ms_one_call": lat.get("one_call"),
                                          "latency_ms_one_call_pinned_caller_images": lat.get("one_call_pinned_input"),
                                          "cpu_ref_ms": cpu_ms_c1, "ratio": cpu_ms_c1 / lat["graph"],
                                          "how": "C++ shim (shim/ORBextractor.cc) driven like Frame::Frame: two std::threads + ComputeStereoMatches, 200 frames "
                                                 "(latency_ms: the headline; *_one_call: ExtractStereoGPU, both eyes and the matcher queued from one thread — a "
                                                 "five-line change in Frame.cc); cpu_ref_ms = the reference's own code with its own threading on the same pair"}
            except Exception as ex:      # no g++ on the box: the python-thread number in configs.C1.e2e stands
                line["c1_single_pair"] = {"unavailable": str(ex)[:200]}
        emit(line)

    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
