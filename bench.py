#!/usr/bin/env python
"""bench.py — stereo frames/s of the IV-SLAM stereo front-end (ORB extract L+R + stereo match) on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  * workload (BASELINE.json configs[2], metric shape): a batch of 1241x376 synthetic stereo pairs per GPU, ORBextractor
    nFeatures=2000, 8 levels, scale 1.2, iniTh 20 / minTh 7, introspection off, mbf/maxD of KITTI00-02.yaml.
    One "step" = one pass of the hot path over the rank's whole batch (default 1024 pairs per GPU, weak scaling:
    frames are independent, every GPU owns its own contiguous frame batch, no collective on the data path).
  * `value`     device-resident: level-0 images already in HBM when the timed region starts (inputs of one step are
                ~0.95 GB per GPU, far larger than the 126 MB L2, so nothing is cache-warm between steps).
  * `e2e`       the same metric through the public API with HOST (pinned) buffers: H2D of every image, kernels, D2H of
                keypoints / descriptors / counts / uRight / depth inside the timed region.
  * `roofline`  for the dominant kernel (largest share of device time, measured live with CUDA events bracketing each
                launch on the launching stream during the timed region): algorithmic bytes per launch / mean duration
                against the measured HBM peak (MEASURED_PEAKS.json).
  * `cpu_baseline` the CPU oracle (a port of the reference path, oracle/ivslam_oracle.cpp) on a bounded sample.
`--impl reference` times the reference's CPU implementation of the path (the oracle port: the reference itself cannot be
compiled in this image) on all host threads, frame-parallel, same metric/config.
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


def level_sizes(w, h, nlevels=8, sf=1.2):
    s, out = np.float32(1.0), []
    for l in range(nlevels):
        if l:
            s = np.float32(float(s) * float(np.float32(sf)))
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def algorithmic_bytes(w, h):
    """Per-IMAGE algorithmic bytes of each kernel; n = 2000 kept, n_raw = 3200 raw corners."""
    lv = level_sizes(w, h)
    P = sum(a * b for a, b in lv)
    L0, L7 = lv[0][0] * lv[0][1], lv[-1][0] * lv[-1][1]
    n, n_raw = 2000, 3200
    return {
        "k_resize_level": 2 * P - L0 - L7,                 # read level l-1, write level l, l = 1..7 (all 7 launches)
        "k_fast_cells": P + 8 * n_raw,                     # read every level once, write the corner lists (K2/K3 of SURVEY §8d)
        "k_gauss7": 2 * P,                                 # read P, write blurred P
        "k_level_select": 4 * n_raw + 8 * n,
        "k_orient_describe": min(961 * n, P) + min(1369 * n, P) + 32 * n + 28 * n,
        "k_stereo_match": 2.4e6,                           # per PAIR (B_stereo)
        "k_stereo_median": 8 * n,
    }, P


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


def cpu_frontend_fps(imgsL, imgsR, workers, reps=1):
    from oracle import oracle_lib as O
    params = {k: KITTI[k] for k in ("nfeatures", "scaleFactor", "nlevels", "iniThFAST", "minThFAST")}
    O.stereo_batch(params, imgsL[:min(len(imgsL), workers)], imgsR[:min(len(imgsL), workers)], KITTI["mbf"], KITTI["maxD"], workers)   # warm
    t = time.perf_counter()
    for _ in range(reps):
        nL, nM = O.stereo_batch(params, imgsL, imgsR, KITTI["mbf"], KITTI["maxD"], workers)
    dt = time.perf_counter() - t
    return len(imgsL) * reps / dt, dt, int(nL.sum()), int(nM.sum())


def run_reference(args, rank, world, emit):
    """Reference arm: the reference's CPU path (oracle port) on all host threads, frame-parallel; rank 0 only."""
    if rank != 0:
        return
    from iv_slam_b200 import synthetic as S
    workers = os.cpu_count() or 1
    sample = max(workers * 2, args.ref_sample)
    L, R = S.make_stereo_batch(KITTI["w"], KITTI["h"], sample, 100, distinct=min(16, sample))
    for _ in range(args.warmup):
        cpu_frontend_fps(L[:workers], R[:workers], workers)
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        cpu_frontend_fps(L, R, workers)
        tot += sample
    dt = time.perf_counter() - t0
    fps = tot / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C3: KITTI-shape 1241x376 stereo pairs, nFeatures 2000, 8 levels, scale 1.2, iniTh 20, minTh 7, introspection off",
                       "pairs_per_step": sample, "note": "CPU reference path: host cores only, all threads, frame-parallel"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": workers, "kind": "port",
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
    params = {k: KITTI[k] for k in ("nfeatures", "scaleFactor", "nlevels", "iniThFAST", "minThFAST")}
    # this rank's contiguous frame range [f0, f1) of the global synthetic sequence of world*B pairs (weak scaling)
    f0, f1 = sharding.frame_range(rank, world, world * B)
    assert f1 - f0 == B
    Lh, Rh = S.make_stereo_batch(W, H, B, 100 + f0, distinct=args.distinct)
    pinL, pinR = api.PinnedArray(Lh.shape, np.uint8), api.PinnedArray(Rh.shape, np.uint8)
    pinL.array[...] = Lh
    pinR.array[...] = Rh

    fe = StereoFrontend(params, W, H, args.chunk, args.slots, device=local, share_kernel_stream=args.e2e_streams == "shared")
    out = fe.alloc_outputs(B, pinned=True)

    # ------------------------------------------------------------------ device-resident arm (`value`)
    a = (params["nfeatures"], params["scaleFactor"], params["nlevels"], params["iniThFAST"], params["minThFAST"])
    resL, resR = api.ORBextractor(*a, False, device=local), api.ORBextractor(*a, False, device=local)
    resR.share_stream(resL)     # one kernel stream: left and right extraction run back to back, not interleaved
    resL.upload(pinL.array)
    resR.upload(pinR.array)
    resL.sync(), resR.sync()

    def resident_step():
        resL.run()
        resR.run()
        rc = api.lib().ivg_stereo_match_batch(resL._h, resR._h, KITTI["mbf"], KITTI["maxD"], None, None, resL.cap, 0)
        assert rc == 0, rc

    for _ in range(max(args.warmup, 3)):
        resident_step()
    resL.sync(), resR.sync()
    launches0 = resL.launch_count() + resR.launch_count()
    resL.profile_enable(True), resR.profile_enable(True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    resL.timer_start()
    for _ in range(args.steps):
        resident_step()
    resL.timer_stop()          # left stream: its last op (stereo) waits for the right stream
    ms_total = resL.timer_ms()
    resR.sync()
    barrier()
    clocks = sampler.result()
    launches = resL.launch_count() + resR.launch_count() - launches0
    prof = {}
    for k, (ms, cnt) in resL.profile_read().items():
        prof[k] = [ms, cnt]
    for k, (ms, cnt) in resR.profile_read().items():
        prof[k][0] += ms
        prof[k][1] += cnt
    resL.profile_enable(False), resR.profile_enable(False)

    ms_step = sharding.reduce_max(ms_total, "cuda") / args.steps       # max over ranks
    value = world * B / (ms_step * 1e-3)

    # ------------------------------------------------------------------ end-to-end arm (`e2e`)
    for _ in range(2):
        fe.process(pinL.array, pinR.array, out, KITTI["mbf"], KITTI["maxD"])
        fe.finish()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fe.process(pinL.array, pinR.array, out, KITTI["mbf"], KITTI["maxD"])
    fe.finish()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = world * B * args.steps / sharding.reduce_max(e2e_s, "cuda")
    h2d = int(pinL.array.nbytes + pinR.array.nbytes)
    d2h = int(sum(out[k].nbytes for k in ("kL", "dL", "nL", "kR", "dR", "nR", "uRight", "depth")))
    n_kp = int(out["nL"].sum())
    n_match = int((out["uRight"] >= 0).sum())

    if rank == 0:
        peak, peak_kind = measured_peak()
        alg, P = algorithmic_bytes(W, H)
        tot_ms = sum(v[0] for v in prof.values()) or 1.0
        prof = {k: v for k, v in prof.items() if v[1]}      # kernels that did not launch in this workload (e.g. the N4 prologue)
        shares = {k: v[0] / tot_ms for k, v in prof.items()}
        dom = max(prof, key=lambda k: prof[k][0])
        units = B      # images (extractor kernels: one launch per eye) or pairs (stereo kernels) per launch
        per_launch_ms = prof[dom][0] / max(prof[dom][1], 1)
        if dom == "k_resize_level":
            bytes_per_launch = alg[dom] * units / 7.0      # 7 launches per eye: per-launch bytes = total / 7
        elif dom == "k_stereo_match":
            bytes_per_launch = alg[dom] * units / 2.0      # index + match launches are accounted under one id
        else:
            bytes_per_launch = alg[dom] * units
        traffic = None
        try:     # measured DRAM bytes per image of the same kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "ncu_dram_bytes_per_image.json")) as f:
                per_img = json.load(f).get(dom)
            if per_img:
                traffic = per_img * units
        except Exception:
            pass
        achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
        pair_bytes = 21.9e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "C3: batch of KITTI-shape 1241x376 stereo pairs, nFeatures 2000, 8 levels, scale 1.2, iniTh 20, minTh 7, introspection off",
                       "pairs_per_gpu_per_step": B, "chunk_pairs": args.chunk, "slots": args.slots,
                       "l2": "inputs of one step (%.0f MB per GPU) exceed the 126 MB L2; no flush needed" % (h2d / 1e6),
                       "parallelism": "frame-parallel, %d independent rank(s), no collective" % world,
                       "numa": ("rank 0 bound to %d CPUs local to its GPU" % len(numa_cpus)) if numa_cpus else "no binding"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_kind, "bytes_per_launch": bytes_per_launch,
                         "ms_per_launch": per_launch_ms, "share_of_device_time": shares[dom],
                         "whole_pipeline_GBps": pair_bytes * value / world / 1e9},
            "kernel_shares": {k: round(v, 4) for k, v in shares.items()},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "check": {"keypoints_per_pair_left": n_kp / B, "stereo_matches_per_pair": n_match / B},
        }
        # drop-in usage (one frame at a time, two extractor objects driven from two host threads, then the matcher)
        lL, lR = api.ORBextractor(*a, False, device=local), api.ORBextractor(*a, False, device=local)
        def one_pair(i):
            tl = threading.Thread(target=lambda: lL(Lh[i % B]))
            tr = threading.Thread(target=lambda: lR(Rh[i % B]))
            tl.start(); tr.start(); tl.join(); tr.join()
            api.compute_stereo_matches(lL, lR, KITTI["mbf"], KITTI["maxD"])
        for i in range(5):
            one_pair(i)
        t1 = time.perf_counter()
        for i in range(50):
            one_pair(i)
        line["single_pair_latency_ms"] = 1e3 * (time.perf_counter() - t1) / 50
        lL.close(); lR.close()
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N=1 only
            ncpu = os.cpu_count() or 1
            n_s = min(args.cpu_sample, B)
            fps_all, dt_all, _, _ = cpu_frontend_fps(Lh[:n_s], Rh[:n_s], ncpu)
            from oracle import oracle_lib as O
            eL, eR = O.OracleExtractor(*a, False), O.OracleExtractor(*a, False)
            O.stereo_frame(eL, eR, Lh[0], Rh[0], None, KITTI["mbf"], KITTI["maxD"], threads=2)
            t1 = time.perf_counter()
            reps = 20
            for i in range(reps):
                O.stereo_frame(eL, eR, Lh[i % B], Rh[i % B], None, KITTI["mbf"], KITTI["maxD"], threads=2)
            fps_ref_threads = reps / (time.perf_counter() - t1)
            line["cpu_baseline"] = {"value": fps_all, "unit": UNIT, "cores": ncpu, "kind": "port",
                                    "sample": "first %d pairs of the workload, frame-parallel over %d threads, %.1f s wall" % (n_s, ncpu, dt_all),
                                    "reference_threading_2plus1": {"value": fps_ref_threads, "cores": 2,
                                                                   "note": "one frame at a time, 2 extraction threads + matching, as src/Frame.cc:115-125,:193"}}
        emit(line)

    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
