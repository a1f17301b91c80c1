"""MultiGpuStereoFrontend — frame-parallel stereo front-end over several GPUs of one node from ONE process.

Frames are independent: GPU g owns the contiguous frame range [g*T/G, (g+1)*T/G) of a batch (sharding.frame_range),
runs the whole hot path on it and copies its keypoint / descriptor / uRight / depth records straight into ITS SLICE of one
pinned host result array — the "gather" of the north star is the D2H copy itself, there is no collective and no second copy.
One host thread per GPU drives that GPU's StereoFrontend (the ctypes calls release the GIL, so the threads overlap).
bench.py's multi-rank arm (one process per GPU under torchrun, the driver's contract) uses the same per-GPU pipeline; this
class is the single-process form a SLAM front-end would embed.
"""
import threading

import numpy as np

from . import api, sharding
from .frontend import StereoFrontend


class MultiGpuStereoFrontend:
    def __init__(self, params, width, height, devices, chunk, slots=2, introspection=False):
        self.devices = list(devices)
        self.fes = [StereoFrontend(params, width, height, chunk, slots, device=d, introspection=introspection) for d in self.devices]
        self.cap = self.fes[0].cap
        self.seconds = [0.0] * len(self.devices)      # wall seconds each GPU's thread spent in its last process() call

    def alloc_outputs(self, n):
        """ONE set of pinned (portable: visible to every device) result arrays for all n frames."""
        return self.fes[0].alloc_outputs(n, pinned=True)

    def process(self, imgsL, imgsR, out, mbf, maxD, costs=None):
        """Blocking: every GPU processes its frame range; results land in `out` at the frames' own offsets."""
        import time
        n, G = imgsL.shape[0], len(self.fes)
        errs = []

        def work(g):
            try:
                s, e = sharding.frame_range(g, G, n)
                if e == s:
                    return
                t0 = time.perf_counter()
                view = {k: v[s:e] for k, v in out.items() if isinstance(v, np.ndarray)}
                self.fes[g].process(imgsL[s:e], imgsR[s:e], view, mbf, maxD, None if costs is None else costs[s:e])
                self.fes[g].finish()
                self.seconds[g] = time.perf_counter() - t0
            except Exception as ex:      # surfaced by the caller's thread
                errs.append(ex)
        th = [threading.Thread(target=work, args=(g,)) for g in range(G)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]

    def close(self):
        for fe in self.fes:
            fe.close()
        self.fes = []
