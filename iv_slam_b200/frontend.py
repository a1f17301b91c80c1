"""StereoFrontend — the orchestration of the reference's stereo Frame constructor for batches of frames.

Per stereo frame the reference runs ORBextractor::operator() for the left and right image on two threads and then
Frame::ComputeStereoMatches (introspective_ORB_SLAM/src/Frame.cc:115-125, :193).  Here a batch of frames is split into
chunks; each chunk goes to one of `slots` (a left + right extractor handle, i.e. two CUDA streams), so that the H2D
copy of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap.  Every call below only enqueues work;
`finish()` waits.  Frames are independent: no collective, no cross-frame state.
"""
import numpy as np

from . import api


class StereoFrontend:
    def __init__(self, params, width, height, chunk, slots=3, device=0, introspection=False, share_kernel_stream=True):
        """params: dict(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST); chunk = stereo pairs per launch group."""
        self.params, self.w, self.h, self.chunk, self.device = params, width, height, chunk, device
        a = (params["nfeatures"], params["scaleFactor"], params["nlevels"], params["iniThFAST"], params["minThFAST"])
        self.slots = []
        for _ in range(slots):
            left = api.ORBextractor(*a, introspection, device=device)     # right eye never weighted (src/Tracking.cc:182-183)
            right = api.ORBextractor(*a, False, device=device)
            left.reserve(width, height, chunk)
            right.reserve(width, height, chunk)
            # one kernel stream for everything on this device: kernels of different chunks / eyes never co-run (that
            # costs ~20 % at these sizes); the per-handle copy streams keep H2D/D2H overlapped with the kernels
            owner = self.slots[0][0] if self.slots else left
            if share_kernel_stream:
                if left is not owner:
                    left.share_stream(owner)
                right.share_stream(owner)
            self.slots.append((left, right))
        self.cap = self.slots[0][0].cap

    def alloc_outputs(self, n, pinned=True):
        """Host result arrays for n frames: cv::KeyPoint-layout records, descriptors, counts, mvuRight, mvDepth."""
        shapes = dict(kL=((n, self.cap), api.KP_DTYPE), dL=((n, self.cap, 32), np.uint8), nL=((n,), np.int32),
                      kR=((n, self.cap), api.KP_DTYPE), dR=((n, self.cap, 32), np.uint8), nR=((n,), np.int32),
                      uRight=((n, self.cap), np.float32), depth=((n, self.cap), np.float32))
        out, keep = {}, []
        for k, (shp, dt) in shapes.items():
            if pinned:
                p = api.PinnedArray(shp, dt)
                keep.append(p)
                out[k] = p.array
            else:
                out[k] = np.zeros(shp, dt)
        out["_pinned"] = keep
        return out

    def process(self, imgsL, imgsR, out, mbf, maxD, costs=None):
        """End-to-end: H2D of every chunk, kernels, D2H of (keypoints, descriptors, counts, uRight, depth). Async; call finish()."""
        n = imgsL.shape[0]
        for ci, s in enumerate(range(0, n, self.chunk)):
            e = min(s + self.chunk, n)
            left, right = self.slots[ci % len(self.slots)]
            left.upload(imgsL[s:e], None if costs is None else costs[s:e])
            right.upload(imgsR[s:e])
            left.run()
            right.run()
            left.download(out["kL"][s:e], out["dL"][s:e], out["nL"][s:e])
            right.download(out["kR"][s:e], out["dR"][s:e], out["nR"][s:e])
            api.compute_stereo_matches_batch(left, right, mbf, maxD, out["uRight"][s:e], out["depth"][s:e], sync=False)

    def finish(self):
        for left, right in self.slots:
            left.sync()
            right.sync()

    def launch_count(self):
        return sum(l.launch_count() + r.launch_count() for l, r in self.slots)

    def close(self):
        for l, r in reversed(self.slots):       # the stream owner (slot 0, left) goes last
            r.close()
            l.close()
        self.slots = []
