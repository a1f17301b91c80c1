"""ctypes binding of the C ABI (include/ivslam_gpu.h) — the host-side mirror of the reference interface.

`ORBextractor` has the reference class's constructor arguments and call semantics
(introspective_ORB_SLAM/include/ORBextractor.h:54-128): `kps, desc = ex(image, mask)`, `GetLevels()`,
`GetScaleFactors()`, `mvImagePyramid`-style level access.  `compute_stereo_matches(left, right, mbf, maxD)` is
Frame::ComputeStereoMatches (src/Frame.cc:758-932) on what the two extractors hold on the device.

There is NO CPU fallback: importing works anywhere (so the build check can import the package), but creating an
extractor without the compiled library or without an sm_100 GPU raises.
"""
import atexit
import ctypes as C
import os
import weakref

import numpy as np

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_HERE = os.path.dirname(os.path.abspath(__file__))
# IVSLAM_GPU_LIB: developer override to A/B a differently-built libivslam_gpu.so (same C ABI); never a fallback
LIB_PATH = os.environ.get("IVSLAM_GPU_LIB") or os.path.join(_HERE, "lib", "libivslam_gpu.so")

IVG_OK = 0
KERNEL_NAMES = ("k_resize_level", "k_fast_cells", "k_gauss7", "k_level_select", "k_orient_describe", "k_stereo_match",
                "k_stereo_median", "k_prologue", "k_proj_candidates", "k_proj_resolve")


class IvgError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        L = _lib
        msg = L.ivg_strerror(status).decode() if L else str(status)
        detail = L.ivg_last_cuda_error().decode() if (L and status == -4) else ""
        super().__init__("%s: %s (%d) %s" % (what, msg, status, detail))


_lib = None


def lib():
    """Loads iv_slam_b200/lib/libivslam_gpu.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libivslam_gpu.so is not built (%s); run `make -C iv_slam_b200/csrc` — there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32p, f32p = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    sz = C.c_size_t
    sig = {
        "ivg_strerror": (C.c_char_p, [C.c_int]),
        "ivg_last_cuda_error": (C.c_char_p, []),
        "ivg_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, i32p, i32p]),
        "ivg_extractor_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
        "ivg_extractor_destroy": (None, [vp]),
        "ivg_extractor_set_mode": (C.c_int, [vp, C.c_int]),
        "ivg_extractor_reserve": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
        "ivg_get_levels": (C.c_int, [vp]),
        "ivg_get_scale_factor": (C.c_float, [vp]),
        "ivg_get_scale_table": (C.c_int, [vp, C.c_int, vp]),
        "ivg_get_features_per_level": (C.c_int, [vp, vp]),
        "ivg_max_keypoints": (C.c_int, [vp]),
        "ivg_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, sz, vp, sz, vp, vp, C.c_int, i32p]),
        "ivg_extract_batch": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, sz, sz, vp, sz, sz, vp, vp, C.c_int, vp]),
        "ivg_extract_stereo": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, sz, vp, sz, vp, vp, i32p, vp, vp, i32p, C.c_float, C.c_float, vp, vp, C.c_int]),
        "ivg_upload_batch": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, sz, sz, vp, sz, sz]),
        "ivg_upload_batch_device": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, sz, sz, vp, sz, sz]),
        "ivg_upload_batch_device_cost_f32": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, sz, sz, vp, sz, sz]),
        "ivg_set_rectify_maps": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, sz]),
        "ivg_search_by_projection_last": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp] + [C.c_float] * 9 +
                                          [C.c_int, C.c_float, C.c_int, vp, C.c_int, i32p]),
        "ivg_search_by_bow": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, C.c_float, C.c_int, vp, C.c_int, i32p]),
        "ivg_search_by_projection_map": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp] + [C.c_float] * 6 + [vp, C.c_int, i32p]),
        "ivg_upload_batch_raw": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, sz, sz, C.c_int, C.c_int, vp, sz, sz]),
        "ivg_run_batch": (C.c_int, [vp]),
        "ivg_download_batch": (C.c_int, [vp, vp, vp, C.c_int, vp]),
        "ivg_sync": (C.c_int, [vp]),
        "ivg_share_stream": (C.c_int, [vp, vp]),
        "ivg_set_batch": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "ivg_device_input": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(sz)]),
        "ivg_level_size": (C.c_int, [vp, C.c_int, i32p, i32p]),
        "ivg_get_pyramid_level": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, sz]),
        "ivg_get_level_keypoints": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, i32p]),
        "ivg_stereo_match": (C.c_int, [vp, vp, C.c_float, C.c_float, vp, vp, C.c_int]),
        "ivg_stereo_match_batch": (C.c_int, [vp, vp, C.c_float, C.c_float, vp, vp, C.c_int, C.c_int]),
        "ivg_stereo_match_keypoints": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_float, C.c_float, vp, vp]),
        "ivg_compute_pyramid": (C.c_int, [vp, vp, C.c_int, C.c_int, sz]),
        "ivg_frame_postprocess_batch": (C.c_int, [vp, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, vp, C.c_int, C.c_int]),
        "ivg_timer_start": (C.c_int, [vp]),
        "ivg_timer_stop": (C.c_int, [vp]),
        "ivg_timer_elapsed_ms": (C.c_int, [vp, f32p]),
        "ivg_launch_count": (C.c_longlong, [vp]),
        "ivg_host_alloc": (C.c_int, [C.POINTER(vp), sz]),
        "ivg_host_free": (C.c_int, [vp]),
        "ivg_flush_l2": (C.c_int, [vp, sz]),
        "ivg_set_graph_mode": (C.c_int, [vp, C.c_int]),
        "ivg_profile_enable": (C.c_int, [vp, C.c_int]),
        "ivg_profile_read": (C.c_int, [vp, vp, vp]),
        "ivg_debug_nth_element": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, vp]),
        "ivg_debug_nth_element_block": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, vp]),
        "ivg_debug_force_config": (C.c_int, [vp, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    L._signatures = sig
    _lib = L
    return L


def exported_symbols():
    """Names declared in include/ivslam_gpu.h that the binding expects (used by the CPU-side ABI test)."""
    lib()
    return sorted(_lib._signatures)


def _ck(status, what):
    if status != IVG_OK:
        raise IvgError(status, what)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_info(device=0):
    name = C.create_string_buffer(128)
    sm, cnt = C.c_int(), C.c_int()
    rc = lib().ivg_device_info(device, name, 128, C.byref(sm), C.byref(cnt))
    return rc, name.value.decode(), sm.value, cnt.value


_live = weakref.WeakSet()      # extractors and pinned buffers still open: closed in a safe order before the CUDA runtime unloads


def _close_all():
    objs = list(_live)
    for o in objs:          # handles that borrow a stream first, then stream owners, then pinned memory
        if isinstance(o, ORBextractor) and getattr(o, "_stream_owner", None) is not None:
            o.close()
    for o in objs:
        if isinstance(o, ORBextractor):
            o.close()
    for o in objs:
        if isinstance(o, PinnedArray):
            o.free()


atexit.register(_close_all)


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (for real async H2D/D2H in the batch path)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        self.ptr = C.c_void_p()
        _ck(lib().ivg_host_alloc(C.byref(self.ptr), max(self.nbytes, 1)), "ivg_host_alloc")
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)
        _live.add(self)

    def free(self):
        if self.ptr:
            self.array = None
            lib().ivg_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ORBextractor:
    """ORB_SLAM2::ORBextractor on a B200. Same constructor arguments as the reference (ORBextractor.h:57-58)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, enableIntrospection=False, device=0):
        self._h = C.c_void_p()
        _ck(lib().ivg_extractor_create(C.byref(self._h), device, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                                       int(bool(enableIntrospection))), "ivg_extractor_create")
        self.nfeatures, self.nlevels, self.device = nfeatures, nlevels, device
        self.params = dict(nfeatures=nfeatures, scaleFactor=scaleFactor, nlevels=nlevels, iniThFAST=iniThFAST, minThFAST=minThFAST,
                           introspection=bool(enableIntrospection))
        self.mode = 0
        self.cap = lib().ivg_max_keypoints(self._h)
        self._batch = 0
        self._stream_owner = None
        self._inflight = []          # host arrays of asynchronous copies, kept alive until sync()
        _live.add(self)

    def close(self):
        if getattr(self, "_h", None):
            lib().ivg_extractor_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_keypoint_mode(self, mode):
        """0: ComputeKeyPointsOld (the reference's live path, default); 1: ComputeKeyPointsOctTree (dead code there)."""
        _ck(lib().ivg_extractor_set_mode(self._h, int(mode)), "ivg_extractor_set_mode")
        self.mode = int(mode)
        self.cap = lib().ivg_max_keypoints(self._h)

    # -- getters of the reference class (ORBextractor.h:69-91)
    def GetLevels(self):
        return lib().ivg_get_levels(self._h)

    def GetScaleFactor(self):
        return lib().ivg_get_scale_factor(self._h)

    def _table(self, which):
        out = np.zeros(self.nlevels, np.float32)
        _ck(lib().ivg_get_scale_table(self._h, which, _p(out)), "ivg_get_scale_table")
        return out

    def GetScaleFactors(self):
        return self._table(0)

    def GetInverseScaleFactors(self):
        return self._table(1)

    def GetScaleSigmaSquares(self):
        return self._table(2)

    def GetInverseScaleSigmaSquares(self):
        return self._table(3)

    def features_per_level(self):
        out = np.zeros(self.nlevels, np.int32)
        _ck(lib().ivg_get_features_per_level(self._h, _p(out)), "ivg_get_features_per_level")
        return out

    def reserve(self, width, height, max_batch):
        _ck(lib().ivg_extractor_reserve(self._h, width, height, max_batch), "ivg_extractor_reserve")

    # -- operator()
    def __call__(self, image, mask=None):
        """(keypoints[KP_DTYPE], descriptors[N,32]) for one 8-bit gray image; mask = IV-SLAM cost-map or None."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
        if mask is not None:
            assert mask.dtype == np.uint8 and mask.shape == image.shape and mask.strides[1] == 1
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int(0)
        _ck(lib().ivg_extract(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], _p(mask),
                              mask.strides[0] if mask is not None else 0, _p(kps), _p(desc), self.cap, C.byref(n)), "ivg_extract")
        self._batch = 1
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, masks=None):
        """images: [n,H,W] u8 C-contiguous. Returns (kps[n,cap], desc[n,cap,32], counts[n])."""
        assert images.dtype == np.uint8 and images.ndim == 3 and images.flags.c_contiguous
        n, H, W = images.shape
        if masks is not None:
            assert masks.dtype == np.uint8 and masks.shape == images.shape and masks.flags.c_contiguous
        kps = np.zeros((n, self.cap), KP_DTYPE)
        desc = np.zeros((n, self.cap, 32), np.uint8)
        cnt = np.zeros(n, np.int32)
        _ck(lib().ivg_extract_batch(self._h, n, _p(images), W, H, W, H * W, _p(masks), W, H * W, _p(kps), _p(desc), self.cap, _p(cnt)),
            "ivg_extract_batch")
        self._batch = n
        return kps, desc, cnt

    # -- split phases (bench / pipelining)
    def upload(self, images, masks=None):
        """Asynchronous H2D of [n,H,W] u8 frames (rows may be strided, pixels must be contiguous).  The arrays are referenced
        until sync(): the DMA may still be reading them when this returns."""
        assert images.dtype == np.uint8 and images.ndim == 3 and images.strides[2] == 1
        n, H, W = images.shape
        if masks is not None:
            assert masks.dtype == np.uint8 and masks.shape == images.shape and masks.strides[2] == 1
        self._inflight.extend((images, masks))
        _ck(lib().ivg_upload_batch(self._h, n, _p(images), W, H, images.strides[1], images.strides[0], _p(masks),
                                   masks.strides[1] if masks is not None else 0, masks.strides[0] if masks is not None else 0), "ivg_upload_batch")
        self._batch = n

    def set_rectify_maps(self, mapx, mapy):
        """N4: the CV_32FC1 maps of cv::initUndistortRectifyMap (stereo_kitti.cc:284-343); None, None clears them."""
        if mapx is None or mapy is None:
            _ck(lib().ivg_set_rectify_maps(self._h, None, None, 0, 0, 0), "ivg_set_rectify_maps")
            return
        mapx, mapy = np.ascontiguousarray(mapx, np.float32), np.ascontiguousarray(mapy, np.float32)
        assert mapx.ndim == 2 and mapx.shape == mapy.shape
        _ck(lib().ivg_set_rectify_maps(self._h, _p(mapx), _p(mapy), mapx.shape[1], mapx.shape[0], mapx.shape[1]), "ivg_set_rectify_maps")

    def upload_raw(self, frames, rgb=False, masks=None):
        """N4: raw camera frames [n, H, W] or [n, H, W, 3|4] (u8): remap (if maps are set) + cvtColor on the device."""
        assert frames.dtype == np.uint8 and frames.ndim in (3, 4)
        n, H, W = frames.shape[:3]
        cn = 1 if frames.ndim == 3 else frames.shape[3]
        assert frames.strides[2] == cn and (cn == 1 or frames.strides[3] == 1)
        if masks is not None:
            assert masks.dtype == np.uint8 and masks.shape == (n, H, W) and masks.strides[2] == 1
        _ck(lib().ivg_upload_batch_raw(self._h, n, _p(frames), W, H, frames.strides[1], frames.strides[0], cn, int(bool(rgb)), _p(masks),
                                       masks.strides[1] if masks is not None else 0, masks.strides[0] if masks is not None else 0), "ivg_upload_batch_raw")
        self._batch = n

    def extract_raw(self, frame, rgb=False, mask=None):
        """remap + cvtColor + operator() for one raw frame: (keypoints, descriptors)."""
        self.upload_raw(frame[None], rgb, None if mask is None else mask[None])
        self.run()
        kps = np.zeros((1, self.cap), KP_DTYPE)
        desc = np.zeros((1, self.cap, 32), np.uint8)
        cnt = np.zeros(1, np.int32)
        self.download(kps, desc, cnt)
        self.sync()
        return kps[0, :cnt[0]].copy(), desc[0, :cnt[0]].copy()

    def search_by_projection_last(self, world_pos, desc, octave, angle, flags, Rcw, tcw, cam, bounds, mode=0, th=7.0,
                                  check_orientation=True, index=0):
        """N2: ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) on frame `index` (needs frame_postprocess).
        cam = (fx, fy, cx, cy, mbf); bounds = (minX, maxX, minY, maxY).  Returns (match[cap] int32, nmatches)."""
        f32, i32, u8 = np.float32, np.int32, np.uint8
        world_pos, desc = np.ascontiguousarray(world_pos, f32), np.ascontiguousarray(desc, u8)
        octave, angle, flags = np.ascontiguousarray(octave, i32), np.ascontiguousarray(angle, f32), np.ascontiguousarray(flags, u8)
        Rcw, tcw = np.ascontiguousarray(Rcw, f32), np.ascontiguousarray(tcw, f32)
        n = flags.size
        assert world_pos.shape == (n, 3) and desc.shape == (n, 32) and octave.size == n and angle.size == n and Rcw.size == 9 and tcw.size == 3
        match = np.zeros(self.cap, i32)
        nm = C.c_int(0)
        _ck(lib().ivg_search_by_projection_last(self._h, index, n, _p(world_pos), _p(desc), _p(octave), _p(angle), _p(flags), _p(Rcw), _p(tcw),
                                                *[float(v) for v in cam], *[float(v) for v in bounds], int(mode), float(th),
                                                int(bool(check_orientation)), _p(match), self.cap, C.byref(nm)), "ivg_search_by_projection_last")
        return match, nm.value

    def search_by_projection_map(self, proj, view_cos, level, desc, flags, bounds, cur_blocked=None, th=1.0, nnratio=0.8, index=0):
        """N2: ORBmatcher::SearchByProjection(F, vpMapPoints, th).  proj[n,3] = (mTrackProjX, mTrackProjY, mTrackProjXR)."""
        f32, i32, u8 = np.float32, np.int32, np.uint8
        proj, view_cos, level = np.ascontiguousarray(proj, f32), np.ascontiguousarray(view_cos, f32), np.ascontiguousarray(level, i32)
        desc, flags = np.ascontiguousarray(desc, u8), np.ascontiguousarray(flags, u8)
        n = flags.size
        assert proj.shape == (n, 3) and desc.shape == (n, 32) and view_cos.size == n and level.size == n
        if cur_blocked is not None:
            cur_blocked = np.ascontiguousarray(cur_blocked, u8)
            assert cur_blocked.size >= self.cap
        match = np.zeros(self.cap, i32)
        nm = C.c_int(0)
        _ck(lib().ivg_search_by_projection_map(self._h, index, n, _p(proj), _p(view_cos), _p(level), _p(desc), _p(flags), _p(cur_blocked),
                                               *[float(v) for v in bounds], float(th), float(nnratio), _p(match), self.cap, C.byref(nm)),
            "ivg_search_by_projection_map")
        return match, nm.value

    def search_by_bow(self, desc, angle, flags, node_slot, node_start, node_idx, nnratio=0.7, check_orientation=True, index=0):
        """N2: ORBmatcher::SearchByBoW(pKF, F, matches) on frame `index`; points in the reference's traversal order."""
        f32, i32, u8 = np.float32, np.int32, np.uint8
        desc, angle, flags = np.ascontiguousarray(desc, u8), np.ascontiguousarray(angle, f32), np.ascontiguousarray(flags, u8)
        node_slot, node_start, node_idx = (np.ascontiguousarray(a, i32) for a in (node_slot, node_start, node_idx))
        n = flags.size
        assert desc.shape == (n, 32) and angle.size == n and node_slot.size == n and node_start.size >= 1
        match = np.zeros(self.cap, i32)
        nm = C.c_int(0)
        _ck(lib().ivg_search_by_bow(self._h, index, n, _p(desc), _p(angle), _p(flags), _p(node_slot), node_start.size - 1, _p(node_start),
                                    _p(node_idx), float(nnratio), int(bool(check_orientation)), _p(match), self.cap, C.byref(nm)), "ivg_search_by_bow")
        return match, nm.value

    def upload_device(self, n, width, height, d_images, d_masks=None):
        """Frames already in device memory (integer device pointers to n contiguous HxW u8 frames)."""
        _ck(lib().ivg_upload_batch_device(self._h, n, C.c_void_p(d_images), width, height, width, width * height,
                                          C.c_void_p(d_masks) if d_masks else None, width, width * height), "ivg_upload_batch_device")
        self._batch = n

    def upload_device_cost_f32(self, n, width, height, d_images, d_costs_f32):
        """Device-resident u8 frames plus FLOAT cost-maps as the introspection CNN emits them (converted like (t * 255).to(uint8))."""
        _ck(lib().ivg_upload_batch_device_cost_f32(self._h, n, C.c_void_p(d_images), width, height, width, width * height,
                                                   C.c_void_p(d_costs_f32), width, width * height), "ivg_upload_batch_device_cost_f32")
        self._batch = n

    def run(self):
        _ck(lib().ivg_run_batch(self._h), "ivg_run_batch")

    def download(self, kps, desc, cnt):
        """Asynchronous D2H into caller arrays (referenced until sync())."""
        assert kps.dtype == KP_DTYPE and desc.dtype == np.uint8 and cnt.dtype == np.int32
        assert kps.flags.c_contiguous and desc.flags.c_contiguous and cnt.flags.c_contiguous
        self._inflight.extend((kps, desc, cnt))
        _ck(lib().ivg_download_batch(self._h, _p(kps), _p(desc), kps.shape[-1], _p(cnt)), "ivg_download_batch")

    def sync(self):
        _ck(lib().ivg_sync(self._h), "ivg_sync")
        self._inflight.clear()

    def share_stream(self, owner):
        """Run this extractor's kernels on `owner`'s stream (no kernel overlap between the two; copies still overlap)."""
        _ck(lib().ivg_share_stream(self._h, owner._h), "ivg_share_stream")
        self._stream_owner = owner      # keep the owner alive

    def compute_pyramid(self, image):
        _ck(lib().ivg_compute_pyramid(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0]), "ivg_compute_pyramid")
        self._batch = 1

    # -- mvImagePyramid / mvQualityImagePyramid access
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _ck(lib().ivg_level_size(self._h, level, C.byref(w), C.byref(h)), "ivg_level_size")
        return w.value, h.value

    def level(self, level, which=0, index=0):
        """which: 0 mvImagePyramid[level], 1 blurred working copy, 2 mvQualityImagePyramid[level]."""
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _ck(lib().ivg_get_pyramid_level(self._h, index, level, which, _p(out), out.strides[0]), "ivg_get_pyramid_level")
        return out

    def level_keypoints(self, level, index=0):
        cap = self.cap
        x, y, r = (np.zeros(cap, np.float32) for _ in range(3))
        n = C.c_int(0)
        _ck(lib().ivg_get_level_keypoints(self._h, index, level, _p(x), _p(y), _p(r), cap, C.byref(n)), "ivg_get_level_keypoints")
        return x[:n.value].copy(), y[:n.value].copy(), r[:n.value].copy()

    def frame_postprocess(self, minX, maxX, minY, maxY):
        """mvKeyQualScore + AssignFeaturesToGrid for every frame of the last batch -> (qual[n,cap], gridStart[n,3073], gridIdx[n,cap])."""
        n = self._batch
        qual = np.zeros((n, self.cap), np.float32)
        gs = np.zeros((n, 64 * 48 + 1), np.int32)
        gi = np.zeros((n, self.cap), np.int32)
        _ck(lib().ivg_frame_postprocess_batch(self._h, minX, maxX, minY, maxY, _p(qual), _p(gs), _p(gi), self.cap, 1), "ivg_frame_postprocess_batch")
        return qual, gs, gi

    # -- measurement
    def timer_start(self):
        _ck(lib().ivg_timer_start(self._h), "ivg_timer_start")

    def timer_stop(self):
        _ck(lib().ivg_timer_stop(self._h), "ivg_timer_stop")

    def timer_ms(self):
        ms = C.c_float()
        _ck(lib().ivg_timer_elapsed_ms(self._h, C.byref(ms)), "ivg_timer_elapsed_ms")
        return ms.value

    def launch_count(self):
        return lib().ivg_launch_count(self._h)

    def profile_enable(self, on=True):
        _ck(lib().ivg_profile_enable(self._h, int(on)), "ivg_profile_enable")

    def profile_read(self):
        """{kernel name: (total ms, launches)} since profile_enable."""
        ms = np.zeros(len(KERNEL_NAMES), np.float64)
        cnt = np.zeros(len(KERNEL_NAMES), np.int64)
        _ck(lib().ivg_profile_read(self._h, _p(ms), _p(cnt)), "ivg_profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNEL_NAMES)}

    def debug_force_config(self, mode):
        """Test hook: 0 automatic, 1 always the throughput kernels, 2 the one-frame kernels wherever they apply."""
        _ck(lib().ivg_debug_force_config(self._h, int(mode)), "ivg_debug_force_config")

    def set_graph_mode(self, on=True):
        """Replay the kernel sequence of a run as one CUDA graph (single-frame latency)."""
        _ck(lib().ivg_set_graph_mode(self._h, int(on)), "ivg_set_graph_mode")

    def flush_l2(self, nbytes=256 << 20):
        _ck(lib().ivg_flush_l2(self._h, nbytes), "ivg_flush_l2")


def compute_stereo_matches(left, right, mbf, maxD):
    """Frame::ComputeStereoMatches for frame 0 of the last extraction -> (mvuRight[N], mvDepth[N]) sized to left.cap."""
    u = np.empty(left.cap, np.float32)
    d = np.empty(left.cap, np.float32)
    _ck(lib().ivg_stereo_match(left._h, right._h, mbf, maxD, _p(u), _p(d), left.cap), "ivg_stereo_match")
    return u, d


def extract_stereo(left, right, image_left, image_right, mbf, maxD, mask_left=None):
    """Both eyes and the matcher in one call from one thread (ivg_extract_stereo) ->
    (kpsL, descL, kpsR, descR, mvuRight[nL], mvDepth[nL])."""
    for im in (image_left, image_right):
        assert im.dtype == np.uint8 and im.ndim == 2 and im.strides[1] == 1
    assert image_left.shape == image_right.shape and image_left.strides == image_right.strides
    if mask_left is not None:
        assert mask_left.dtype == np.uint8 and mask_left.shape == image_left.shape and mask_left.strides[1] == 1
    cap = max(left.cap, right.cap)
    kL, kR = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dL, dR = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    u, d = np.empty(cap, np.float32), np.empty(cap, np.float32)
    nL, nR = C.c_int(0), C.c_int(0)
    _ck(lib().ivg_extract_stereo(left._h, right._h, _p(image_left), _p(image_right), image_left.shape[1], image_left.shape[0],
                                 image_left.strides[0], _p(mask_left), mask_left.strides[0] if mask_left is not None else 0,
                                 _p(kL), _p(dL), C.byref(nL), _p(kR), _p(dR), C.byref(nR), mbf, maxD, _p(u), _p(d), cap), "ivg_extract_stereo")
    left._batch = right._batch = 1
    return kL[:nL.value].copy(), dL[:nL.value].copy(), kR[:nR.value].copy(), dR[:nR.value].copy(), u[:nL.value].copy(), d[:nL.value].copy()


def compute_stereo_matches_batch(left, right, mbf, maxD, uRight=None, depth=None, sync=True):
    n = left._batch
    if uRight is None:
        uRight = np.empty((n, left.cap), np.float32)
        depth = np.empty((n, left.cap), np.float32)
    assert uRight.dtype == np.float32 and depth.dtype == np.float32 and uRight.flags.c_contiguous and depth.flags.c_contiguous
    if not sync:
        left._inflight.extend((uRight, depth))      # the D2H may still be writing them when this returns
    _ck(lib().ivg_stereo_match_batch(left._h, right._h, mbf, maxD, _p(uRight), _p(depth), uRight.shape[-1], int(sync)), "ivg_stereo_match_batch")
    return uRight, depth


def compute_stereo_matches_keypoints(left, right, kL, dL, kR, dR, mbf, maxD):
    """The matcher on caller-supplied keypoints/descriptors against the pyramids resident in the two extractors."""
    kL = np.ascontiguousarray(kL, KP_DTYPE)
    kR = np.ascontiguousarray(kR, KP_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8)
    dR = np.ascontiguousarray(dR, np.uint8)
    u = np.full(kL.size, -1, np.float32)
    d = np.full(kL.size, -1, np.float32)
    _ck(lib().ivg_stereo_match_keypoints(left._h, right._h, _p(kL), kL.size, _p(dL), _p(kR), kR.size, _p(dR), mbf, maxD, _p(u), _p(d)),
        "ivg_stereo_match_keypoints")
    return u, d


def debug_nth_element(keys, nth, device=0, block=False):
    """Permutation produced by the GPU's warp-parallel (or, block=True, CTA-parallel) nth_element replay (test hook)."""
    keys = np.ascontiguousarray(keys, np.uint32)
    order = np.zeros(keys.size, np.uint32)
    fn = lib().ivg_debug_nth_element_block if block else lib().ivg_debug_nth_element
    _ck(fn(device, _p(keys), keys.size, int(nth), _p(order)), "ivg_debug_nth_element")
    return order
