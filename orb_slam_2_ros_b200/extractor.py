"""ORBextractor — host-side mirror of ORB_SLAM2::ORBextractor (reference orb_slam2/include/ORBextractor.h:46-107)
over the C ABI of liborb_b200.so.  Same constructor arguments, same call (image, mask) -> (keypoints,
descriptors), same getters, and `mvImagePyramid` stays readable after a call (ORBextractor.h:85).
All computation happens in the CUDA library; there is no CPU path here."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, lib, ptr

EDGE_THRESHOLD = 19


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, device=0, max_batch=1):
        self._h = C.c_void_p()
        check(lib().orb_create(C.byref(self._h), nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device, max_batch))
        self.nfeatures, self.scaleFactor, self.nlevels = nfeatures, scaleFactor, nlevels
        self.iniThFAST, self.minThFAST, self.device, self.max_batch = iniThFAST, minThFAST, device, max_batch
        sc = np.zeros(nlevels, np.float32); isc = sc.copy(); s2 = sc.copy(); is2 = sc.copy()
        per = np.zeros(nlevels, np.int32)
        check(lib().orb_get_tables(self._h, ptr(sc), ptr(isc), ptr(s2), ptr(is2), ptr(per)))
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2 = sc, isc, s2, is2
        self.mnFeaturesPerLevel = per
        self.max_keypoints = lib().orb_max_keypoints(self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().orb_destroy(self._h)
            self._h = None

    __del__ = close

    # getters (ORBextractor.h:63-83)
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self.mvScaleFactor.copy()
    def GetInverseScaleFactors(self): return self.mvInvScaleFactor.copy()
    def GetScaleSigmaSquares(self): return self.mvLevelSigma2.copy()
    def GetInverseScaleSigmaSquares(self): return self.mvInvLevelSigma2.copy()

    def __call__(self, image, mask=None, rgb=False):
        """operator()(image, mask, keypoints, descriptors) — mask is ignored like in the reference
        (ORBextractor.h:58).  Returns (keypoints: KP_DTYPE[n], descriptors: uint8[n, 32]).
        A (h, w, 3|4) uint8 image is the camera frame BEFORE Tracking's cvtColor (Tracking.cc:181-204): the BGR(A) /
        RGB(A) (rgb=True, like mbRGB) -> gray conversion is fused into the level-0 pyramid copy on the device."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)   # ORBextractor.cc:1086-1087
        if image.dtype == np.uint8 and image.ndim == 3 and image.shape[2] in (3, 4):
            image = np.ascontiguousarray(image)
            h, w, ch = image.shape
            fmt = {(3, False): _lib.PIX_BGR8, (3, True): _lib.PIX_RGB8, (4, False): _lib.PIX_BGRA8, (4, True): _lib.PIX_RGBA8}[(ch, bool(rgb))]
            cap = self.max_keypoints
            kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); n = np.zeros(1, np.int32)
            check(lib().orb_extract_batch_pix(self._h, ptr(image), fmt, 1, w, h, image.strides[0], image.strides[0] * h, ptr(kps),
                                              ptr(desc), cap, ptr(n)))
            return kps[:n[0]].copy(), desc[:n[0]].copy()
        if image.dtype != np.uint8 or image.ndim != 2:
            raise TypeError("image must be CV_8UC1 (uint8, 2-D)")       # the reference asserts (ORBextractor.cc:1090)
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        cap = self.max_keypoints
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int32(0)
        check(lib().orb_extract(self._h, ptr(image), w, h, image.strides[0], ptr(kps), ptr(desc), cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        """images: uint8 [F, H, W] (host).  Returns a list of (keypoints, descriptors) per frame."""
        images = np.ascontiguousarray(images, np.uint8)
        F, h, w = images.shape
        cap = self.max_keypoints
        kps = np.zeros((F, cap), KP_DTYPE)
        desc = np.zeros((F, cap, 32), np.uint8)
        n = np.zeros(F, np.int32)
        check(lib().orb_extract_batch(self._h, ptr(images), F, w, h, w, w * h, ptr(kps), ptr(desc), cap, ptr(n)))
        return [(kps[f, :n[f]].copy(), desc[f, :n[f]].copy()) for f in range(F)]

    # ---- mvImagePyramid (ORBextractor.h:85): interior ROI views of the bordered level buffers ----
    def level_dims(self, level):
        w, h = C.c_int32(), C.c_int32()
        check(lib().orb_level_dims(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid_level_bordered(self, level, frame=0):
        w, h = self.level_dims(level)
        buf = np.zeros((h + 2 * EDGE_THRESHOLD, w + 2 * EDGE_THRESHOLD), np.uint8)
        check(lib().orb_pyramid_level(self._h, frame, level, ptr(buf), buf.strides[0]))
        return buf

    @property
    def mvImagePyramid(self):
        out = []
        for l in range(self.nlevels):
            b = self.pyramid_level_bordered(l)
            out.append(b[EDGE_THRESHOLD:-EDGE_THRESHOLD, EDGE_THRESHOLD:-EDGE_THRESHOLD])  # step = w + 38, like the Mat ROI
        return out

    # ---- stage taps for the parity tests ----
    def debug_blurred(self, level, frame=0):
        w, h = self.level_dims(level)
        buf = np.zeros((h, w), np.uint8)
        check(lib().orb_debug_blurred(self._h, frame, level, ptr(buf), w))
        return buf

    def debug_raw_corners(self, level, frame=0):
        n = C.c_int32(0)
        check(lib().orb_debug_raw_corners(self._h, frame, level, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.float32)
        check(lib().orb_debug_raw_corners(self._h, frame, level, ptr(out), n.value, C.byref(n)))
        return out[:n.value]

    def debug_tie_counts(self, frame=0):
        t = np.zeros(self.nlevels, np.int32)
        check(lib().orb_debug_tie_counts(self._h, frame, ptr(t)))
        return t

    # ---- device-resident form (bench / GPU callers): raw device pointers, async on `stream` ----
    def set_stream(self, cuda_stream_handle):
        check(lib().orb_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def extract_batch_device(self, d_imgs, nframes, w, h, row_stride, frame_stride, d_kps, d_desc, cap, d_n):
        check(lib().orb_extract_batch_device(self._h, C.c_void_p(d_imgs), nframes, w, h, row_stride, frame_stride,
                                             C.c_void_p(d_kps), C.c_void_p(d_desc), cap, C.c_void_p(d_n)))

    def sync(self):
        check(lib().orb_sync(self._h))

    def launch_count(self):
        return lib().orb_launch_count(self._h)

    # ---- per-stage device timers (CUDA events recorded by the library around every stage) ----
    STAGES = ("pyramid", "fast_cells", "quadtree", "blur", "orient_describe")

    def profile_enable(self, on=True):
        check(lib().orb_profile_enable(self._h, int(on)))

    def profile_read(self, reset=True):
        """-> ({stage: accumulated ms}, calls, frames)"""
        ms = (C.c_double * len(self.STAGES))()
        calls, frames = C.c_int64(0), C.c_int64(0)
        check(lib().orb_profile_read(self._h, ms, C.byref(calls), C.byref(frames), int(reset)))
        return dict(zip(self.STAGES, list(ms))), calls.value, frames.value
