"""ctypes front-end of the rh_* harness (oracle/ref_shim/harness.cpp).  TEST INFRASTRUCTURE ONLY.

The same C entry points exist in two shared libraries:
  * oracle/_ref/liborb_ref.so   — the reference's own, unmodified ORBextractor.cc / ORBmatcher.cc / Frame.cc / DBoW2
                                  compiled over the OpenCV stand-in (arm "reference"): the checker;
  * tests/_build/liborb_dropin.so — the same harness and the same reference Frame.cc / headers with this repo's drop-in
                                  ORBextractor.cc / ORBmatcher.cc / Frame::ComputeStereoMatches over liborb_b200.so
                                  (arm "b200"): the thing being checked.
`Harness(path)` binds one of them.  /root/reference is only needed to BUILD them (in the authoring container); the GPU
box uses the prebuilt files.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "liborb_ref.so")
DROPIN_SO = os.path.join(os.path.dirname(_HERE), "tests", "_build", "liborb_dropin.so")

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


class Harness:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        # RTLD_LOCAL (ctypes default): the two arms define the same C++ symbols (ORB_SLAM2::Frame, ...) and must not
        # see each other's; both are linked -Bsymbolic as well.
        self.L = L = C.CDLL(path)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.rh_arm.restype = C.c_char_p
        L.rh_extractor_create.restype = vp
        L.rh_extractor_create.argtypes = [i32, f32, i32, i32, i32]
        L.rh_extractor_destroy.argtypes = [vp]
        L.rh_extractor_tables.argtypes = [vp] * 5
        L.rh_extract.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32]
        L.rh_level_dims.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
        L.rh_get_level.argtypes = [vp, i32, vp]
        L.rh_frame_mono.restype = vp
        L.rh_frame_mono.argtypes = [vp, vp, i32, i32, i32, vp, f32, f32]
        L.rh_frame_stereo.restype = vp
        L.rh_frame_stereo.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, f32, f32]
        L.rh_frame_from_arrays.restype = vp
        L.rh_frame_from_arrays.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, f32, f32]
        L.rh_frame_copy.restype = vp
        L.rh_frame_copy.argtypes = [vp]
        L.rh_frame_destroy.argtypes = [vp]
        L.rh_frame_n.argtypes = [vp]
        L.rh_frame_get.argtypes = [vp] * 6
        L.rh_frame_right_n.argtypes = [vp]
        L.rh_frame_get_right.argtypes = [vp] * 3
        L.rh_frame_bounds.argtypes = [vp]
        L.rh_frame_set_pose.argtypes = [vp, vp]
        L.rh_frame_features_in_area.argtypes = [vp, f32, f32, f32, i32, i32, vp, i32]
        L.rh_frame_set_featvec.argtypes = [vp, vp, vp, vp, i32]
        L.rh_points_create.restype = vp
        L.rh_points_create.argtypes = [i32] + [vp] * 7
        L.rh_points_destroy.argtypes = [vp]
        L.rh_points_state.argtypes = [vp] * 4
        L.rh_points_track_state.argtypes = [vp] * 7
        L.rh_frame_set_points.argtypes = [vp] * 3
        L.rh_frame_get_points.argtypes = [vp] * 3
        L.rh_frame_set_outliers.argtypes = [vp, vp]
        L.rh_keyframe_create.restype = vp
        L.rh_keyframe_create.argtypes = [vp]
        L.rh_keyframe_destroy.argtypes = [vp]
        L.rh_keyframe_set_pose.argtypes = [vp, vp]
        L.rh_keyframe_set_bad.argtypes = [vp, i32]
        L.rh_keyframe_set_points.argtypes = [vp, vp, vp, i32]
        L.rh_keyframe_get_points.argtypes = [vp] * 3
        L.rh_descriptor_distance.argtypes = [vp, vp]
        L.rh_matcher_constants.argtypes = [vp] * 3
        L.rh_search_local_points.argtypes = [f32, i32, vp, vp, vp, i32, f32, i32]
        L.rh_search_last_frame.argtypes = [f32, i32, vp, vp, f32, i32]
        L.rh_search_reloc.argtypes = [f32, i32, vp, vp, vp, vp, i32, f32, i32]
        L.rh_search_loop.argtypes = [f32, i32, vp, vp, vp, vp, i32, vp, i32]
        L.rh_search_bow_kf_frame.argtypes = [f32, i32, vp, vp, vp, vp]
        L.rh_search_bow_kf_kf.argtypes = [f32, i32, vp, vp, vp, vp]
        L.rh_search_initialization.argtypes = [f32, i32, vp, vp, vp, vp, i32]
        L.rh_search_triangulation.argtypes = [f32, i32, vp, vp, vp, vp, i32, i32]
        L.rh_search_sim3.argtypes = [f32, i32, vp, vp, vp, vp, f32, vp, vp, f32]
        L.rh_fuse.argtypes = [f32, i32, vp, vp, vp, i32, f32]
        L.rh_fuse_sim3.argtypes = [f32, i32, vp, vp, vp, vp, i32, f32, vp]
        L.rh_voc_load_text.restype = vp
        L.rh_voc_load_text.argtypes = [C.c_char_p]
        L.rh_voc_destroy.argtypes = [vp]
        L.rh_voc_size.argtypes = [vp]
        L.rh_frame_compute_bow.argtypes = [vp, vp]
        L.rh_frame_get_bow.argtypes = [vp, vp, vp, i32]
        L.rh_frame_get_featvec.argtypes = [vp, vp, vp, vp, i32, i32]
        L.rh_voc_score.restype = C.c_double
        L.rh_voc_score.argtypes = [vp, vp, vp]
        L.rh_probe_gemm.argtypes = [vp, i32, i32, vp, i32, i32, vp, vp]
        L.rh_probe_norm.restype = C.c_double
        L.rh_probe_norm.argtypes = [vp, i32]
        L.rh_probe_dot.restype = C.c_double
        L.rh_probe_dot.argtypes = [vp, vp, i32]
        self.arm = L.rh_arm().decode()

    # ---- extractor ----
    def extractor(self, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7):
        return Extractor(self, nfeatures, scale, nlevels, ini, mn)

    def reset_calibration(self):
        self.L.rh_reset_calibration()

    def bounds(self):
        b = np.zeros(4, np.float32)
        self.L.rh_frame_bounds(_p(b))
        return b

    def descriptor_distance(self, a, b):
        a, b = _u8(a), _u8(b)
        return self.L.rh_descriptor_distance(_p(a), _p(b))

    def constants(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self.L.rh_matcher_constants(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def points(self, n, pos=None, normal=None, desc=None, nobs=None, bad=None, min_dist=None, max_dist=None):
        return Points(self, n, pos, normal, desc, nobs, bad, min_dist, max_dist)


class Extractor:
    def __init__(self, H, nfeatures, scale, nlevels, ini, mn):
        self.H, self.nlevels, self.nfeatures = H, nlevels, nfeatures
        self.h = H.L.rh_extractor_create(nfeatures, scale, nlevels, ini, mn)

    def __del__(self):
        if getattr(self, "h", None):
            self.H.L.rh_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        out = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        self.H.L.rh_extractor_tables(self.h, *[_p(a) for a in out])
        return out

    def extract(self, img):
        img = _u8(img)
        h, w = img.shape
        cap = self.nfeatures * 2 + 64
        while True:
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = self.H.L.rh_extract(self.h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap)
            if n >= 0:
                return kps[:n].copy(), desc[:n].copy()
            cap = -n

    def level_dims(self, level):
        w, h = C.c_int32(), C.c_int32()
        if self.H.L.rh_level_dims(self.h, level, C.byref(w), C.byref(h)) != 0:
            return None
        return w.value, h.value

    def level(self, level):
        w, h = self.level_dims(level)
        out = np.zeros((h + 38, w + 38), np.uint8)
        self.H.L.rh_get_level(self.h, level, _p(out))
        return out


class Points:
    def __init__(self, H, n, pos, normal, desc, nobs, bad, min_dist, max_dist):
        self.H, self.n = H, n
        keep = [_f32(pos), _f32(normal), _u8(desc), _i32(nobs), _u8(bad), _f32(min_dist), _f32(max_dist)]
        self.h = H.L.rh_points_create(n, *[_p(a) for a in keep])

    def __del__(self):
        if getattr(self, "h", None):
            self.H.L.rh_points_destroy(self.h)
            self.h = None

    def state(self):
        rep, nobs, bad = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32), np.zeros(self.n, np.uint8)
        self.H.L.rh_points_state(self.h, _p(rep), _p(nobs), _p(bad))
        return rep, nobs, bad

    def track_state(self):
        iv = np.zeros(self.n, np.uint8)
        px, py, pxr, vc = (np.zeros(self.n, np.float32) for _ in range(4))
        lv = np.zeros(self.n, np.int32)
        self.H.L.rh_points_track_state(self.h, _p(iv), _p(px), _p(py), _p(pxr), _p(lv), _p(vc))
        return dict(in_view=iv, proj_x=px, proj_y=py, proj_xr=pxr, level=lv, view_cos=vc)


class Frame:
    """A reference `Frame` object living inside one harness library."""

    def __init__(self, H, handle):
        self.H, self.h = H, handle
        self.N = H.L.rh_frame_n(handle)

    def __del__(self):
        if getattr(self, "h", None):
            self.H.L.rh_frame_destroy(self.h)
            self.h = None

    @staticmethod
    def mono(H, ex, img, K=(500.0, 500.0, 320.0, 240.0), bf=40.0, th_depth=35.0):
        img = _u8(img)
        h, w = img.shape
        k = np.asarray(K, np.float32)
        return Frame(H, H.L.rh_frame_mono(ex.h, _p(img), w, h, img.strides[0], _p(k), bf, th_depth))

    @staticmethod
    def stereo(H, exl, exr, iml, imr, K=(718.856, 718.856, 607.1928, 185.2157), bf=386.1448, th_depth=35.0):
        iml, imr = _u8(iml), _u8(imr)
        h, w = iml.shape
        assert iml.strides[0] == imr.strides[0]
        k = np.asarray(K, np.float32)
        return Frame(H, H.L.rh_frame_stereo(exl.h, exr.h, _p(iml), _p(imr), w, h, iml.strides[0], _p(k), bf, th_depth))

    @staticmethod
    def from_arrays(H, ex, kps, desc, w, h, u_right=None, depth=None, K=(500.0, 500.0, 320.0, 240.0), bf=40.0, th_depth=35.0):
        kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        desc = _u8(desc)
        k = np.asarray(K, np.float32)
        ur, dp = _f32(u_right), _f32(depth)
        return Frame(H, H.L.rh_frame_from_arrays(ex.h, _p(kps), _p(desc), _p(ur), _p(dp), len(kps), w, h, _p(k), bf, th_depth))

    def copy(self):
        return Frame(self.H, self.H.L.rh_frame_copy(self.h))

    def get(self):
        n = self.N
        kps, kun = np.zeros(n, KP_DTYPE), np.zeros(n, KP_DTYPE)
        desc = np.zeros((n, 32), np.uint8)
        ur, dp = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.H.L.rh_frame_get(self.h, _p(kps), _p(kun), _p(desc), _p(ur), _p(dp))
        return dict(kps=kps, kps_un=kun, desc=desc, u_right=ur, depth=dp)

    def get_right(self):
        n = self.H.L.rh_frame_right_n(self.h)
        kps, desc = np.zeros(n, KP_DTYPE), np.zeros((n, 32), np.uint8)
        self.H.L.rh_frame_get_right(self.h, _p(kps), _p(desc))
        return kps, desc

    def set_pose(self, Tcw):
        t = _f32(Tcw).reshape(16)
        self.H.L.rh_frame_set_pose(self.h, _p(t))

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(max(self.N, 1), np.int32)
        n = self.H.L.rh_frame_features_in_area(self.h, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()

    def set_featvec(self, node, start, feat):
        node, start, feat = _i32(node), _i32(start), _i32(feat)
        self.H.L.rh_frame_set_featvec(self.h, _p(node), _p(start), _p(feat), len(node))

    def compute_bow(self, voc):
        """Frame::ComputeBoW with this vocabulary -> (words, values), (node, start, feat)"""
        self.H.L.rh_frame_compute_bow(self.h, voc.h)
        n = max(self.N, 1)
        w, v = np.zeros(n, np.int32), np.zeros(n, np.float64)
        nb = self.H.L.rh_frame_get_bow(self.h, _p(w), _p(v), n)
        node, start, feat = np.zeros(n, np.int32), np.zeros(n + 1, np.int32), np.zeros(n, np.int32)
        nf = self.H.L.rh_frame_get_featvec(self.h, _p(node), _p(start), _p(feat), n, n)
        assert nb <= n and nf >= 0
        return (w[:nb].copy(), v[:nb].copy()), (node[:nf].copy(), start[:nf + 1].copy(), feat[:start[nf]].copy())

    def set_points(self, pts, idx):
        idx = _i32(idx)
        assert len(idx) == self.N
        self.H.L.rh_frame_set_points(self.h, pts.h, _p(idx))

    def get_points(self, pts):
        idx = np.zeros(self.N, np.int32)
        self.H.L.rh_frame_get_points(self.h, pts.h, _p(idx))
        return idx

    def set_outliers(self, o):
        o = _u8(o)
        self.H.L.rh_frame_set_outliers(self.h, _p(o))


class Vocabulary:
    """ORBVocabulary::loadFromTextFile of one harness library (the reference's DBoW2 template, or this repo's GPU vocabulary)."""

    def __init__(self, H, path):
        self.H = H
        self.h = H.L.rh_voc_load_text(path.encode())
        if not self.h:
            raise RuntimeError("loadFromTextFile failed: %s" % path)

    def __del__(self):
        if getattr(self, "h", None):
            self.H.L.rh_voc_destroy(self.h)
            self.h = None

    def size(self):
        return self.H.L.rh_voc_size(self.h)

    def score(self, f1, f2):
        return self.H.L.rh_voc_score(self.h, f1.h, f2.h)


class KeyFrame:
    def __init__(self, frame):
        self.H, self.N = frame.H, frame.N
        self.h = self.H.L.rh_keyframe_create(frame.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.H.L.rh_keyframe_destroy(self.h)
            self.h = None

    def set_pose(self, Tcw):
        t = _f32(Tcw).reshape(16)
        self.H.L.rh_keyframe_set_pose(self.h, _p(t))

    def set_bad(self, bad):
        self.H.L.rh_keyframe_set_bad(self.h, int(bad))

    def set_points(self, pts, idx, observe=True):
        idx = _i32(idx)
        assert len(idx) == self.N
        self.H.L.rh_keyframe_set_points(self.h, pts.h, _p(idx), int(observe))

    def get_points(self, pts):
        idx = np.zeros(self.N, np.int32)
        self.H.L.rh_keyframe_get_points(self.h, pts.h, _p(idx))
        return idx


class Matcher:
    """ORBmatcher(nnratio, checkOri) — every method constructs the reference class and calls the reference signature."""

    def __init__(self, H, nnratio=0.6, check_ori=True):
        self.H, self.r, self.o = H, float(nnratio), int(check_ori)

    def search_local_points(self, F, pts, idx, th=3.0, run_frustum=True):
        idx = _i32(idx)
        return self.H.L.rh_search_local_points(self.r, self.o, F.h, pts.h, _p(idx), len(idx), th, int(run_frustum))

    def search_last_frame(self, cur, last, th, mono):
        return self.H.L.rh_search_last_frame(self.r, self.o, cur.h, last.h, th, int(mono))

    def search_reloc(self, cur, kf, pts, found_idx, th, orb_dist):
        found_idx = _i32(found_idx)
        return self.H.L.rh_search_reloc(self.r, self.o, cur.h, kf.h, pts.h, _p(found_idx), len(found_idx), th, orb_dist)

    def search_loop(self, kf, Scw, pts, idx, matched, th):
        idx, s = _i32(idx), _f32(Scw).reshape(16)
        matched = np.ascontiguousarray(matched, dtype=np.int32).copy()
        n = self.H.L.rh_search_loop(self.r, self.o, kf.h, _p(s), pts.h, _p(idx), len(idx), _p(matched), th)
        return n, matched

    def search_bow_kf_frame(self, kf, F, pts):
        m = np.zeros(F.N, np.int32)
        n = self.H.L.rh_search_bow_kf_frame(self.r, self.o, kf.h, F.h, pts.h, _p(m))
        return n, m

    def search_bow_kf_kf(self, kf1, kf2, pts):
        m = np.zeros(kf1.N, np.int32)
        n = self.H.L.rh_search_bow_kf_kf(self.r, self.o, kf1.h, kf2.h, pts.h, _p(m))
        return n, m

    def search_initialization(self, F1, F2, prev_xy, window=10):
        prev = np.ascontiguousarray(prev_xy, dtype=np.float32).copy()
        m = np.zeros(F1.N, np.int32)
        n = self.H.L.rh_search_initialization(self.r, self.o, F1.h, F2.h, _p(prev), _p(m), window)
        return n, m, prev

    def search_triangulation(self, kf1, kf2, F12, only_stereo=False):
        f = _f32(F12).reshape(9)
        cap = max(kf1.N, 1)
        pairs = np.zeros((cap, 2), np.int32)
        n = self.H.L.rh_search_triangulation(self.r, self.o, kf1.h, kf2.h, _p(f), _p(pairs), cap, int(only_stereo))
        return n, pairs[:n].copy()

    def search_sim3(self, kf1, kf2, pts, matches12, s12, R12, t12, th):
        m = np.ascontiguousarray(matches12, dtype=np.int32).copy()
        r, t = _f32(R12).reshape(9), _f32(t12).reshape(3)
        n = self.H.L.rh_search_sim3(self.r, self.o, kf1.h, kf2.h, pts.h, _p(m), s12, _p(r), _p(t), th)
        return n, m

    def fuse(self, kf, pts, idx, th=3.0):
        idx = _i32(idx)
        return self.H.L.rh_fuse(self.r, self.o, kf.h, pts.h, _p(idx), len(idx), th)

    def fuse_sim3(self, kf, Scw, pts, idx, th):
        idx, s = _i32(idx), _f32(Scw).reshape(16)
        rep = np.full(len(idx), -1, np.int32)
        n = self.H.L.rh_fuse_sim3(self.r, self.o, kf.h, _p(s), pts.h, _p(idx), len(idx), th, _p(rep))
        return n, rep


_ref = None


def ref():
    """The reference arm (oracle/_ref/liborb_ref.so); builds it when /root/reference is present."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO) or os.path.exists("/root/reference/orb_slam2/src/ORBextractor.cc"):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _ref = Harness(REF_SO)
    return _ref


def available():
    return os.path.exists(REF_SO) or os.path.exists("/root/reference/orb_slam2/src/ORBextractor.cc")
