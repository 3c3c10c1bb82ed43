"""ctypes front-end of the CPU oracle (oracle/_build/liborb_oracle.so).

TEST INFRASTRUCTURE ONLY — see oracle/orb_oracle.h.  Importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never from orb_slam_2_ros_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborb_oracle.so")

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
TOP2_DTYPE = np.dtype([("best_dist", "<i4"), ("best_idx", "<i4"), ("second_dist", "<i4"), ("second_idx", "<i4")])

MODE_TRACK_LAST = 0
MODE_LOCAL_POINTS = 1
MODE_INITIALIZATION = 2


class SearchParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("th_dist", C.c_int32), ("nn_ratio", C.c_float),
                ("check_orientation", C.c_int32)]


class StereoParams(C.Structure):
    _fields_ = [("bf", C.c_float), ("b", C.c_float)]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only, a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in
            ("orb_oracle_extract.cpp", "orb_oracle_match.cpp", "orb_oracle_bow.cpp", "orb_oracle.h", "orb_pattern_31.inc", "Makefile")]
    if (not force) and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.orc_extractor_create.restype = vp
        L.orc_extractor_create.argtypes = [i32, f32, i32, i32, i32]
        L.orc_extractor_destroy.argtypes = [vp]
        L.orc_extractor_tables.argtypes = [vp] + [vp] * 6
        L.orc_extract.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32]
        L.orc_level_dims.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
        L.orc_get_level.argtypes = [vp, i32, vp]
        L.orc_get_blurred.argtypes = [vp, i32, vp]
        L.orc_raw_corner_count.argtypes = [vp, i32]
        L.orc_get_raw_corners.argtypes = [vp, i32, vp, i32]
        L.orc_level_kp_count.argtypes = [vp, i32]
        L.orc_get_level_kps.argtypes = [vp, i32, vp, i32]
        L.orc_get_stats.argtypes = [vp, vp]
        L.orc_count_near_half_taps.argtypes = [vp, f32]
        L.orc_resize_linear_u8.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32]
        L.orc_border_reflect101.argtypes = [vp, i32, i32, i32, vp, i32, i32]
        L.orc_fast9_16.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32]
        L.orc_gaussian7x7_s2.argtypes = [vp, i32, i32, i32, vp, i32]
        L.orc_cvt_gray.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32]
        L.orc_fast_atan2.restype = f32
        L.orc_fast_atan2.argtypes = [f32, f32]
        L.orc_cv_round_f.argtypes = [f32]
        L.orc_ic_angle.restype = f32
        L.orc_ic_angle.argtypes = [vp, i32]
        L.orc_brief_descriptor.argtypes = [vp, i32, f32, vp]
        L.orc_sincos_range.argtypes = [C.c_uint32, C.c_longlong, vp, vp]
        L.orc_descriptor_distance.argtypes = [vp, vp]
        L.orc_hamming_top2.argtypes = [vp, i32, vp, i32, vp]
        L.orc_hamming_top2_csr.argtypes = [vp, i32, vp, vp, vp, vp]
        L.orc_grid_create.restype = vp
        L.orc_grid_create.argtypes = [vp, i32, f32, f32, f32, f32]
        L.orc_grid_destroy.argtypes = [vp]
        L.orc_grid_query.argtypes = [vp, f32, f32, f32, i32, i32, vp, i32]
        L.orc_search_by_projection_ex.argtypes = [C.POINTER(SearchParams), vp, vp, vp, vp, i32, vp, i32] + [vp] * 13
        L.orc_match_bruteforce.argtypes = [vp, vp, i32, vp, vp, i32, i32, f32, i32, vp]
        L.orc_stereo_match.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, C.POINTER(StereoParams), vp, vp, vp]
        f64 = C.c_double
        L.orc_search_for_triangulation.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, f32, f32, vp, vp, i32, i32, vp]
        L.orc_sim3_search_one_way.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.orc_sim3_search_one_way.restype = None
        L.orc_distinctive_descriptors.argtypes = [vp, vp, i32, vp]
        L.orc_fuse_search.argtypes = [vp, vp, vp, vp, vp, i32] + [vp] * 9
        L.orc_fuse_search.restype = None
        L.orc_distinctive_descriptors.restype = None
        L.orc_search_by_bow.argtypes = [vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, f32, i32, vp, vp]
        L.orc_voc_create.restype = vp
        L.orc_voc_create.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.orc_voc_load_text.restype = vp
        L.orc_voc_load_text.argtypes = [C.c_char_p]
        L.orc_voc_destroy.argtypes = [vp]
        L.orc_voc_info.argtypes = [vp] + [C.POINTER(i32)] * 6
        L.orc_voc_export.argtypes = [vp] * 5
        L.orc_bow_transform_features.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.orc_bow_transform.argtypes = [vp, vp, i32, i32, C.POINTER(i32), vp, vp, C.POINTER(i32), vp, vp, vp]
        L.orc_bow_score_l1.restype = f64
        L.orc_bow_score_l1.argtypes = [vp, vp, i32, vp, vp, i32]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class Extractor:
    """Oracle ORBextractor (reference ORBextractor.cc:416-479, 1083-1149)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = lib().orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        sc = np.zeros(nlevels, np.float32); isc = sc.copy(); s2 = sc.copy(); is2 = sc.copy()
        per = np.zeros(nlevels, np.int32); um = np.zeros(16, np.int32)
        lib().orc_extractor_tables(self._h, _p(sc), _p(isc), _p(s2), _p(is2), _p(per), _p(um))
        self.scale_factors, self.inv_scale_factors = sc, isc
        self.level_sigma2, self.inv_level_sigma2 = s2, is2
        self.features_per_level, self.umax = per, um

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_extractor_destroy(self._h)
            self._h = None

    def extract(self, img):
        img = _u8(img)
        h, w = img.shape
        cap = self.nfeatures + 8 * self.nlevels + 64
        while True:
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = lib().orc_extract(self._h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap)
            if n <= -1000000:
                raise ValueError("image too small for the reference cell grid")
            if n < 0:
                cap = -n
                continue
            return kps[:n].copy(), desc[:n].copy()

    def level_dims(self, l):
        w, h = C.c_int32(), C.c_int32()
        lib().orc_level_dims(self._h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, l):
        """Bordered pyramid buffer of level l: (h+38, w+38)."""
        w, h = self.level_dims(l)
        out = np.zeros((h + 38, w + 38), np.uint8)
        lib().orc_get_level(self._h, l, _p(out))
        return out

    def blurred(self, l):
        w, h = self.level_dims(l)
        out = np.zeros((h, w), np.uint8)
        if lib().orc_get_blurred(self._h, l, _p(out)) != 0:
            return None
        return out

    def raw_corners(self, l):
        n = lib().orc_raw_corner_count(self._h, l)
        out = np.zeros(n, KP_DTYPE)
        lib().orc_get_raw_corners(self._h, l, _p(out), n)
        return out

    def level_kps(self, l):
        n = lib().orc_level_kp_count(self._h, l)
        out = np.zeros(n, KP_DTYPE)
        lib().orc_get_level_kps(self._h, l, _p(out), n)
        return out

    def stats(self):
        s = np.zeros((self.nlevels, 4), np.int32)
        lib().orc_get_stats(self._h, _p(s))
        return s

    def near_half_taps(self):
        return lib().orc_count_near_half_taps(self._h, 1e-4)


# ---- primitives -------------------------------------------------------------------------------------
def resize_linear(src, dw, dh):
    src = _u8(src)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def border_reflect101(src, b=19):
    src = _u8(src)
    h, w = src.shape
    dst = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    lib().orc_border_reflect101(_p(src), w, h, src.strides[0], _p(dst), b, w + 2 * b)
    return dst


def fast(img, threshold, nms=True):
    img = _u8(img)
    h, w = img.shape
    cap = max(16, w * h)
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_fast9_16(_p(img), w, h, img.strides[0], threshold, int(nms), _p(out), cap)
    return out[:n].copy()


def gaussian_blur(src):
    src = _u8(src)
    h, w = src.shape
    dst = np.zeros((h, w), np.uint8)
    lib().orc_gaussian7x7_s2(_p(src), w, h, src.strides[0], _p(dst), w)
    return dst


def cvt_gray(img, rgb=False):
    """cvtColor(img, BGR2GRAY / RGB2GRAY / BGRA2GRAY / RGBA2GRAY) with OpenCV 4.13.0 arithmetic; img: uint8 (h, w, 3|4)."""
    img = _u8(img)
    h, w, ch = img.shape
    dst = np.zeros((h, w), np.uint8)
    lib().orc_cvt_gray(_p(img), w, h, img.strides[0], ch, int(rgb), _p(dst), w)
    return dst


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def ic_angle(img, x, y):
    img = _u8(img)
    ptr = img.ctypes.data + y * img.strides[0] + x
    return lib().orc_ic_angle(C.c_void_p(ptr), img.strides[0])


def brief_descriptor(img, x, y, angle_deg):
    img = _u8(img)
    ptr = img.ctypes.data + y * img.strides[0] + x
    d = np.zeros(32, np.uint8)
    lib().orc_brief_descriptor(C.c_void_p(ptr), img.strides[0], float(angle_deg), _p(d))
    return d


def sincos_range(first_bits, n):
    a = np.zeros(n, np.float32); b = np.zeros(n, np.float32)
    lib().orc_sincos_range(first_bits, n, _p(a), _p(b))
    return a, b


# ---- matcher -----------------------------------------------------------------------------------------
def descriptor_distance(a, b):
    return lib().orc_descriptor_distance(_p(_u8(a)), _p(_u8(b)))


def hamming_top2(q, db):
    q, db = _u8(q), _u8(db)
    out = np.zeros(len(q), TOP2_DTYPE)
    lib().orc_hamming_top2(_p(q), len(q), _p(db), len(db), _p(out))
    return out


def hamming_top2_csr(q, db, off, idx):
    q, db = _u8(q), _u8(db)
    off = np.ascontiguousarray(off, np.int32); idx = np.ascontiguousarray(idx, np.int32)
    out = np.zeros(len(q), TOP2_DTYPE)
    lib().orc_hamming_top2_csr(_p(q), len(q), _p(db), _p(off), _p(idx), _p(out))
    return out


class Grid:
    """Frame::mGrid (reference Frame.cc:239-256) + GetFeaturesInArea (Frame.cc:354-412)."""

    def __init__(self, kps_un, min_x, min_y, max_x, max_y):
        self.kps = np.ascontiguousarray(kps_un, KP_DTYPE)
        self._h = lib().orc_grid_create(_p(self.kps), len(self.kps), min_x, min_y, max_x, max_y)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_grid_destroy(self._h)
            self._h = None

    def query(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(max(1, len(self.kps)), np.int32)
        n = lib().orc_grid_query(self._h, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()


def search_by_projection(mode, grid, desc, u_right, taken, q_u, q_v, q_radius, q_min_level, q_max_level, q_desc,
                         q_ur=None, q_er_max=None, q_angle=None, q_valid=None, q_obs=None, th_dist=100,
                         nn_ratio=0.9, check_orientation=True):
    """Returns (nmatches, match_of_query, target_query); `taken` (uint8[n]) is updated in place."""
    n, nq = len(grid.kps), len(q_u)
    f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    i = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
    b = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    desc, q_desc = _u8(desc), _u8(q_desc)
    u_right, q_u, q_v, q_radius, q_ur, q_er_max, q_angle = map(f, (u_right, q_u, q_v, q_radius, q_ur, q_er_max, q_angle))
    q_min_level, q_max_level = i(q_min_level), i(q_max_level)
    q_valid, q_obs = b(q_valid), b(q_obs)
    if q_angle is None:
        q_angle = np.zeros(nq, np.float32)
    if u_right is not None and (q_ur is None or q_er_max is None):
        raise ValueError("q_ur and q_er_max are required with u_right")
    assert taken.dtype == np.uint8 and taken.flags.c_contiguous
    prm = SearchParams(mode, th_dist, nn_ratio, int(check_orientation))
    moq = np.zeros(nq, np.int32)
    tq = np.zeros(n, np.int32)
    nm = lib().orc_search_by_projection_ex(C.byref(prm), grid._h, _p(grid.kps), _p(desc), _p(u_right), n, _p(taken),
                                           nq, _p(q_u), _p(q_v), _p(q_radius), _p(q_min_level), _p(q_max_level),
                                           _p(q_desc), _p(q_ur), _p(q_er_max), _p(q_angle), _p(q_valid), _p(q_obs),
                                           _p(moq), _p(tq))
    return nm, moq, tq


def match_bruteforce(desc1, angle1, desc2, angle2, th_dist=50, nn_ratio=0.6, check_orientation=True):
    desc1, desc2 = _u8(desc1), _u8(desc2)
    angle1 = np.ascontiguousarray(angle1, np.float32); angle2 = np.ascontiguousarray(angle2, np.float32)
    m = np.zeros(len(desc1), np.int32)
    nm = lib().orc_match_bruteforce(_p(desc1), _p(angle1), len(desc1), _p(desc2), _p(angle2), len(desc2), th_dist,
                                    nn_ratio, int(check_orientation), _p(m))
    return nm, m


def stereo_match(ex_left, ex_right, kps_l, desc_l, kps_r, desc_r, bf, b):
    kps_l = np.ascontiguousarray(kps_l, KP_DTYPE); kps_r = np.ascontiguousarray(kps_r, KP_DTYPE)
    desc_l, desc_r = _u8(desc_l), _u8(desc_r)
    n = len(kps_l)
    ur = np.zeros(n, np.float32); depth = np.zeros(n, np.float32); sad = np.zeros(n, np.int32)
    prm = StereoParams(bf, b)
    kept = lib().orc_stereo_match(ex_left._h, ex_right._h, _p(kps_l), _p(desc_l), n, _p(kps_r), _p(desc_r), len(kps_r),
                                  C.byref(prm), _p(ur), _p(depth), _p(sad))
    return kept, ur, depth, sad


class Vocabulary:
    """ORBVocabulary (DBoW2 TemplatedVocabulary<FORB>) restated on the CPU."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("vocabulary could not be created / loaded")
        self._h = handle
        v = [C.c_int32() for _ in range(6)]
        lib().orc_voc_info(self._h, *[C.byref(x) for x in v])
        self.k, self.L, self.n_nodes, self.n_words, self.scoring, self.weighting = [int(x.value) for x in v]

    @classmethod
    def from_arrays(cls, k, L, scoring, weighting, parent, is_leaf, desc, weight):
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = _u8(is_leaf); desc = _u8(desc)
        weight = np.ascontiguousarray(weight, np.float64)
        return cls(lib().orc_voc_create(k, L, scoring, weighting, len(parent), _p(parent), _p(is_leaf), _p(desc), _p(weight)))

    @classmethod
    def load_text(cls, path):
        return cls(lib().orc_voc_load_text(os.fsencode(path)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_voc_destroy(self._h)
            self._h = None

    def export(self):
        parent = np.zeros(self.n_nodes, np.int32); leaf = np.zeros(self.n_nodes, np.uint8)
        desc = np.zeros((self.n_nodes, 32), np.uint8); weight = np.zeros(self.n_nodes, np.float64)
        lib().orc_voc_export(self._h, _p(parent), _p(leaf), _p(desc), _p(weight))
        return parent, leaf, desc, weight

    def transform_features(self, desc, levelsup=4):
        desc = _u8(desc); n = len(desc)
        word = np.zeros(n, np.int32); weight = np.zeros(n, np.float64); node = np.zeros(n, np.int32)
        lib().orc_bow_transform_features(self._h, _p(desc), n, levelsup, _p(word), _p(weight), _p(node))
        return word, weight, node

    def transform(self, desc, levelsup=4):
        """-> (bow_word, bow_value), (fv_node, fv_start, fv_feat)"""
        desc = _u8(desc); n = len(desc)
        bw = np.zeros(n, np.int32); bv = np.zeros(n, np.float64)
        fn = np.zeros(n, np.int32); fs = np.zeros(n + 1, np.int32); ff = np.zeros(n, np.int32)
        nb, nf = C.c_int32(), C.c_int32()
        lib().orc_bow_transform(self._h, _p(desc), n, levelsup, C.byref(nb), _p(bw), _p(bv), C.byref(nf), _p(fn), _p(fs), _p(ff))
        nb, nf = nb.value, nf.value
        return (bw[:nb].copy(), bv[:nb].copy()), (fn[:nf].copy(), fs[:nf + 1].copy(), ff[:fs[nf]].copy())


def bow_score_l1(a, b):
    (w1, v1), (w2, v2) = a, b
    w1 = np.ascontiguousarray(w1, np.int32); w2 = np.ascontiguousarray(w2, np.int32)
    v1 = np.ascontiguousarray(v1, np.float64); v2 = np.ascontiguousarray(v2, np.float64)
    return float(lib().orc_bow_score_l1(_p(w1), _p(v1), len(w1), _p(w2), _p(v2), len(w2)))


def search_by_bow(desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2, th_dist=50, strict=False, nn_ratio=0.6,
                  check_orientation=True):
    """fvX = (node, start, feat) CSR feature vectors.  -> (match12, match21, nmatches)"""
    desc1 = _u8(desc1); desc2 = _u8(desc2)
    n1, n2 = len(desc1), len(desc2)
    angle1 = np.ascontiguousarray(angle1, np.float32); angle2 = np.ascontiguousarray(angle2, np.float32)
    valid1 = None if valid1 is None else _u8(valid1)
    valid2 = None if valid2 is None else _u8(valid2)
    f1 = [np.ascontiguousarray(x, np.int32) for x in fv1]
    f2 = [np.ascontiguousarray(x, np.int32) for x in fv2]
    m12 = np.zeros(n1, np.int32); m21 = np.zeros(n2, np.int32)
    nm = lib().orc_search_by_bow(_p(desc1), _p(angle1), _p(valid1), n1, _p(f1[0]), _p(f1[1]), _p(f1[2]), len(f1[0]),
                                 _p(desc2), _p(angle2), _p(valid2), n2, _p(f2[0]), _p(f2[1]), _p(f2[2]), len(f2[0]),
                                 th_dist, 1 if strict else 0, nn_ratio, 1 if check_orientation else 0, _p(m12), _p(m21))
    return m12, m21, int(nm)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def search_for_triangulation(kps1, desc1, has_mp1, u_right1, fv1, kps2, desc2, has_mp2, u_right2, fv2, F12, ex, ey,
                             scale_factors, level_sigma2, only_stereo=False, check_orientation=True):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:659-825) -> (match12, nmatches)"""
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    desc1 = _u8(desc1); desc2 = _u8(desc2)
    has_mp1 = None if has_mp1 is None else _u8(has_mp1); has_mp2 = None if has_mp2 is None else _u8(has_mp2)
    u_right1, u_right2 = _f32(u_right1), _f32(u_right2)
    f1 = [np.ascontiguousarray(x, np.int32) for x in fv1]; f2 = [np.ascontiguousarray(x, np.int32) for x in fv2]
    F12 = _f32(F12).reshape(9); sf = _f32(scale_factors); ls = _f32(level_sigma2)
    m12 = np.zeros(len(kps1), np.int32)
    nm = lib().orc_search_for_triangulation(_p(kps1), _p(desc1), _p(has_mp1), _p(u_right1), len(kps1), _p(f1[0]), _p(f1[1]), _p(f1[2]),
                                            len(f1[0]), _p(kps2), _p(desc2), _p(has_mp2), _p(u_right2), len(kps2), _p(f2[0]), _p(f2[1]),
                                            _p(f2[2]), len(f2[0]), _p(F12), ex, ey, _p(sf), _p(ls), int(only_stereo),
                                            int(check_orientation), _p(m12))
    return m12, int(nm)


def sim3_search_one_way(grid, desc, q_u, q_v, q_radius, q_level, q_desc, q_valid=None):
    nq = len(q_u)
    desc = _u8(desc); q_desc = _u8(q_desc)
    q_u, q_v, q_radius = _f32(q_u), _f32(q_v), _f32(q_radius)
    q_level = np.ascontiguousarray(q_level, np.int32)
    q_valid = None if q_valid is None else _u8(q_valid)
    m = np.zeros(nq, np.int32)
    lib().orc_sim3_search_one_way(grid._h, _p(grid.kps), _p(desc), nq, _p(q_u), _p(q_v), _p(q_radius), _p(q_level), _p(q_desc), _p(q_valid), _p(m))
    return m


def search_by_sim3(grid1, desc1, grid2, desc2, q12, q21):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1104-1328) on arrays.  q12 = dict(u, v, radius, level, desc, valid): the
    map points of keyframe 1 projected into keyframe 2 (one entry per keypoint of keyframe 1), q21 the reverse.
    -> (match12 after the agreement check, nFound)"""
    m1 = sim3_search_one_way(grid2, desc2, q12["u"], q12["v"], q12["radius"], q12["level"], q12["desc"], q12.get("valid"))
    m2 = sim3_search_one_way(grid1, desc1, q21["u"], q21["v"], q21["radius"], q21["level"], q21["desc"], q21.get("valid"))
    out = np.full(len(m1), -1, np.int32)
    for i1, idx2 in enumerate(m1):          # OM:1282-1299
        if idx2 >= 0 and m2[idx2] == i1:
            out[i1] = idx2
    return out, int((out >= 0).sum())


def distinctive_descriptors(desc, off):
    desc = _u8(desc); off = np.ascontiguousarray(off, np.int32)
    best = np.zeros(len(off) - 1, np.int32)
    lib().orc_distinctive_descriptors(_p(desc), _p(off), len(off) - 1, _p(best))
    return best


def fuse_search(grid, desc, u_right, inv_level_sigma2, q_u, q_v, q_ur, q_radius, q_level, q_desc, q_valid=None):
    """search half of ORBmatcher::Fuse -> (best_idx, best_dist)"""
    nq = len(q_u)
    desc = _u8(desc); q_desc = _u8(q_desc)
    u_right = _f32(u_right); inv = _f32(inv_level_sigma2)
    q_u, q_v, q_ur, q_radius = _f32(q_u), _f32(q_v), _f32(q_ur), _f32(q_radius)
    q_level = np.ascontiguousarray(q_level, np.int32)
    q_valid = None if q_valid is None else _u8(q_valid)
    bi = np.zeros(nq, np.int32); bd = np.zeros(nq, np.int32)
    lib().orc_fuse_search(grid._h, _p(grid.kps), _p(desc), _p(u_right), _p(inv), nq, _p(q_u), _p(q_v), _p(q_ur), _p(q_radius), _p(q_level),
                          _p(q_desc), _p(q_valid), _p(bi), _p(bd))
    return bi, bd


# ---- map-file record payloads -----------------------------------------------------------------------
# Plain `struct` restatement of the reference's serialize() bodies for cv::Mat and cv::KeyPoint
# (orb_slam2/include/BoostArchiver.h:46-91) inside a boost binary archive opened with no_header (System.cc:627, primitives
# raw, native endian).  Parity unpinned: no boost in this image, so no file written by the reference binary to compare with.
def mat_record_encode(mat, elem_type):
    import struct
    mat = np.ascontiguousarray(mat)                       # BoostArchiver.h:64-66: clone when not continuous
    rows, cols = mat.shape
    es = mat.dtype.itemsize
    return struct.pack("<iiQQ", cols, rows, es, elem_type) + mat.tobytes()   # :68-75 cols, rows, elem_size, elem_type, data


def mat_record_decode(buf, offset=0):
    import struct
    cols, rows, es, et = struct.unpack_from("<iiQQ", buf, offset)            # :82-85
    n = cols * rows * es                                                     # :88
    data = np.frombuffer(buf, np.uint8, n, offset + 24)
    return rows, cols, es, et, data, 24 + n


def keypoint_records_encode(kps):
    import struct
    out = bytearray()
    for k in kps:   # :49-57: angle, class_id, octave, response, response, pt.x, pt.y (no size)
        out += struct.pack("<fiiffff", float(k["angle"]), int(k["class_id"]), int(k["octave"]), float(k["response"]),
                           float(k["response"]), float(k["x"]), float(k["y"]))
    return bytes(out)


def keypoint_records_decode(buf, n, offset=0):
    import struct
    kps = np.zeros(n, KP_DTYPE)                            # cv::KeyPoint(): size = 0, never restored
    for i in range(n):
        a, cid, octv, r1, r2, x, y = struct.unpack_from("<fiiffff", buf, offset + 28 * i)
        kps[i]["angle"], kps[i]["class_id"], kps[i]["octave"] = a, cid, octv
        kps[i]["response"] = r1
        kps[i]["response"] = r2                            # the second `ar & kf.response` overwrites the first
        kps[i]["x"], kps[i]["y"] = x, y
    return kps
