"""ORBmatcher — host-side mirror of ORB_SLAM2::ORBmatcher (reference orb_slam2/include/ORBmatcher.h:37-103)
over the C ABI.  The reference methods take Frame / KeyFrame / MapPoint objects; here the same routines take
the arrays those objects hold (what each loop reads), see include/orb_b200.h."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, TOP2_DTYPE, SearchParams, check, lib, ptr

TH_HIGH = 100      # ORBmatcher.cc:37
TH_LOW = 50        # ORBmatcher.cc:38
HISTO_LENGTH = 30  # ORBmatcher.cc:39
MODE_TRACK_LAST, MODE_LOCAL_POINTS, MODE_INITIALIZATION = 0, 1, 2


def _f(a): return None if a is None else np.ascontiguousarray(a, np.float32)
def _i(a): return None if a is None else np.ascontiguousarray(a, np.int32)
def _b(a): return None if a is None else np.ascontiguousarray(a, np.uint8)


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = TH_HIGH, TH_LOW, HISTO_LENGTH

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio, self.mbCheckOrientation, self.device = nnratio, checkOri, device

    @staticmethod
    def DescriptorDistance(a, b, device=0):
        """ORBmatcher.cc:1649-1665.  One 256-bit popcount stays on the host (a device round trip per pair would stall
        MapPoint::ComputeDistinctiveDescriptors' O(N^2) loop, MapPoint.cc:332); the GPU serves the batched entry points."""
        x = np.bitwise_xor(np.asarray(a, np.uint8).reshape(32), np.asarray(b, np.uint8).reshape(32))
        return int(np.unpackbits(x).sum())

    def SearchByProjection(self, mode, kps_un, desc, bounds, taken, q_u, q_v, q_radius, q_min_level, q_max_level,
                           q_desc, u_right=None, q_ur=None, q_er_max=None, q_angle=None, q_valid=None, q_obs=None,
                           th_dist=TH_HIGH):
        """ORBmatcher.cc:45-129 (mode LOCAL_POINTS) / :1330-1472 (mode TRACK_LAST).
        Returns (nmatches, match_of_query[nq], target_query[n]); `taken` (uint8[n]) is updated in place."""
        kps_un = np.ascontiguousarray(kps_un, KP_DTYPE)
        n, nq = len(kps_un), len(q_u)
        desc, q_desc = _b(desc), _b(q_desc)
        u_right, q_u, q_v, q_radius, q_ur, q_er_max, q_angle = map(_f, (u_right, q_u, q_v, q_radius, q_ur, q_er_max, q_angle))
        q_min_level, q_max_level, q_valid, q_obs = _i(q_min_level), _i(q_max_level), _b(q_valid), _b(q_obs)
        assert taken.dtype == np.uint8 and taken.flags.c_contiguous and len(taken) == n
        prm = SearchParams(mode, th_dist, self.mfNNratio, int(self.mbCheckOrientation), *map(float, bounds))
        moq = np.full(nq, -1, np.int32); tq = np.full(n, -1, np.int32); nm = C.c_int32(0)
        check(lib().orb_search_by_projection(self.device, C.byref(prm), ptr(kps_un), ptr(desc), ptr(u_right), n, ptr(taken), nq,
                                             ptr(q_u), ptr(q_v), ptr(q_radius), ptr(q_min_level), ptr(q_max_level), ptr(q_desc),
                                             ptr(q_ur), ptr(q_er_max), ptr(q_angle), ptr(q_valid), ptr(q_obs), ptr(moq), ptr(tq),
                                             C.byref(nm)))
        return nm.value, moq, tq

    def SearchForInitialization(self, kps1_un, desc1, kps2_un, desc2, bounds, vbPrevMatched, windowSize=100):
        """ORBmatcher::SearchForInitialization (ORBmatcher.cc:406-521): F1 keypoints of octave 0 search F2 in a window
        around vbPrevMatched.  Returns (nmatches, vnMatches12[n1]) and updates vbPrevMatched (float32 [n1, 2]) in place."""
        kps1_un = np.ascontiguousarray(kps1_un, KP_DTYPE); kps2_un = np.ascontiguousarray(kps2_un, KP_DTYPE)
        n1 = len(kps1_un)
        pm = np.ascontiguousarray(vbPrevMatched, np.float32)
        zeros = np.zeros(n1, np.int32)
        nm, m12, _m21 = self.SearchByProjection(MODE_INITIALIZATION, kps2_un, desc2, bounds, np.zeros(len(kps2_un), np.uint8),
                                                pm[:, 0].copy(), pm[:, 1].copy(), np.full(n1, windowSize, np.float32), zeros, zeros,
                                                desc1, q_angle=kps1_un["angle"], q_valid=(kps1_un["octave"] <= 0).astype(np.uint8),
                                                th_dist=TH_LOW)
        ok = m12 >= 0                                            # ORBmatcher.cc:515-518
        vbPrevMatched[ok, 0] = kps2_un["x"][m12[ok]]; vbPrevMatched[ok, 1] = kps2_un["y"][m12[ok]]
        return nm, m12

    def MatchBruteForce(self, desc1, angle1, desc2, angle2, th_dist=TH_LOW):
        """SearchByBoW inner loop (ORBmatcher.cc:196-252) over one node holding both frames' keypoints."""
        desc1, desc2, angle1, angle2 = _b(desc1), _b(desc2), _f(angle1), _f(angle2)
        m = np.full(len(desc1), -1, np.int32); nm = C.c_int32(0)
        check(lib().orb_match_bruteforce(self.device, ptr(desc1), ptr(angle1), len(desc1), ptr(desc2), ptr(angle2), len(desc2),
                                         th_dist, self.mfNNratio, int(self.mbCheckOrientation), ptr(m), C.byref(nm)))
        return nm.value, m


    def SearchByBoW(self, desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2, keyframe_pair=False, th_dist=TH_LOW):
        """ORBmatcher::SearchByBoW: KeyFrame -> Frame (ORBmatcher.cc:160-289; valid2 = None) or, keyframe_pair=True,
        KeyFrame -> KeyFrame (ORBmatcher.cc:524-657: strict '<' on TH_LOW, targets need a good map point too).
        fvX = (node, start, feat), the FeatureVector ORBVocabulary.transform returns.  -> nmatches, match12, match21"""
        desc1, desc2, angle1, angle2 = _b(desc1), _b(desc2), _f(angle1), _f(angle2)
        valid1, valid2 = _b(valid1), _b(valid2)
        f1 = [_i(x) for x in fv1]; f2 = [_i(x) for x in fv2]
        n1, n2 = len(desc1), len(desc2)
        m12 = np.full(n1, -1, np.int32); m21 = np.full(n2, -1, np.int32); nm = C.c_int32(0)
        check(lib().orb_search_by_bow(self.device, ptr(desc1), ptr(angle1), ptr(valid1), n1, ptr(f1[0]), ptr(f1[1]), ptr(f1[2]), len(f1[0]),
                                      ptr(desc2), ptr(angle2), ptr(valid2), n2, ptr(f2[0]), ptr(f2[1]), ptr(f2[2]), len(f2[0]),
                                      th_dist, 1 if keyframe_pair else 0, self.mfNNratio, int(self.mbCheckOrientation), ptr(m12), ptr(m21),
                                      C.byref(nm)))
        return nm.value, m12, m21

    def SearchForTriangulation(self, kps1, desc1, has_mp1, u_right1, fv1, kps2, desc2, has_mp2, u_right2, fv2, F12, ex, ey,
                               scale_factors, level_sigma2, bOnlyStereo=False):
        """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:659-825).  -> nmatches, vMatches12 (vMatchedPairs = its
        non-negative entries in index order, :812-822)"""
        kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
        desc1, desc2 = _b(desc1), _b(desc2)
        has_mp1, has_mp2, u_right1, u_right2 = _b(has_mp1), _b(has_mp2), _f(u_right1), _f(u_right2)
        f1 = [_i(x) for x in fv1]; f2 = [_i(x) for x in fv2]
        F12 = _f(F12).reshape(9); sf = _f(scale_factors); ls = _f(level_sigma2)
        m12 = np.full(len(kps1), -1, np.int32); nm = C.c_int32(0)
        check(lib().orb_search_for_triangulation(self.device, ptr(kps1), ptr(desc1), ptr(has_mp1), ptr(u_right1), len(kps1), ptr(f1[0]), ptr(f1[1]),
                                                 ptr(f1[2]), len(f1[0]), ptr(kps2), ptr(desc2), ptr(has_mp2), ptr(u_right2), len(kps2), ptr(f2[0]),
                                                 ptr(f2[1]), ptr(f2[2]), len(f2[0]), ptr(F12), ex, ey, ptr(sf), ptr(ls), len(sf), int(bOnlyStereo),
                                                 int(self.mbCheckOrientation), ptr(m12), C.byref(nm)))
        return nm.value, m12

    def SearchBySim3(self, kps1_un, desc1, bounds1, kps2_un, desc2, bounds2, q12, q21, th_dist=TH_HIGH):
        """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1104-1328).  q12 / q21 = dict(u, v, radius, level, desc[, valid]) with one
        entry per keypoint of keyframe 1 / 2.  -> nFound, match12"""
        kps1_un = np.ascontiguousarray(kps1_un, KP_DTYPE); kps2_un = np.ascontiguousarray(kps2_un, KP_DTYPE)
        desc1, desc2 = _b(desc1), _b(desc2)
        b1, b2 = _f(bounds1), _f(bounds2)
        a = [(_f(q["u"]), _f(q["v"]), _f(q["radius"]), _i(q["level"]), _b(q["desc"]), _b(q.get("valid"))) for q in (q12, q21)]
        m12 = np.full(len(kps1_un), -1, np.int32); nf = C.c_int32(0)
        check(lib().orb_search_by_sim3(self.device, ptr(kps1_un), ptr(desc1), len(kps1_un), ptr(b1), ptr(kps2_un), ptr(desc2), len(kps2_un), ptr(b2),
                                       *[ptr(x) for x in a[0]], *[ptr(x) for x in a[1]], th_dist, ptr(m12), C.byref(nf)))
        return nf.value, m12


    def FuseSearch(self, kps_un, desc, u_right, bounds, inv_level_sigma2, q_u, q_v, q_ur, q_radius, q_level, q_desc, q_valid=None):
        """The search half of ORBmatcher::Fuse (ORBmatcher.cc:827-977; inv_level_sigma2=None: the Sim3 overload :979-1102).
        -> (best_idx, best_dist); the reference fuses point i into keypoint best_idx[i] when best_dist[i] <= TH_LOW."""
        kps_un = np.ascontiguousarray(kps_un, KP_DTYPE)
        desc, q_desc, q_valid = _b(desc), _b(q_desc), _b(q_valid)
        u_right, inv, b = _f(u_right), _f(inv_level_sigma2), _f(bounds)
        q_u, q_v, q_ur, q_radius, q_level = _f(q_u), _f(q_v), _f(q_ur), _f(q_radius), _i(q_level)
        nq = len(q_u)
        bi = np.full(nq, -1, np.int32); bd = np.full(nq, 256, np.int32)
        check(lib().orb_fuse_search(self.device, ptr(kps_un), ptr(desc), ptr(u_right), len(kps_un), ptr(b), ptr(inv), 0 if inv is None else len(inv),
                                    nq, ptr(q_u), ptr(q_v), ptr(q_ur), ptr(q_radius), ptr(q_level), ptr(q_desc), ptr(q_valid), ptr(bi), ptr(bd)))
        return bi, bd

def distinctive_descriptors(desc, off, device=0):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:288-361) for many map points: rows [off[p], off[p+1]) of desc are
    the observations of point p.  -> (best_idx, best_desc)"""
    desc = _b(desc); off = _i(off)
    npts = len(off) - 1
    best = np.full(npts, -1, np.int32); bd = np.zeros((npts, 32), np.uint8)
    check(lib().orb_distinctive_descriptors(device, ptr(desc), ptr(off), npts, ptr(best), ptr(bd)))
    return best, bd

def hamming_top2(q, db, device=0):
    """Brute-force best / second-best (ORBmatcher.cc:202-227 update rule) of each query over db (host arrays)."""
    q, db = _b(q), _b(db)
    out = np.zeros(len(q), TOP2_DTYPE)
    check(lib().orb_hamming_top2(device, ptr(q), len(q), ptr(db), len(db), ptr(out)))
    return out


def hamming_top2_csr(q, db, cand_off, cand_idx, device=0):
    """Best / second-best of each query over its own candidate list (CSR): candidates of query i are
    cand_idx[cand_off[i]:cand_off[i+1]] (rows of db); ties keep the first candidate in list order."""
    q, db = _b(q), _b(db)
    cand_off, cand_idx = _i(cand_off), _i(cand_idx)
    out = np.zeros(len(q), TOP2_DTYPE)
    check(lib().orb_hamming_top2_csr(device, ptr(q), len(q), ptr(db), len(db), ptr(cand_off), ptr(cand_idx), ptr(out)))
    return out


def top2_merge(parts):
    """parts: TOP2_DTYPE [nparts, nq] (e.g. the all-gathered per-shard results) -> merged [nq]."""
    parts = np.ascontiguousarray(parts, TOP2_DTYPE)
    nparts, nq = parts.shape
    out = np.zeros(nq, TOP2_DTYPE)
    check(lib().orb_top2_merge(ptr(parts), nparts, nq, ptr(out)))
    return out


def top2_merge_device(d_parts, nparts, nq, d_out, device=0, stream=0):
    """Device-side merge (raw device pointers, asynchronous on `stream`): d_parts = TOP2_DTYPE[nparts][nq]."""
    check(lib().orb_top2_merge_device(device, C.c_void_p(d_parts), nparts, nq, C.c_void_p(d_out), C.c_void_p(stream)))


class DescriptorDB:
    """One device-resident shard of a descriptor database (BASELINE config 5)."""

    def __init__(self, capacity_rows, index_base=0, device=0):
        self._h = C.c_void_p()
        check(lib().orb_db_create(C.byref(self._h), device, capacity_rows, index_base))
        self.device, self.index_base = device, index_base

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().orb_db_destroy(self._h)
            self._h = None

    __del__ = close

    def add(self, desc):
        desc = _b(desc)
        check(lib().orb_db_add(self._h, ptr(desc), len(desc)))

    def add_device(self, d_ptr, nrows):
        check(lib().orb_db_add_device(self._h, C.c_void_p(d_ptr), nrows))

    def __len__(self):
        return lib().orb_db_size(self._h)

    def set_stream(self, handle):
        check(lib().orb_db_set_stream(self._h, C.c_void_p(handle)))

    def query_top2(self, q):
        q = _b(q)
        out = np.zeros(len(q), TOP2_DTYPE)
        check(lib().orb_db_query_top2(self._h, ptr(q), len(q), ptr(out)))
        return out

    def query_top2_device(self, d_q, nq, d_out):
        check(lib().orb_db_query_top2_device(self._h, C.c_void_p(d_q), nq, C.c_void_p(d_out)))

    def launch_count(self):
        return lib().orb_db_launch_count(self._h)

    def profile_enable(self, on=True):
        check(lib().orb_db_profile_enable(self._h, int(on)))

    def profile_read(self, reset=True):
        """-> (search kernel ms, merge kernel ms, calls), accumulated since the last reset"""
        a, b, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        check(lib().orb_db_profile_read(self._h, C.byref(a), C.byref(b), C.byref(n), int(reset)))
        return a.value, b.value, n.value


def match_bruteforce_batch_device(npairs, d_kps1, d_desc1, d_n1, cap1, d_kps2, d_desc2, d_n2, cap2, d_match12, d_nmatches, th_dist=TH_LOW,
                                  nn_ratio=0.6, check_orientation=True, device=0, stream=0):
    """orb_match_bruteforce for a batch of frame pairs, raw device pointers (outputs of two extract_batch_device calls)."""
    vp = C.c_void_p
    check(lib().orb_match_bruteforce_batch_device(device, npairs, vp(d_kps1), vp(d_desc1), vp(d_n1), cap1, vp(d_kps2), vp(d_desc2), vp(d_n2), cap2,
                                                  th_dist, nn_ratio, int(check_orientation), vp(d_match12), vp(d_nmatches), vp(stream)))


def search_by_projection_batch_device(mode, npairs, batch, bounds, th_dist=TH_HIGH, nn_ratio=0.9, check_orientation=True, device=0, stream=0):
    """orb_search_by_projection for a batch of (target frame, query set) pairs; batch = _lib.SearchBatch of device pointers."""
    prm = SearchParams(mode, th_dist, nn_ratio, int(check_orientation), *map(float, bounds))
    check(lib().orb_search_by_projection_batch_device(device, C.byref(prm), npairs, C.byref(batch), C.c_void_p(stream)))


def shard_unique_id():
    """128 bytes identifying one sharded-database communicator: create on rank 0, hand to every rank (any transport)."""
    buf = np.zeros(128, np.uint8)
    check(lib().orb_shard_unique_id(ptr(buf)))
    return buf


class ShardedDescriptorDB(DescriptorDB):
    """BASELINE config 5 as a library object: this rank's contiguous slice of a descriptor database that is sharded over `world`
    GPUs.  query_top2_device / query_top2 return the exact GLOBAL best / second-best on every rank (per-shard search, NCCL
    all-gather of the 24-byte results and device merge all happen inside liborb_b200.so on the shard's stream)."""

    def __init__(self, capacity_rows, index_base, rank, world, unique_id, device=0):
        self._h = C.c_void_p()
        uid = None if unique_id is None else np.ascontiguousarray(unique_id, np.uint8)
        check(lib().orb_db_create_sharded(C.byref(self._h), device, capacity_rows, index_base, rank, world, ptr(uid)))
        self.device, self.index_base, self.rank, self.world = device, index_base, rank, world

    def query_top2_device(self, d_q, nq, d_out):
        check(lib().orb_db_query_top2_sharded(self._h, C.c_void_p(d_q), nq, C.c_void_p(d_out)))

    def query_top2(self, q):
        q = _b(q)
        out = np.zeros(len(q), TOP2_DTYPE)
        check(lib().orb_db_query_top2_sharded_host(self._h, ptr(q), len(q), ptr(out)))
        return out
