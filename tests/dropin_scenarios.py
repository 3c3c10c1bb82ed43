"""Scenarios for the rh_* harness (oracle/ref_shim/harness.cpp): every function builds the SAME world (frames, keyframes,
map points, poses) inside one harness library from seeded numpy data, calls one reference-signature routine and returns
everything a caller could observe afterwards as plain numpy arrays.  tests/test_gpu_dropin.py runs them on the reference
arm (oracle/_ref: the reference's own code) and on the drop-in arm (this repo's host sources over liborb_b200.so) and
compares; tests/test_ref_pin.py compares the reference arm with the oracle restatement."""
import numpy as np

from oracle import orb_ref
from orb_slam_2_ros_b200 import synth

K_TUM = (517.3, 516.5, 318.6, 255.3)
K_KITTI = (718.856, 718.856, 607.1928, 185.2157)
BF_KITTI = 386.1448


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def pose(rx=0.0, ry=0.0, rz=0.0, t=(0.0, 0.0, 0.0)):
    """4x4 float32 Tcw from small Euler angles (rad) and a translation."""
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T.astype(np.float32)


class World:
    """Two extracted frames A / B of one harness library plus a map-point set built from A's keypoints."""

    def __init__(self, H, seed, w=640, h=480, nfeatures=1000, K=K_TUM, bf=40.0, shift=(3, -2), stereo_fraction=0.0):
        self.H, self.seed, self.w, self.h, self.K, self.bf = H, seed, w, h, K, bf
        H.reset_calibration()
        self.ex = H.extractor(nfeatures, 1.2, 8, 20, 7)
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = self.ex.tables()
        self.imgA = synth.synth_frame(seed, w, h)
        self.imgB = synth.shifted_frame(self.imgA, shift[0], shift[1], seed + 1000)
        rng = np.random.default_rng(seed)
        fa = orb_ref.Frame.mono(H, self.ex, self.imgA, K, bf)
        fb = orb_ref.Frame.mono(H, self.ex, self.imgB, K, bf)
        self.a, self.b = fa.get(), fb.get()
        if stereo_fraction > 0:   # re-create both frames from arrays with synthetic right coordinates / depths
            fa = self._with_stereo(fa, self.a, rng, stereo_fraction)
            fb = self._with_stereo(fb, self.b, rng, stereo_fraction)
            self.a, self.b = fa.get(), fb.get()
        self.FA, self.FB = fa, fb
        self.rng = rng
        self._points()

    def _with_stereo(self, F, d, rng, frac):
        n = F.N
        z = rng.uniform(2.0, 12.0, n).astype(np.float32)
        has = rng.random(n) < frac
        ur = np.where(has, d["kps"]["x"] - np.float32(self.bf) / z, np.float32(-1)).astype(np.float32)
        dep = np.where(has, z, np.float32(-1)).astype(np.float32)
        return orb_ref.Frame.from_arrays(self.H, self.ex, d["kps"], d["desc"], self.w, self.h, ur, dep, self.K, self.bf)

    def _points(self):
        """One map point per keypoint of A, back-projected with camera A at the identity pose."""
        fx, fy, cx, cy = self.K
        a, rng, n = self.a, self.rng, self.FA.N
        z = np.where(a["depth"] > 0, a["depth"], rng.uniform(2.0, 12.0, n)).astype(np.float32)
        self.z = z
        X = np.stack([(a["kps_un"]["x"] - cx) / fx * z, (a["kps_un"]["y"] - cy) / fy * z, z], 1).astype(np.float32)
        dist = np.linalg.norm(X, axis=1)
        lvl = a["kps_un"]["octave"]
        max_d = (dist * self.scale[lvl]).astype(np.float32)                 # MapPoint::UpdateNormalAndDepth (MapPoint.cc:427-436)
        min_d = (max_d / self.scale[-1]).astype(np.float32)
        self.pos, self.normal, self.min_d, self.max_d = X, _unit(X).astype(np.float32), min_d, max_d
        self.nobs = rng.integers(0, 4, n).astype(np.int32)
        self.bad = (rng.random(n) < 0.03).astype(np.uint8)
        # the point's representative descriptor: A's descriptor with a few flipped bits
        desc = a["desc"].copy()
        flips = rng.integers(0, 256, (n, 6))
        mask = rng.random((n, 6)) < 0.5
        for k in range(6):
            rows = np.nonzero(mask[:, k])[0]
            desc[rows, flips[rows, k] // 8] ^= (1 << (flips[rows, k] % 8)).astype(np.uint8)
        self.pdesc = desc
        self.pts = self.H.points(n, X, self.normal, desc, self.nobs, self.bad, min_d, max_d)

    def grid_featvec(self, d, cell=64, shift=(0, 0)):
        """A FeatureVector stand-in: node = coarse image cell of the keypoint (so corresponding keypoints of A and B mostly
        share a node), features ascending inside a node like DBoW2's addFeature produces."""
        x = d["kps"]["x"] - shift[0]
        y = d["kps"]["y"] - shift[1]
        node = (np.floor(y / cell).astype(np.int64) * 64 + np.floor(x / cell).astype(np.int64)).astype(np.int32)
        order = np.lexsort((np.arange(len(node)), node))
        nodes, counts = np.unique(node, return_counts=True)
        start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        return nodes.astype(np.int32), start, order.astype(np.int32)


# ------------------------------------------------------------------------------------------------------------------------
def extraction(H, seed, w, h, nfeatures, nlevels):
    ex = H.extractor(nfeatures, 1.2, nlevels, 20, 7)
    img = synth.synth_frame(seed, w, h)
    kps, desc = ex.extract(img)
    out = dict(kps=kps, desc=desc, tables=np.stack(ex.tables()))
    for l in range(nlevels):
        out["level%d" % l] = ex.level(l)
    return out


def stereo_frame(H, seed, w=1241, h=376, nfeatures=2000, half_pixel=False):
    H.reset_calibration()
    exl, exr = H.extractor(nfeatures, 1.2, 8, 20, 7), H.extractor(nfeatures, 1.2, 8, 20, 7)
    left, right = synth.synth_stereo_pair(seed, w, h, half_pixel=half_pixel)[:2]
    F = orb_ref.Frame.stereo(H, exl, exr, left, right, K_KITTI, BF_KITTI)
    d = F.get()
    kr, dr = F.get_right()
    return dict(kps=d["kps"], kps_un=d["kps_un"], desc=d["desc"], u_right=d["u_right"], depth=d["depth"], kps_right=kr, desc_right=dr,
                bounds=H.bounds())


def features_in_area(H, seed):
    W = World(H, seed)
    rng = np.random.default_rng(seed + 5)
    out = []
    for _ in range(200):
        x, y, r = rng.uniform(-20, 660), rng.uniform(-20, 500), rng.uniform(1, 60)
        lo = int(rng.integers(-1, 6)); hi = int(rng.integers(-1, 8))
        out.append(W.FB.features_in_area(x, y, r, lo, hi))
    return dict(counts=np.array([len(o) for o in out]), idx=np.concatenate(out) if out else np.zeros(0, np.int32))


def track_last(H, seed, mono, th, motion, nnratio=0.9, check_ori=True, stereo_fraction=0.0, prefill=False):
    """TrackWithMotionModel: SearchByProjection(CurrentFrame, LastFrame, th, bMono)."""
    W = World(H, seed, stereo_fraction=stereo_fraction)
    rng = np.random.default_rng(seed + 11)
    idx = np.where(rng.random(W.FA.N) < 0.85, np.arange(W.FA.N), -1).astype(np.int32)
    W.FA.set_points(W.pts, idx)
    W.FA.set_outliers((rng.random(W.FA.N) < 0.05).astype(np.uint8))
    W.FA.set_pose(pose())
    tz = {"none": 0.0, "forward": -0.4, "backward": 0.4}[motion]      # tlc.z = -(Rcw^T tcw).z compared with mb = bf / fx
    W.FB.set_pose(pose(0.002, -0.003, 0.01, (0.01, -0.005, tz)))
    if prefill:   # some current keypoints already hold points: with and without observations
        cur = np.where(rng.random(W.FB.N) < 0.1, rng.integers(0, W.pts.n, W.FB.N), -1).astype(np.int32)
        W.FB.set_points(W.pts, cur)
    n = orb_ref.Matcher(H, nnratio, check_ori).search_last_frame(W.FB, W.FA, th, mono)
    return dict(n=np.int64(n), cur_points=W.FB.get_points(W.pts))


def local_points(H, seed, th, nnratio=0.8, stereo_fraction=0.0):
    """Tracking::SearchLocalPoints: isInFrustum on every point, then SearchByProjection(Frame, vpMapPoints, th)."""
    W = World(H, seed, stereo_fraction=stereo_fraction)
    rng = np.random.default_rng(seed + 12)
    W.FB.set_pose(pose(0.001, 0.002, -0.004, (0.02, 0.01, -0.05)))
    cur = np.where(rng.random(W.FB.N) < 0.15, rng.integers(0, W.pts.n, W.FB.N), -1).astype(np.int32)
    W.FB.set_points(W.pts, cur)
    order = rng.permutation(W.pts.n).astype(np.int32)
    n = orb_ref.Matcher(H, nnratio, True).search_local_points(W.FB, W.pts, order, th)
    ts = W.pts.track_state()
    return dict(n=np.int64(n), cur_points=W.FB.get_points(W.pts), in_view=ts["in_view"], proj_x=ts["proj_x"], proj_y=ts["proj_y"],
                proj_xr=ts["proj_xr"], level=ts["level"], view_cos=ts["view_cos"])


def reloc(H, seed, th, orb_dist, nnratio=0.9, check_ori=True):
    """Relocalization refinement: SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)."""
    W = World(H, seed)
    rng = np.random.default_rng(seed + 13)
    idx = np.where(rng.random(W.FA.N) < 0.9, np.arange(W.FA.N), -1).astype(np.int32)
    W.FA.set_pose(pose())
    kf = orb_ref.KeyFrame(W.FA)
    kf.set_points(W.pts, idx)
    W.FB.set_pose(pose(-0.002, 0.001, 0.003, (0.015, 0.0, 0.02)))
    cur = np.where(rng.random(W.FB.N) < 0.12, rng.integers(0, W.pts.n, W.FB.N), -1).astype(np.int32)   # keypoints that already hold ANY point
    W.FB.set_points(W.pts, cur)
    found = rng.choice(W.pts.n, W.pts.n // 5, replace=False).astype(np.int32)                           # sAlreadyFound
    n = orb_ref.Matcher(H, nnratio, check_ori).search_reloc(W.FB, kf, W.pts, found, th, orb_dist)
    return dict(n=np.int64(n), cur_points=W.FB.get_points(W.pts))


def loop_projection(H, seed, th, scale=1.0):
    """LoopClosing::ComputeSim3: SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)."""
    W = World(H, seed)
    rng = np.random.default_rng(seed + 14)
    W.FB.set_pose(pose())
    kf = orb_ref.KeyFrame(W.FB)
    kf.set_points(W.pts, np.full(W.FB.N, -1, np.int32))
    matched = np.where(rng.random(W.FB.N) < 0.2, rng.integers(0, W.pts.n, W.FB.N), -1).astype(np.int32)
    Scw = pose(0.002, -0.001, 0.002, (0.01, 0.01, -0.02)).astype(np.float64)
    Scw[:3, :] *= scale
    order = rng.permutation(W.pts.n).astype(np.int32)
    n, m = orb_ref.Matcher(H, 0.75, True).search_loop(kf, Scw.astype(np.float32), W.pts, order, matched, th)
    return dict(n=np.int64(n), matched=m)


def bow_kf_frame(H, seed, nnratio=0.75, check_ori=True):
    W = World(H, seed)
    rng = np.random.default_rng(seed + 15)
    W.FA.set_featvec(*W.grid_featvec(W.a))
    W.FB.set_featvec(*W.grid_featvec(W.b, shift=(3, -2)))
    W.FA.set_pose(pose())
    kf = orb_ref.KeyFrame(W.FA)
    kf.set_points(W.pts, np.where(rng.random(W.FA.N) < 0.8, np.arange(W.FA.N), -1).astype(np.int32))
    n, m = orb_ref.Matcher(H, nnratio, check_ori).search_bow_kf_frame(kf, W.FB, W.pts)
    return dict(n=np.int64(n), matches=m)


def bow_kf_kf(H, seed, nnratio=0.75, check_ori=True):
    W = World(H, seed)
    rng = np.random.default_rng(seed + 16)
    W.FA.set_featvec(*W.grid_featvec(W.a))
    W.FB.set_featvec(*W.grid_featvec(W.b, shift=(3, -2)))
    W.FA.set_pose(pose()); W.FB.set_pose(pose())
    kf1, kf2 = orb_ref.KeyFrame(W.FA), orb_ref.KeyFrame(W.FB)
    kf1.set_points(W.pts, np.where(rng.random(W.FA.N) < 0.8, np.arange(W.FA.N), -1).astype(np.int32))
    kf2.set_points(W.pts, np.where(rng.random(W.FB.N) < 0.8, rng.integers(0, W.pts.n, W.FB.N), -1).astype(np.int32), observe=False)
    n, m = orb_ref.Matcher(H, nnratio, check_ori).search_bow_kf_kf(kf1, kf2, W.pts)
    return dict(n=np.int64(n), matches12=m)


def initialization(H, seed, window=100, nnratio=0.9, check_ori=True):
    W = World(H, seed, nfeatures=2000)
    prev = np.stack([W.a["kps_un"]["x"], W.a["kps_un"]["y"]], 1)
    n, m, prev2 = orb_ref.Matcher(H, nnratio, check_ori).search_initialization(W.FA, W.FB, prev, window)
    return dict(n=np.int64(n), matches12=m, prev=prev2)


def _fundamental(K, T1, T2):
    """F12 as LocalMapping::ComputeF12 builds it (LocalMapping.cc:594-612), in float64 then rounded: only an input here."""
    fx, fy, cx, cy = K
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    R1, t1, R2, t2 = T1[:3, :3].astype(np.float64), T1[:3, 3].astype(np.float64), T2[:3, :3].astype(np.float64), T2[:3, 3].astype(np.float64)
    R12 = R1 @ R2.T
    t12 = -R12 @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    return (np.linalg.inv(Km).T @ tx @ R12 @ np.linalg.inv(Km)).astype(np.float32)


def triangulation(H, seed, only_stereo=False, stereo_fraction=0.0, check_ori=True):
    W = World(H, seed, stereo_fraction=stereo_fraction)
    rng = np.random.default_rng(seed + 17)
    W.FA.set_featvec(*W.grid_featvec(W.a))
    W.FB.set_featvec(*W.grid_featvec(W.b, shift=(3, -2)))
    T1, T2 = pose(), pose(0.001, -0.002, 0.001, (-0.25, 0.01, 0.02))
    W.FA.set_pose(T1); W.FB.set_pose(T2)
    kf1, kf2 = orb_ref.KeyFrame(W.FA), orb_ref.KeyFrame(W.FB)
    kf1.set_points(W.pts, np.where(rng.random(W.FA.N) < 0.3, np.arange(W.FA.N), -1).astype(np.int32))
    kf2.set_points(W.pts, np.where(rng.random(W.FB.N) < 0.3, rng.integers(0, W.pts.n, W.FB.N), -1).astype(np.int32), observe=False)
    F12 = _fundamental(W.K, T1, T2) if seed % 5 else np.zeros((3, 3), np.float32)   # all-zero F12: den == 0 -> every pair rejected (ORBmatcher.cc:149-150)
    n, pairs = orb_ref.Matcher(H, 0.6, check_ori).search_triangulation(kf1, kf2, F12, only_stereo)
    return dict(n=np.int64(n), pairs=pairs)


def sim3(H, seed, th=7.5, s12=1.0):
    W = World(H, seed)
    rng = np.random.default_rng(seed + 18)
    W.FA.set_pose(pose()); W.FB.set_pose(pose())
    kf1, kf2 = orb_ref.KeyFrame(W.FA), orb_ref.KeyFrame(W.FB)
    i1 = np.where(rng.random(W.FA.N) < 0.85, np.arange(W.FA.N), -1).astype(np.int32)
    kf1.set_points(W.pts, i1)
    # keyframe 2 observes (mostly other) points located where ITS keypoints back-project
    fx, fy, cx, cy = W.K
    z2 = rng.uniform(2.0, 12.0, W.FB.N).astype(np.float32)
    X2 = np.stack([(W.b["kps_un"]["x"] - cx) / fx * z2, (W.b["kps_un"]["y"] - cy) / fy * z2, z2], 1).astype(np.float32)
    d2 = np.linalg.norm(X2, axis=1)
    max2 = (d2 * W.scale[W.b["kps_un"]["octave"]]).astype(np.float32)
    pts2 = H.points(W.FB.N, X2, _unit(X2).astype(np.float32), W.b["desc"], np.ones(W.FB.N, np.int32), (rng.random(W.FB.N) < 0.03).astype(np.uint8),
                    (max2 / W.scale[-1]).astype(np.float32), max2)
    # one PointSet must hold both: rebuild a joint set [A points | B points]
    n1, n2 = W.pts.n, W.FB.N
    joint = H.points(n1 + n2, np.concatenate([W.pos, X2]), np.concatenate([W.normal, _unit(X2).astype(np.float32)]),
                     np.concatenate([W.pdesc, W.b["desc"]]), np.concatenate([W.nobs, np.ones(n2, np.int32)]),
                     np.concatenate([W.bad, (rng.random(n2) < 0.03).astype(np.uint8)]), np.concatenate([W.min_d, (max2 / W.scale[-1]).astype(np.float32)]),
                     np.concatenate([W.max_d, max2]))
    del pts2
    kf1.set_points(joint, i1)
    i2 = np.where(rng.random(n2) < 0.85, n1 + np.arange(n2), -1).astype(np.int32)
    kf2.set_points(joint, i2)
    m12 = np.where(rng.random(W.FA.N) < 0.05, n1 + rng.integers(0, n2, W.FA.N), -1).astype(np.int32)   # already matched
    R12 = pose(0.001, 0.002, -0.001)[:3, :3]
    t12 = np.array([0.004, -0.003, 0.01], np.float32)
    n, m = orb_ref.Matcher(H, 0.75, True).search_sim3(kf1, kf2, joint, m12, s12, R12, t12, th)
    return dict(n=np.int64(n), matches12=m)


def fuse(H, seed, th=3.0, stereo_fraction=0.0):
    W = World(H, seed, stereo_fraction=stereo_fraction)
    rng = np.random.default_rng(seed + 19)
    W.FB.set_pose(pose(0.001, -0.001, 0.002, (0.01, 0.0, -0.01)))
    kf = orb_ref.KeyFrame(W.FB)
    held = np.where(rng.random(W.FB.N) < 0.4, rng.integers(0, W.pts.n, W.FB.N), -1).astype(np.int32)
    kf.set_points(W.pts, held)
    cand = rng.permutation(W.pts.n).astype(np.int32)
    cand[rng.random(len(cand)) < 0.05] = -1                        # NULL entries are skipped (ORBmatcher.cc:843-844)
    n = orb_ref.Matcher(H, 0.6, True).fuse(kf, W.pts, cand, th)
    rep, nobs, bad = W.pts.state()
    return dict(n=np.int64(n), kf_points=kf.get_points(W.pts), replaced_by=rep, nobs=nobs, bad=bad)


def fuse_sim3(H, seed, th=4.0, scale=1.0):
    W = World(H, seed)
    rng = np.random.default_rng(seed + 20)
    W.FB.set_pose(pose())
    kf = orb_ref.KeyFrame(W.FB)
    held = np.where(rng.random(W.FB.N) < 0.4, rng.integers(0, W.pts.n, W.FB.N), -1).astype(np.int32)
    kf.set_points(W.pts, held, observe=False)
    Scw = pose(0.001, 0.001, -0.002, (0.0, 0.01, 0.01)).astype(np.float64)
    Scw[:3, :] *= scale
    cand = rng.permutation(W.pts.n).astype(np.int32)
    n, rep = orb_ref.Matcher(H, 0.6, True).fuse_sim3(kf, Scw.astype(np.float32), W.pts, cand, th)
    _, nobs, bad = W.pts.state()
    return dict(n=np.int64(n), replace=rep, kf_points=kf.get_points(W.pts), nobs=nobs, bad=bad)


def dense_ties(H, seed, n=1500, clusters=40):
    """from-arrays frames with many near-identical descriptors packed into small image regions: long chains of the
    'already matched' rule and distance ties for SearchByProjection(Cur, Last) and SearchForInitialization."""
    rng = np.random.default_rng(seed)
    H.reset_calibration()
    ex = H.extractor(1000, 1.2, 8, 20, 7)
    scale = ex.tables()[0]
    base = rng.integers(0, 256, (clusters, 32), dtype=np.uint8)
    c = rng.integers(0, clusters, n)
    cxy = rng.uniform(60, 420, (clusters, 2))
    off = rng.uniform(-25, 25, (n, 2))

    def frame(jitter):
        kps = np.zeros(n, orb_ref.KP_DTYPE)
        kps["x"] = np.round(cxy[c, 0] + off[:, 0] + jitter * rng.uniform(-1, 1, n)).astype(np.float32) + 100
        kps["y"] = np.round(cxy[c, 1] + off[:, 1] + jitter * rng.uniform(-1, 1, n)).astype(np.float32)
        kps["octave"] = rng.integers(0, 3, n)
        kps["size"] = 31 * scale[kps["octave"]]
        kps["angle"] = (rng.integers(0, 4, n) * 12.0 + rng.uniform(0, 3, n)).astype(np.float32)
        kps["response"] = rng.integers(20, 120, n)
        kps["class_id"] = -1
        d = base[c].copy()
        flip = rng.random(n) < 0.5
        d[flip, rng.integers(0, 32, flip.sum())] ^= np.uint8(1)     # half of them differ from their cluster centre by one bit
        return kps, d

    ka, da = frame(0.0)
    kb, db = frame(3.0)
    kb["octave"] = ka["octave"]; kb["size"] = ka["size"]
    FA = orb_ref.Frame.from_arrays(H, ex, ka, da, 640, 480, K=K_TUM)
    FB = orb_ref.Frame.from_arrays(H, ex, kb, db, 640, 480, K=K_TUM)
    fx, fy, cx, cy = K_TUM
    z = rng.uniform(3, 9, n).astype(np.float32)
    X = np.stack([(ka["x"] - cx) / fx * z, (ka["y"] - cy) / fy * z, z], 1).astype(np.float32)
    pts = H.points(n, X, _unit(X).astype(np.float32), da, rng.integers(0, 3, n).astype(np.int32), None, None, None)
    FA.set_points(pts, np.arange(n, dtype=np.int32))
    FA.set_pose(pose()); FB.set_pose(pose())
    out = {}
    m = orb_ref.Matcher(H, 0.9, True)
    out["last_n"] = np.int64(m.search_last_frame(FB, FA, 15.0, True))
    out["last_points"] = FB.get_points(pts)
    prev = np.stack([ka["x"], ka["y"]], 1)
    ni, m12, prev2 = m.search_initialization(FA, FB, prev, 60)
    out["init_n"], out["init_m12"], out["init_prev"] = np.int64(ni), m12, prev2
    return out


def write_voc_file(path, seed=3, k=8, L=3):
    """A regular synthetic vocabulary in the reference's text format WITHOUT a trailing newline: the reference's
    `while(!f.eof())` loader (TemplatedVocabulary.h:1396-1420) would turn the empty string after a final newline into a
    phantom extra child of the root (INTEGRATION.md pin (v))."""
    parent, is_leaf, desc, weight = synth.synth_vocabulary(seed, k, L)
    synth.write_vocabulary_text(path, k, L, 0, 0, parent, is_leaf, desc, weight)
    txt = open(path).read().rstrip("\n")
    open(path, "w").write(txt)
    return parent, is_leaf, desc, weight


def bag_of_words(H, seed, voc_path):
    """ORBVocabulary::loadFromTextFile + Frame::ComputeBoW on two extracted frames + score + SearchByBoW over the REAL
    feature vectors."""
    W = World(H, seed)
    voc = orb_ref.Vocabulary(H, voc_path)
    (wa, va), fva = W.FA.compute_bow(voc)
    (wb, vb), fvb = W.FB.compute_bow(voc)
    out = dict(nwords=np.int64(voc.size()), words_a=wa, values_a=va, node_a=fva[0], start_a=fva[1], feat_a=fva[2],
               words_b=wb, values_b=vb, node_b=fvb[0], start_b=fvb[1], feat_b=fvb[2], score=np.float64(voc.score(W.FA, W.FB)))
    rng = np.random.default_rng(seed + 30)
    W.FA.set_pose(pose())
    kf = orb_ref.KeyFrame(W.FA)
    kf.set_points(W.pts, np.where(rng.random(W.FA.N) < 0.85, np.arange(W.FA.N), -1).astype(np.int32))
    n, m = orb_ref.Matcher(H, 0.75, True).search_bow_kf_frame(kf, W.FB, W.pts)
    out["bow_n"], out["bow_matches"] = np.int64(n), m
    return out
