"""GPU parity of the LocalMapping / LoopClosing matchers of SURVEY.md §8f N3 against the CPU oracle, through the C ABI:
ORBmatcher::SearchForTriangulation (ORBmatcher.cc:659-825), ORBmatcher::SearchBySim3 (:1104-1328) and the batched
MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:288-361).  Index outputs are compared exactly."""
import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    from orb_slam_2_ros_b200 import ORBextractor, ORBVocabulary
    ex = ORBextractor(1500, 1.2, 8, 20, 7)
    a = synth.synth_frame(12, 752, 480)
    shift = (3, -5)                                   # rows, cols
    b = np.roll(a, shift, axis=(0, 1))
    rng = np.random.default_rng(2)
    noise = rng.random(b.shape) < 0.015
    b = np.where(noise, rng.integers(0, 256, b.shape), b).astype(np.uint8)
    k1, d1 = ex(a)
    k2, d2 = ex(b)
    P = synth.synth_vocabulary(33, k=10, L=5)
    voc = ORBVocabulary.from_arrays(10, 5, 0, 0, *P)
    (_, fv1), (_, fv2) = voc.transform_batch([d1, d2], 4)
    scale = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    return dict(k1=k1, d1=d1, k2=k2, d2=d2, fv1=fv1, fv2=fv2, shift=shift, scale=scale.astype(np.float32), w=752, h=480, voc=voc)


@pytest.mark.parametrize("only_stereo,check_ori,mono", [(False, True, True), (False, True, False), (True, True, False), (False, False, False)])
def test_search_for_triangulation_vs_oracle(oracle, pair, only_stereo, check_ori, mono):
    from orb_slam_2_ros_b200 import ORBmatcher
    p = pair
    rng = np.random.default_rng(11)
    n1, n2 = len(p["k1"]), len(p["k2"])
    has1 = (rng.random(n1) < 0.3).astype(np.uint8); has2 = (rng.random(n2) < 0.3).astype(np.uint8)
    ur1 = None if mono else np.where(rng.random(n1) < 0.6, p["k1"]["x"] - 5, -1).astype(np.float32)
    ur2 = None if mono else np.where(rng.random(n2) < 0.6, p["k2"]["x"] - 5, -1).astype(np.float32)
    dy, dx = p["shift"]
    F12 = np.array([[0, 0, dy], [0, 0, -dx], [-dy, dx, 0]], np.float32) * np.float32(0.013)
    F12 += rng.normal(0, 3e-7, (3, 3)).astype(np.float32)            # slightly off so that the chi-square gate bites
    sigma2 = (p["scale"] * p["scale"]).astype(np.float32)
    counts = []
    for ex_, ey_ in ((1e6, -1e6), (300.0, 200.0)):                      # epipole far away / inside the image (gate active when mono)
        m = ORBmatcher(0.6, check_ori)
        nm, m12 = m.SearchForTriangulation(p["k1"], p["d1"], has1, ur1, p["fv1"], p["k2"], p["d2"], has2, ur2, p["fv2"], F12, ex_, ey_,
                                           p["scale"], sigma2, bOnlyStereo=only_stereo)
        o12, onm = oracle.search_for_triangulation(p["k1"], p["d1"], has1, ur1, p["fv1"], p["k2"], p["d2"], has2, ur2, p["fv2"], F12,
                                                   ex_, ey_, p["scale"], sigma2, only_stereo, check_ori)
        assert nm == onm and np.array_equal(m12, o12)
        counts.append(nm)
    assert counts[0] > 60, counts


def test_search_for_triangulation_ties_take_the_last_candidate(oracle):
    """Equal distances: the reference's `dist > bestDist` gate keeps the LAST candidate (ORBmatcher.cc:741)."""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    k1 = np.zeros(1, KP_DTYPE); k2 = np.zeros(5, KP_DTYPE)
    k1["x"], k1["y"] = 100, 100
    k2["x"] = [110, 120, 130, 140, 150]; k2["y"] = 100
    d1 = np.zeros((1, 32), np.uint8); d2 = np.zeros((5, 32), np.uint8)
    d2[0, 0] = 0x07; d2[1, 0] = 0x01; d2[2, 0] = 0x02; d2[3, 0] = 0x04; d2[4, 0] = 0xFF   # distances 3 1 1 1 8
    fv = lambda n: (np.array([5], np.int32), np.array([0, n], np.int32), np.arange(n, dtype=np.int32))
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)   # horizontal epipolar lines
    sc = np.ones(8, np.float32)
    nm, m12 = ORBmatcher(0.6, False).SearchForTriangulation(k1, d1, None, None, fv(1), k2, d2, None, None, fv(5), F12, 1e6, 1e6, sc, sc)
    o12, onm = oracle.search_for_triangulation(k1, d1, None, None, fv(1), k2, d2, None, None, fv(5), F12, 1e6, 1e6, sc, sc, False, False)
    assert nm == onm == 1 and m12.tolist() == o12.tolist() == [3]


def test_search_by_sim3_vs_oracle(oracle, pair):
    from orb_slam_2_ros_b200 import ORBmatcher
    p = pair
    rng = np.random.default_rng(4)
    w, h = p["w"], p["h"]
    bounds = np.array([0, 0, w, h], np.float32)
    dy, dx = p["shift"]

    def queries(k, d, sx, sy):
        n = len(k)
        lvl = np.clip(k["octave"] + rng.integers(-1, 2, n), 0, 7).astype(np.int32)
        return dict(u=(k["x"] + sx + rng.normal(0, 1.5, n)).astype(np.float32), v=(k["y"] + sy + rng.normal(0, 1.5, n)).astype(np.float32),
                    radius=(np.float32(7.5) * p["scale"][lvl]).astype(np.float32), level=lvl, desc=d,
                    valid=(rng.random(n) < 0.85).astype(np.uint8))
    q12 = queries(p["k1"], p["d1"], dx, dy)
    q21 = queries(p["k2"], p["d2"], -dx, -dy)
    g1 = oracle.Grid(p["k1"], 0, 0, w, h); g2 = oracle.Grid(p["k2"], 0, 0, w, h)
    o12, onf = oracle.search_by_sim3(g1, p["d1"], g2, p["d2"], q12, q21)
    nf, m12 = ORBmatcher().SearchBySim3(p["k1"], p["d1"], bounds, p["k2"], p["d2"], bounds, q12, q21)
    assert nf == onf and nf > 100 and np.array_equal(m12, o12)
    # without validity masks
    q12.pop("valid"); q21.pop("valid")
    o12, onf = oracle.search_by_sim3(g1, p["d1"], g2, p["d2"], q12, q21)
    nf, m12 = ORBmatcher().SearchBySim3(p["k1"], p["d1"], bounds, p["k2"], p["d2"], bounds, q12, q21)
    assert nf == onf and np.array_equal(m12, o12)


def test_distinctive_descriptors_vs_oracle(oracle):
    from orb_slam_2_ros_b200.matcher import distinctive_descriptors
    rng = np.random.default_rng(6)
    sizes = [0, 1, 2, 3, 5, 17, 31, 32, 33, 64, 100, 257, 700] + rng.integers(1, 40, 300).tolist()
    rows, off = [], [0]
    for n in sizes:
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.repeat(base[None], n, 0)
        flips = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
        d = d ^ (flips * (rng.random((n, 1)) < 0.8)).astype(np.uint8)      # 20 % exact copies of the base: ties
        rows.append(d); off.append(off[-1] + n)
    desc = np.concatenate(rows); off = np.asarray(off, np.int32)
    best, bd = distinctive_descriptors(desc, off)
    obest = oracle.distinctive_descriptors(desc, off)
    assert np.array_equal(best, obest)
    for p_, b in enumerate(best):
        if b >= 0:
            assert np.array_equal(bd[p_], desc[off[p_] + b])


@pytest.mark.parametrize("stereo,gate", [(True, True), (False, True), (True, False)])
def test_fuse_search_vs_oracle(oracle, pair, stereo, gate):
    """Search half of ORBmatcher::Fuse (ORBmatcher.cc:827-977) / its Sim3 overload (:979-1102, no reprojection gate)."""
    from orb_slam_2_ros_b200 import ORBmatcher
    p = pair
    rng = np.random.default_rng(17)
    w, h = p["w"], p["h"]
    bounds = np.array([0, 0, w, h], np.float32)
    k2, d2, k1, d1 = p["k2"], p["d2"], p["k1"], p["d1"]
    n1 = len(k1)
    dy, dx = p["shift"]
    ur2 = np.where(rng.random(len(k2)) < 0.6, k2["x"] - rng.uniform(2, 40, len(k2)), -1).astype(np.float32) if stereo else None
    lvl = np.clip(k1["octave"] + rng.integers(-1, 2, n1), 0, 7).astype(np.int32)
    q_u = (k1["x"] + dx + rng.normal(0, 1.2, n1)).astype(np.float32); q_v = (k1["y"] + dy + rng.normal(0, 1.2, n1)).astype(np.float32)
    q_ur = (q_u - rng.uniform(2, 40, n1)).astype(np.float32)
    q_r = (np.float32(3.0) * p["scale"][lvl]).astype(np.float32)
    valid = (rng.random(n1) < 0.9).astype(np.uint8)
    inv_sigma2 = (1.0 / (p["scale"] * p["scale"])).astype(np.float32) if gate else None
    grid = oracle.Grid(k2, 0, 0, w, h)
    obi, obd = oracle.fuse_search(grid, d2, ur2, inv_sigma2, q_u, q_v, q_ur, q_r, lvl, d1, valid)
    bi, bd = ORBmatcher().FuseSearch(k2, d2, ur2, bounds, inv_sigma2, q_u, q_v, q_ur, q_r, lvl, d1, valid)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd)
    assert ((bd <= 50) & (bi >= 0)).sum() > 100
