"""CPU checks of the oracle restatements of the LocalMapping / LoopClosing matchers against
independent numpy formulations on small inputs."""
import numpy as np

from oracle.orb_oracle import KP_DTYPE


def _hamming(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def test_distinctive_descriptors_vs_numpy(oracle):
    rng = np.random.default_rng(1)
    sizes = [0, 1, 2, 3, 4, 9, 33, 40]
    rows, off = [], [0]
    for n in sizes:
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.repeat(base[None], n, 0) ^ (rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & 0x33)
        rows.append(d.astype(np.uint8)); off.append(off[-1] + n)
    desc = np.concatenate(rows)
    best = oracle.distinctive_descriptors(desc, np.asarray(off, np.int32))
    for p, n in enumerate(sizes):
        if n == 0:
            assert best[p] == -1
            continue
        d = desc[off[p]:off[p + 1]]
        D = np.array([[_hamming(d[i], d[j]) for j in range(n)] for i in range(n)])
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert best[p] == int(np.argmin(med))            # argmin = first minimum (MapPoint.cc:349-353)


def test_triangulation_gates_and_tie_rule(oracle):
    k1 = np.zeros(2, KP_DTYPE); k2 = np.zeros(6, KP_DTYPE)
    k1["x"], k1["y"] = [100, 100], [100, 300]
    k2["x"] = [110, 120, 130, 140, 150, 160]; k2["y"] = [100, 100, 100, 100, 100, 140]
    d1 = np.zeros((2, 32), np.uint8); d2 = np.zeros((6, 32), np.uint8)
    d2[0, 0] = 0x07; d2[1, 0] = 0x01; d2[2, 0] = 0x02; d2[3, 0] = 0x04; d2[4, 0] = 0xFF; d2[5, 0] = 0x00
    fv1 = (np.array([5]), np.array([0, 2]), np.array([0, 1])); fv2 = (np.array([5]), np.array([0, 6]), np.arange(6))
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)       # epipolar line of (x, y): y2 = y
    sc = np.ones(8, np.float32)
    m12, nm = oracle.search_for_triangulation(k1, d1, None, None, fv1, k2, d2, None, None, fv2, F12, 1e6, 1e6, sc, sc, False, False)
    # query 0: distance-0 candidate 5 is 40 px off the line (rejected); candidates 1,2,3 tie at 1 -> the last one wins
    # query 1: nothing within sqrt(3.84) px of y = 300
    assert m12.tolist() == [3, -1] and nm == 1
    # epipole next to candidate 3 (monocular): it is skipped, the tie goes to candidate 2
    m12, nm = oracle.search_for_triangulation(k1, d1, None, None, fv1, k2, d2, None, None, fv2, F12, 141.0, 101.0, sc, sc, False, False)
    assert m12.tolist() == [2, -1]
    # has_mp on the query / the target
    m12, nm = oracle.search_for_triangulation(k1, d1, np.array([1, 0], np.uint8), None, fv1, k2, d2, None, None, fv2, F12, 1e6, 1e6, sc, sc, False, False)
    assert nm == 0
    m12, nm = oracle.search_for_triangulation(k1, d1, None, None, fv1, k2, d2, np.array([0, 0, 1, 1, 0, 0], np.uint8), None, fv2, F12, 1e6, 1e6, sc, sc, False, False)
    assert m12.tolist() == [1, -1]


def test_sim3_agreement(oracle):
    rng = np.random.default_rng(2)
    n = 50
    k = np.zeros(n, KP_DTYPE)
    k["x"] = rng.uniform(20, 600, n); k["y"] = rng.uniform(20, 440, n); k["octave"] = rng.integers(0, 3, n)
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    g = oracle.Grid(k, 0, 0, 640, 480)
    q = dict(u=k["x"], v=k["y"], radius=np.full(n, 5, np.float32), level=k["octave"].astype(np.int32), desc=d)
    m12, nf = oracle.search_by_sim3(g, d, g, d, q, q)               # identical keyframes: everybody finds itself
    assert nf == n and np.array_equal(m12, np.arange(n))
    q2 = dict(q, valid=(np.arange(n) % 2).astype(np.uint8))
    m12, nf = oracle.search_by_sim3(g, d, g, d, q2, q)
    assert nf == n // 2 and (m12[::2] == -1).all()


def test_fuse_search_gates(oracle):
    k = np.zeros(4, KP_DTYPE)
    k["x"] = [100, 101, 103, 100]; k["y"] = [100, 100, 100, 100]; k["octave"] = [0, 0, 0, 2]
    d = np.zeros((4, 32), np.uint8); d[0, 0] = 0x0F; d[1, 0] = 0x01; d[2, 0] = 0x00; d[3, 0] = 0x00
    g = oracle.Grid(k, 0, 0, 640, 480)
    inv = np.ones(8, np.float32)
    q = dict(q_u=np.array([100.0]), q_v=np.array([100.0]), q_ur=np.array([90.0]), q_radius=np.array([5.0]), q_level=np.array([0]),
             q_desc=np.zeros((1, 32), np.uint8))
    # keypoint 2 has distance 0 but e2 = 9 > 5.99; keypoint 3 is at the wrong level: keypoint 1 (distance 1) wins
    bi, bd = oracle.fuse_search(g, d, None, inv, **q)
    assert bi.tolist() == [1] and bd.tolist() == [1]
    # without the gate (Sim3 overload) keypoint 2 wins
    bi, bd = oracle.fuse_search(g, d, None, None, **q)
    assert bi.tolist() == [2] and bd.tolist() == [0]
    # stereo keypoints: the right-coordinate error enters e2 (er = 90 - 87 = 3 -> e2 = 1 + 9 = 10 > 7.8 for keypoint 1)
    ur = np.array([90.0, 87.0, -1.0, -1.0], np.float32)
    bi, bd = oracle.fuse_search(g, d, ur, inv, **q)
    assert bi.tolist() == [0] and bd.tolist() == [4]
