"""GPU parity: Hamming search, windowed / brute-force matching and stereo matching (through the C ABI) against
the CPU oracle.  All integer / index results are compared bit-exactly; the stereo floats bit-exactly too."""
import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth

pytestmark = pytest.mark.gpu


def _same_top2(gpu, ora, base=0):
    assert np.array_equal(gpu["best_dist"], ora["best_dist"])
    assert np.array_equal(gpu["second_dist"], ora["second_dist"])
    assert np.array_equal(gpu["best_idx"], np.where(ora["best_idx"] >= 0, ora["best_idx"].astype(np.int64) + base, -1))
    assert np.array_equal(gpu["second_idx"], np.where(ora["second_idx"] >= 0, ora["second_idx"].astype(np.int64) + base, -1))


@pytest.mark.parametrize("nq,ndb", [(700, 5000), (1, 1), (3, 2), (1500, 777), (2100, 40000), (5, 0), (64, 100001)])
def test_hamming_top2_vs_oracle(oracle, nq, ndb):
    from orb_slam_2_ros_b200 import hamming_top2
    rng = np.random.default_rng(nq * 31 + ndb)
    db = rng.integers(0, 256, (ndb, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    if ndb >= 100:   # exact duplicates, near duplicates and ties
        db[ndb // 2] = db[3]; db[ndb - 1] = db[3]
        q[0] = db[3]
        q[min(1, nq - 1)] = db[10] ^ np.uint8(1)
        db[50:60] = db[50]
        q[min(2, nq - 1)] = db[50]
    _same_top2(hamming_top2(q, db), oracle.hamming_top2(q, db))


def test_db_shards_merge(oracle):
    """Config-5 shape in miniature: a database sharded into 3 device-resident shards, replicated queries,
    per-shard top-2, exact merge == single-database oracle; planted queries find their rows."""
    from orb_slam_2_ros_b200 import DescriptorDB, top2_merge
    total, nq = 30000, 400
    db = synth.synth_descriptors(99, 0, total)
    q, planted, flips = synth.synth_queries(99, total, nq)
    bounds = [0, 9000, 21000, total]
    parts = []
    for s in range(3):
        shard = DescriptorDB(bounds[s + 1] - bounds[s], index_base=bounds[s])
        # counter-based generator: the shard builds its own slice
        shard.add(synth.synth_descriptors(99, bounds[s], bounds[s + 1] - bounds[s]))
        assert len(shard) == bounds[s + 1] - bounds[s]
        parts.append(shard.query_top2(q))
        shard.close()
    merged = top2_merge(np.stack(parts))
    _same_top2(merged, oracle.hamming_top2(q, db))
    ok = planted >= 0
    assert np.array_equal(merged["best_idx"][ok], planted[ok])
    assert np.array_equal(merged["best_dist"][ok], flips[ok])


@pytest.fixture(scope="module")
def frame_pair(oracle):
    a = synth.synth_frame(40)
    b = synth.shifted_frame(a, 3, -2, 40)
    ex = oracle.Extractor(1000)
    ka, da = ex.extract(a)
    kb, db = ex.extract(b)
    return ka, da, kb, db, ex.scale_factors.copy()


@pytest.mark.parametrize("variant", ["mono", "stereo", "nonblocking", "backward_levels"])
def test_search_by_projection_track_last(oracle, frame_pair, variant):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1330-1472)."""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200.matcher import MODE_TRACK_LAST
    ka, da, kb, db, sf = frame_pair
    rng = np.random.default_rng(3)
    n, nq = len(kb), len(ka)
    q_u = (ka["x"] + np.float32(3)).astype(np.float32); q_v = (ka["y"] - np.float32(2)).astype(np.float32)
    th = 15.0 if variant != "stereo" else 7.0
    q_radius = (np.float32(th) * sf[ka["octave"]]).astype(np.float32)
    if variant == "backward_levels":
        q_min, q_max = np.zeros(nq, np.int32), ka["octave"].astype(np.int32)
    else:
        q_min, q_max = ka["octave"] - 1, ka["octave"] + 1
    q_valid = (rng.random(nq) > 0.1).astype(np.uint8)
    q_obs = None
    u_right = q_ur = q_er = None
    if variant == "stereo":
        u_right = np.where(rng.random(n) < 0.7, kb["x"] - rng.uniform(1, 40, n), -1).astype(np.float32)
        q_ur = (q_u - rng.uniform(1, 40, nq)).astype(np.float32)
        q_er = q_radius.copy()
    if variant == "nonblocking":
        q_obs = (rng.random(nq) > 0.5).astype(np.uint8)
    taken0 = (rng.random(n) < 0.05).astype(np.uint8)
    bounds = (0.0, 0.0, 640.0, 480.0)
    t_o, t_g = taken0.copy(), taken0.copy()
    grid = oracle.Grid(kb, *bounds)
    nm_o, moq_o, tq_o = oracle.search_by_projection(oracle.MODE_TRACK_LAST, grid, db, u_right, t_o, q_u, q_v, q_radius, q_min,
                                                    q_max, da, q_ur, q_er, ka["angle"], q_valid, q_obs, th_dist=100,
                                                    nn_ratio=0.9, check_orientation=True)
    nm_g, moq_g, tq_g = ORBmatcher(0.9, True).SearchByProjection(MODE_TRACK_LAST, kb, db, bounds, t_g, q_u, q_v, q_radius, q_min,
                                                                 q_max, da, u_right, q_ur, q_er, ka["angle"], q_valid, q_obs,
                                                                 th_dist=100)
    assert nm_o > 100, "test input produced too few matches (%d)" % nm_o
    assert nm_g == nm_o
    assert np.array_equal(moq_g, moq_o)
    assert np.array_equal(tq_g, tq_o)
    assert np.array_equal(t_g, t_o)


def test_search_by_projection_local_points(oracle, frame_pair):
    """ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) (ORBmatcher.cc:45-129): best/second, same-octave ratio."""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200.matcher import MODE_LOCAL_POINTS
    ka, da, kb, db, sf = frame_pair
    rng = np.random.default_rng(4)
    n, nq = len(kb), len(ka)
    q_u = (ka["x"] + np.float32(3) + rng.uniform(-1, 1, nq)).astype(np.float32)
    q_v = (ka["y"] - np.float32(2) + rng.uniform(-1, 1, nq)).astype(np.float32)
    r = np.where(rng.random(nq) < 0.5, 2.5, 4.0).astype(np.float32) * np.float32(3.0)   # RadiusByViewingCos * th
    q_radius = (r * sf[ka["octave"]]).astype(np.float32)
    q_min, q_max = ka["octave"] - 1, ka["octave"]
    taken0 = np.zeros(n, np.uint8)
    bounds = (0.0, 0.0, 640.0, 480.0)
    t_o, t_g = taken0.copy(), taken0.copy()
    grid = oracle.Grid(kb, *bounds)
    for ratio in (0.8, 0.6):
        nm_o, moq_o, tq_o = oracle.search_by_projection(oracle.MODE_LOCAL_POINTS, grid, db, None, t_o, q_u, q_v, q_radius, q_min,
                                                        q_max, da, th_dist=100, nn_ratio=ratio, check_orientation=False)
        nm_g, moq_g, tq_g = ORBmatcher(ratio, False).SearchByProjection(MODE_LOCAL_POINTS, kb, db, bounds, t_g, q_u, q_v, q_radius,
                                                                      q_min, q_max, da, th_dist=100)
        assert nm_g == nm_o and nm_o > 50
        assert np.array_equal(moq_g, moq_o) and np.array_equal(tq_g, tq_o) and np.array_equal(t_g, t_o)


def test_grid_window_order(oracle, frame_pair):
    """GetFeaturesInArea candidate ORDER (ix, iy, insertion) decides ties: check through equal-distance candidates."""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200.matcher import MODE_TRACK_LAST
    _, _, kb, db, _ = frame_pair
    n = len(kb)
    same = np.tile(db[0], (n, 1))              # every target has the same descriptor: all distances tie
    q_u = kb["x"][:300].copy(); q_v = kb["y"][:300].copy()
    q_radius = np.full(300, 40, np.float32)
    q_min = np.full(300, -1, np.int32); q_max = np.full(300, -1, np.int32)
    bounds = (0.0, 0.0, 640.0, 480.0)
    t_o, t_g = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    grid = oracle.Grid(kb, *bounds)
    nm_o, moq_o, tq_o = oracle.search_by_projection(oracle.MODE_TRACK_LAST, grid, same, None, t_o, q_u, q_v, q_radius, q_min, q_max,
                                                    same[:300], q_angle=np.zeros(300, np.float32), check_orientation=False)
    nm_g, moq_g, tq_g = ORBmatcher(0.9, False).SearchByProjection(MODE_TRACK_LAST, kb, same, bounds, t_g, q_u, q_v, q_radius, q_min,
                                                                  q_max, same[:300], q_angle=np.zeros(300, np.float32))
    assert nm_g == nm_o == 300
    assert np.array_equal(moq_g, moq_o) and np.array_equal(tq_g, tq_o)


@pytest.mark.parametrize("ratio,th", [(0.6, 50), (0.9, 100)])
def test_match_bruteforce(oracle, frame_pair, ratio, th):
    from orb_slam_2_ros_b200 import ORBmatcher
    ka, da, kb, db, _ = frame_pair
    nm_o, m_o = oracle.match_bruteforce(da, ka["angle"], db, kb["angle"], th, ratio, True)
    nm_g, m_g = ORBmatcher(ratio, True).MatchBruteForce(da, ka["angle"], db, kb["angle"], th)
    assert nm_o > 50
    assert nm_g == nm_o and np.array_equal(m_g, m_o)


def test_match_bruteforce_collisions(oracle):
    """Many queries compete for few targets: exercises the masked fallback scan of the resolve kernel."""
    from orb_slam_2_ros_b200 import ORBmatcher
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, (6, 32), dtype=np.uint8)
    d2 = np.concatenate([base[rng.integers(0, 6, 40)] ^ (rng.random((40, 32)) < 0.02).astype(np.uint8),
                         rng.integers(0, 256, (25, 32), dtype=np.uint8)])
    d1 = base[rng.integers(0, 6, 300)] ^ (rng.random((300, 32)) < 0.03).astype(np.uint8)
    a1 = rng.uniform(0, 360, 300).astype(np.float32); a2 = rng.uniform(0, 360, len(d2)).astype(np.float32)
    for ratio in (0.95, 1.5):
        nm_o, m_o = oracle.match_bruteforce(d1, a1, d2, a2, 100, ratio, False)
        nm_g, m_g = ORBmatcher(ratio, False).MatchBruteForce(d1, a1, d2, a2, 100)
        assert nm_g == nm_o and np.array_equal(m_g, m_o)
    # fewer targets than the top-K list length
    nm_o, m_o = oracle.match_bruteforce(d1, a1, d2[:3], a2[:3], 100, 1.5, False)
    nm_g, m_g = ORBmatcher(1.5, False).MatchBruteForce(d1, a1, d2[:3], a2[:3], 100)
    assert nm_g == nm_o and np.array_equal(m_g, m_o)


def test_stereo_matches(oracle):
    """Frame::ComputeStereoMatches (Frame.cc:502-676) on a KITTI-shape pair, nFeatures=2000."""
    from orb_slam_2_ros_b200 import ORBextractor, compute_stereo_matches
    left, right, _ = synth.synth_stereo_pair(2)
    bf, fx = np.float32(386.1448), np.float32(718.856)
    b = np.float32(bf / fx)
    exl, exr = ORBextractor(2000), ORBextractor(2000)
    kl, dl = exl(left)
    kr, dr = exr(right)
    oel, oer = oracle.Extractor(2000), oracle.Extractor(2000)
    okl, odl = oel.extract(left)
    okr, odr = oer.extract(right)
    assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
    kept_o, ur_o, dep_o, _ = oracle.stereo_match(oel, oer, okl, odl, okr, odr, float(bf), float(b))
    kept_g, ur_g, dep_g = compute_stereo_matches(exl, exr, kl, dl, kr, dr, float(bf), float(b))
    assert kept_o > 200, "stereo test input too weak (%d matches)" % kept_o
    assert kept_g == kept_o
    assert np.array_equal(ur_g.view(np.uint32), ur_o.view(np.uint32))
    assert np.array_equal(dep_g.view(np.uint32), dep_o.view(np.uint32))


def test_search_contention_chunks_and_empty(oracle, frame_pair):
    """The sequential "already matched" walk under stress: 2600 queries (three shared-memory chunks of 1024) that all
    project onto a few dozen targets, so most top-8 lists run dry and the full candidate lists are scanned; both
    modes; plus empty query / target sets."""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200.matcher import MODE_LOCAL_POINTS, MODE_TRACK_LAST
    ka, da, kb, db, sf = frame_pair
    rng = np.random.default_rng(17)
    n = len(kb)
    hot = rng.choice(n, 40, replace=False)                      # contested targets
    nq = 2600
    tgt = hot[rng.integers(0, len(hot), nq)]
    q_u = (kb["x"][tgt] + rng.uniform(-2, 2, nq)).astype(np.float32)
    q_v = (kb["y"][tgt] + rng.uniform(-2, 2, nq)).astype(np.float32)
    q_radius = np.full(nq, 25.0, np.float32)
    q_min, q_max = np.full(nq, -1, np.int32), np.full(nq, -1, np.int32)   # no level filter
    q_desc = db[tgt] ^ (rng.random((nq, 32)) < 0.03).astype(np.uint8)
    q_angle = kb["angle"][tgt].copy()
    q_obs = (rng.random(nq) > 0.2).astype(np.uint8)
    bounds = (0.0, 0.0, 640.0, 480.0)
    grid = oracle.Grid(kb, *bounds)
    for mode, omode, ratio in ((MODE_TRACK_LAST, oracle.MODE_TRACK_LAST, 0.9), (MODE_LOCAL_POINTS, oracle.MODE_LOCAL_POINTS, 0.8)):
        t_o, t_g = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        nm_o, moq_o, tq_o = oracle.search_by_projection(omode, grid, db, None, t_o, q_u, q_v, q_radius, q_min, q_max, q_desc,
                                                        q_angle=q_angle, q_obs=q_obs, th_dist=100, nn_ratio=ratio,
                                                        check_orientation=True)
        nm_g, moq_g, tq_g = ORBmatcher(ratio, True).SearchByProjection(mode, kb, db, bounds, t_g, q_u, q_v, q_radius, q_min, q_max,
                                                                     q_desc, q_angle=q_angle, q_obs=q_obs, th_dist=100)
        assert nm_o > 20
        assert nm_g == nm_o and np.array_equal(moq_g, moq_o) and np.array_equal(tq_g, tq_o) and np.array_equal(t_g, t_o)
    # empty sets: no queries, no targets
    m = ORBmatcher(0.9, True)
    nm, moq, tq = m.SearchByProjection(MODE_TRACK_LAST, kb, db, bounds, np.zeros(n, np.uint8), q_u[:0], q_v[:0], q_radius[:0],
                                       q_min[:0], q_max[:0], q_desc[:0], q_angle=q_angle[:0])
    assert nm == 0 and len(moq) == 0 and np.all(tq == -1)
    nm, moq, tq = m.SearchByProjection(MODE_TRACK_LAST, kb[:0], db[:0], bounds, np.zeros(0, np.uint8), q_u[:5], q_v[:5], q_radius[:5],
                                       q_min[:5], q_max[:5], q_desc[:5], q_angle=q_angle[:5])
    assert nm == 0 and np.all(moq == -1)


def test_match_bruteforce_large_and_empty(oracle):
    """More queries than one shared-memory chunk of the resolve kernel (1024), a non-multiple-of-8 target count (row
    pitch padding of the distance matrix), and empty sides."""
    from orb_slam_2_ros_b200 import ORBmatcher
    rng = np.random.default_rng(23)
    d2 = rng.integers(0, 256, (1501, 32), dtype=np.uint8)
    pick = rng.integers(0, len(d2), 2300)
    d1 = d2[pick] ^ (rng.random((2300, 32)) < 0.04).astype(np.uint8)
    a1 = rng.uniform(0, 360, len(d1)).astype(np.float32); a2 = rng.uniform(0, 360, len(d2)).astype(np.float32)
    for ratio, ori in ((0.9, True), (0.6, False)):
        nm_o, m_o = oracle.match_bruteforce(d1, a1, d2, a2, 80, ratio, ori)
        nm_g, m_g = ORBmatcher(ratio, ori).MatchBruteForce(d1, a1, d2, a2, 80)
        assert nm_o > 100 and nm_g == nm_o and np.array_equal(m_g, m_o)
    nm, m = ORBmatcher(0.9, True).MatchBruteForce(d1[:0], a1[:0], d2, a2, 80)
    assert nm == 0 and len(m) == 0
    nm, m = ORBmatcher(0.9, True).MatchBruteForce(d1[:7], a1[:7], d2[:0], a2[:0], 80)
    assert nm == 0 and np.all(m == -1)


@pytest.mark.parametrize("ratio,window", [(0.9, 100), (0.6, 30)])
def test_search_for_initialization(oracle, frame_pair, ratio, window):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:406-521): level-0 keypoints only, large windows, targets are
    stolen by strictly better matches and the previous owner loses its match, ratio test, rotation histogram that
    also counts the stale assignments."""
    from orb_slam_2_ros_b200 import ORBmatcher
    ka, da, kb, db, _ = frame_pair
    bounds = (0.0, 0.0, 640.0, 480.0)
    n1 = len(ka)
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)          # Tracking.cc: mvbPrevMatched = initial keypoints
    grid = oracle.Grid(kb, *bounds)
    zeros = np.zeros(n1, np.int32)
    valid = (ka["octave"] == 0).astype(np.uint8)
    nm_o, m12_o, m21_o = oracle.search_by_projection(oracle.MODE_INITIALIZATION, grid, db, None, np.zeros(len(kb), np.uint8),
                                                     prev[:, 0].copy(), prev[:, 1].copy(), np.full(n1, window, np.float32), zeros,
                                                     zeros, da, q_angle=ka["angle"], q_valid=valid, th_dist=50, nn_ratio=ratio,
                                                     check_orientation=True)
    prev_g = prev.copy()
    nm_g, m12_g = ORBmatcher(ratio, True).SearchForInitialization(ka, da, kb, db, bounds, prev_g, window)
    assert nm_o > 30, "test input produced too few matches (%d)" % nm_o
    assert nm_g == nm_o and np.array_equal(m12_g, m12_o)
    assert (m12_g >= 0).sum() == nm_g and np.all(ka["octave"][m12_g >= 0] == 0)
    ok = m12_o >= 0
    assert np.array_equal(prev_g[ok, 0], kb["x"][m12_o[ok]]) and np.array_equal(prev_g[~ok], prev[~ok])


def test_search_for_initialization_stealing(oracle):
    """Few targets, many near-duplicate queries: nearly every accepted match steals a target from an earlier query."""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    rng = np.random.default_rng(29)
    n2, n1 = 60, 900
    kb = np.zeros(n2, KP_DTYPE)
    kb["x"] = rng.uniform(100, 540, n2).astype(np.float32); kb["y"] = rng.uniform(100, 380, n2).astype(np.float32)
    kb["angle"] = rng.uniform(0, 360, n2).astype(np.float32); kb["octave"] = 0
    db = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    t = rng.integers(0, n2, n1)
    ka = np.zeros(n1, KP_DTYPE)
    ka["x"] = kb["x"][t] + rng.uniform(-5, 5, n1).astype(np.float32); ka["y"] = kb["y"][t] + rng.uniform(-5, 5, n1).astype(np.float32)
    ka["angle"] = (kb["angle"][t] + rng.uniform(-3, 3, n1)).astype(np.float32) % np.float32(360)
    ka["octave"] = (rng.random(n1) < 0.1).astype(np.int32)             # 10 % of the queries are not on level 0
    da = db[t] ^ (rng.random((n1, 32)) < rng.uniform(0.0, 0.05, (n1, 1))).astype(np.uint8)
    bounds = (0.0, 0.0, 640.0, 480.0)
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    grid = oracle.Grid(kb, *bounds)
    zeros = np.zeros(n1, np.int32)
    for ori in (True, False):
        nm_o, m12_o, _ = oracle.search_by_projection(oracle.MODE_INITIALIZATION, grid, db, None, np.zeros(n2, np.uint8), prev[:, 0].copy(),
                                                     prev[:, 1].copy(), np.full(n1, 100, np.float32), zeros, zeros, da, q_angle=ka["angle"],
                                                     q_valid=(ka["octave"] == 0).astype(np.uint8), th_dist=50, nn_ratio=0.9,
                                                     check_orientation=ori)
        nm_g, m12_g = ORBmatcher(0.9, ori).SearchForInitialization(ka, da, kb, db, bounds, prev.copy(), 100)
        assert nm_g == nm_o and np.array_equal(m12_g, m12_o)
        assert nm_o <= n2


def test_hamming_top2_csr(oracle):
    """Candidate-list (CSR) best / second: ragged lists incl. empty ones, duplicates inside a list, ties resolved by
    list order (not by index) — the inner loop of SearchForTriangulation / SearchBySim3 style routines."""
    from orb_slam_2_ros_b200 import hamming_top2_csr
    rng = np.random.default_rng(31)
    ndb, nq = 3000, 500
    db = rng.integers(0, 256, (ndb, 32), dtype=np.uint8)
    db[100:110] = db[100]                                              # equal descriptors => distance ties
    q = db[rng.integers(0, ndb, nq)] ^ (rng.random((nq, 32)) < 0.05).astype(np.uint8)
    q[3] = db[100]
    lens = rng.integers(0, 200, nq); lens[::17] = 0; lens[5] = 1; lens[6] = 1500
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    idx = rng.integers(0, ndb, off[-1]).astype(np.int32)
    idx[off[3]:off[3] + 6] = [109, 103, 100, 2999, 103, 0]              # ties: first in list order (109) must win
    g = hamming_top2_csr(q, db, off, idx)
    o = oracle.hamming_top2_csr(q, db, off, idx)
    assert np.array_equal(g["best_dist"], o["best_dist"]) and np.array_equal(g["second_dist"], o["second_dist"])
    assert np.array_equal(g["best_idx"], o["best_idx"]) and np.array_equal(g["second_idx"], o["second_idx"])
    if lens[3] >= 6:
        assert g["best_idx"][3] == 109 and g["best_dist"][3] == 0
    assert np.all(g["best_idx"][lens == 0] == -1) and np.all(g["best_dist"][lens == 0] == 256)


def test_top2_merge_device_equals_host(oracle):
    """orb_top2_merge_device (the kernel that follows the NCCL all-gather in bench.py) against the host merge and the
    single-database oracle, incl. ties on equal distance across shards and empty shards."""
    import torch
    from orb_slam_2_ros_b200 import DescriptorDB, top2_merge, top2_merge_device
    from orb_slam_2_ros_b200._lib import TOP2_DTYPE
    total, nq = 12000, 300
    db = synth.synth_descriptors(5, 0, total)
    db[7000] = db[10]; db[11999] = db[10]                       # duplicates across shards: lowest global index wins
    q, _, _ = synth.synth_queries(5, total, nq)
    q[0] = db[10]
    cuts = [0, 4000, 4000, 9000, total]                         # one empty shard
    parts = []
    for s in range(4):
        sh = DescriptorDB(max(cuts[s + 1] - cuts[s], 1), index_base=cuts[s])
        sh.add(db[cuts[s]:cuts[s + 1]])
        parts.append(sh.query_top2(q))
        sh.close()
    parts = np.stack(parts)
    host = top2_merge(parts)
    d_parts = torch.from_numpy(parts.view(np.uint8).reshape(4, nq, TOP2_DTYPE.itemsize).copy()).cuda()
    d_out = torch.zeros((nq, TOP2_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    top2_merge_device(d_parts.data_ptr(), 4, nq, d_out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dev = d_out.cpu().numpy().view(TOP2_DTYPE).reshape(nq)
    for f in ("best_dist", "second_dist", "best_idx", "second_idx"):
        assert np.array_equal(dev[f], host[f]), f
    _same_top2(dev, oracle.hamming_top2(q, db))
    assert dev["best_idx"][0] == 10 and dev["second_idx"][0] == 7000


def _flip_first_bits(desc, j):
    """a copy of the 32-byte descriptor with its first j bits inverted (Hamming distance j)."""
    out = desc.copy()
    bits = np.unpackbits(out, bitorder="little")
    bits[:j] ^= 1
    return np.packbits(bits, bitorder="little")


@pytest.mark.gpu
def test_resolve_rounds_long_dependency_chain(oracle):
    """Adversarial input for the parallel fixed-point resolve: identical queries, targets at distances 0, 1, 2, ...  The
    sequential rule makes query i take target i (every earlier target is gone), so the dependency chain is as long as the
    match count and query i only becomes final in round i + 1; beyond 8 taken targets every top-8 list is exhausted and
    the full-row / full-candidate-list scans run.  Brute force (ratio 0.99: the chain ends where i >= 0.99 (i + 1)) and
    TRACK_LAST / LOCAL_POINTS projection searches (all targets in every query's window)."""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    from orb_slam_2_ros_b200.matcher import MODE_LOCAL_POINTS, MODE_TRACK_LAST
    rng = np.random.default_rng(5)
    q0 = rng.integers(0, 256, 32, dtype=np.uint8)
    n1, n2 = 150, 200
    d1 = np.repeat(q0[None], n1, 0)
    d2 = np.stack([_flip_first_bits(q0, j) for j in range(n2)])
    perm = rng.permutation(n2)                                   # target order must not matter
    d2p = d2[perm]
    a1 = np.zeros(n1, np.float32); a2 = np.zeros(n2, np.float32)
    for ratio, th in ((0.99, 100), (0.9, 100), (1.5, 60)):
        nm_o, m_o = oracle.match_bruteforce(d1, a1, d2p, a2, th, ratio, True)
        nm_g, m_g = ORBmatcher(ratio, True).MatchBruteForce(d1, a1, d2p, a2, th)
        assert nm_g == nm_o and np.array_equal(m_g, m_o)
        if ratio == 0.99:
            assert nm_o == 99 and np.array_equal(perm[m_o[:99]], np.arange(99))   # the chain: query i -> distance i
    # projection searches: all targets within 10 px of every query
    kb = np.zeros(n2, KP_DTYPE)
    kb["x"] = 300 + rng.uniform(-10, 10, n2); kb["y"] = 200 + rng.uniform(-10, 10, n2)
    kb["octave"] = rng.integers(0, 3, n2); kb["size"] = 31; kb["class_id"] = -1
    bounds = (0.0, 0.0, 640.0, 480.0)
    grid = oracle.Grid(kb, *bounds)
    q_u = np.full(n1, 300, np.float32); q_v = np.full(n1, 200, np.float32); q_r = np.full(n1, 40, np.float32)
    q_min, q_max = np.full(n1, -1, np.int32), np.full(n1, -1, np.int32)
    q_obs = (np.arange(n1) % 7 != 3).astype(np.uint8)            # some queries do not block their target
    for mode, omode, ratio in ((MODE_TRACK_LAST, oracle.MODE_TRACK_LAST, 0.9), (MODE_LOCAL_POINTS, oracle.MODE_LOCAL_POINTS, 0.99)):
        t_o, t_g = np.zeros(n2, np.uint8), np.zeros(n2, np.uint8)
        t_o[perm[:3]] = 1; t_g[perm[:3]] = 1                     # the three closest targets were taken before the call
        nm_o, moq_o, tq_o = oracle.search_by_projection(omode, grid, d2p, None, t_o, q_u, q_v, q_r, q_min, q_max, d1, q_angle=a1,
                                                        q_obs=q_obs, th_dist=100, nn_ratio=ratio, check_orientation=False)
        nm_g, moq_g, tq_g = ORBmatcher(ratio, False).SearchByProjection(mode, kb, d2p, bounds, t_g, q_u, q_v, q_r, q_min, q_max, d1,
                                                                      q_angle=a1, q_obs=q_obs, th_dist=100)
        assert nm_o > 30
        assert nm_g == nm_o and np.array_equal(moq_g, moq_o) and np.array_equal(tq_g, tq_o) and np.array_equal(t_g, t_o)


@pytest.mark.gpu
def test_search_for_initialization_steal_chains_in_parallel_rounds(oracle):
    """SearchForInitialization with 1-4 takers per target in decreasing-distance order (every later taker steals), the
    regime the parallel rounds resolve with their per-target taker lists; plus a target whose FIFTH taker forces the
    fallback to the sequential walk."""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    rng = np.random.default_rng(31)
    bounds = (0.0, 0.0, 640.0, 480.0)
    for max_takers in (4, 5):
        n2 = 120
        kb = np.zeros(n2, KP_DTYPE)
        gx, gy = np.meshgrid(np.arange(12), np.arange(10))
        kb["x"] = (40 + 48 * gx.ravel()).astype(np.float32); kb["y"] = (30 + 44 * gy.ravel()).astype(np.float32)   # well separated
        kb["angle"] = rng.uniform(0, 360, n2).astype(np.float32); kb["octave"] = 0
        db = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        takers = rng.integers(1, max_takers + 1, n2)
        takers[7] = max_takers
        rows = []
        for t in range(n2):
            flips = np.sort(rng.choice(np.arange(1, 30), takers[t], replace=False))[::-1]     # strictly decreasing distances
            for fl in flips:
                rows.append((t, int(fl)))
        order = rng.permutation(len(rows))
        # keep the per-target order of decreasing distance while interleaving the targets
        per_t = {t: [r for r in rows if r[0] == t] for t in range(n2)}
        seq = []
        for i in order:
            t = rows[i][0]
            if per_t[t]:
                seq.append(per_t[t].pop(0))
        n1 = len(seq)
        ka = np.zeros(n1, KP_DTYPE)
        da = np.zeros((n1, 32), np.uint8)
        for i, (t, fl) in enumerate(seq):
            ka["x"][i] = kb["x"][t] + rng.uniform(-3, 3); ka["y"][i] = kb["y"][t] + rng.uniform(-3, 3)
            ka["angle"][i] = kb["angle"][t]
            da[i] = _flip_first_bits(db[t], fl)
        prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
        grid = oracle.Grid(kb, *bounds)
        zeros = np.zeros(n1, np.int32)
        for ori in (False, True):
            nm_o, m12_o, _ = oracle.search_by_projection(oracle.MODE_INITIALIZATION, grid, db, None, np.zeros(n2, np.uint8), prev[:, 0].copy(),
                                                         prev[:, 1].copy(), np.full(n1, 20, np.float32), zeros, zeros, da, q_angle=ka["angle"],
                                                         q_valid=np.ones(n1, np.uint8), th_dist=50, nn_ratio=0.9, check_orientation=ori)
            nm_g, m12_g = ORBmatcher(0.9, ori).SearchForInitialization(ka, da, kb, db, bounds, prev.copy(), 20)
            assert nm_g == nm_o and np.array_equal(m12_g, m12_o)
            if not ori:
                assert nm_o == n2                                                    # every target ends up owned by its LAST taker
                last = {t: i for i, (t, fl) in enumerate(seq)}
                assert all(m12_o[i] == t for t, i in last.items()) and (m12_o >= 0).sum() == n2


def test_complementary_descriptor_distance_256(oracle):
    """A distance of 256 (complementary descriptors) never becomes best or second best in the reference (dist < 256 is false,
    ORBmatcher.cc:217-226), so the ratio test sees second = 256, not 255: with best = 153 and ratio 0.6 that decides the match
    (153 < 153.6 but not < 153.0).  Top-K fast path and the full-row fallback must agree with the oracle."""
    from orb_slam_2_ros_b200 import ORBmatcher
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    near = q.copy()
    bits = rng.choice(256, 153, replace=False)
    for bidx in bits:
        near[0, bidx // 8] ^= np.uint8(1 << (bidx % 8))
    comp = q ^ np.uint8(255)
    for targets in (np.concatenate([comp, near]), np.concatenate([near, comp]), np.concatenate([comp] * 12 + [near] + [comp] * 5)):
        a1, a2 = np.zeros(1, np.float32), np.zeros(len(targets), np.float32)
        nm_o, m_o = oracle.match_bruteforce(q, a1, targets, a2, 200, 0.6, False)
        nm_g, m_g = ORBmatcher(0.6, False).MatchBruteForce(q, a1, targets, a2, 200)
        assert nm_o == 1 and nm_g == nm_o and np.array_equal(m_g, m_o)
    # all targets complementary: no match at all
    nm_g, m_g = ORBmatcher(0.6, False).MatchBruteForce(q, np.zeros(1, np.float32), np.concatenate([comp] * 3), np.zeros(3, np.float32), 256)
    nm_o, m_o = oracle.match_bruteforce(q, np.zeros(1, np.float32), np.concatenate([comp] * 3), np.zeros(3, np.float32), 256, 0.6, False)
    assert nm_g == nm_o == 0 and np.array_equal(m_g, m_o)


@pytest.mark.parametrize("style", ["relocalisation_64", "relocalisation_100", "loop_closure"])
def test_search_by_projection_reloc_and_loop_parameterisations(oracle, frame_pair, style):
    """The array-level form of the two remaining SearchByProjection overloads, GPU vs oracle:
    relocalisation refinement (ORBmatcher.cc:1474-1601): best only <= ORBdist (64 / 100), ANY pre-assigned keypoint hidden
    (taken = mvpMapPoints[i2] != NULL), a non-empty sAlreadyFound (q_valid), levels [l-1, l+1], no stereo gate, rotation histogram;
    loop closure (ORBmatcher.cc:291-404): best only <= TH_LOW, pre-filled vpMatched hides keypoints, levels [l-1, l], no histogram.
    (tests/test_gpu_dropin.py runs the same two overloads through the reference's own signatures against the reference's own code.)"""
    from orb_slam_2_ros_b200 import ORBmatcher
    from orb_slam_2_ros_b200.matcher import MODE_TRACK_LAST
    ka, da, kb, db, sf = frame_pair
    rng = np.random.default_rng(23)
    n, nq = len(kb), len(ka)
    q_u = (ka["x"] + np.float32(3)).astype(np.float32); q_v = (ka["y"] - np.float32(2)).astype(np.float32)
    loop = style == "loop_closure"
    th, th_dist, ori = (10.0, 50, False) if loop else ((10.0 if style.endswith("100") else 3.0), int(style.split("_")[1]), True)
    q_radius = (np.float32(th) * sf[ka["octave"]]).astype(np.float32)
    q_min, q_max = ka["octave"] - 1, ka["octave"] + (0 if loop else 1)
    q_valid = (rng.random(nq) > 0.2).astype(np.uint8)            # sAlreadyFound / spAlreadyFound / bad points
    taken0 = (rng.random(n) < 0.15).astype(np.uint8)             # keypoints that already hold a point
    bounds = (0.0, 0.0, 640.0, 480.0)
    t_o, t_g = taken0.copy(), taken0.copy()
    grid = oracle.Grid(kb, *bounds)
    nm_o, moq_o, tq_o = oracle.search_by_projection(oracle.MODE_TRACK_LAST, grid, db, None, t_o, q_u, q_v, q_radius, q_min, q_max, da, None, None,
                                                    ka["angle"], q_valid, None, th_dist=th_dist, nn_ratio=0.9, check_orientation=ori)
    nm_g, moq_g, tq_g = ORBmatcher(0.9, ori).SearchByProjection(MODE_TRACK_LAST, kb, db, bounds, t_g, q_u, q_v, q_radius, q_min, q_max, da, None, None,
                                                                None, ka["angle"], q_valid, None, th_dist=th_dist)
    assert nm_o > 50, "test input produced too few matches (%d)" % nm_o
    assert nm_g == nm_o and np.array_equal(moq_g, moq_o) and np.array_equal(tq_g, tq_o) and np.array_equal(t_g, t_o)
    assert not np.any((taken0 == 1) & (tq_g >= 0))               # a hidden keypoint never receives a match
