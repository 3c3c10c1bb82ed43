"""GPU parity of the batched, device-resident pair entry points (per-pair matching / stereo batches):
orb_stereo_match_batch_device, orb_match_bruteforce_batch_device and orb_search_by_projection_batch_device against the
oracle, pair by pair."""
import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth

pytestmark = pytest.mark.gpu
KP = None


def _dev_batch_extract(ex, frames):
    """frames: uint8 [P, h, w] host -> (device tensors kps [P, cap, 7 x f32 as bytes], desc, n) after extract_batch_device"""
    import torch
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    P, h, w = frames.shape
    cap = ex.max_keypoints
    d_img = torch.from_numpy(frames).cuda()
    d_kps = torch.zeros((P, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((P, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(P, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_img.data_ptr(), P, w, h, w, w * h, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
    ex.sync()
    return d_img, d_kps, d_desc, d_n, cap


@pytest.mark.parametrize("w,h,nf", [(1241, 376, 2000), (640, 240, 800)])
def test_stereo_match_batch_device(oracle, w, h, nf):
    import torch
    from orb_slam_2_ros_b200 import ORBextractor
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    from orb_slam_2_ros_b200.stereo import ComputeStereoMatchesBatchDevice
    P = 3
    pairs = [synth.synth_stereo_pair(10 + p, w, h)[:2] for p in range(P)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    exl, exr = ORBextractor(nf, max_batch=P), ORBextractor(nf, max_batch=P)
    _, kl, dl, nl, cap = _dev_batch_extract(exl, L)
    _, kr, dr, nr, _ = _dev_batch_extract(exr, R)
    bf, fx = np.float32(386.1448), np.float32(718.856)
    b = np.float32(bf / fx)
    ur = torch.zeros((P, cap), dtype=torch.float32, device="cuda"); dep = torch.zeros_like(ur)
    nm = torch.zeros(P, dtype=torch.int32, device="cuda")
    ComputeStereoMatchesBatchDevice(exl, exr, P, kl.data_ptr(), dl.data_ptr(), nl.data_ptr(), kr.data_ptr(), dr.data_ptr(), nr.data_ptr(), cap,
                                    float(bf), float(b), ur.data_ptr(), dep.data_ptr(), nm.data_ptr())
    exl.sync()
    ur, dep, nm, nl_h = ur.cpu().numpy(), dep.cpu().numpy(), nm.cpu().numpy(), nl.cpu().numpy()
    for p in range(P):
        oel, oer = oracle.Extractor(nf), oracle.Extractor(nf)
        okl, odl = oel.extract(L[p]); okr, odr = oer.extract(R[p])
        assert nl_h[p] == len(okl)
        gk = kl[p].cpu().numpy().view(KP_DTYPE).reshape(-1)[:len(okl)]
        assert gk.tobytes() == okl.tobytes()
        kept_o, ur_o, dep_o, _ = oracle.stereo_match(oel, oer, okl, odl, okr, odr, float(bf), float(b))
        assert kept_o > 50
        assert nm[p] == kept_o
        n = len(okl)
        assert np.array_equal(ur[p, :n].view(np.uint32), ur_o.view(np.uint32))
        assert np.array_equal(dep[p, :n].view(np.uint32), dep_o.view(np.uint32))
        assert np.all(ur[p, n:] == -1) and np.all(dep[p, n:] == -1)


def _pairs(P, seed0=20):
    A = np.stack([synth.synth_frame(seed0 + p, 640, 480) for p in range(P)])
    B = np.stack([synth.shifted_frame(A[p], 3, -2, seed0 + 100 + p) for p in range(P)])
    return A, B


def test_match_bruteforce_batch_device(oracle):
    """BASELINE config 2, batched: 1000 x 1000 brute force + ratio test + rotation histogram per frame pair."""
    import torch
    from orb_slam_2_ros_b200 import ORBextractor
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    from orb_slam_2_ros_b200.matcher import match_bruteforce_batch_device
    P = 4
    A, B = _pairs(P)
    exa, exb = ORBextractor(1000, max_batch=P), ORBextractor(1000, max_batch=P)
    _, ka, da, na, cap = _dev_batch_extract(exa, A)
    _, kb, db, nb, _ = _dev_batch_extract(exb, B)
    m12 = torch.full((P, cap), -7, dtype=torch.int32, device="cuda"); nm = torch.zeros(P, dtype=torch.int32, device="cuda")
    for th, ratio, ori in ((50, 0.6, True), (100, 0.9, False)):
        match_bruteforce_batch_device(P, ka.data_ptr(), da.data_ptr(), na.data_ptr(), cap, kb.data_ptr(), db.data_ptr(), nb.data_ptr(), cap,
                                      m12.data_ptr(), nm.data_ptr(), th, ratio, ori, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        m12h, nmh = m12.cpu().numpy(), nm.cpu().numpy()
        for p in range(P):
            n1, n2 = int(na[p]), int(nb[p])
            k1 = ka[p].cpu().numpy().view(KP_DTYPE).reshape(-1)[:n1]; k2 = kb[p].cpu().numpy().view(KP_DTYPE).reshape(-1)[:n2]
            d1 = da[p].cpu().numpy()[:n1]; d2 = db[p].cpu().numpy()[:n2]
            nm_o, m_o = oracle.match_bruteforce(d1, k1["angle"], d2, k2["angle"], th, ratio, ori)
            assert nm_o > 50
            assert nmh[p] == nm_o
            assert np.array_equal(m12h[p, :n1], m_o)
            assert np.all(m12h[p, n1:] == -1)


@pytest.mark.parametrize("mode_name", ["track_last", "local_points"])
def test_search_by_projection_batch_device(oracle, mode_name):
    """BASELINE config 2, batched: SearchByProjection windowed matching of frame A's keypoints in frame B, P pairs per call."""
    import torch
    from orb_slam_2_ros_b200 import ORBextractor
    from orb_slam_2_ros_b200._lib import KP_DTYPE, SearchBatch
    from orb_slam_2_ros_b200.matcher import MODE_LOCAL_POINTS, MODE_TRACK_LAST, search_by_projection_batch_device
    P = 4
    A, B = _pairs(P, 40)
    exa, exb = ORBextractor(1000, max_batch=P), ORBextractor(1000, max_batch=P)
    _, ka, da, na, cap = _dev_batch_extract(exa, A)
    _, kb, db, nb, _ = _dev_batch_extract(exb, B)
    sf = exa.GetScaleFactors()
    rng = np.random.default_rng(9)
    kah = ka.cpu().numpy().view(KP_DTYPE).reshape(P, cap)
    kbh = kb.cpu().numpy().view(KP_DTYPE).reshape(P, cap)
    nah, nbh = na.cpu().numpy(), nb.cpu().numpy()
    th = 15.0 if mode_name == "track_last" else 4.0
    q_u = (kah["x"] + np.float32(3)).astype(np.float32); q_v = (kah["y"] - np.float32(2)).astype(np.float32)
    octv = np.clip(kah["octave"], 0, 7)
    q_radius = (np.float32(th) * sf[octv]).astype(np.float32)
    q_min = (octv - 1).astype(np.int32); q_max = (octv + (1 if mode_name == "track_last" else 0)).astype(np.int32)
    q_valid = (rng.random((P, cap)) > 0.1).astype(np.uint8)
    q_obs = (rng.random((P, cap)) > 0.3).astype(np.uint8)
    u_right = np.where(rng.random((P, cap)) < 0.6, kbh["x"] - rng.uniform(1, 40, (P, cap)), -1).astype(np.float32)
    q_ur = (q_u - rng.uniform(1, 40, (P, cap))).astype(np.float32)
    q_er = q_radius.copy()
    q_angle = kah["angle"].astype(np.float32).copy()
    taken0 = (rng.random((P, cap)) < 0.05).astype(np.uint8)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t = dict(q_u=dev(q_u), q_v=dev(q_v), q_radius=dev(q_radius), q_min=dev(q_min), q_max=dev(q_max), q_valid=dev(q_valid), q_obs=dev(q_obs),
             u_right=dev(u_right), q_ur=dev(q_ur), q_er=dev(q_er), q_angle=dev(q_angle), taken=dev(taken0),
             moq=torch.full((P, cap), -7, dtype=torch.int32, device="cuda"), tq=torch.full((P, cap), -7, dtype=torch.int32, device="cuda"),
             nm=torch.zeros(P, dtype=torch.int32, device="cuda"))
    b = SearchBatch(kb.data_ptr(), db.data_ptr(), t["u_right"].data_ptr(), nb.data_ptr(), cap, t["taken"].data_ptr(), na.data_ptr(), cap,
                    t["q_u"].data_ptr(), t["q_v"].data_ptr(), t["q_radius"].data_ptr(), t["q_min"].data_ptr(), t["q_max"].data_ptr(), da.data_ptr(),
                    t["q_ur"].data_ptr(), t["q_er"].data_ptr(), t["q_angle"].data_ptr(), t["q_valid"].data_ptr(), t["q_obs"].data_ptr(),
                    t["moq"].data_ptr(), t["tq"].data_ptr(), t["nm"].data_ptr())
    mode = MODE_TRACK_LAST if mode_name == "track_last" else MODE_LOCAL_POINTS
    omode = oracle.MODE_TRACK_LAST if mode_name == "track_last" else oracle.MODE_LOCAL_POINTS
    bounds = (0.0, 0.0, 640.0, 480.0)
    search_by_projection_batch_device(mode, P, b, bounds, 100, 0.8, True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    moq, tq, nm, taken = t["moq"].cpu().numpy(), t["tq"].cpu().numpy(), t["nm"].cpu().numpy(), t["taken"].cpu().numpy()
    dah, dbh = da.cpu().numpy(), db.cpu().numpy()
    for p in range(P):
        n, nq = int(nbh[p]), int(nah[p])
        grid = oracle.Grid(kbh[p, :n], *bounds)
        t_o = taken0[p, :n].copy()
        nm_o, moq_o, tq_o = oracle.search_by_projection(omode, grid, dbh[p, :n], u_right[p, :n], t_o, q_u[p, :nq], q_v[p, :nq], q_radius[p, :nq],
                                                        q_min[p, :nq], q_max[p, :nq], dah[p, :nq], q_ur[p, :nq], q_er[p, :nq], q_angle[p, :nq],
                                                        q_valid[p, :nq], q_obs[p, :nq], th_dist=100, nn_ratio=0.8, check_orientation=True)
        assert nm_o > 100
        assert nm[p] == nm_o
        assert np.array_equal(moq[p, :nq], moq_o) and np.all(moq[p, nq:] == -1)
        assert np.array_equal(tq[p, :n], tq_o) and np.all(tq[p, n:] == -1)
        assert np.array_equal(taken[p, :n], t_o)


def test_batch_edge_cases_empty_pairs_and_candidate_overflow(oracle):
    """A pair without keypoints (count 0 on the device) yields no matches and leaves the other pairs untouched; windows so large
    that a pair's candidate arena (160 per query on average) overflows report nmatches = -1 for that pair instead of a wrong
    result."""
    import torch
    from orb_slam_2_ros_b200 import ORBextractor
    from orb_slam_2_ros_b200._lib import KP_DTYPE, SearchBatch
    from orb_slam_2_ros_b200.matcher import MODE_TRACK_LAST, match_bruteforce_batch_device, search_by_projection_batch_device
    P = 3
    A, B = _pairs(P, 60)
    exa, exb = ORBextractor(1000, max_batch=P), ORBextractor(1000, max_batch=P)
    _, ka, da, na, cap = _dev_batch_extract(exa, A)
    _, kb, db, nb, _ = _dev_batch_extract(exb, B)
    na2, nb2 = na.clone(), nb.clone()
    na2[1] = 0            # pair 1: no queries
    nb2[2] = 0            # pair 2: no targets
    m12 = torch.full((P, cap), -7, dtype=torch.int32, device="cuda"); nm = torch.full((P,), -7, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    match_bruteforce_batch_device(P, ka.data_ptr(), da.data_ptr(), na2.data_ptr(), cap, kb.data_ptr(), db.data_ptr(), nb2.data_ptr(), cap,
                                  m12.data_ptr(), nm.data_ptr(), 50, 0.6, True, stream=st)
    torch.cuda.synchronize()
    nmh, m = nm.cpu().numpy(), m12.cpu().numpy()
    assert nmh[1] == 0 and nmh[2] == 0 and np.all(m[1] == -1) and np.all(m[2] == -1)
    k1 = ka[0].cpu().numpy().view(KP_DTYPE).reshape(-1)[:int(na[0])]; k2 = kb[0].cpu().numpy().view(KP_DTYPE).reshape(-1)[:int(nb[0])]
    nm_o, m_o = oracle.match_bruteforce(da[0].cpu().numpy()[:len(k1)], k1["angle"], db[0].cpu().numpy()[:len(k2)], k2["angle"], 50, 0.6, True)
    assert nmh[0] == nm_o and np.array_equal(m[0, :len(k1)], m_o)
    # windowed search: empty pairs, then absurd windows
    kah = ka.cpu().numpy().view(KP_DTYPE).reshape(P, cap)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for radius, expect_overflow in ((15.0, False), (5000.0, True)):
        q_u, q_v = dev(kah["x"].astype(np.float32)), dev(kah["y"].astype(np.float32))
        q_r = torch.full((P, cap), radius, dtype=torch.float32, device="cuda")
        lo = torch.full((P, cap), -1, dtype=torch.int32, device="cuda"); hi = torch.full((P, cap), -1, dtype=torch.int32, device="cuda")
        ang = dev(kah["angle"].astype(np.float32))
        taken = torch.zeros((P, cap), dtype=torch.uint8, device="cuda")
        moq = torch.full((P, cap), -7, dtype=torch.int32, device="cuda"); tq = torch.full((P, cap), -7, dtype=torch.int32, device="cuda")
        nms = torch.full((P,), -7, dtype=torch.int32, device="cuda")
        b = SearchBatch(kb.data_ptr(), db.data_ptr(), None, nb2.data_ptr(), cap, taken.data_ptr(), na2.data_ptr(), cap, q_u.data_ptr(), q_v.data_ptr(),
                        q_r.data_ptr(), lo.data_ptr(), hi.data_ptr(), da.data_ptr(), None, None, ang.data_ptr(), None, None, moq.data_ptr(), tq.data_ptr(),
                        nms.data_ptr())
        search_by_projection_batch_device(MODE_TRACK_LAST, P, b, (0.0, 0.0, 640.0, 480.0), 100, 0.9, True, stream=st)
        torch.cuda.synchronize()
        n = nms.cpu().numpy()
        assert n[1] == 0 and n[2] == 0
        assert (n[0] == -1) if expect_overflow else (n[0] > 100)


@pytest.mark.parametrize("th,ratio", [(50, 0.6), (100, 0.9), (100, 0.75)])
def test_match_bruteforce_batch_dense_near_duplicates(oracle, th, ratio):
    """The batched brute-force kernel lists, per query, only the targets whose distance can still change a decision (<= the
    `track` limit derived from th_dist / nn_ratio) and at most 8 of them.  Descriptor sets made of a few prototypes with a
    handful of flipped bits give every query dozens of such targets, many of them taken by earlier queries: the list is
    truncated, the resolve step must fall back to recomputing the row, and the result must still be the oracle's."""
    import torch
    from orb_slam_2_ros_b200.matcher import match_bruteforce_batch_device
    rng = np.random.default_rng(1234 + th)
    P, cap = 3, 640
    n1s, n2s = [600, 640, 37], [640, 500, 41]
    d1 = np.zeros((P, cap, 32), np.uint8); d2 = np.zeros((P, cap, 32), np.uint8)
    a1 = rng.uniform(0, 360, (P, cap)).astype(np.float32); a2 = (a1 + rng.normal(0, 20, (P, cap))).astype(np.float32) % np.float32(360)
    for p in range(P):
        protos = rng.integers(0, 256, (12, 32), dtype=np.uint8)
        def noisy(n):
            out = protos[rng.integers(0, 12, n)].copy()
            bits = np.unpackbits(out, axis=1)
            for i in range(n):
                k = int(rng.integers(0, 40))
                bits[i, rng.choice(256, k, replace=False)] ^= 1
            return np.packbits(bits, axis=1)
        d1[p, :n1s[p]] = noisy(n1s[p]); d2[p, :n2s[p]] = noisy(n2s[p])
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    k1 = np.zeros((P, cap), KP_DTYPE); k2 = np.zeros((P, cap), KP_DTYPE)
    k1["angle"] = a1; k2["angle"] = a2
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tk1, tk2 = dev(k1.view(np.uint8).reshape(P, cap, -1)), dev(k2.view(np.uint8).reshape(P, cap, -1))
    td1, td2 = dev(d1), dev(d2)
    tn1, tn2 = dev(np.array(n1s, np.int32)), dev(np.array(n2s, np.int32))
    m12 = torch.full((P, cap), -7, dtype=torch.int32, device="cuda"); nm = torch.zeros(P, dtype=torch.int32, device="cuda")
    match_bruteforce_batch_device(P, tk1.data_ptr(), td1.data_ptr(), tn1.data_ptr(), cap, tk2.data_ptr(), td2.data_ptr(), tn2.data_ptr(), cap,
                                  m12.data_ptr(), nm.data_ptr(), th, ratio, True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    m, n = m12.cpu().numpy(), nm.cpu().numpy()
    for p in range(P):
        nm_o, m_o = oracle.match_bruteforce(d1[p, :n1s[p]], a1[p, :n1s[p]], d2[p, :n2s[p]], a2[p, :n2s[p]], th, ratio, True)
        assert n[p] == nm_o, (p, n[p], nm_o)
        assert np.array_equal(m[p, :n1s[p]], m_o)
        assert np.all(m[p, n1s[p]:] == -1)
