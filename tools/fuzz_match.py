"""Randomised parity campaign for the matching half of the path: brute force, the three SearchByProjection modes,
SearchForInitialization and ComputeStereoMatches (C ABI) against the CPU oracle on random keypoint clouds, clustered
descriptors (ties and near ties), random radii / level ranges / thresholds / ratios / image bounds / taken masks.

    python tools/fuzz_match.py --seconds 240 --seed 1 [--log gpurun_out/fuzz_match.log]

A mismatch prints the case (reproducible from --seed and the case number).  Exit code = number of failures."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam_2_ros_b200 import ORBextractor, ORBmatcher, compute_stereo_matches, synth   # noqa: E402
from orb_slam_2_ros_b200._lib import KP_DTYPE                                              # noqa: E402
from orb_slam_2_ros_b200.matcher import MODE_LOCAL_POINTS, MODE_TRACK_LAST                 # noqa: E402
from oracle import orb_oracle as O                                                         # noqa: E402


def clustered(rng, n, centres, p):
    """n descriptors around `centres` (m x 32): each bit of a copy flips with probability p (p = 0: exact duplicates)"""
    if n == 0:
        return np.zeros((0, 32), np.uint8)
    pick = rng.integers(0, len(centres), n)
    flips = np.packbits(rng.random((n, 256)) < p, axis=1)
    return centres[pick] ^ flips


def cloud(rng, n, bounds, nlev):
    k = np.zeros(n, KP_DTYPE)
    x0, y0, x1, y1 = bounds
    style = rng.integers(3)
    if style == 0:      # uniform
        k["x"] = rng.uniform(x0, x1, n); k["y"] = rng.uniform(y0, y1, n)
    elif style == 1:    # a few dense blobs (long grid cells, contention)
        m = max(1, int(rng.integers(1, 12)))
        cx = rng.uniform(x0, x1, m); cy = rng.uniform(y0, y1, m)
        t = rng.integers(0, m, n)
        k["x"] = np.clip(cx[t] + rng.normal(0, 6, n), x0, np.nextafter(np.float32(x1), np.float32(0)))
        k["y"] = np.clip(cy[t] + rng.normal(0, 6, n), y0, np.nextafter(np.float32(y1), np.float32(0)))
    else:               # integer coordinates (coincident keypoints, window edges hit exactly)
        k["x"] = rng.integers(int(np.ceil(x0)), int(x1), n); k["y"] = rng.integers(int(np.ceil(y0)), int(y1), n)
    k["octave"] = rng.integers(0, nlev, n)
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k["size"] = 31; k["class_id"] = -1
    return k


def eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b))


def case_bruteforce(rng):
    n1, n2 = int(rng.choice([0, 1, 5, 100, 1000, 2300])), int(rng.choice([0, 1, 3, 9, 64, 1000, 1501, 3000]))
    if rng.random() < 0.5:
        n1, n2 = int(rng.integers(0, 2500)), int(rng.integers(0, 2500))
    cen = rng.integers(0, 256, (int(rng.integers(1, 400)), 32), dtype=np.uint8)
    p = float(rng.choice([0.0, 0.01, 0.03, 0.08, 0.3]))
    d1, d2 = clustered(rng, n1, cen, p), clustered(rng, n2, cen, p)
    a1 = rng.uniform(0, 360, n1).astype(np.float32); a2 = rng.uniform(0, 360, n2).astype(np.float32)
    if rng.random() < 0.3:
        a2 = (a1[rng.integers(0, max(n1, 1), n2)] + rng.uniform(-4, 4, n2)).astype(np.float32) % np.float32(360) if n1 else a2
    # th >= 256 is undefined in the reference (best = 256 with index -1 would be accepted)
    th = int(rng.choice([30, 50, 64, 100, 200, 255])); ratio = float(rng.choice([0.5, 0.6, 0.75, 0.9, 1.0, 1.5])); ori = bool(rng.integers(2))
    desc = "bruteforce n1=%d n2=%d centres=%d p=%.2f th=%d ratio=%.2f ori=%d" % (n1, n2, len(cen), p, th, ratio, ori)
    nm_o, m_o = O.match_bruteforce(d1, a1, d2, a2, th, ratio, ori)
    nm_g, m_g = ORBmatcher(ratio, ori).MatchBruteForce(d1, a1, d2, a2, th)
    return desc, nm_o == nm_g and eq(m_o, m_g), "matches %d vs %d" % (nm_g, nm_o)


def case_projection(rng):
    w, h = int(rng.integers(200, 2000)), int(rng.integers(150, 1100))
    pad = float(rng.choice([0, 0, 7.5, 30.25]))
    bounds = (-pad, -pad * 0.5, w + pad, h + pad)
    nlev = int(rng.integers(1, 9))
    n, nq = int(rng.choice([0, 1, 50, 1000, 2000, 4000])), int(rng.choice([0, 1, 30, 1000, 2600, 5000]))
    kb = cloud(rng, n, bounds, nlev)
    cen = rng.integers(0, 256, (int(rng.integers(1, 300)), 32), dtype=np.uint8)
    p = float(rng.choice([0.0, 0.02, 0.05, 0.15]))
    db = clustered(rng, n, cen, p)
    if n and rng.random() < 0.7:      # queries near targets
        t = rng.integers(0, n, nq)
        jit = float(rng.choice([0, 1, 4, 20]))
        q_u = (kb["x"][t] + rng.uniform(-jit, jit, nq)).astype(np.float32); q_v = (kb["y"][t] + rng.uniform(-jit, jit, nq)).astype(np.float32)
        q_desc = db[t] ^ np.packbits(rng.random((nq, 256)) < p, axis=1)
        q_angle = (kb["angle"][t] + rng.uniform(-5, 5, nq)).astype(np.float32) % np.float32(360)
        q_oct = kb["octave"][t].astype(np.int32)
    else:                             # queries anywhere, also outside the image bounds
        q_u = rng.uniform(-50, w + 50, nq).astype(np.float32); q_v = rng.uniform(-50, h + 50, nq).astype(np.float32)
        q_desc = clustered(rng, nq, cen, p); q_angle = rng.uniform(0, 360, nq).astype(np.float32)
        q_oct = rng.integers(0, nlev, nq).astype(np.int32)
    sf = np.float32(1.2) ** np.arange(nlev, dtype=np.float32)
    q_radius = (np.float32(rng.choice([1.0, 3.0, 7.0, 15.0, 40.0, 120.0])) * sf[q_oct]).astype(np.float32)
    lev = int(rng.integers(4))
    if lev == 0:
        q_min, q_max = q_oct - 1, q_oct + 1
    elif lev == 1:
        q_min, q_max = np.full(nq, -1, np.int32), np.full(nq, -1, np.int32)
    elif lev == 2:
        q_min, q_max = np.zeros(nq, np.int32), q_oct
    else:
        q_min, q_max = q_oct, np.full(nq, -1, np.int32)
    q_valid = (rng.random(nq) > 0.1).astype(np.uint8) if rng.random() < 0.7 else None
    q_obs = (rng.random(nq) > 0.4).astype(np.uint8) if rng.random() < 0.5 else None
    stereo = rng.random() < 0.3
    u_right = q_ur = q_er = None
    if stereo:
        u_right = np.where(rng.random(n) < 0.7, kb["x"] - rng.uniform(1, 40, n), -1).astype(np.float32)
        q_ur = (q_u - rng.uniform(1, 40, nq)).astype(np.float32); q_er = q_radius.copy()
    taken0 = (rng.random(n) < float(rng.choice([0, 0.05, 0.5]))).astype(np.uint8)
    track = bool(rng.integers(2))
    mode, omode = (MODE_TRACK_LAST, O.MODE_TRACK_LAST) if track else (MODE_LOCAL_POINTS, O.MODE_LOCAL_POINTS)
    th = int(rng.choice([50, 64, 100, 255])); ratio = float(rng.choice([0.6, 0.8, 0.9, 1.0])); ori = bool(rng.integers(2))
    desc = "projection %s %dx%d pad=%.2f n=%d nq=%d levels=%d levmode=%d stereo=%d th=%d ratio=%.1f ori=%d valid=%d obs=%d" % (
        "track_last" if track else "local_points", w, h, pad, n, nq, nlev, lev, stereo, th, ratio, ori, q_valid is not None, q_obs is not None)
    t_o, t_g = taken0.copy(), taken0.copy()
    grid = O.Grid(kb, *bounds)
    nm_o, moq_o, tq_o = O.search_by_projection(omode, grid, db, u_right, t_o, q_u, q_v, q_radius, q_min, q_max, q_desc, q_ur, q_er, q_angle,
                                               q_valid, q_obs, th_dist=th, nn_ratio=ratio, check_orientation=ori)
    nm_g, moq_g, tq_g = ORBmatcher(ratio, ori).SearchByProjection(mode, kb, db, bounds, t_g, q_u, q_v, q_radius, q_min, q_max, q_desc, u_right,
                                                                  q_ur, q_er, q_angle, q_valid, q_obs, th_dist=th)
    ok = nm_o == nm_g and eq(moq_o, moq_g) and eq(tq_o, tq_g) and eq(t_o, t_g)
    return desc, ok, "matches %d vs %d, moq diff %d, tq diff %d, taken diff %d" % (
        nm_g, nm_o, int((moq_o != moq_g).sum()), int((tq_o != tq_g).sum()), int((t_o != t_g).sum()))


def case_initialization(rng):
    w, h = int(rng.integers(200, 1400)), int(rng.integers(150, 900))
    bounds = (0.0, 0.0, float(w), float(h))
    n1, n2 = int(rng.choice([0, 1, 60, 900, 2000])), int(rng.choice([0, 1, 40, 60, 1000, 2000]))
    kb = cloud(rng, n2, bounds, 3)
    kb["octave"] = (rng.random(n2) < 0.15).astype(np.int32)
    cen = rng.integers(0, 256, (int(rng.integers(1, 200)), 32), dtype=np.uint8)
    p = float(rng.choice([0.0, 0.02, 0.06]))
    db = clustered(rng, n2, cen, p)
    ka = np.zeros(n1, KP_DTYPE)
    if n2:
        t = rng.integers(0, n2, n1)
        ka["x"] = np.clip(kb["x"][t] + rng.uniform(-8, 8, n1), 0, w - 1).astype(np.float32)
        ka["y"] = np.clip(kb["y"][t] + rng.uniform(-8, 8, n1), 0, h - 1).astype(np.float32)
        ka["angle"] = (kb["angle"][t] + rng.uniform(-3, 3, n1)).astype(np.float32) % np.float32(360)
        da = db[t] ^ np.packbits(rng.random((n1, 256)) < p, axis=1)
    else:
        ka = cloud(rng, n1, bounds, 1); da = clustered(rng, n1, cen, p)
    ka["octave"] = (rng.random(n1) < 0.1).astype(np.int32)
    window = int(rng.choice([10, 30, 100, 300])); ratio = float(rng.choice([0.6, 0.9, 1.0])); ori = bool(rng.integers(2))
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    desc = "initialization %dx%d n1=%d n2=%d centres=%d p=%.2f window=%d ratio=%.1f ori=%d" % (w, h, n1, n2, len(cen), p, window, ratio, ori)
    zeros = np.zeros(n1, np.int32)
    grid = O.Grid(kb, *bounds)
    nm_o, m12_o, _ = O.search_by_projection(O.MODE_INITIALIZATION, grid, db, None, np.zeros(n2, np.uint8), prev[:, 0].copy(), prev[:, 1].copy(),
                                            np.full(n1, window, np.float32), zeros, zeros, da, q_angle=ka["angle"],
                                            q_valid=(ka["octave"] == 0).astype(np.uint8), th_dist=50, nn_ratio=ratio, check_orientation=ori)
    nm_g, m12_g = ORBmatcher(ratio, ori).SearchForInitialization(ka, da, kb, db, bounds, prev.copy(), window)
    return desc, nm_o == nm_g and eq(m12_o, m12_g), "matches %d vs %d, m12 diff %d" % (nm_g, nm_o, int((m12_o != m12_g).sum()))


def case_stereo(rng):
    w, h = int(rng.integers(300, 1500)), int(rng.integers(250, 520))   # level 7 of 8 (h / 3.58) must hold a 30-px cell inside the 16-px margins
    nf = int(rng.choice([300, 1000, 2000, 3000])); half = bool(rng.integers(2))
    seed = int(rng.integers(1 << 30))
    dmax = int(rng.choice([20, 60, 150]))
    left, right = synth.synth_stereo_pair(seed, w, h, dmax=dmax, half_pixel=half)[:2]
    if rng.random() < 0.2:      # unrelated right image: few / spurious matches
        right = synth.synth_frame(seed + 1, w, h)
    fx = float(rng.uniform(300, 900)); b = float(rng.choice([0.05, 0.12, 0.5372, 1.0])); bf = float(np.float32(fx * b))
    desc = "stereo %dx%d nf=%d half_pixel=%d dmax=%d seed=%d bf=%.3f b=%.4f" % (w, h, nf, half, dmax, seed, bf, b)
    exl, exr = ORBextractor(nf), ORBextractor(nf)
    kl, dl = exl(left); kr, dr = exr(right)
    oel, oer = O.Extractor(nf), O.Extractor(nf)
    okl, odl = oel.extract(left); okr, odr = oer.extract(right)
    if kl.tobytes() != okl.tobytes() or kr.tobytes() != okr.tobytes() or not eq(dl, odl) or not eq(dr, odr):
        return desc, False, "extraction differs"
    kept_o, ur_o, dep_o, _ = O.stereo_match(oel, oer, okl, odl, okr, odr, bf, b)
    kept_g, ur_g, dep_g = compute_stereo_matches(exl, exr, kl, dl, kr, dr, bf, b)
    ok = kept_o == kept_g and eq(ur_o.view(np.uint32), ur_g.view(np.uint32)) and eq(dep_o.view(np.uint32), dep_g.view(np.uint32))
    return desc, ok, "kept %d vs %d, u_right diff %d" % (kept_g, kept_o, int((ur_o.view(np.uint32) != ur_g.view(np.uint32)).sum()))


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def case_bruteforce_batch(rng):
    """orb_match_bruteforce_batch_device: P pairs of random sizes (incl. empty sides) in arenas of `cap` rows"""
    import torch
    from orb_slam_2_ros_b200.matcher import match_bruteforce_batch_device
    P = int(rng.integers(1, 7)); cap = int(rng.choice([8, 100, 640, 1200, 2048]))
    n1s = [int(rng.integers(0, cap + 1)) if rng.random() < 0.9 else 0 for _ in range(P)]
    n2s = [int(rng.integers(0, cap + 1)) if rng.random() < 0.9 else 0 for _ in range(P)]
    d1 = rng.integers(0, 256, (P, cap, 32), dtype=np.uint8); d2 = rng.integers(0, 256, (P, cap, 32), dtype=np.uint8)   # rows beyond n: garbage that must be ignored
    k1 = np.zeros((P, cap), KP_DTYPE); k2 = np.zeros((P, cap), KP_DTYPE)
    k1["angle"] = rng.uniform(0, 360, (P, cap)); k2["angle"] = rng.uniform(0, 360, (P, cap))
    pflip = float(rng.choice([0.0, 0.01, 0.04, 0.1]))
    for p in range(P):
        cen = rng.integers(0, 256, (int(rng.integers(1, 300)), 32), dtype=np.uint8)
        d1[p, :n1s[p]] = clustered(rng, n1s[p], cen, pflip); d2[p, :n2s[p]] = clustered(rng, n2s[p], cen, pflip)
        if n1s[p] and rng.random() < 0.5:
            k2["angle"][p, :n2s[p]] = (k1["angle"][p, rng.integers(0, n1s[p], n2s[p])] + rng.uniform(-4, 4, n2s[p])) % 360
    th = int(rng.choice([30, 50, 64, 100, 200, 255])); ratio = float(rng.choice([0.5, 0.6, 0.75, 0.9, 1.0, 1.5])); ori = bool(rng.integers(2))
    desc = "bruteforce_batch P=%d cap=%d n1=%s n2=%s p=%.2f th=%d ratio=%.2f ori=%d" % (P, cap, n1s, n2s, pflip, th, ratio, ori)
    tk1, tk2 = _dev(k1.view(np.uint8).reshape(P, cap, -1)), _dev(k2.view(np.uint8).reshape(P, cap, -1))
    td1, td2, tn1, tn2 = _dev(d1), _dev(d2), _dev(np.array(n1s, np.int32)), _dev(np.array(n2s, np.int32))
    m12 = torch.full((P, cap), -7, dtype=torch.int32, device="cuda"); nm = torch.full((P,), -7, dtype=torch.int32, device="cuda")
    match_bruteforce_batch_device(P, tk1.data_ptr(), td1.data_ptr(), tn1.data_ptr(), cap, tk2.data_ptr(), td2.data_ptr(), tn2.data_ptr(), cap,
                                  m12.data_ptr(), nm.data_ptr(), th, ratio, ori, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    m, n = m12.cpu().numpy(), nm.cpu().numpy()
    for p in range(P):
        nm_o, m_o = O.match_bruteforce(d1[p, :n1s[p]], k1["angle"][p, :n1s[p]], d2[p, :n2s[p]], k2["angle"][p, :n2s[p]], th, ratio, ori)
        if n[p] != nm_o or not eq(m[p, :n1s[p]], m_o) or not np.all(m[p, n1s[p]:] == -1):
            return desc, False, "pair %d: matches %d vs %d, m12 diff %d" % (p, n[p], nm_o, int((m[p, :n1s[p]] != m_o).sum()))
    return desc, True, ""


def case_projection_batch(rng):
    """orb_search_by_projection_batch_device: P (target frame, query set) pairs, random counts, both modes, optional arrays absent"""
    import torch
    from orb_slam_2_ros_b200._lib import SearchBatch
    from orb_slam_2_ros_b200.matcher import search_by_projection_batch_device
    P = int(rng.integers(1, 6)); cap = int(rng.choice([16, 300, 1200, 2048])); qcap = int(rng.choice([16, 300, 1200, 2600]))
    w, h = int(rng.integers(200, 1600)), int(rng.integers(150, 1000))
    bounds = (0.0, 0.0, float(w), float(h)); nlev = int(rng.integers(1, 9))
    ns = [int(rng.integers(0, cap + 1)) if rng.random() < 0.9 else 0 for _ in range(P)]
    nqs = [int(rng.integers(0, qcap + 1)) if rng.random() < 0.9 else 0 for _ in range(P)]
    kb = np.zeros((P, cap), KP_DTYPE); db = rng.integers(0, 256, (P, cap, 32), dtype=np.uint8)
    q_u = rng.uniform(0, w, (P, qcap)).astype(np.float32); q_v = rng.uniform(0, h, (P, qcap)).astype(np.float32)
    q_desc = rng.integers(0, 256, (P, qcap, 32), dtype=np.uint8); q_angle = rng.uniform(0, 360, (P, qcap)).astype(np.float32)
    q_oct = rng.integers(0, nlev, (P, qcap)).astype(np.int32)
    pflip = float(rng.choice([0.0, 0.02, 0.06]))
    for p in range(P):
        kb[p, :ns[p]] = cloud(rng, ns[p], bounds, nlev)
        cen = rng.integers(0, 256, (int(rng.integers(1, 200)), 32), dtype=np.uint8)
        db[p, :ns[p]] = clustered(rng, ns[p], cen, pflip)
        if ns[p] and nqs[p]:
            t = rng.integers(0, ns[p], nqs[p])
            q_u[p, :nqs[p]] = kb["x"][p, t] + rng.uniform(-3, 3, nqs[p]); q_v[p, :nqs[p]] = kb["y"][p, t] + rng.uniform(-3, 3, nqs[p])
            q_desc[p, :nqs[p]] = db[p, t] ^ np.packbits(rng.random((nqs[p], 256)) < pflip, axis=1)
            q_angle[p, :nqs[p]] = (kb["angle"][p, t] + rng.uniform(-5, 5, nqs[p])) % 360
            q_oct[p, :nqs[p]] = kb["octave"][p, t]
    sf = np.float32(1.2) ** np.arange(nlev, dtype=np.float32)
    q_radius = (np.float32(rng.choice([2.0, 7.0, 15.0, 30.0])) * sf[q_oct]).astype(np.float32)
    q_min = (q_oct - 1).astype(np.int32); q_max = (q_oct + 1).astype(np.int32)
    if rng.random() < 0.3:
        q_min[:] = -1; q_max[:] = -1
    use_valid, use_obs, stereo = rng.random() < 0.6, rng.random() < 0.5, rng.random() < 0.3
    q_valid = (rng.random((P, qcap)) > 0.1).astype(np.uint8); q_obs = (rng.random((P, qcap)) > 0.4).astype(np.uint8)
    u_right = np.where(rng.random((P, cap)) < 0.7, kb["x"] - rng.uniform(1, 40, (P, cap)), -1).astype(np.float32)
    q_ur = (q_u - rng.uniform(1, 40, (P, qcap))).astype(np.float32); q_er = q_radius.copy()
    taken0 = (rng.random((P, cap)) < float(rng.choice([0, 0.05, 0.4]))).astype(np.uint8)
    track = bool(rng.integers(2))
    mode, omode = (MODE_TRACK_LAST, O.MODE_TRACK_LAST) if track else (MODE_LOCAL_POINTS, O.MODE_LOCAL_POINTS)
    th = int(rng.choice([50, 64, 100, 255])); ratio = float(rng.choice([0.6, 0.8, 0.9, 1.0])); ori = bool(rng.integers(2))
    desc = "projection_batch %s P=%d cap=%d qcap=%d %dx%d n=%s nq=%s stereo=%d valid=%d obs=%d th=%d ratio=%.1f ori=%d" % (
        "track_last" if track else "local_points", P, cap, qcap, w, h, ns, nqs, stereo, use_valid, use_obs, th, ratio, ori)
    t = dict(kb=_dev(kb.view(np.uint8).reshape(P, cap, -1)), db=_dev(db), n=_dev(np.array(ns, np.int32)), nq=_dev(np.array(nqs, np.int32)),
             q_u=_dev(q_u), q_v=_dev(q_v), q_r=_dev(q_radius), q_min=_dev(q_min), q_max=_dev(q_max), q_desc=_dev(q_desc), q_angle=_dev(q_angle),
             q_valid=_dev(q_valid), q_obs=_dev(q_obs), u_right=_dev(u_right), q_ur=_dev(q_ur), q_er=_dev(q_er), taken=_dev(taken0),
             moq=torch.full((P, qcap), -7, dtype=torch.int32, device="cuda"), tq=torch.full((P, cap), -7, dtype=torch.int32, device="cuda"),
             nm=torch.full((P,), -7, dtype=torch.int32, device="cuda"))
    opt = lambda k, on: t[k].data_ptr() if on else None
    b = SearchBatch(t["kb"].data_ptr(), t["db"].data_ptr(), opt("u_right", stereo), t["n"].data_ptr(), cap, t["taken"].data_ptr(), t["nq"].data_ptr(), qcap,
                    t["q_u"].data_ptr(), t["q_v"].data_ptr(), t["q_r"].data_ptr(), t["q_min"].data_ptr(), t["q_max"].data_ptr(), t["q_desc"].data_ptr(),
                    opt("q_ur", stereo), opt("q_er", stereo), t["q_angle"].data_ptr(), opt("q_valid", use_valid), opt("q_obs", use_obs),
                    t["moq"].data_ptr(), t["tq"].data_ptr(), t["nm"].data_ptr())
    search_by_projection_batch_device(mode, P, b, bounds, th, ratio, ori, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    moq, tq, nm, taken = t["moq"].cpu().numpy(), t["tq"].cpu().numpy(), t["nm"].cpu().numpy(), t["taken"].cpu().numpy()
    for p in range(P):
        n, nq = ns[p], nqs[p]
        grid = O.Grid(kb[p, :n], *bounds)
        t_o = taken0[p, :n].copy()
        nm_o, moq_o, tq_o = O.search_by_projection(omode, grid, db[p, :n], u_right[p, :n] if stereo else None, t_o, q_u[p, :nq], q_v[p, :nq],
                                                   q_radius[p, :nq], q_min[p, :nq], q_max[p, :nq], q_desc[p, :nq], q_ur[p, :nq] if stereo else None,
                                                   q_er[p, :nq] if stereo else None, q_angle[p, :nq], q_valid[p, :nq] if use_valid else None,
                                                   q_obs[p, :nq] if use_obs else None, th_dist=th, nn_ratio=ratio, check_orientation=ori)
        if nm[p] == -1:
            continue        # candidate arena overflow, reported (not a wrong result): tests/test_gpu_batch_pairs.py covers the contract
        if nm[p] != nm_o or not eq(moq[p, :nq], moq_o) or not eq(tq[p, :n], tq_o) or not eq(taken[p, :n], t_o):
            return desc, False, "pair %d: matches %d vs %d, moq diff %d, tq diff %d" % (p, nm[p], nm_o, int((moq[p, :nq] != moq_o).sum()),
                                                                                       int((tq[p, :n] != tq_o).sum()))
    return desc, True, ""


def case_stereo_batch(rng):
    """orb_stereo_match_batch_device on P pairs extracted by extract_batch_device (device-resident hand-off)"""
    import torch
    from orb_slam_2_ros_b200.stereo import ComputeStereoMatchesBatchDevice
    P = int(rng.integers(1, 5))
    w, h = int(rng.integers(300, 1400)), int(rng.integers(250, 500))
    nf = int(rng.choice([300, 1000, 2000])); half = bool(rng.integers(2)); seed = int(rng.integers(1 << 30))
    pairs = [synth.synth_stereo_pair(seed + p, w, h, dmax=int(rng.choice([20, 60, 150])), half_pixel=half)[:2] for p in range(P)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    fx = float(rng.uniform(300, 900)); b = float(rng.choice([0.05, 0.12, 0.5372])); bf = float(np.float32(fx * b))
    desc = "stereo_batch P=%d %dx%d nf=%d half_pixel=%d seed=%d bf=%.3f b=%.4f" % (P, w, h, nf, half, seed, bf, b)
    exl, exr = ORBextractor(nf, max_batch=P), ORBextractor(nf, max_batch=P)
    cap = exl.max_keypoints

    def dev_extract(ex, frames):
        d_img = _dev(frames)
        d_kps = torch.zeros((P, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
        d_desc = torch.zeros((P, cap, 32), dtype=torch.uint8, device="cuda"); d_n = torch.zeros(P, dtype=torch.int32, device="cuda")
        ex.extract_batch_device(d_img.data_ptr(), P, w, h, w, w * h, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
        ex.sync()
        return d_img, d_kps, d_desc, d_n
    _, kl, dl, nl = dev_extract(exl, L)
    _, kr, dr, nr = dev_extract(exr, R)
    ur = torch.zeros((P, cap), dtype=torch.float32, device="cuda"); dep = torch.zeros_like(ur); nm = torch.zeros(P, dtype=torch.int32, device="cuda")
    ComputeStereoMatchesBatchDevice(exl, exr, P, kl.data_ptr(), dl.data_ptr(), nl.data_ptr(), kr.data_ptr(), dr.data_ptr(), nr.data_ptr(), cap,
                                    bf, b, ur.data_ptr(), dep.data_ptr(), nm.data_ptr())
    exl.sync()
    ur, dep, nm = ur.cpu().numpy(), dep.cpu().numpy(), nm.cpu().numpy()
    for p in range(P):
        oel, oer = O.Extractor(nf), O.Extractor(nf)
        okl, odl = oel.extract(L[p]); okr, odr = oer.extract(R[p])
        kept_o, ur_o, dep_o, _ = O.stereo_match(oel, oer, okl, odl, okr, odr, bf, b)
        n = len(okl)
        if nm[p] != kept_o or not eq(ur[p, :n].view(np.uint32), ur_o.view(np.uint32)) or not eq(dep[p, :n].view(np.uint32), dep_o.view(np.uint32)):
            return desc, False, "pair %d: kept %d vs %d" % (p, nm[p], kept_o)
    return desc, True, ""


CASES = [case_bruteforce, case_projection, case_projection, case_initialization, case_stereo,
         case_bruteforce_batch, case_projection_batch, case_projection_batch, case_stereo_batch, case_bruteforce_batch]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--log", default="")
    ap.add_argument("--kinds", default="", help="only the case functions whose name contains this substring (e.g. batch)")
    a = ap.parse_args()
    O.build()
    fh = open(a.log, "w") if a.log else None

    def log(s):
        print(s, flush=True)
        if fh:
            fh.write(s + "\n"); fh.flush()

    t0 = time.time()
    fails = n = case = 0
    while time.time() - t0 < a.seconds or a.only >= 0:
        if a.only < 0 or case == a.only:
            rng = np.random.default_rng([a.seed, case])
            fn = CASES[case % len(CASES)]
            if a.kinds not in fn.__name__:
                case += 1
                continue
            try:
                desc, ok, why = fn(rng)
            except Exception as e:       # an error return of the library on a case the oracle handles is a finding too
                desc, ok, why = fn.__name__, False, "exception %s" % str(e)[:200]
            n += 1
            if ok:
                log("ok   case %d: %s" % (case, desc))
            else:
                fails += 1
                log("MISMATCH case %d: %s -> %s" % (case, desc, why))
        if a.only >= 0 and case >= a.only:
            break
        case += 1
    log("fuzz_match: %d cases, %d mismatches, %.0f s" % (n, fails, time.time() - t0))
    return fails


if __name__ == "__main__":
    sys.exit(min(main(), 100))
