#!/usr/bin/env python
"""Drives the latency-path entry points a few times (single-frame extract, brute-force match, SearchByProjection,
stereo) so that `ncu --metrics gpu__time_duration.sum` lists their kernels.  Also prints host-side wall times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam_2_ros_b200 import ORBextractor, ORBmatcher, compute_stereo_matches, synth  # noqa: E402
from orb_slam_2_ros_b200.matcher import MODE_TRACK_LAST  # noqa: E402

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def t(fn, name):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(REPS):
        fn()
    print("%-28s %.3f ms" % (name, (time.perf_counter() - t0) / REPS * 1e3))


img = synth.synth_frame(0)
ex = ORBextractor(1000)
t(lambda: ex(img), "extract 640x480")
ka, da = ex(img)
kb, db = ex(synth.shifted_frame(img, 3, -2, 0))
m = ORBmatcher(0.6, True)
t(lambda: m.MatchBruteForce(da, ka["angle"], db, kb["angle"], 50), "bruteforce 1000x1000")
sf = ex.mvScaleFactor
q_u = (ka["x"] + np.float32(3)).astype(np.float32); q_v = (ka["y"] - np.float32(2)).astype(np.float32)
q_r = (np.float32(15.0) * sf[ka["octave"]]).astype(np.float32)
m9 = ORBmatcher(0.9, True)
t(lambda: m9.SearchByProjection(MODE_TRACK_LAST, kb, db, (0.0, 0.0, 640.0, 480.0), np.zeros(len(kb), np.uint8), q_u, q_v, q_r,
                                ka["octave"] - 1, ka["octave"] + 1, da, q_angle=ka["angle"], th_dist=100), "search_by_projection")
left, right, _ = synth.synth_stereo_pair(2, 1241, 376)
exl, exr = ORBextractor(2000), ORBextractor(2000)
kl, dl = exl(left); kr, dr = exr(right)
t(lambda: exl(left), "extract 1241x376 (2000)")
t(lambda: compute_stereo_matches(exl, exr, kl, dl, kr, dr, 386.1448, 0.53716), "stereo match only")
