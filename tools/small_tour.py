"""A small tour of the library (seconds on a GPU box): single frames and a 9-frame batch at two sizes with both FAST variants
(ORB_B200_FAST_DUAL), a flat frame, brute-force match, SearchByProjection, stereo.  Meant for a quick run under a memory checker
where one is available (compute-sanitizer is closed on the build pool) or as a plain does-it-run check."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam_2_ros_b200 import ORBextractor, ORBmatcher, compute_stereo_matches, synth   # noqa: E402
from orb_slam_2_ros_b200.matcher import MODE_TRACK_LAST                                    # noqa: E402

for mode in ("0", "1"):
    os.environ["ORB_B200_FAST_DUAL"] = mode
    for (w, h, nf) in ((320, 240, 500), (645, 487, 1000)):
        imgs = synth.synth_batch(7, 9, w, h, unique=2)
        imgs[1] = 90                                      # a flat frame: every cell empty at both thresholds
        ex = ORBextractor(nf, 1.2, 6, 20, 7, max_batch=9)
        res = ex.extract_batch(imgs)
        k, d = ex(imgs[0])
        print("mode", mode, w, h, [len(r[0]) for r in res[:3]], len(k), flush=True)
a = synth.synth_frame(1, 640, 480); b = synth.shifted_frame(a, 3, -2, 5)
ex = ORBextractor(1000)
ka, da = ex(a); kb, db = ex(b)
nm, m = ORBmatcher(0.6, True).MatchBruteForce(da, ka["angle"], db, kb["angle"], 50)
print("bruteforce", nm, flush=True)
nq = len(ka)
r = np.full(nq, 15, np.float32)
nm, moq, tq = ORBmatcher(0.9, True).SearchByProjection(MODE_TRACK_LAST, kb, db, (0.0, 0.0, 640.0, 480.0), np.zeros(len(kb), np.uint8), ka["x"] + 3, ka["y"] - 2, r,
                                                       ka["octave"] - 1, ka["octave"] + 1, da, q_angle=ka["angle"], th_dist=100)
print("projection", nm, flush=True)
left, right = synth.synth_stereo_pair(2, 620, 300)[:2]
exl, exr = ORBextractor(1000), ORBextractor(1000)
kl, dl = exl(left); kr, dr = exr(right)
kept, ur, dep = compute_stereo_matches(exl, exr, kl, dl, kr, dr, 386.1448, 0.5372)
print("stereo", kept, flush=True)
