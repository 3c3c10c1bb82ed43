"""single-call latency of orb_match_bruteforce on two extracted frames (host -> host), median of 300 calls"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam_2_ros_b200 import ORBextractor, synth
from orb_slam_2_ros_b200.matcher import ORBmatcher
a = synth.synth_frame(1); b = synth.shifted_frame(a, 3, -2, 99)
ex = ORBextractor(1000)
ka, da = ex(a); kb, db = ex(b)
m = ORBmatcher(0.6, True)
for _ in range(20): m.MatchBruteForce(da, ka["angle"], db, kb["angle"], 50)
ts = []
for _ in range(300):
    t0 = time.perf_counter(); r = m.MatchBruteForce(da, ka["angle"], db, kb["angle"], 50); ts.append(time.perf_counter() - t0)
print("bruteforce %d x %d: median %.1f us, min %.1f us, matches %d" % (len(da), len(db), 1e6 * np.median(ts), 1e6 * min(ts), r[0]))
