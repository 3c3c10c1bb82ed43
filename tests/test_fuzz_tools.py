"""The randomised-campaign tools (tools/fuzz_parity.py, tools/fuzz_match.py) must keep producing valid cases: their generators are
run here against the oracle alone (no GPU) — image contents and extractor parameters the oracle accepts, matcher inputs the oracle
answers without faulting — so that a GPU campaign never fails for a reason of the tool itself."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_image_generators_and_parameters(oracle):
    import fuzz_parity
    rng = np.random.default_rng(3)
    for kind in sorted(set(fuzz_parity.KINDS)):
        img = fuzz_parity.make_image(rng, kind, 333, 251)
        assert img.dtype == np.uint8 and img.shape == (251, 333), kind
        k, d = oracle.Extractor(300, 1.2, 4, 20, 7).extract(img)
        assert d.shape == (len(k), 32)
        if kind in ("synth", "noise", "saltpepper"):
            assert len(k) > 100, kind


def test_matcher_case_generators_against_oracle_only(oracle, monkeypatch):
    """Every single-call case function with the library half replaced by the oracle itself: exercises the generators and the
    oracle's handling of empty sets, coincident keypoints, duplicates, thresholds up to 255 and padded bounds."""
    import fuzz_match as fz
    O = fz.O

    class OracleMatcher:
        def __init__(self, ratio, ori):
            self.r, self.o = ratio, ori

        def MatchBruteForce(self, d1, a1, d2, a2, th):
            return O.match_bruteforce(d1, a1, d2, a2, th, self.r, self.o)

        def SearchByProjection(self, mode, kb, db, bounds, t, q_u, q_v, q_r, q_min, q_max, q_desc, u_right, q_ur, q_er, q_angle, q_valid, q_obs, th_dist):
            om = O.MODE_TRACK_LAST if mode == fz.MODE_TRACK_LAST else O.MODE_LOCAL_POINTS
            return O.search_by_projection(om, O.Grid(kb, *bounds), db, u_right, t, q_u, q_v, q_r, q_min, q_max, q_desc, q_ur, q_er, q_angle, q_valid,
                                          q_obs, th_dist=th_dist, nn_ratio=self.r, check_orientation=self.o)

        def SearchForInitialization(self, ka, da, kb, db, bounds, prev, window):
            n1 = len(ka); z = np.zeros(n1, np.int32)
            nm, m12, _ = O.search_by_projection(O.MODE_INITIALIZATION, O.Grid(kb, *bounds), db, None, np.zeros(len(kb), np.uint8), prev[:, 0].copy(),
                                                prev[:, 1].copy(), np.full(n1, window, np.float32), z, z, da, q_angle=ka["angle"],
                                                q_valid=(ka["octave"] == 0).astype(np.uint8), th_dist=50, nn_ratio=self.r, check_orientation=self.o)
            return nm, m12

    monkeypatch.setattr(fz, "ORBmatcher", OracleMatcher)
    for case in range(45):
        fn = (fz.case_bruteforce, fz.case_projection, fz.case_initialization)[case % 3]
        desc, ok, why = fn(np.random.default_rng([5, case]))
        assert ok, "%s: %s" % (desc, why)
