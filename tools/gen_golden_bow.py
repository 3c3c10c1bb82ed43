#!/usr/bin/env python
"""Generates tests/golden/bow_k6_L3.npz: a small synthetic vocabulary tree (reference text format fields), 300 query
descriptors and the oracle's transform outputs (per-feature word / weight / node, BowVector, FeatureVector, levelsup 2).
The reference ships no vocabulary (ORBvoc.txt is absent from the checkout) and no fixtures for this path; the oracle's
restatement is cross-checked by tests/test_oracle_bow.py (hand-computed tree + independent numpy restatement)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orb_oracle as oo  # noqa: E402
from orb_slam_2_ros_b200 import synth  # noqa: E402


def main():
    P = synth.synth_vocabulary(3, k=6, L=3, irregular=True, p_stop=0.05)
    v = oo.Vocabulary.from_arrays(6, 3, 0, 0, *P)
    rng = np.random.default_rng(5)
    leaves = np.nonzero(P[1])[0]
    q = P[2][rng.choice(leaves, 300)].copy()
    q ^= rng.integers(0, 256, (300, 32), dtype=np.uint8) & rng.integers(0, 256, (300, 32), dtype=np.uint8) & rng.integers(0, 256, (300, 32), dtype=np.uint8)
    w, wt, nd = v.transform_features(q, 2)
    (bw, bv), (fn, fs, ff) = v.transform(q, 2)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bow_k6_L3.npz"), parent=P[0], is_leaf=P[1], desc=P[2], weight=P[3],
                        q=q, word=w, wt=wt, node=nd, bow_word=bw, bow_value=bv, fv_node=fn, fv_start=fs, fv_feat=ff)
    print(len(P[0]), "nodes", len(bw), "words", len(fn), "feature-vector nodes")


if __name__ == "__main__":
    main()
