"""GPU parity of the bag-of-words transform (csrc/orb_bow.cu, through the C ABI) against the CPU oracle
(oracle/orb_oracle_bow.cpp, DBoW2 TemplatedVocabulary.h:1140-1272): word ids, node ids, the double-precision weights
of the BowVector and the FeatureVector lists are all compared bit-exactly."""
import os

import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bow_k6_L3.npz")


def _queries(P, n, seed):
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(P[1])[0]
    q = P[2][rng.choice(leaves, n)].copy()
    flip = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
    q ^= flip                                   # ~12 % of the bits
    q[: n // 8] = rng.integers(0, 256, (n // 8, 32), dtype=np.uint8)   # some unrelated descriptors
    return q


def _same_vectors(gpu, ora):
    (bw, bv), (fn, fs, ff) = gpu
    (obw, obv), (ofn, ofs, off) = ora
    assert np.array_equal(bw, obw)
    assert np.array_equal(bv.view(np.uint64), obv.view(np.uint64))      # bit-exact doubles
    assert np.array_equal(fn, ofn) and np.array_equal(fs, ofs) and np.array_equal(ff, off)


@pytest.mark.parametrize("k,L,irregular,scoring,weighting", [
    (10, 4, False, 0, 0), (10, 5, True, 0, 0), (6, 3, True, 1, 0), (4, 6, True, 5, 1), (20, 2, False, 0, 2), (9, 3, True, 3, 3),
    (18, 3, True, 2, 0)])
def test_transform_vs_oracle(oracle, k, L, irregular, scoring, weighting):
    from orb_slam_2_ros_b200 import ORBVocabulary
    P = synth.synth_vocabulary(k * 100 + L, k=k, L=L, irregular=irregular, p_stop=0.08, p_dup=0.05)
    ov = oracle.Vocabulary.from_arrays(k, L, scoring, weighting, *P)
    gv = ORBVocabulary.from_arrays(k, L, scoring, weighting, *P)
    assert (gv.n_nodes, gv.n_words, gv.k, gv.L) == (ov.n_nodes, ov.n_words, k, L)
    q = _queries(P, 1500, seed=k + L)
    for levelsup in (0, 2, 4, L, L + 3):
        w, wt, nid = gv.transform_features(q, levelsup)
        ow, owt, onid = ov.transform_features(q, levelsup)
        assert np.array_equal(w, ow) and np.array_equal(wt.view(np.uint64), owt.view(np.uint64)) and np.array_equal(nid, onid)
        _same_vectors(gv.transform(q, levelsup), ov.transform(q, levelsup))


def test_batch_ragged_frames(oracle):
    """Several frames in one call: empty frames, one descriptor, the per-frame maximum of 8192."""
    from orb_slam_2_ros_b200 import ORBVocabulary
    P = synth.synth_vocabulary(7, k=10, L=4, irregular=True, p_stop=0.05)
    ov = oracle.Vocabulary.from_arrays(10, 4, 0, 0, *P)
    gv = ORBVocabulary.from_arrays(10, 4, 0, 0, *P)
    sizes = [1000, 0, 1, 2047, 1024, 1025, 8192, 3, 0]
    frames = [_queries(P, max(n, 8), seed=100 + i)[:n] for i, n in enumerate(sizes)]
    res = gv.transform_batch(frames, 4)
    for fr, r in zip(frames, res):
        _same_vectors(r, ov.transform(fr, 4))
    with pytest.raises(Exception):
        gv.transform(_queries(P, 8193, seed=5), 4)


def test_all_words_stopped_and_empty_vocabulary(oracle):
    from orb_slam_2_ros_b200 import ORBVocabulary
    P = list(synth.synth_vocabulary(9, k=5, L=2))
    P[3] = np.zeros_like(P[3])                                      # every weight 0: nothing is added (w > 0, :1170)
    gv = ORBVocabulary.from_arrays(5, 2, 0, 0, *P)
    (bw, bv), (fn, fs, ff) = gv.transform(_queries(P, 200, 1), 1)
    assert len(bw) == 0 and len(fn) == 0 and fs.tolist() == [0]
    ev = ORBVocabulary.from_arrays(10, 6, 0, 0, np.zeros(1, np.int32), np.zeros(1, np.uint8), np.zeros((1, 32), np.uint8), np.zeros(1))
    assert ev.empty()
    (bw, bv), (fn, fs, ff) = ev.transform(np.zeros((10, 32), np.uint8), 4)
    assert len(bw) == 0 and len(fn) == 0


def test_load_text_and_golden(oracle, tmp_path):
    from orb_slam_2_ros_b200 import ORBVocabulary
    g = np.load(GOLD)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, 6, 3, 0, 0, g["parent"], g["is_leaf"], g["desc"], g["weight"])
    gv = ORBVocabulary.loadFromTextFile(path)
    assert gv.n_nodes == len(g["parent"]) and gv.size() == int(g["is_leaf"].sum())
    w, wt, nid = gv.transform_features(g["q"], 2)
    assert np.array_equal(w, g["word"]) and np.array_equal(wt, g["wt"]) and np.array_equal(nid, g["node"])
    (bw, bv), (fn, fs, ff) = gv.transform(g["q"], 2)
    assert np.array_equal(bw, g["bow_word"]) and np.array_equal(bv.view(np.uint64), g["bow_value"].view(np.uint64))
    assert np.array_equal(fn, g["fv_node"]) and np.array_equal(fs, g["fv_start"]) and np.array_equal(ff, g["fv_feat"])
    with pytest.raises(Exception):
        ORBVocabulary.loadFromTextFile(str(tmp_path / "missing.txt"))


def test_orbvoc_shape_extracted_frame(oracle):
    """ORBvoc shape (k = 10, L = 6, 1.1 M nodes) with the descriptors of an extracted frame, levelsup = 4
    (Frame::ComputeBoW, Frame.cc:428-435)."""
    from orb_slam_2_ros_b200 import ORBextractor, ORBVocabulary
    P = synth.synth_vocabulary(11, k=10, L=6)
    assert len(P[0]) == 1111111
    ov = oracle.Vocabulary.from_arrays(10, 6, 0, 0, *P)
    gv = ORBVocabulary.from_arrays(10, 6, 0, 0, *P)
    kps, desc = ORBextractor(1000, 1.2, 8, 20, 7)(synth.synth_frame(0, 640, 480))
    _same_vectors(gv.transform(desc, 4), ov.transform(desc, 4))
    q = _queries(P, 4000, seed=3)
    _same_vectors(gv.transform(q, 4), ov.transform(q, 4))


@pytest.mark.parametrize("keyframe_pair,check_ori,ratio", [(False, True, 0.7), (True, True, 0.75), (False, False, 0.9), (True, False, 0.6)])
def test_search_by_bow_vs_oracle(oracle, keyframe_pair, check_ori, ratio):
    """ORBmatcher::SearchByBoW (ORBmatcher.cc:160-289 / 524-657) over the FeatureVectors of two extracted frames."""
    from orb_slam_2_ros_b200 import ORBextractor, ORBmatcher, ORBVocabulary
    P = synth.synth_vocabulary(21, k=10, L=5)
    gv = ORBVocabulary.from_arrays(10, 5, 0, 0, *P)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    a = synth.synth_frame(4, 640, 480)
    b = np.roll(a, (2, -3), axis=(0, 1))
    rng = np.random.default_rng(8)
    noise = rng.random(b.shape) < 0.02
    b = np.where(noise, rng.integers(0, 256, b.shape), b).astype(np.uint8)
    k1, d1 = ex(a)
    k2, d2 = ex(b)
    for levelsup in (4, 3, 5):      # 5 = root: one node holding everything
        (_, fv1), (_, fv2) = gv.transform_batch([d1, d2], levelsup)
        valid1 = (rng.random(len(d1)) < 0.8).astype(np.uint8)
        valid2 = (rng.random(len(d2)) < 0.9).astype(np.uint8) if keyframe_pair else None
        m = ORBmatcher(ratio, check_ori)
        nm, m12, m21 = m.SearchByBoW(d1, k1["angle"], valid1, fv1, d2, k2["angle"], valid2, fv2, keyframe_pair=keyframe_pair)
        o12, o21, onm = oracle.search_by_bow(d1, k1["angle"], valid1, fv1, d2, k2["angle"], valid2, fv2, 50, keyframe_pair, ratio, check_ori)
        assert nm == onm and np.array_equal(m12, o12) and np.array_equal(m21, o21)
        assert levelsup == 5 or nm > 50
    # degenerate inputs
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    nm, m12, m21 = ORBmatcher(0.7, True).SearchByBoW(d1, k1["angle"], None, fv1, d2, k2["angle"], None, empty)
    assert nm == 0 and (m12 == -1).all() and (m21 == -1).all()
