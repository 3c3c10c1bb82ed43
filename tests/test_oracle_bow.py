"""CPU checks of the bag-of-words oracle (oracle/orb_oracle_bow.cpp): a hand-computed miniature tree, an independent
pure-numpy restatement of DBoW2's transform (TemplatedVocabulary.h:1140-1272), the text format round trip and the
committed golden vectors."""
import os

import numpy as np

from orb_slam_2_ros_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bow_k6_L3.npz")


def _numpy_transform(P, L, q, levelsup, tf=True, norm="L1"):
    """Independent restatement with python loops + dict (small cases only)."""
    parent, is_leaf, desc, weight = P
    n = len(parent)
    children = [[] for _ in range(n)]
    for i in range(1, n):
        children[parent[i]].append(i)
    word_of = np.zeros(n, np.int64)
    word_of[is_leaf == 1] = np.arange(int(is_leaf.sum()))
    bits = np.unpackbits(desc, axis=1)
    bow, fv = {}, {}
    words, nodes = [], []
    for i, d in enumerate(q):
        db = np.unpackbits(d)
        cur, lvl, nid, nid_set = 0, 0, 0, (L - levelsup) <= 0
        while True:
            lvl += 1
            ch = children[cur]
            dist = [(int((bits[c] != db).sum())) for c in ch]
            cur = ch[int(np.argmin(dist))]          # argmin = first minimum
            if lvl == L - levelsup:
                nid, nid_set = cur, True
            if not children[cur]:
                break
        if not nid_set:
            nid = cur
        w = float(weight[cur])
        words.append(int(word_of[cur])); nodes.append(nid)
        if w > 0:
            wid = int(word_of[cur])
            if wid in bow:
                if tf:
                    bow[wid] += w
            else:
                bow[wid] = w
            fv.setdefault(nid, []).append(i)
    keys = sorted(bow)
    vals = [bow[k] for k in keys]
    if norm == "L1":
        s = 0.0
        for v in vals:
            s += abs(v)
        if s > 0:
            vals = [v / s for v in vals]
    return np.asarray(words), np.asarray(nodes), np.asarray(keys), np.asarray(vals), fv


def test_tiny_tree_by_hand(oracle):
    # root -> A(1), B(2); A -> a1(3), a2(4); B is a leaf.  k=2, L=2.
    parent = np.array([0, 0, 0, 1, 1], np.int32)
    leaf = np.array([0, 0, 1, 1, 1], np.uint8)
    desc = np.zeros((5, 32), np.uint8)
    desc[1] = 0x00; desc[2] = 0xFF; desc[3, 0] = 0x0F; desc[4, 0] = 0xF0
    weight = np.array([0, 0, 2.5, 1.0, 3.0])
    v = oracle.Vocabulary.from_arrays(2, 2, 0, 0, parent, leaf, desc, weight)
    assert (v.n_nodes, v.n_words) == (5, 3)
    q = np.zeros((4, 32), np.uint8)
    q[0, 0] = 0x0F          # -> A -> a1 (word 1)
    q[1, 0] = 0xF0          # -> A -> a2 (word 2)
    q[2] = 0xFF             # -> B (word 0), leaf above level 2
    q[3, 0] = 0x3C          # A; a1 and a2 tie at distance 4 -> first child a1
    w, wt, nid = v.transform_features(q, 0)
    assert w.tolist() == [1, 2, 0, 1] and wt.tolist() == [1.0, 3.0, 2.5, 1.0] and nid.tolist() == [3, 4, 2, 3]
    w, wt, nid = v.transform_features(q, 1)
    assert nid.tolist() == [1, 1, 2, 1]
    w, wt, nid = v.transform_features(q, 2)
    assert nid.tolist() == [0, 0, 0, 0]
    (bw, bv), (fn, fs, ff) = v.transform(q, 1)
    assert bw.tolist() == [0, 1, 2] and np.array_equal(bv, np.array([2.5, 2.0, 3.0]) / 7.5)
    assert fn.tolist() == [1, 2] and fs.tolist() == [0, 3, 4] and ff.tolist() == [0, 1, 3, 2]


def test_oracle_vs_numpy_restatement(oracle):
    for seed, irregular in ((1, False), (2, True)):
        P = synth.synth_vocabulary(seed, k=5, L=3, irregular=irregular, p_stop=0.1, p_dup=0.1)
        v = oracle.Vocabulary.from_arrays(5, 3, 0, 0, *P)
        rng = np.random.default_rng(seed)
        leaves = np.nonzero(P[1])[0]
        q = P[2][rng.choice(leaves, 120)].copy()
        q[:60] ^= rng.integers(0, 256, (60, 32), dtype=np.uint8) & rng.integers(0, 256, (60, 32), dtype=np.uint8) & 0x11
        for levelsup in (0, 1, 2, 4):
            words, nodes, keys, vals, fv = _numpy_transform(P, 3, q, levelsup)
            w, wt, nid = v.transform_features(q, levelsup)
            assert np.array_equal(w, words) and np.array_equal(nid, nodes)
            (bw, bv), (fn, fs, ff) = v.transform(q, levelsup)
            assert np.array_equal(bw, keys) and np.array_equal(bv, vals)
            assert fn.tolist() == sorted(fv)
            for j, node in enumerate(fn):
                assert ff[fs[j]:fs[j + 1]].tolist() == fv[int(node)]


def test_text_round_trip(oracle, tmp_path):
    P = synth.synth_vocabulary(3, k=6, L=3, irregular=True, p_stop=0.05)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, 6, 3, 0, 0, *P)
    a = oracle.Vocabulary.from_arrays(6, 3, 0, 0, *P)
    b = oracle.Vocabulary.load_text(path)
    assert (a.n_nodes, a.n_words) == (b.n_nodes, b.n_words)
    for x, y in zip(a.export(), b.export()):
        assert np.array_equal(x, y)
    # a trailing blank line does not create a phantom node (pin (v))
    with open(path, "a") as f:
        f.write("\n")
    assert oracle.Vocabulary.load_text(path).n_nodes == a.n_nodes


def test_golden(oracle):
    g = np.load(GOLD)
    v = oracle.Vocabulary.from_arrays(6, 3, 0, 0, g["parent"], g["is_leaf"], g["desc"], g["weight"])
    w, wt, nid = v.transform_features(g["q"], 2)
    assert np.array_equal(w, g["word"]) and np.array_equal(wt, g["wt"]) and np.array_equal(nid, g["node"])
    (bw, bv), (fn, fs, ff) = v.transform(g["q"], 2)
    assert np.array_equal(bw, g["bow_word"]) and np.array_equal(bv.view(np.uint64), g["bow_value"].view(np.uint64))
    assert np.array_equal(fn, g["fv_node"]) and np.array_equal(fs, g["fv_start"]) and np.array_equal(ff, g["fv_feat"])


def test_score_l1(oracle):
    a = (np.array([1, 4, 9]), np.array([0.5, 0.25, 0.25]))
    b = (np.array([4, 9, 11]), np.array([0.5, 0.1, 0.4]))
    expect = -((abs(0.25 - 0.5) - 0.25 - 0.5) + (abs(0.25 - 0.1) - 0.25 - 0.1)) / 2.0
    assert oracle.bow_score_l1(a, b) == expect
    assert oracle.bow_score_l1(a, a) == 1.0


def test_search_by_bow_single_node_equals_bruteforce(oracle):
    """With one node holding every keypoint, SearchByBoW is the brute-force inner loop (ORBmatcher.cc:196-252)."""
    rng = np.random.default_rng(3)
    d2 = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    d1 = d2[rng.permutation(300)[:200]].copy()
    d1[:, 0] ^= rng.integers(0, 4, 200, dtype=np.uint8)
    a1 = rng.uniform(0, 360, 200).astype(np.float32); a2 = rng.uniform(0, 360, 300).astype(np.float32)
    fv1 = (np.array([7]), np.array([0, 200]), np.arange(200))
    fv2 = (np.array([7]), np.array([0, 300]), np.arange(300))
    m12, m21, nm = oracle.search_by_bow(d1, a1, None, fv1, d2, a2, None, fv2, 50, False, 0.6, True)
    nm2, b12 = oracle.match_bruteforce(d1, a1, d2, a2, 50, 0.6, True)
    assert nm == nm2 and np.array_equal(m12, b12)
    assert all(m21[m12[i]] == i for i in range(200) if m12[i] >= 0)
    # disjoint node ids: nothing is compared
    fv2b = (np.array([8]), np.array([0, 300]), np.arange(300))
    assert oracle.search_by_bow(d1, a1, None, fv1, d2, a2, None, fv2b)[2] == 0
