"""The C++ host mirror (orb_slam_2_ros_b200/host: ORB_SLAM2::ORBextractor, ORBmatcher, ComputeStereoMatches with the
reference's signatures over the C ABI) driven by a C++ program the way Frame.cc drives the reference classes —
two extractor instances on two std::threads for stereo (Frame.cc:79-82) — and compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth
from orb_slam_2_ros_b200._lib import KP_DTYPE

pytestmark = pytest.mark.gpu
LIBDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orb_slam_2_ros_b200", "lib")
EXE = os.path.join(LIBDIR, "host_check")


def _read(path, w, h, stereo):
    raw = open(path, "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype, count, off)
        off += a.nbytes
        return a
    n = int(take(np.int32, 1)[0])
    kps = take(KP_DTYPE, n); desc = take(np.uint8, n * 32).reshape(n, 32)
    step = int(take(np.int32, 1)[0]); lvl0 = take(np.uint8, w * h).reshape(h, w)
    out = dict(kps=kps, desc=desc, step=step, lvl0=lvl0)
    if stereo:
        nr = int(take(np.int32, 1)[0])
        out["kps_r"] = take(KP_DTYPE, nr); out["desc_r"] = take(np.uint8, nr * 32).reshape(nr, 32)
        out["nmatches"] = int(take(np.int32, 1)[0])
        out["u_right"] = take(np.float32, n); out["depth"] = take(np.float32, n)
    assert off == len(raw)
    return out


def test_cpp_extractor_shim_matches_oracle(oracle, tmp_path):
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    w, h = 640, 480
    img = synth.synth_frame(21, w, h)
    (tmp_path / "l.raw").write_bytes(img.tobytes())
    subprocess.check_call([EXE, str(w), str(h), "1000", "8", str(tmp_path / "l.raw"), str(tmp_path / "o.bin")])
    got = _read(tmp_path / "o.bin", w, h, False)
    okps, odesc = oracle.Extractor(1000, 1.2, 8, 20, 7).extract(img)
    assert len(got["kps"]) == len(okps)
    assert np.array_equal(np.ascontiguousarray(got["kps"]).view(np.uint8), np.ascontiguousarray(okps).view(np.uint8))
    assert np.array_equal(got["desc"], odesc)
    assert got["step"] == w + 38 and np.array_equal(got["lvl0"], img)       # mvImagePyramid[0]: ROI with step w+38


def test_cpp_stereo_shim_matches_oracle(oracle, tmp_path):
    w, h = 1241, 376
    left, right, _ = synth.synth_stereo_pair(4, w, h)
    (tmp_path / "l.raw").write_bytes(left.tobytes()); (tmp_path / "r.raw").write_bytes(right.tobytes())
    bf, b = 386.1448, 0.53716
    subprocess.check_call([EXE, str(w), str(h), "2000", "8", str(tmp_path / "l.raw"), str(tmp_path / "o.bin"),
                           str(tmp_path / "r.raw"), repr(bf), repr(b)])
    got = _read(tmp_path / "o.bin", w, h, True)
    oL, oR = oracle.Extractor(2000, 1.2, 8, 20, 7), oracle.Extractor(2000, 1.2, 8, 20, 7)
    kl, dl = oL.extract(left); kr, dr = oR.extract(right)
    assert np.array_equal(np.ascontiguousarray(got["kps"]).view(np.uint8), np.ascontiguousarray(kl).view(np.uint8))
    assert np.array_equal(np.ascontiguousarray(got["kps_r"]).view(np.uint8), np.ascontiguousarray(kr).view(np.uint8))
    assert np.array_equal(got["desc"], dl) and np.array_equal(got["desc_r"], dr)
    kept, ur, depth, _sad = oracle.stereo_match(oL, oR, kl, dl, kr, dr, np.float32(bf), np.float32(b))
    assert got["nmatches"] == kept and kept > 100
    assert np.array_equal(got["u_right"].view(np.uint32), ur.view(np.uint32))
    assert np.array_equal(got["depth"].view(np.uint32), depth.view(np.uint32))


def test_cpp_vocabulary_and_search_by_bow(oracle, tmp_path):
    """ORBVocabulary::loadFromTextFile + transform (Frame::ComputeBoW) + ORBmatcher::SearchByBoW through the C++ mirror."""
    w, h = 640, 480
    a = synth.synth_frame(31, w, h)
    b = np.roll(a, (1, 2), axis=(0, 1))
    P = synth.synth_vocabulary(5, k=10, L=5)
    voc_path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(voc_path, 10, 5, 0, 0, *P)
    (tmp_path / "a.raw").write_bytes(a.tobytes()); (tmp_path / "b.raw").write_bytes(b.tobytes())
    subprocess.check_call([EXE, "bow", voc_path, str(w), str(h), str(tmp_path / "a.raw"), str(tmp_path / "b.raw"), str(tmp_path / "o.bin")])
    raw = open(tmp_path / "o.bin", "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        x = np.frombuffer(raw, dtype, count, off)
        off += x.nbytes
        return x
    ov = oracle.Vocabulary.load_text(voc_path)
    ex = oracle.Extractor(1000, 1.2, 8, 20, 7)
    frames = []
    for img in (a, b):
        okps, odesc = ex.extract(img)
        n = int(take(np.int32, 1)[0]); desc = take(np.uint8, n * 32).reshape(n, 32)
        assert np.array_equal(desc, odesc)
        (obw, obv), (ofn, ofs, off_) = ov.transform(odesc, 4)
        nb = int(take(np.int32, 1)[0])
        rec = take(np.dtype([("w", "<u4"), ("v", "<f8")]), nb)
        assert np.array_equal(rec["w"], obw) and np.array_equal(rec["v"].view(np.uint64), obv.view(np.uint64))
        nf = int(take(np.int32, 1)[0])
        assert nf == len(ofn)
        for j in range(nf):
            node, cnt = take(np.uint32, 1)[0], int(take(np.int32, 1)[0])
            feats = take(np.uint32, cnt)
            assert node == ofn[j] and np.array_equal(feats, off_[ofs[j]:ofs[j + 1]])
        frames.append((okps, odesc, (ofn, ofs, off_)))
    nm = int(take(np.int32, 1)[0])
    m12 = take(np.int32, len(frames[0][1])); m21 = take(np.int32, len(frames[1][1]))
    assert off == len(raw)
    o12, o21, onm = oracle.search_by_bow(frames[0][1], frames[0][0]["angle"], None, frames[0][2], frames[1][1], frames[1][0]["angle"], None,
                                         frames[1][2], 50, False, np.float32(0.7), True)
    assert nm == onm and nm > 100 and np.array_equal(m12, o12) and np.array_equal(m21, o21)
