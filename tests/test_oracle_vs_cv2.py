"""CPU: pin every OpenCV primitive restated in the oracle against the cv2 wheel (4.13.0 in this image).
Skipped when cv2 is not importable; the committed golden files (test_oracle_golden.py) carry the same pin."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from orb_slam_2_ros_b200 import synth  # noqa: E402

pytestmark = pytest.mark.skipif(not cv2.__version__.startswith("4.13"), reason="pins are for OpenCV 4.13.x")


def test_resize_chain_and_border(oracle):
    for (w, h) in [(640, 480), (752, 480), (1241, 376), (101, 77)]:
        prev = synth.synth_frame(11, w, h)
        for l in range(1, 6):
            dw, dh = int(round(w / 1.2 ** l)), int(round(h / 1.2 ** l))
            ref = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(oracle.resize_linear(prev, dw, dh), ref)
            assert np.array_equal(oracle.border_reflect101(ref, 19),
                                  cv2.copyMakeBorder(ref, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
            prev = ref


def test_gaussian_blur(oracle):
    rng = np.random.default_rng(3)
    for (w, h) in [(179, 134), (64, 48), (9, 8), (640, 480)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(oracle.gaussian_blur(img), ref)


@pytest.mark.parametrize("t", [20, 7])
def test_fast_cells(oracle, t):
    rng = np.random.default_rng(4)
    fd = cv2.FastFeatureDetector_create(t, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    img = synth.synth_frame(3)
    nk = 0
    for it in range(150):
        cw, ch = int(rng.integers(5, 45)), int(rng.integers(5, 45))
        x0, y0 = int(rng.integers(0, 640 - cw)), int(rng.integers(0, 480 - ch))
        cell = np.ascontiguousarray(img[y0:y0 + ch, x0:x0 + cw])
        if it % 3 == 0:
            cell = rng.integers(0, 256, size=(ch, cw)).astype(np.uint8)
        if it % 3 == 1:
            cell = (cell // 16 * 16).astype(np.uint8)   # quantised: forces score ties
        ref = [(k.pt[0], k.pt[1], k.response) for k in fd.detect(cell)]
        mine = [(float(k["x"]), float(k["y"]), float(k["response"])) for k in oracle.fast(cell, t, True)]
        assert ref == mine
        nk += len(ref)
    assert nk > 500


def test_fast_atan2(oracle):
    rng = np.random.default_rng(5)
    ys = rng.integers(-400000, 400000, size=20000)
    xs = rng.integers(-400000, 400000, size=20000)
    ys[:100] = 0
    xs[50:150] = 0
    ys[200:300] = xs[200:300]
    for y, x in zip(ys, xs):
        assert cv2.fastAtan2(float(y), float(x)) == oracle.fast_atan2(float(y), float(x))


def test_full_pipeline_vs_cv2_harness(oracle):
    from oracle import cv2_harness as H
    img = synth.synth_frame(21, 320, 240)
    k2, d2, _ = H.extract(img, 500, 1.2, 6)
    k1, d1 = oracle.Extractor(500, 1.2, 6).extract(img)
    assert len(k1) == len(k2)
    for f in k1.dtype.names:
        assert np.array_equal(k1[f].view(np.uint32), k2[f].view(np.uint32)), f
    assert np.array_equal(d1, d2)


def test_cvt_gray_vs_cv2(oracle):
    """cvtColor(.., *2GRAY) of Tracking::GrabImage* (Tracking.cc:181-204): the oracle's integer recipe equals cv2 4.13.0."""
    import cv2
    rng = np.random.default_rng(12)
    for ch, rgb, code in ((3, False, cv2.COLOR_BGR2GRAY), (3, True, cv2.COLOR_RGB2GRAY), (4, False, cv2.COLOR_BGRA2GRAY),
                          (4, True, cv2.COLOR_RGBA2GRAY)):
        img = rng.integers(0, 256, (97, 131, ch), dtype=np.uint8)
        img[:8, :8] = 255; img[8:16, :8] = 0                     # saturated corners
        assert np.array_equal(oracle.cvt_gray(img, rgb), cv2.cvtColor(img, code)), (ch, rgb)
