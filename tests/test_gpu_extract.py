"""GPU parity: the CUDA extractor (through the C ABI) against the CPU oracle and the committed golden vectors.
Bit-exact on keypoint coordinates, octaves, responses, sizes, descriptors, pyramid, blur and raw corners;
the float angle is compared bit-exactly too (tolerance allowed by north_star: 1e-4 rad)."""
import glob
import hashlib
import os

import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("bow_"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_kps_equal(a, b, what=""):
    assert len(a) == len(b), "%s count %d vs %d" % (what, len(a), len(b))
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), "%s field %s" % (what, f)
    ang_bad = int((a["angle"].view(np.uint32) != b["angle"].view(np.uint32)).sum())
    if ang_bad:   # report, then apply the stated tolerance (1e-4 rad)
        d = np.abs(a["angle"].astype(np.float64) - b["angle"].astype(np.float64))
        d = np.minimum(d, 360 - d)
        assert np.deg2rad(d.max()) < 1e-4, "%s angle off by %g deg" % (what, d.max())
    return ang_bad


@pytest.fixture(scope="module")
def ORB():
    from orb_slam_2_ros_b200 import ORBextractor
    return ORBextractor


CASES = [(0, 640, 480, 1000, 8), (1, 640, 480, 1000, 8), (2, 640, 480, 1000, 8), (3, 752, 480, 1000, 8),
         (4, 1241, 376, 2000, 8), (5, 640, 480, 1200, 8), (6, 320, 240, 500, 6), (7, 160, 120, 300, 4),
         (8, 645, 487, 1000, 8), (9, 1920, 1080, 3000, 8), (10, 1280, 720, 1500, 8), (11, 1023, 767, 2000, 10)]


@pytest.mark.parametrize("seed,w,h,nf,nl", CASES)
def test_extract_matches_oracle_all_stages(oracle, ORB, seed, w, h, nf, nl):
    img = synth.synth_frame(seed, w, h)
    ex = ORB(nf, 1.2, nl, 20, 7)
    kps, desc = ex(img)
    oex = oracle.Extractor(nf, 1.2, nl, 20, 7)
    okps, odesc = oex.extract(img)
    # stage by stage first, so a failure points at the first wrong stage
    for l in range(nl):
        assert ex.level_dims(l) == oex.level_dims(l)
        assert np.array_equal(ex.pyramid_level_bordered(l), oex.level(l)), "pyramid level %d" % l
    for l in range(nl):
        raw = ex.debug_raw_corners(l)
        oraw = oex.raw_corners(l)
        oraw = np.stack([oraw["x"], oraw["y"], oraw["response"]], 1) if len(oraw) else np.zeros((0, 3), np.float32)
        assert raw.shape == oraw.shape, "raw corner count level %d: %s vs %s" % (l, raw.shape, oraw.shape)
        assert np.array_equal(raw, oraw), "raw corners level %d" % l
    for l in range(nl):
        ob = oex.blurred(l)
        if ob is not None:
            assert np.array_equal(ex.debug_blurred(l), ob), "blur level %d" % l
    assert_kps_equal(kps, okps, "final")
    flips = int(np.unpackbits(desc ^ odesc).sum())
    assert flips == 0, "%d descriptor bit flips" % flips
    assert np.array_equal(ex.debug_tie_counts(), oex.stats()[:, 2])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_extract_matches_golden(ORB, path):
    g = np.load(path)
    w, h, nf, nl = int(g["w"]), int(g["h"]), int(g["nfeatures"]), int(g["nlevels"])
    img = synth.synth_frame(int(g["seed"]), w, h)
    assert sha(img) == str(g["image_sha"])
    ex = ORB(nf, 1.2, nl, 20, 7)
    kps, desc = ex(img)
    assert_kps_equal(kps, g["kps"], "golden")
    assert np.array_equal(desc, g["desc"])
    for l in range(nl):
        assert sha(ex.pyramid_level_bordered(l)) == str(g["L%d_bordered_sha" % l])
        if str(g["L%d_blurred_sha" % l]):
            assert sha(ex.debug_blurred(l)) == str(g["L%d_blurred_sha" % l])


def test_batch_equals_single_and_strided_input(oracle, ORB):
    imgs = synth.synth_batch(100, 6, 640, 480, unique=3)
    exb = ORB(1000, max_batch=4)      # 6 frames through a 4-frame arena: two chunks
    res = exb.extract_batch(imgs)
    oex = oracle.Extractor(1000)
    for f in range(6):
        ok, od = oex.extract(imgs[f])
        assert_kps_equal(res[f][0], ok, "frame %d" % f)
        assert np.array_equal(res[f][1], od)
    # a non-contiguous (strided) single image, like a cv::Mat ROI
    big = np.zeros((500, 700), np.uint8)
    big[10:490, 30:670] = imgs[1]
    k, d = exb(big[10:490, 30:670])
    ok, od = oex.extract(imgs[1])
    assert_kps_equal(k, ok, "strided")
    assert np.array_equal(d, od)


def test_pipelined_batch_chunks_pinned_and_pageable(oracle, ORB):
    """orb_extract_batch cuts a batch into 64-frame chunks flowing through H2D / kernel / D2H streams: 150 frames
    through a 100-frame arena = chunks of 64 + 36, then 50; once with pageable numpy buffers (staged through the
    context's pinned mirrors) and once with pinned buffers (direct DMA, strided 2-D copies into the caller's rows)."""
    import ctypes as C
    import torch
    from orb_slam_2_ros_b200 import _lib
    w, h, n = 320, 240, 150
    imgs = synth.synth_batch(300, n, w, h, unique=5)
    ex = ORB(500, 1.2, 6, 20, 7, max_batch=100)
    oex = oracle.Extractor(500, 1.2, 6, 20, 7)
    want = [oex.extract(imgs[f]) for f in range(5)]        # frames >= 5 are circular shifts: check those through frame 0..4 too
    res = ex.extract_batch(imgs)
    assert len(res) == n
    for f in range(5):
        assert_kps_equal(res[f][0], want[f][0], "pageable frame %d" % f)
        assert np.array_equal(res[f][1], want[f][1])
    # every frame against single-frame extraction on the same context (same kernels, different arena slots)
    for f in (5, 63, 64, 99, 100, 149):
        k1, d1 = ex(imgs[f])
        assert_kps_equal(res[f][0], k1, "frame %d" % f)
        assert np.array_equal(res[f][1], d1)
    # pinned in / out with a caller capacity larger than the device row length
    cap = ex.max_keypoints + 37
    t_in = torch.from_numpy(imgs).pin_memory()
    t_k = torch.zeros((n, cap, _lib.KP_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    t_d = torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory()
    n_out = np.zeros(n, np.int32)
    _lib.check(_lib.lib().orb_extract_batch(ex._h, C.c_void_p(t_in.data_ptr()), n, w, h, w, w * h, C.c_void_p(t_k.data_ptr()),
                                            C.c_void_p(t_d.data_ptr()), cap, _lib.ptr(n_out)))
    kp = t_k.numpy().view(_lib.KP_DTYPE).reshape(n, cap)
    for f in range(n):
        assert n_out[f] == len(res[f][0])
        assert np.array_equal(kp[f, :n_out[f]].view(np.uint8), np.ascontiguousarray(res[f][0]).view(np.uint8)), "pinned frame %d" % f
        assert np.array_equal(t_d.numpy()[f, :n_out[f]], res[f][1])


def test_edge_cases(oracle, ORB):
    ex = ORB(500)
    k, d = ex(np.zeros((0, 0), np.uint8))                 # empty image => silent return (ORBextractor.cc:1086)
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.full((480, 640), 90, np.uint8))          # flat image: no corner at any threshold
    assert len(k) == 0
    # image size change on the same context (arena rebuild)
    img = synth.synth_frame(9, 320, 240)
    k, d = ex(img)
    ok, od = oracle.Extractor(500).extract(img)
    assert_kps_equal(k, ok, "resized ctx")
    assert np.array_equal(d, od)
    # too small for the reference's cell grid: explicit error instead of the reference's division by zero
    from orb_slam_2_ros_b200 import _lib
    with pytest.raises(_lib.OrbError) as ei:
        ex(synth.synth_frame(1, 160, 120))
    assert ei.value.code == _lib.ORB_ERR_TOO_SMALL
    # mvImagePyramid is public API (ORBextractor.h:85): interior ROI with step w+38
    ex(synth.synth_frame(0))
    pyr = ex.mvImagePyramid
    assert pyr[0].shape == (480, 640) and pyr[0].strides[0] == 640 + 38
    assert np.array_equal(pyr[0], synth.synth_frame(0))


def test_noise_image_many_corners(oracle, ORB):
    """Pure noise: tens of thousands of raw corners per level, quota-limited everywhere (deep quadtree)."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    k, d = ORB(1000)(img)
    ok, od = oracle.Extractor(1000).extract(img)
    assert_kps_equal(k, ok, "noise")
    assert np.array_equal(d, od)


def test_sincos_pin_sampled(oracle):
    """pin (iii): (float)cos((double)x), (float)sin((double)x) on the device equal the host libm on sampled
    ranges of fp32 bit patterns in [0, 2*pi] (the exhaustive sweep is tools/check_sincos_exhaustive.py)."""
    import ctypes as C
    from orb_slam_2_ros_b200 import _lib
    L = _lib.lib()
    hi = np.array([6.2831855], np.float32).view(np.uint32)[0]
    rng = np.random.default_rng(0)
    starts = [0, 1, int(hi) - 200000] + [int(s) for s in rng.integers(0x30000000, int(hi) - 200000, size=12)]
    for s in starts:
        n = 200000
        a = np.zeros(n, np.float32); b = np.zeros(n, np.float32)
        _lib.check(L.orb_debug_sincos_range(0, C.c_uint32(s), n, _lib.ptr(a), _lib.ptr(b)))
        oa, ob = oracle.sincos_range(s, n)
        assert np.array_equal(a.view(np.uint32), oa.view(np.uint32)), "cos mismatch in range %x" % s
        assert np.array_equal(b.view(np.uint32), ob.view(np.uint32)), "sin mismatch in range %x" % s


def test_config4_shape_large_batch_properties(oracle, ORB):
    """BASELINE config 4 shape (752x480 EuRoC frames, nFeatures=1000) at a batch that spans several pipeline chunks:
    sampled frames equal the oracle, duplicated frames give identical results wherever they sit in the batch, counts stay
    within the reference's N .. N + 3 per level bound, and descriptors of a frame never depend on its neighbours."""
    w, h, n = 752, 480, 200
    base = [synth.synth_frame(500 + i, w, h) for i in range(4)]
    idx = np.random.default_rng(1).integers(0, 4, n)
    idx[:4] = [0, 1, 2, 3]
    imgs = np.stack([base[i] for i in idx])
    ex = ORB(1000, max_batch=128)                          # 200 frames through a 128-frame arena: 64 + 64, then 64 + 8
    res = ex.extract_batch(imgs)
    oex = oracle.Extractor(1000)
    want = [oex.extract(b) for b in base]
    for f in range(n):
        k, d = res[f]
        ok, od = want[idx[f]]
        assert len(k) == len(ok) and 1000 <= len(k) <= 1000 + 3 * 8, "frame %d: %d keypoints" % (f, len(k))
        assert np.array_equal(np.ascontiguousarray(k).view(np.uint8), np.ascontiguousarray(ok).view(np.uint8)), "frame %d" % f
        assert np.array_equal(d, od), "frame %d" % f


@pytest.mark.parametrize("ch,rgb", [(3, False), (3, True), (4, False), (4, True)])
def test_colour_input_fused_cvtcolor(oracle, ORB, ch, rgb):
    """the camera frame before Tracking's cvtColor goes straight to the extractor; level 0 of the pyramid
    must equal cvtColor's gray image (OpenCV 4.13.0 arithmetic) and everything downstream the oracle on that gray image."""
    rng = np.random.default_rng(40 + ch + rgb)
    gray = synth.synth_frame(60 + ch, 640, 480).astype(np.int32)
    col = np.stack([np.clip(gray + rng.integers(-25, 26, gray.shape), 0, 255) for _ in range(ch)], 2).astype(np.uint8)
    want_gray = oracle.cvt_gray(col, rgb)
    ex = ORB(1000)
    k, d = ex(col, rgb=rgb)
    assert np.array_equal(ex.mvImagePyramid[0], want_gray)
    ok, od = oracle.Extractor(1000).extract(want_gray)
    assert_kps_equal(k, ok, "colour %d %s" % (ch, rgb))
    assert np.array_equal(d, od)


@pytest.mark.parametrize("nf,sf,nl,ini,mn,w,h", [
    (1000, 1.5, 5, 20, 7, 640, 480),      # scale factor > 4/3: the generic resize kernel
    (800, 2.0, 3, 20, 7, 752, 480),       # octave pyramid
    (1200, 1.1, 10, 25, 10, 640, 480),    # ten shallow levels, other FAST thresholds
    (5000, 1.2, 8, 20, 7, 1241, 376),     # large quota: quadtree shared memory above 48 KB
    (300, 1.2, 2, 40, 5, 400, 300),       # few features, wide threshold gap
    (1000, 1.2, 8, 12, 12, 640, 480),     # iniThFAST == minThFAST (the retry never changes anything)
])
def test_other_extractor_parameters(oracle, ORB, nf, sf, nl, ini, mn, w, h):
    """Parameters outside the TUM / KITTI / EuRoC defaults: every code path that depends on them (generic resize,
    level count, quotas, thresholds, quadtree capacity) against the oracle."""
    img = synth.synth_frame(70 + nl, w, h)
    ex = ORB(nf, sf, nl, ini, mn)
    oex = oracle.Extractor(nf, sf, nl, ini, mn)
    k, d = ex(img)
    ok, od = oex.extract(img)
    for l in range(nl):
        assert np.array_equal(ex.pyramid_level_bordered(l), oex.level(l)), "pyramid level %d" % l
    assert_kps_equal(k, ok, "params %s" % ((nf, sf, nl, ini, mn),))
    assert np.array_equal(d, od)
    # the same through the throughput shapes (8-row resize threads, 256-thread quadtree CTAs): a batch of 9 frames
    exb = ORB(nf, sf, nl, ini, mn, max_batch=9)
    res = exb.extract_batch(np.stack([img] * 9))
    for f in (0, 8):
        assert_kps_equal(res[f][0], ok, "batched frame %d" % f)
        assert np.array_equal(res[f][1], od)


@pytest.mark.parametrize("w,h,pad,shift", [(640, 480, 0, 0), (640, 480, 4, 4), (640, 480, 3, 1), (333, 251, 0, 0), (333, 251, 5, 2), (334, 250, 2, 0)])
def test_device_resident_input_alignments(oracle, ORB, w, h, pad, shift):
    """orb_extract_batch_device on caller-owned device frames: the level-0 pass reads 16-byte aligned rows as vectors, 4-byte aligned
    rows as words and anything else byte by byte, and builds the 19-px border from the same rows — all three must give the oracle's
    bordered level 0 and keypoints (row stride = w + pad, first pixel `shift` bytes into the allocation)."""
    import torch
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    F = 3
    imgs = synth.synth_batch(300 + w + pad, F, w, h, unique=F)
    stride = w + pad
    buf = torch.zeros(shift + F * h * stride + 64, dtype=torch.uint8, device="cuda")
    view = buf[shift:shift + F * h * stride].view(F, h, stride)
    view[:, :, :w] = torch.from_numpy(imgs).cuda()
    ex = ORB(500, max_batch=F)
    cap = ex.max_keypoints
    d_kps = torch.zeros((F, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((F, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(buf.data_ptr() + shift, F, w, h, stride, h * stride, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
    ex.sync()
    n = d_n.cpu().numpy()
    oex = oracle.Extractor(500)
    for f in range(F):
        ok, od = oex.extract(imgs[f])
        kps = d_kps[f, :n[f]].cpu().numpy().view(KP_DTYPE).reshape(-1)
        assert_kps_equal(kps, ok, "frame %d" % f)
        assert np.array_equal(d_desc[f, :n[f]].cpu().numpy(), od)
    # frame slot 0 of the arena still holds frame 0's pyramid: compare the bordered levels (the fused border pass)
    oex.extract(imgs[0])
    for l in range(8):
        assert np.array_equal(ex.pyramid_level_bordered(l), oex.level(l)), "bordered level %d" % l


@pytest.mark.parametrize("w,h,nf", [(1920, 1080, 2000), (645, 487, 800), (333, 251, 400), (1241, 376, 1500), (2600, 300, 1500)])
def test_batch_throughput_kernels_odd_and_wide_shapes(oracle, ORB, w, h, nf):
    """Batches of >= 8 frames take the throughput kernels (pyramid levels staged by bulk copies, one CTA per 16 rows x a segment
    of <= 256 bordered words): wide images need several segments per row, odd sizes end in partial words / partial strips.
    Every frame of the batch must give the oracle's keypoints and descriptors, frame 0 also its bordered and blurred levels."""
    F = 9
    imgs = synth.synth_batch(700 + w, F, w, h, unique=3)
    ex = ORB(nf, max_batch=F)
    res = ex.extract_batch(imgs)
    oex = oracle.Extractor(nf)
    for f in (1, 4, 8, 0):          # frame 0 last: the oracle then holds frame 0's levels
        ok, od = oex.extract(imgs[f])
        assert_kps_equal(res[f][0], ok, "frame %d" % f)
        assert np.array_equal(res[f][1], od)
    for l in range(8):
        assert np.array_equal(ex.pyramid_level_bordered(l), oex.level(l)), "bordered level %d" % l
        ob = oex.blurred(l)
        if ob is not None:
            assert np.array_equal(ex.debug_blurred(l), ob), "blur level %d" % l


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_fast_threshold_modes_give_identical_corners(oracle, ORB, mode, monkeypatch):
    """The FAST kernel evaluates the quick test at iniThFAST only (0: empty cells are walked again at minThFAST), at both thresholds
    in one walk (1), or per strip position by the emptiness seen by the previous frame (2, the default): all three must give the
    oracle's raw corner lists and keypoints — on the headline scene, on a sparse scene (most cells empty at iniThFAST), on flat
    and low-contrast frames, and when the frames of a batch alternate between dense and sparse (stale hints)."""
    monkeypatch.setenv("ORB_B200_FAST_DUAL", mode)
    dense = synth.synth_batch(900, 3, 640, 480, unique=3)
    sparse = synth.synth_batch(910, 3, 640, 480, unique=3, noise=2, n_rect=40, n_tri=20)
    flat = np.full((480, 640), 90, np.uint8); flat[200:215, 300:330] = 110      # one low-contrast blob: found at minThFAST only
    imgs = np.stack([dense[0], sparse[0], flat, dense[1], sparse[1], dense[2], sparse[2], flat, dense[0], sparse[0]])
    for ini, mn in ((20, 7), (12, 12), (40, 3)):
        ex = ORB(1000, 1.2, 8, ini, mn, max_batch=len(imgs))
        oex = oracle.Extractor(1000, 1.2, 8, ini, mn)
        for rep in range(2):                                # second round: hints left by the last frames of the first
            res = ex.extract_batch(imgs)
            for f in (0, 1, 2, 4, 7, 9):
                ok, od = oex.extract(imgs[f])
                assert_kps_equal(res[f][0], ok, "mode %s th %d/%d frame %d" % (mode, ini, mn, f))
                assert np.array_equal(res[f][1], od)
        for img in (sparse[1], dense[2], flat):            # single frames: raw corner lists level by level
            k, d = ex(img)
            ok, od = oex.extract(img)
            assert_kps_equal(k, ok, "single")
            for l in range(8):
                raw, oraw = ex.debug_raw_corners(l), oex.raw_corners(l)
                oraw = np.stack([oraw["x"], oraw["y"], oraw["response"]], 1) if len(oraw) else np.zeros((0, 3), np.float32)
                assert np.array_equal(raw, oraw), "raw corners level %d" % l
