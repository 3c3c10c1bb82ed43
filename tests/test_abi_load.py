"""CPU: the C-ABI library loads and exports every symbol include/orb_b200.h declares; no compute without a GPU."""
import os
import re

import numpy as np
import pytest

from orb_slam_2_ros_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_if_missing():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def test_header_symbols_exported():
    _build_if_missing()
    hdr = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(orb_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(declared) == sorted(_lib.EXPORTS)


def test_tables_without_gpu_match_reference_constructor():
    """orb_create / orb_get_tables need no device (ORBextractor.cc:416-455)."""
    _build_if_missing()
    from orb_slam_2_ros_b200 import ORBextractor
    e = ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.mnFeaturesPerLevel.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert e.GetLevels() == 8 and abs(e.GetScaleFactor() - 1.2) < 1e-9
    assert np.allclose(e.GetScaleFactors() * e.GetInverseScaleFactors(), 1, atol=1e-6)
    assert ORBextractor(2000).mnFeaturesPerLevel.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]


def test_tables_equal_oracle(oracle):
    _build_if_missing()
    from orb_slam_2_ros_b200 import ORBextractor
    for nf, sf, nl in [(1000, 1.2, 8), (1200, 1.2, 8), (2000, 1.2, 8), (500, 1.5, 5), (3000, 1.1, 12)]:
        e = ORBextractor(nf, sf, nl)
        o = oracle.Extractor(nf, sf, nl)
        assert np.array_equal(e.mvScaleFactor.view(np.uint32), o.scale_factors.view(np.uint32))
        assert np.array_equal(e.mvInvScaleFactor.view(np.uint32), o.inv_scale_factors.view(np.uint32))
        assert np.array_equal(e.mvLevelSigma2.view(np.uint32), o.level_sigma2.view(np.uint32))
        assert np.array_equal(e.mvInvLevelSigma2.view(np.uint32), o.inv_level_sigma2.view(np.uint32))
        assert np.array_equal(e.mnFeaturesPerLevel, o.features_per_level)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails loudly (ORB_ERR_NO_DEVICE)."""
    _build_if_missing()
    if _lib.lib().orb_device_count() > 0:
        pytest.skip("a GPU is visible")
    from orb_slam_2_ros_b200 import ORBextractor, hamming_top2
    with pytest.raises(_lib.OrbError) as ei:
        ORBextractor()(np.zeros((480, 640), np.uint8))
    assert ei.value.code == _lib.ORB_ERR_NO_DEVICE
    with pytest.raises(_lib.OrbError):
        hamming_top2(np.zeros((2, 32), np.uint8), np.zeros((3, 32), np.uint8))


def test_top2_merge_host(oracle):
    """orb_top2_merge is host code: sharded results merge to the single-shard oracle answer."""
    _build_if_missing()
    from orb_slam_2_ros_b200 import top2_merge
    rng = np.random.default_rng(7)
    db = rng.integers(0, 256, (4000, 32), dtype=np.uint8)
    db[1234] = db[77]; db[3000] = db[77]
    q = np.concatenate([db[[77, 5, 3999]], rng.integers(0, 256, (20, 32), dtype=np.uint8)])
    full = oracle.hamming_top2(q, db)
    parts = np.zeros((4, len(q)), _lib.TOP2_DTYPE)
    for s in range(4):
        o = oracle.hamming_top2(q, db[s * 1000:(s + 1) * 1000])
        parts[s]["best_dist"] = o["best_dist"]; parts[s]["second_dist"] = o["second_dist"]
        parts[s]["best_idx"] = np.where(o["best_idx"] >= 0, o["best_idx"] + s * 1000, -1)
        parts[s]["second_idx"] = np.where(o["second_idx"] >= 0, o["second_idx"] + s * 1000, -1)
    for order in ([0, 1, 2, 3], [3, 1, 0, 2]):
        m = top2_merge(parts[order])
        for f in ("best_dist", "second_dist", "best_idx", "second_idx"):
            assert np.array_equal(m[f], full[f]), f
