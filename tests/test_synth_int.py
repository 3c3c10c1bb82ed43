"""the integer-only synthetic frame generator is byte-identical in C (tools/synth_int.c, built here with gcc)
and Python (synth.synth_frame_int), and produces frames with the three required regions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("synth") / "libsynth_int.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "synth_int.c")])
    L = C.CDLL(so)
    L.synth_frame_int.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.synth_frame_int.restype = None
    return L


@pytest.mark.parametrize("seed,w,h", [(0, 640, 480), (7, 752, 480), (123456, 1241, 376), (3, 160, 120)])
def test_c_and_python_generators_agree(clib, seed, w, h):
    py = synth.synth_frame_int(seed, w, h)
    c = np.zeros((h, w), np.uint8)
    clib.synth_frame_int(seed, w, h, 380, 8, c.ctypes.data_as(C.c_void_p))
    assert np.array_equal(py, c)
    # perfectly flat patch, and a low-contrast patch whose steps lie in (7, 20] before noise
    assert np.all(py[h - h // 4:, w - w // 4:] == 128)
    assert py.std() > 30


def test_oracle_extracts_from_the_integer_frame(oracle):
    img = synth.synth_frame_int(1, 640, 480)
    ex = oracle.Extractor(1000, 1.2, 8, 20, 7)
    kps, desc = ex.extract(img)
    st = ex.stats()
    assert len(kps) >= 900
    assert st[:, 3].sum() > 0          # some cells needed the minThFAST retry
