#!/usr/bin/env python
"""Exhaustive check of parity pin (iii) on a GPU box: for EVERY fp32 bit pattern x in [0, 2*pi] the
device's (float)cos((double)x), (float)sin((double)x) — the `sincos` call of orient_describe_kernel — equal the
host libm's.  ~1.09e9 patterns, a few minutes.  Usage (on the GPU box): python tools/check_sincos_exhaustive.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import orb_oracle  # noqa: E402
from orb_slam_2_ros_b200 import _lib  # noqa: E402


def main():
    L = _lib.lib()
    hi = int(np.array([6.2831855], np.float32).view(np.uint32)[0])
    step = 1 << 24
    bad = 0
    for s in range(0, hi + 1, step):
        n = min(step, hi + 1 - s)
        a = np.zeros(n, np.float32); b = np.zeros(n, np.float32)
        _lib.check(L.orb_debug_sincos_range(0, C.c_uint32(s), n, _lib.ptr(a), _lib.ptr(b)))
        oa, ob = orb_oracle.sincos_range(s, n)
        bad += int((a.view(np.uint32) != oa.view(np.uint32)).sum() + (b.view(np.uint32) != ob.view(np.uint32)).sum())
        print("patterns %#010x..%#010x  mismatches so far: %d" % (s, s + n - 1, bad), flush=True)
    print("exhaustive sincos pin:", "OK" if bad == 0 else "%d MISMATCHES" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
