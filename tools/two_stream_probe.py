"""Does running two half-batches on two streams (two extractor contexts) beat one full batch on one stream for the device-resident
path?  Usage: [ORB_B200_CARVEOUT=58] python tools/two_stream_probe.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam_2_ros_b200 import ORBextractor, synth  # noqa: E402
from orb_slam_2_ros_b200._lib import KP_DTYPE  # noqa: E402

B, W, H = 512, 640, 480
frames = torch.from_numpy(synth.synth_batch(0, B, W, H, unique=16)).cuda()


def setup(nctx):
    per = B // nctx
    out = []
    for i in range(nctx):
        ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=per)
        st = torch.cuda.Stream()
        ex.set_stream(st.cuda_stream)
        cap = ex.max_keypoints
        k = torch.zeros((per, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
        d = torch.zeros((per, cap, 32), dtype=torch.uint8, device="cuda")
        n = torch.zeros(per, dtype=torch.int32, device="cuda")
        out.append((ex, st, k, d, n, cap, frames[i * per:(i + 1) * per]))
    return out


def run(ctxs, steps=10):
    def once():
        for ex, st, k, d, n, cap, fr in ctxs:
            ex.extract_batch_device(fr.data_ptr(), fr.shape[0], W, H, W, W * H, k.data_ptr(), d.data_ptr(), cap, n.data_ptr())
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


for nctx in (1, 2, 4):
    c = setup(nctx)
    ms = run(c)
    print("contexts/streams %d: %.3f ms per %d frames = %.0f frames/s" % (nctx, ms, B, B / ms * 1e3))
    del c
