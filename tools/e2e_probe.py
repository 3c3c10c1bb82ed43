#!/usr/bin/env python
"""Why is e2e (134 k frames/s) below both the PCIe bound (172 k) and the device-resident rate (180 k)?
A: device-resident extraction alone; B: the same while another stream copies host->device continuously;
C: the same while both copy directions run; D: two contexts extracting concurrently (device-resident, 256 frames each)."""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam_2_ros_b200 import ORBextractor, synth
from orb_slam_2_ros_b200._lib import KP_DTYPE

W, H, B = 640, 480, 512
dev = torch.device("cuda", 0)
frames = torch.from_numpy(synth.synth_batch(0, B, W, H, unique=16, noise=8)).pin_memory()
d_frames = frames.to(dev)


def make(batch):
    ex = ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_batch=batch)
    cap = ex.max_keypoints
    st = torch.cuda.Stream(dev)
    ex.set_stream(st.cuda_stream)
    return ex, st, torch.zeros((batch, cap, 28), dtype=torch.uint8, device=dev), torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev), \
        torch.zeros(batch, dtype=torch.int32, device=dev), cap


def run(ex, st, k, d, n, cap, batch, steps, off=0):
    for _ in range(steps):
        ex.extract_batch_device(d_frames.data_ptr() + off * W * H, batch, W, H, W, W * H, k.data_ptr(), d.data_ptr(), cap, n.data_ptr())
    st.synchronize()


stop = False


def copier(h2d, d2h):
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dst = torch.empty_like(d_frames)
    ho = torch.empty(34_000_000, dtype=torch.uint8).pin_memory(); do = torch.empty(34_000_000, dtype=torch.uint8, device=dev)
    while not stop:
        if h2d:
            with torch.cuda.stream(s1):
                dst.copy_(frames, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
        s1.synchronize(); s2.synchronize()


a = make(B)
run(*a, B, 3)
for name, h2d, d2h in (("A alone", 0, 0), ("B + H2D", 1, 0), ("C + H2D + D2H", 1, 1)):
    stop = False
    th = threading.Thread(target=copier, args=(h2d, d2h)) if (h2d or d2h) else None
    if th:
        th.start(); time.sleep(0.2)
    t = time.perf_counter(); run(*a, B, 10); dt = time.perf_counter() - t
    stop = True
    if th:
        th.join()
    print("%-16s %.0f frames/s" % (name, 10 * B / dt))
b1, b2 = make(256), make(256)
run(*b1, 256, 2); run(*b2, 256, 2, 256)
t = time.perf_counter()
t1 = threading.Thread(target=run, args=(*b1, 256, 10)); t2 = threading.Thread(target=run, args=(*b2, 256, 10, 256))
t1.start(); t2.start(); t1.join(); t2.join()
print("%-16s %.0f frames/s" % ("D 2 ctx x 256", 20 * 256 / (time.perf_counter() - t)))
c = make(64)
run(*c, 64, 3)
t = time.perf_counter(); run(*c, 64, 40); print("%-16s %.0f frames/s" % ("E 1 ctx x 64", 40 * 64 / (time.perf_counter() - t)))
cs = [make(64) for _ in range(4)]
for i, x in enumerate(cs):
    run(*x, 64, 2, 64 * i)
t = time.perf_counter()
ths = [threading.Thread(target=run, args=(*x, 64, 20, 64 * i)) for i, x in enumerate(cs)]
[x.start() for x in ths]; [x.join() for x in ths]
print("%-16s %.0f frames/s" % ("F 4 ctx x 64", 4 * 20 * 64 / (time.perf_counter() - t)))
