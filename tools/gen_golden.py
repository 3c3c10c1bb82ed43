"""Generate tests/golden/*.npz from the cv2-driven harness (oracle/cv2_harness.py).

Run in the build container (needs the cv2 4.13.0 wheel).  The golden files hold, for seeded synthetic
frames, the final keypoints + descriptors and per-level digests of every stage boundary, all produced by
REAL OpenCV primitives under the reference's control flow.  Both the C++ oracle (CPU tests) and the CUDA
path (GPU tests) are checked against them.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cv2_harness as H  # noqa: E402
from orb_slam_2_ros_b200 import synth  # noqa: E402

CASES = [  # (name, seed, w, h, nfeatures, nlevels)
    ("tum_640x480_s0", 0, 640, 480, 1000, 8),
    ("tum_640x480_s7", 7, 640, 480, 1000, 8),
    ("euroc_752x480_s1", 1, 752, 480, 1000, 8),
    ("kitti_1241x376_s2", 2, 1241, 376, 2000, 8),
    ("small_160x120_s3", 3, 160, 120, 300, 4),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import cv2
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, seed, w, h, nf, nl in CASES:
        img = synth.synth_frame(seed, w, h)
        kps, desc, stages = H.extract(img, nf, 1.2, nl)
        d = {"seed": seed, "w": w, "h": h, "nfeatures": nf, "nlevels": nl, "cv2_version": cv2.__version__,
             "image_sha": sha(img), "kps": kps, "desc": desc}
        for l, st in enumerate(stages):
            d["L%d_bordered_sha" % l] = sha(st["bordered"])
            d["L%d_blurred_sha" % l] = sha(st["blurred"]) if st["blurred"] is not None else ""
            d["L%d_raw" % l] = st["raw"]
            d["L%d_kps" % l] = st["level_kps"]
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, len(kps), "keypoints")


if __name__ == "__main__":
    main()
