"""Randomised parity campaign: the CUDA extractor (C ABI) against the CPU oracle over random image shapes, extractor
parameters, batch sizes and pathological image contents (noise, saturated blocks, checkerboards, flat frames, low contrast).

    python tools/fuzz_parity.py --seconds 240 --seed 1 [--log gpurun_out/fuzz.log]

Every case compares keypoints (raw 32-bit fields) and descriptors of every frame of the batch; a mismatch prints the
case's parameters (reproducible from --seed and the case number) and the first differing stage.  Exit code = failures."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam_2_ros_b200 import ORBextractor, synth   # noqa: E402
from oracle import orb_oracle                           # noqa: E402


def make_image(rng, kind, w, h):
    if kind == "synth":
        return synth.synth_frame(int(rng.integers(1 << 30)), w, h)
    if kind == "noise":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "flat":
        img = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        for _ in range(int(rng.integers(0, 4))):     # a few isolated blobs: most cells empty at both thresholds
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            img[y:y + int(rng.integers(1, 9)), x:x + int(rng.integers(1, 9))] = int(rng.integers(0, 256))
        return img
    if kind == "checker":
        p, q = int(rng.integers(1, 10)), int(rng.integers(1, 10))
        lo, hi = sorted(int(v) for v in rng.integers(0, 256, 2))
        yy, xx = np.mgrid[0:h, 0:w]
        return np.where(((xx // p) + (yy // q)) & 1, hi, lo).astype(np.uint8)
    if kind == "saturated":
        img = np.where(rng.integers(0, 2, (h // 8 + 1, w // 8 + 1)) > 0, 255, 0).astype(np.uint8)
        img = np.kron(img, np.ones((8, 8), np.uint8))[:h, :w]
        n = rng.integers(0, int(rng.integers(1, 40)), (h, w))
        return np.where(img > 0, 255 - n, n).astype(np.uint8)
    if kind == "lowcontrast":
        base = int(rng.integers(0, 230))
        step = int(rng.integers(5, 24))
        img = base + step * rng.integers(0, 2, (h // 5 + 1, w // 5 + 1))
        img = np.kron(img, np.ones((5, 5), np.int64))[:h, :w] + rng.integers(0, 3, (h, w))
        return np.clip(img, 0, 255).astype(np.uint8)
    if kind == "saltpepper":
        img = synth.synth_frame(int(rng.integers(1 << 30)), w, h).copy()
        m = rng.random((h, w))
        img[m < 0.02] = 0
        img[m > 0.98] = 255
        return img
    raise ValueError(kind)


KINDS = ["synth", "synth", "noise", "flat", "checker", "saturated", "lowcontrast", "saltpepper"]


def one_case(rng, case, log):
    scale = float(rng.choice([1.1, 1.2, 1.2, 1.2, 1.25, 1.3, 1.5, 2.0]))
    nl = int(rng.integers(1, 11))
    # the smallest level must hold one 30-px cell inside the 16-px margins and be no taller than 2x its width
    while 70 * scale ** (nl - 1) > 1200:
        nl -= 1
    top = scale ** (nl - 1)
    wmin = int(np.ceil(70 * top)) + 2
    w = int(rng.integers(wmin, max(wmin + 1, 2300)))
    if rng.random() < 0.5:
        w = int(rng.integers(wmin, max(wmin + 1, 800)))
    h = int(rng.integers(wmin, max(wmin + 1, min(1300, int(1.9 * w)))))
    if w * h > 2200 * 1200:
        h = max(wmin, 2200 * 1200 // w)
    nf = int(rng.choice([1, 7, 100, 500, 1000, 1000, 2000, 4000]))
    ini = int(rng.integers(2, 60))
    mn = int(rng.integers(1, ini + 1))
    F = int(rng.choice([1, 1, 2, 9, 17]))
    if w * h * F > 12_000_000:
        F = 1
    kind = KINDS[int(rng.integers(len(KINDS)))]
    desc = "case %d: %s %dx%d F=%d nf=%d scale=%.2f levels=%d ini=%d min=%d" % (case, kind, w, h, F, nf, scale, nl, ini, mn)
    imgs = np.stack([make_image(rng, kind, w, h) for _ in range(min(F, 3))])
    if F > 3:
        imgs = np.concatenate([imgs] + [np.roll(imgs[k % 3], 3 * k + 1, axis=1)[None] for k in range(F - 3)])
    try:
        ex = ORBextractor(nf, scale, nl, ini, mn, max_batch=F)
    except Exception as e:      # geometry refused: the oracle must refuse as well (or the limit is documented)
        log("%s -> refused by orb_create: %s" % (desc, str(e)[:100]))
        return 0
    try:
        res = ex.extract_batch(imgs) if F > 1 else [ex(imgs[0])]
    except Exception as e:
        log("%s -> refused at extract: %s" % (desc, str(e)[:120]))
        return 0
    oex = orb_oracle.Extractor(nf, scale, nl, ini, mn)
    check = range(F) if F <= 3 else [0, 1, 2, F - 1, F // 2]
    nk = 0
    for f in check:
        ok, od = oex.extract(imgs[f])
        k, d = res[f]
        nk += len(ok)
        bad = len(k) != len(ok) or k.tobytes() != ok.tobytes() or not np.array_equal(d, od)
        if bad:
            where = "count %d vs %d" % (len(k), len(ok))
            if len(k) == len(ok):
                fields = [n for n in k.dtype.names if not np.array_equal(k[n].view(np.uint32), ok[n].view(np.uint32))]
                where = "fields %s, descriptor rows differing %d" % (fields, int((d != od).any(1).sum()))
            log("MISMATCH %s frame %d: %s" % (desc, f, where))
            return 1
    log("ok   %s (%d keypoints checked)" % (desc, nk))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=1 << 30)
    ap.add_argument("--only", type=int, default=-1, help="run just this case number of the seed")
    ap.add_argument("--log", default="")
    a = ap.parse_args()
    orb_oracle.build()
    fh = open(a.log, "w") if a.log else None

    def log(s):
        print(s, flush=True)
        if fh:
            fh.write(s + "\n"); fh.flush()

    t0 = time.time()
    fails = n = 0
    case = 0
    while case < a.cases and (time.time() - t0 < a.seconds or a.only >= 0):
        rng = np.random.default_rng([a.seed, case])
        if a.only < 0 or case == a.only:
            fails += one_case(rng, case, log)
            n += 1
        if a.only >= 0 and case >= a.only:
            break
        case += 1
    log("fuzz: %d cases, %d mismatches, %.0f s" % (n, fails, time.time() - t0))
    return fails


if __name__ == "__main__":
    sys.exit(min(main(), 100))
