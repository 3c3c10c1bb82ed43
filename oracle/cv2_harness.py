"""Second, independent restatement of ORBextractor::operator() that drives the REAL OpenCV primitives
(cv2 4.13.0 wheel: resize / copyMakeBorder / FastFeatureDetector / GaussianBlur / fastAtan2) with the
reference's control flow (ORBextractor.cc:790-892, 1083-1185) written in plain Python.

TEST INFRASTRUCTURE ONLY.  Purpose: pin the C++ oracle (and, through the committed golden files it
generates, the CUDA path) to the third-party module that owns the arithmetic.  OpenCV is not vendored
by the reference and its version is not pinned there (CMakeLists.txt:31-40); this repo pins 4.13.0.

Slow (pure-Python quadtree and descriptor loops): use on a handful of frames only.
"""
import math

import cv2
import numpy as np

from . import orb_oracle as O

EDGE = 19
PATCH = 31
HALF = 15


def _cv_round(v):
    """cvRound == round-half-to-even on the float32 value."""
    return int(np.rint(np.float32(v)))


class Node:
    __slots__ = ("keys", "UL", "UR", "BL", "BR", "no_more", "seq", "alive")

    def __init__(self):
        self.keys = []
        self.no_more = False
        self.alive = True


def _divide(n):
    f32 = np.float32
    halfX = int(math.ceil(f32(n.UR[0] - n.UL[0]) / f32(2)))
    halfY = int(math.ceil(f32(n.BR[1] - n.UL[1]) / f32(2)))
    c = [Node() for _ in range(4)]
    c[0].UL = n.UL; c[0].UR = (n.UL[0] + halfX, n.UL[1]); c[0].BL = (n.UL[0], n.UL[1] + halfY)
    c[0].BR = (n.UL[0] + halfX, n.UL[1] + halfY)
    c[1].UL = c[0].UR; c[1].UR = n.UR; c[1].BL = c[0].BR; c[1].BR = (n.UR[0], n.UL[1] + halfY)
    c[2].UL = c[0].BL; c[2].UR = c[0].BR; c[2].BL = n.BL; c[2].BR = (c[0].BR[0], n.BL[1])
    c[3].UL = c[2].UR; c[3].UR = c[1].BR; c[3].BL = c[2].BR; c[3].BR = n.BR
    sx, sy = c[0].UR[0], c[0].BR[1]
    for kp in n.keys:
        if kp[0] < sx:
            (c[0] if kp[1] < sy else c[2]).keys.append(kp)
        elif kp[1] < sy:
            c[1].keys.append(kp)
        else:
            c[3].keys.append(kp)
    for ch in c:
        if len(ch.keys) == 1:
            ch.no_more = True
    return c


def distribute_quadtree(keys, minX, maxX, minY, maxY, N):
    """keys: list of (x, y, response) in vToDistributeKeys order.  Python list stands in for std::list:
    index 0 == front.  Tie rule for equal sizes: creation sequence (pin (ii))."""
    f32 = np.float32
    nIni = int(math.floor(f32(maxX - minX) / f32(maxY - minY) + f32(0.5)))  # round(), positive argument
    hX = f32(maxX - minX) / f32(nIni)
    seq = 0
    nodes = []
    for i in range(nIni):
        n = Node()
        n.UL = (int(hX * f32(i)), 0); n.UR = (int(hX * f32(i + 1)), 0)
        n.BL = (n.UL[0], maxY - minY); n.BR = (n.UR[0], maxY - minY)
        n.seq = seq; seq += 1
        nodes.append(n)
    ini = list(nodes)
    for kp in keys:
        ini[int(f32(kp[0]) / hX)].keys.append(kp)
    kept = []
    for n in nodes:
        if len(n.keys) == 1:
            n.no_more = True
            kept.append(n)
        elif len(n.keys) > 1:
            kept.append(n)
    nodes = kept
    finish = False
    while not finish:
        prev = len(nodes)
        front = []      # children, most recently pushed first
        cand = []
        n_expand = 0
        rest = []
        for n in nodes:
            if n.no_more:
                rest.append(n)
                continue
            for ch in _divide(n):
                if ch.keys:
                    ch.seq = seq; seq += 1
                    front.insert(0, ch)
                    if len(ch.keys) > 1:
                        n_expand += 1
                        cand.append(ch)
        nodes = front + rest
        if len(nodes) >= N or len(nodes) == prev:
            finish = True
        elif len(nodes) + n_expand * 3 > N:
            while not finish:
                prev = len(nodes)
                prev_cand = sorted(cand, key=lambda c: (len(c.keys), c.seq))
                cand = []
                for j in range(len(prev_cand) - 1, -1, -1):
                    parent = prev_cand[j]
                    for ch in _divide(parent):
                        if ch.keys:
                            ch.seq = seq; seq += 1
                            nodes.insert(0, ch)
                            if len(ch.keys) > 1:
                                cand.append(ch)
                    nodes.remove(parent)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for kp in n.keys[1:]:
            if kp[2] > best[2]:
                best = kp
        out.append(best)
    return out


def _pattern():
    import os
    txt = open(os.path.join(os.path.dirname(__file__), "orb_pattern_31.inc")).read()
    vals = [int(t) for line in txt.splitlines() if not line.startswith("//") for t in line.replace(",", " ").split()]
    return np.array(vals, np.int32).reshape(512, 2)


PATTERN = _pattern()
UMAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def ic_angle(img, x, y):
    m01 = 0
    m10 = 0
    row = img[y].astype(np.int64)
    for u in range(-HALF, HALF + 1):
        m10 += u * int(row[x + u])
    for v in range(1, HALF + 1):
        d = UMAX[v]
        us = np.arange(-d, d + 1)
        plus = img[y + v, x - d:x + d + 1].astype(np.int64)
        minus = img[y - v, x - d:x + d + 1].astype(np.int64)
        m01 += v * int((plus - minus).sum())
        m10 += int((us * (plus + minus)).sum())
    return cv2.fastAtan2(float(m01), float(m10))


def brief(img, x, y, angle_deg):
    f32 = np.float32
    angle = f32(angle_deg) * f32(np.float32(math.pi) / f32(180.0))
    a = f32(math.cos(float(angle)))
    b = f32(math.sin(float(angle)))
    px = PATTERN[:, 0].astype(np.float32)
    py = PATTERN[:, 1].astype(np.float32)
    fy = (px * b + py * a).astype(np.float32)   # numpy evaluates each op in float32, no FMA
    fx = (px * a - py * b).astype(np.float32)
    iy = np.rint(fy).astype(np.int64) + y
    ix = np.rint(fx).astype(np.int64) + x
    vals = img[iy, ix].astype(np.int32)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
    return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)


def extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """Returns (kps structured array in output order, desc (n,32), per-level dict of stage dumps)."""
    ex = O.Extractor(nfeatures, scale_factor, nlevels, ini_th, min_th)   # constructor tables only
    scale, inv = ex.scale_factors, ex.inv_scale_factors
    quota = ex.features_per_level
    h0, w0 = img.shape
    pyr = []
    for l in range(nlevels):
        w = _cv_round(np.float32(w0) * inv[l]); h = _cv_round(np.float32(h0) * inv[l])
        cur = img if l == 0 else cv2.resize(pyr[l - 1], (w, h), interpolation=cv2.INTER_LINEAR)
        pyr.append(cur)
    fd_ini = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fd_min = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    all_kps, all_desc, stages = [], [], []
    for l in range(nlevels):
        im = pyr[l]
        h, w = im.shape
        bordered = cv2.copyMakeBorder(im, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)
        minB = EDGE - 3
        maxBX, maxBY = w - EDGE + 3, h - EDGE + 3
        f32 = np.float32
        width, height = f32(maxBX - minB), f32(maxBY - minB)
        nCols, nRows = int(width / f32(30)), int(height / f32(30))
        wCell, hCell = int(math.ceil(width / f32(nCols))), int(math.ceil(height / f32(nRows)))
        raw = []
        for i in range(nRows):
            iniY = minB + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minB + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                cell = im[iniY:maxY, iniX:maxX]
                k = fd_ini.detect(cell)
                if not k:
                    k = fd_min.detect(cell)
                for p in k:
                    raw.append((p.pt[0] + j * wCell, p.pt[1] + i * hCell, p.response))
        kept = distribute_quadtree(raw, minB, maxBX, minB, maxBY, int(quota[l])) if raw else []
        lvl = np.zeros(len(kept), O.KP_DTYPE)
        blurred = cv2.GaussianBlur(im, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101) if kept else None
        desc = np.zeros((len(kept), 32), np.uint8)
        for n, (x, y, r) in enumerate(kept):
            xi, yi = int(x) + minB, int(y) + minB
            ang = ic_angle(bordered, xi + EDGE, yi + EDGE)
            lvl[n] = (xi, yi, float(int(f32(PATCH) * scale[l])), ang, r, l, -1)
            desc[n] = brief(blurred, xi, yi, ang)
        out = lvl.copy()
        if l != 0:
            out["x"] = out["x"] * scale[l]
            out["y"] = out["y"] * scale[l]
        all_kps.append(out); all_desc.append(desc)
        stages.append({"bordered": bordered, "blurred": blurred, "raw": np.array(raw, np.float32).reshape(-1, 3),
                       "level_kps": lvl})
    return np.concatenate(all_kps), np.concatenate(all_desc), stages
