"""Seeded synthetic inputs for the ORB front-end (frames, stereo pairs, descriptor databases).

No datasets are reachable from the build or the GPU box, so every test / bench input is generated here
from integer RNG streams (numpy PCG64 / splitmix64): identical on every machine.  Shapes follow
BASELINE.json's configs (640x480 TUM, 752x480 EuRoC, 1241x376 KITTI).

A frame contains: a smooth low-frequency background, a few hundred random filled
rectangles / triangles (corner sources for FAST), a dense-texture patch (quota-limited quadtree), a
perfectly flat patch (cells empty even at minThFAST), a low-contrast patch whose step heights lie in
(7, 20] grey levels (exercises the iniThFAST -> minThFAST retry), and +-noise.
"""
import numpy as np

_M64 = (1 << 64) - 1


def _rng(seed):
    return np.random.Generator(np.random.PCG64(int(seed)))


def _background(rng, w, h):
    """Bilinear upsampling of a coarse random grid, integer arithmetic only."""
    gw, gh = 6, 5
    coarse = rng.integers(40, 216, size=(gh + 1, gw + 1)).astype(np.int64)
    xs = (np.arange(w, dtype=np.int64) * gw * 256) // w
    ys = (np.arange(h, dtype=np.int64) * gh * 256) // h
    xi, xf = xs >> 8, xs & 255
    yi, yf = ys >> 8, ys & 255
    c00 = coarse[yi][:, xi]
    c01 = coarse[yi][:, xi + 1]
    c10 = coarse[yi + 1][:, xi]
    c11 = coarse[yi + 1][:, xi + 1]
    top = c00 * (256 - xf)[None, :] + c01 * xf[None, :]
    bot = c10 * (256 - xf)[None, :] + c11 * xf[None, :]
    return ((top * (256 - yf)[:, None] + bot * yf[:, None]) >> 16).astype(np.int32)


def _fill_triangle(img, pts, val):
    (x0, y0), (x1, y1), (x2, y2) = pts
    xa, xb = max(min(x0, x1, x2), 0), min(max(x0, x1, x2) + 1, img.shape[1])
    ya, yb = max(min(y0, y1, y2), 0), min(max(y0, y1, y2) + 1, img.shape[0])
    if xa >= xb or ya >= yb:
        return
    yy, xx = np.mgrid[ya:yb, xa:xb]
    e0 = (x1 - x0) * (yy - y0) - (y1 - y0) * (xx - x0)
    e1 = (x2 - x1) * (yy - y1) - (y2 - y1) * (xx - x1)
    e2 = (x0 - x2) * (yy - y2) - (y0 - y2) * (xx - x2)
    inside = ((e0 >= 0) & (e1 >= 0) & (e2 >= 0)) | ((e0 <= 0) & (e1 <= 0) & (e2 <= 0))
    img[ya:yb, xa:xb][inside] = val


def synth_frame(seed, w=640, h=480, n_rect=260, n_tri=120, noise=8):
    """One uint8 (h, w) frame; deterministic in (seed, w, h)."""
    rng = _rng(seed)
    img = _background(rng, w, h)
    # random rectangles and triangles with strong contrast
    for _ in range(n_rect):
        x0 = int(rng.integers(0, w)); y0 = int(rng.integers(0, h))
        rw = int(rng.integers(6, 70)); rh = int(rng.integers(6, 70))
        img[y0:y0 + rh, x0:x0 + rw] = int(rng.integers(0, 256))
    for _ in range(n_tri):
        cx = int(rng.integers(0, w)); cy = int(rng.integers(0, h))
        pts = [(cx + int(rng.integers(-40, 41)), cy + int(rng.integers(-40, 41))) for _ in range(3)]
        _fill_triangle(img, pts, int(rng.integers(0, 256)))
    # dense texture patch: 4-px random checker blocks (many strong corners -> quota-limited quadtree)
    tx0, ty0 = int(rng.integers(0, w // 2)), int(rng.integers(0, h // 2))
    tw, th = w // 4, h // 4
    blocks = rng.integers(0, 256, size=(th // 4 + 1, tw // 4 + 1)).astype(np.int32)
    img[ty0:ty0 + th, tx0:tx0 + tw] = np.kron(blocks, np.ones((4, 4), np.int32))[:th, :tw]
    # low-contrast patch: base level +- steps in (7, 20]
    lx0, ly0 = int(rng.integers(w // 2, w - w // 5)), int(rng.integers(0, h - h // 4))
    lw, lh = w // 5, h // 4
    base = int(rng.integers(60, 190))
    low = np.full((lh, lw), base, np.int32)
    for _ in range(40):
        x0 = int(rng.integers(0, lw)); y0 = int(rng.integers(0, lh))
        rw = int(rng.integers(5, 30)); rh = int(rng.integers(5, 30))
        low[y0:y0 + rh, x0:x0 + rw] = base + int(rng.integers(8, 21)) * (1 if rng.integers(0, 2) else -1)
    img[ly0:ly0 + lh, lx0:lx0 + lw] = low
    # noise everywhere except the low-contrast patch (kept clean so steps stay in (7,20]) ...
    nz = rng.integers(-noise, noise + 1, size=(h, w)).astype(np.int32)
    nz[ly0:ly0 + lh, lx0:lx0 + lw] = 0
    img = img + nz
    # ... and a perfectly flat patch
    fx0, fy0 = int(rng.integers(0, w - w // 6)), int(rng.integers(h // 2, h - h // 5))
    img[fy0:fy0 + h // 5, fx0:fx0 + w // 6] = int(rng.integers(30, 226))
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_batch(seed0, n, w=640, h=480, unique=16, noise=8, n_rect=260, n_tri=120):
    """n frames (n, h, w): `unique` generated frames, the rest are circular shifts of those (cheap but all
    distinct), so that a batch does not repeat the same memory image."""
    base = [synth_frame(seed0 + i, w, h, n_rect=n_rect, n_tri=n_tri, noise=noise) for i in range(min(unique, n))]
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        b = base[i % len(base)]
        k = i // len(base)
        out[i] = b if k == 0 else np.roll(b, (7 * k) % h, axis=0)
        if k:
            out[i] = np.roll(out[i], (13 * k) % w, axis=1)
    return out


def shifted_frame(img, dx, dy, seed, renoise_frac=0.02):
    """Second view for frame-to-frame matching: integer translation + a fraction of re-noised pixels."""
    rng = _rng(seed ^ 0x5EED)
    out = np.roll(np.roll(img, dy, axis=0), dx, axis=1).astype(np.int32)
    mask = rng.random(img.shape) < renoise_frac
    out[mask] += rng.integers(-8, 9, size=int(mask.sum()))
    return np.clip(out, 0, 255).astype(np.uint8)


def synth_stereo_pair(seed, w=1241, h=376, dmin=2, dmax=60, half_pixel=False):
    """Left frame + right frame warped by a piece-wise constant integer disparity field (right(x) = left(x+d)).
    half_pixel (second variant): the scene is rendered at 2x horizontal resolution, the right view is
    displaced by an ODD number of half pixels and both views are box-filtered down by 2, so that the true disparities are
    d + 0.5 and the parabola sub-pixel fit of ComputeStereoMatches (Frame.cc:634-641) has something to find."""
    if half_pixel:
        wide = synth_frame(seed, 2 * w, h, n_rect=420, n_tri=200)
        rng = _rng(seed ^ 0x57E2E1)
        disp2 = np.empty((h, 2 * w), np.int32)                       # disparity in half pixels, odd
        nby, nbx = 4, 6
        ys = np.linspace(0, h, nby + 1).astype(int)
        xs = np.linspace(0, 2 * w, nbx + 1).astype(int)
        for by in range(nby):
            for bx in range(nbx):
                disp2[ys[by]:ys[by + 1], xs[bx]:xs[bx + 1]] = 2 * int(rng.integers(dmin, dmax + 1)) + 1
        xx = np.clip(np.arange(2 * w)[None, :] + disp2, 0, 2 * w - 1)
        wide_r = np.take_along_axis(wide, xx, axis=1)
        down = lambda a: ((a[:, 0::2].astype(np.int32) + a[:, 1::2].astype(np.int32) + 1) >> 1)
        left = down(wide).astype(np.uint8)
        nz = rng.integers(-2, 3, size=(h, w))
        right = np.clip(down(wide_r) + nz, 0, 255).astype(np.uint8)
        return left, right, disp2[:, 0::2].astype(np.float32) / 2
    left = synth_frame(seed, w, h, n_rect=420, n_tri=200)
    rng = _rng(seed ^ 0x57E2E0)
    right = np.empty_like(left)
    disp = np.empty((h, w), np.int32)
    # horizontal bands x vertical slabs, each with its own disparity
    nby, nbx = 4, 6
    ys = np.linspace(0, h, nby + 1).astype(int)
    xs = np.linspace(0, w, nbx + 1).astype(int)
    for by in range(nby):
        for bx in range(nbx):
            disp[ys[by]:ys[by + 1], xs[bx]:xs[bx + 1]] = int(rng.integers(dmin, dmax + 1))
    xx = np.arange(w)[None, :] + disp
    xx = np.clip(xx, 0, w - 1)
    right = np.take_along_axis(left, xx, axis=1)
    nz = rng.integers(-2, 3, size=(h, w))
    right = np.clip(right.astype(np.int32) + nz, 0, 255).astype(np.uint8)
    return left, right, disp


# ---- descriptor databases (config 5) ---------------------------------------------------------------
def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_descriptors(seed, start, count):
    """Rows [start, start+count) of an endless descriptor table: word k of row i = splitmix64(seed ^ (4i+k)).
    Counter-based, so every shard can generate its own slice independently.  Returns uint8 (count, 32)."""
    with np.errstate(over="ignore"):
        i = np.arange(start, start + count, dtype=np.uint64)
        ctr = (i[:, None] * np.uint64(4) + np.arange(4, dtype=np.uint64)[None, :]) ^ np.uint64(seed)
        words = _splitmix64(ctr)
    return words.view(np.uint8).reshape(count, 32)


def synth_queries(seed, db_total, nq, n_planted=None, max_flips=40):
    """nq query descriptors: the first n_planted are DB rows (at pseudo-random known indices) with r in
    [0, max_flips] seeded bit flips, the rest are fresh random.  Returns (queries, planted_index, flips)."""
    if n_planted is None:
        n_planted = nq // 2
    rng = _rng(seed ^ 0xABCDEF)
    idx = rng.integers(0, db_total, size=n_planted)
    q = np.empty((nq, 32), np.uint8)
    flips = np.zeros(nq, np.int32)
    for j in range(n_planted):
        row = synth_descriptors(seed, int(idx[j]), 1)[0].copy()
        r = int(rng.integers(0, max_flips + 1))
        bits = rng.choice(256, size=r, replace=False)
        for b in bits:
            row[b >> 3] ^= np.uint8(1 << (b & 7))
        q[j] = row
        flips[j] = r
    q[n_planted:] = rng.integers(0, 256, size=(nq - n_planted, 32), dtype=np.uint8)
    planted = np.full(nq, -1, np.int64)
    planted[:n_planted] = idx
    return q, planted, flips


# ---- vocabulary trees (DBoW2 TemplatedVocabulary<FORB>; ORBvoc.txt is not reachable) -------------------
def synth_vocabulary(seed, k=10, L=6, irregular=False, p_stop=0.0, p_dup=0.02, flips=48):
    """A k-ary, depth-L vocabulary tree in DBoW2's node numbering (children of a node get consecutive ids when the
    node is expanded, the tree is expanded level by level here): returns (parent, is_leaf, desc, weight) in node-id
    order, node 0 = root.  A child's descriptor = its parent's with `flips` random bit flips (so the descent is
    meaningful), with probability p_dup a copy of its previous sibling (exercises the first-minimum tie rule).
    irregular: nodes get 2..k children and become leaves early with probability 0.15; p_stop: fraction of words
    with weight 0 (stopped words).  Weights carry 6 significant digits, like saveToTextFile writes them."""
    rng = _rng(seed ^ 0x5EED0B0)
    if not irregular:
        return _synth_vocabulary_regular(rng, k, L, p_stop, p_dup)
    parent = [0]
    depth = [0]
    descs = [np.zeros(32, np.uint8)]
    frontier = [0]
    leaf = [False]
    for level in range(1, L + 1):
        nxt = []
        for p in frontier:
            nch = int(rng.integers(2, k + 1)) if irregular else k
            base = rng.integers(0, 256, size=32, dtype=np.uint8) if p == 0 else descs[p]
            prev = None
            for c in range(nch):
                if prev is not None and rng.random() < p_dup:
                    d = prev.copy()
                else:
                    d = base.copy()
                    bits = rng.choice(256, size=flips if p else 128, replace=False)
                    np.bitwise_xor.at(d, bits >> 3, (1 << (bits & 7)).astype(np.uint8))
                nid = len(parent)
                parent.append(p); depth.append(level); descs.append(d)
                stop_here = level == L or (irregular and rng.random() < 0.15)
                leaf.append(stop_here)
                if not stop_here:
                    nxt.append(nid)
                prev = d
        frontier = nxt
    n = len(parent)
    parent = np.asarray(parent, np.int32)
    is_leaf = np.asarray(leaf, np.uint8)
    desc = np.stack(descs).astype(np.uint8)
    weight = np.zeros(n, np.float64)
    w = rng.uniform(0.5, 12.0, size=n)
    w = np.asarray([float("%g" % x) for x in w])
    weight[is_leaf == 1] = w[is_leaf == 1]
    if p_stop > 0:
        stop = (rng.random(n) < p_stop) & (is_leaf == 1)
        weight[stop] = 0.0
    return parent, is_leaf, desc, weight


def _synth_vocabulary_regular(rng, k, L, p_stop, p_dup):
    """Full k-ary tree, vectorised level by level (the ORBvoc shape k = 10, L = 6 has 1.1 M nodes): a child = its parent
    with ~25 % of the bits flipped (AND of two random byte planes), or a copy of its previous sibling (p_dup)."""
    parents, descs = [np.zeros(1, np.int32)], [np.zeros((1, 32), np.uint8)]
    first = 1
    prev_ids = np.zeros(1, np.int64)
    prev_desc = rng.integers(0, 256, size=(1, 32), dtype=np.uint8)
    for level in range(1, L + 1):
        m = len(prev_ids)
        par = np.repeat(prev_ids, k)
        d = np.repeat(prev_desc, k, axis=0)
        d ^= rng.integers(0, 256, size=d.shape, dtype=np.uint8) & rng.integers(0, 256, size=d.shape, dtype=np.uint8)
        dup = (rng.random(m * k) < p_dup) & (np.arange(m * k) % k != 0)
        idx = np.nonzero(dup)[0]
        for i in idx:                      # sequential so that runs of duplicates copy the same ancestor
            d[i] = d[i - 1]
        ids = first + np.arange(m * k, dtype=np.int64)
        parents.append(par.astype(np.int32)); descs.append(d)
        first += m * k
        prev_ids, prev_desc = ids, d
    parent = np.concatenate(parents)
    desc = np.concatenate(descs)
    n = len(parent)
    is_leaf = np.zeros(n, np.uint8)
    is_leaf[n - len(prev_ids):] = 1
    weight = np.zeros(n, np.float64)
    w = np.round(rng.uniform(0.5, 12.0, size=len(prev_ids)), 4)
    if p_stop > 0:
        w[rng.random(len(prev_ids)) < p_stop] = 0.0
    weight[n - len(prev_ids):] = w
    return parent, is_leaf, desc, weight


def write_vocabulary_text(path, k, L, scoring, weighting, parent, is_leaf, desc, weight):
    """The reference's text format (TemplatedVocabulary::saveToTextFile, TemplatedVocabulary.h:1447-1467): header
    `k L  scoring weighting`, then one line per node 1..n-1: `parent isLeaf d0 .. d31  weight`."""
    with open(path, "w") as f:
        f.write("%d %d  %d %d\n" % (k, L, scoring, weighting))
        for i in range(1, len(parent)):
            f.write("%d %d %s  %g\n" % (parent[i], 1 if is_leaf[i] else 0, " ".join(str(int(b)) for b in desc[i]), weight[i]))


# ---- integer-only, counter-based frame generator: the same function exists in C (tools/synth_int.c) --------
# No RNG state, no libm, no OpenCV: every random number is splitmix64(seed * 2^32 + counter), so the C and the Python
# implementation produce byte-identical frames (tests/test_synth_int.py) and a C++ caller can regenerate any test input.
def _sm64_scalar(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def synth_frame_int(seed, w=640, h=480, n_rect=380, noise=8):
    """uint8 (h, w).  Content: low-frequency background (integer bilinear blend of a 6x5 grid), n_rect filled
    rectangles (corner sources), a dense-texture patch, a perfectly flat patch, a low-contrast patch with steps in (7, 20],
    +-noise per pixel from a hash of the pixel index."""
    base = (int(seed) & 0xFFFFFFFF) << 32
    ctr = [0]

    def rnd(n):                                   # uniform in [0, n)
        ctr[0] += 1
        return (_sm64_scalar(base + ctr[0]) >> 11) % n
    gw, gh = 6, 5
    coarse = np.array([[40 + rnd(176) for _ in range(gw + 1)] for _ in range(gh + 1)], np.int64)
    xs = (np.arange(w, dtype=np.int64) * gw * 256) // w
    ys = (np.arange(h, dtype=np.int64) * gh * 256) // h
    xi, xf, yi, yf = xs >> 8, xs & 255, ys >> 8, ys & 255
    top = coarse[yi][:, xi] * (256 - xf)[None, :] + coarse[yi][:, xi + 1] * xf[None, :]
    bot = coarse[yi + 1][:, xi] * (256 - xf)[None, :] + coarse[yi + 1][:, xi + 1] * xf[None, :]
    img = ((top * (256 - yf)[:, None] + bot * yf[:, None]) >> 16).astype(np.int64)
    for _ in range(n_rect):
        x0, y0, rw, rh, v = rnd(w), rnd(h), 6 + rnd(64), 6 + rnd(64), rnd(256)
        img[y0:y0 + rh, x0:x0 + rw] = v
    # dense texture: 4x4 checker blocks of random levels in the top-left sixth
    tw, th = w // 3, h // 2
    for by in range(0, th, 4):
        for bx in range(0, tw, 4):
            img[by:by + 4, bx:bx + 4] = rnd(256)
    # perfectly flat patch (bottom-right) and a low-contrast patch (bottom-left) with steps in (7, 20]
    img[h - h // 4:, w - w // 4:] = 128
    lx, ly = w // 4, h // 4
    for by in range(h - ly, h, 12):
        for bx in range(0, lx, 12):
            img[by:by + 12, bx:bx + 12] = 100 + ((bx // 12 + by // 12) & 1) * (8 + rnd(13))
    # noise: hash of (seed, pixel index), uniform in [-noise, +noise]; the flat patch stays flat
    if noise > 0:
        with np.errstate(over="ignore"):
            idx = np.arange(w * h, dtype=np.uint64) + np.uint64((base + (1 << 31)) & _M64)
            nz = ((_splitmix64(idx) >> np.uint64(11)) % np.uint64(2 * noise + 1)).astype(np.int64).reshape(h, w) - noise
        nz[h - h // 4:, w - w // 4:] = 0
        img = img + nz
    return np.clip(img, 0, 255).astype(np.uint8)
