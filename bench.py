#!/usr/bin/env python
"""bench.py — headline measurement of the B200-native ORB front-end.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU restatement on the host cores)

One "step" = one pass of the extractor hot path (pyramid -> per-cell FAST -> quadtree -> blur -> orientation
+ rBRIEF) over one batch of synthetic 640x480 frames per GPU (BASELINE config 1's shape and parameters,
batched the way config 4 shards frames).  The JSON line also carries the second half of BASELINE's metric —
Hamming compares/s of the map-wide top-2 search (config 5: 2000 queries x a 10 M-row database sharded over
the ranks with an all-gather + merge) — under "hamming".

Timed regions:
  value   frames already in HBM, outputs stay in HBM (orb_extract_batch_device), CUDA events, max over ranks
  e2e     HOST buffers in, HOST keypoints/descriptors out through the public call (ORBextractor.extract_batch
          -> orb_extract_batch): pinned staging, H2D, kernels, D2H inside the timed region; the steps are issued by two
          extractor instances on two host threads (value) and by a single caller (one_caller_value)
  roofline    per-stage CUDA-event times recorded by the library inside the timed region of `value`
  cpu_baseline  oracle/ (the CPU restatement = checker) timed on all host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, NFEAT, NLEVELS, SCALE, INI_TH, MIN_TH = 640, 480, 1000, 8, 1.2, 20, 7
METRIC = "ORB extract frames/s @640x480 1k feat"
NQ = 2000
DB_SEED = 7


# ---------------------------------------------------------------------------------------------------------
# algorithmic bytes per frame, unfused-stage model of SURVEY.md §8(d) (stated in DESIGN.md §Roofline)
# ---------------------------------------------------------------------------------------------------------
def level_dims(w, h, nlevels=NLEVELS, scale=SCALE):
    sf = [np.float32(1.0)]
    for _ in range(1, nlevels):
        sf.append(np.float32(np.float64(sf[-1]) * np.float64(scale)))
    out = []
    for l in range(nlevels):
        inv = np.float32(1.0) / sf[l]
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def algorithmic_bytes(w, h):
    dims = level_dims(w, h)
    px = [a * b for a, b in dims]
    bordered = [(a + 38) * (b + 38) for a, b in dims]
    return {
        "pyramid": px[0] + sum(px[:-1]) + sum(bordered),  # read L0, read levels 0..6, write bordered levels
        "fast_cells": sum(px),                             # read every level once
        "blur": 2 * sum(px),                               # read + write every level
    }


# ---------------------------------------------------------------------------------------------------------
def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return
    out["proc"] = p
    for line in p.stdout:
        if stop.is_set():
            break
        f = [x.strip() for x in line.split(",")]
        if len(f) >= 6:
            out.setdefault("rows", []).append(f)
    p.kill()


_CPULIST_CACHE = {}


def _gpu_cpulist(gpu_index):
    if gpu_index in _CPULIST_CACHE:
        return _CPULIST_CACHE[gpu_index]
    out = None
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus:
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]
            with open("/sys/bus/pci/devices/%s/local_cpulist" % bus) as fh:
                out = fh.read().strip()
    except (OSError, ValueError, subprocess.SubprocessError):
        out = None
    _CPULIST_CACHE[gpu_index] = out
    return out


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank (and therefore the pinned host buffers it allocates next: first touch) to the host cores next to its
    GPU.  With 8 ranks on one box every rank otherwise allocates on whichever socket it was started on and half of the
    H2D traffic crosses the socket interconnect.  Returns the cpulist string, or None when sysfs / nvidia-smi say nothing."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        if len(bus.split(":")[0]) == 8:   # nvidia-smi prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bus) as fh:
            cpulist = fh.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        # ranks whose GPUs hang off the same node share that node's cores: give each its own contiguous slice (>= 4 cores) so that
        # the caller threads and the driver's copy threads of different ranks do not pile onto the same cores
        world = int(os.environ.get("WORLD_SIZE", "1"))
        same = [g for g in range(world) if _gpu_cpulist(g) == cpulist] if world > 1 else [gpu_index]
        if len(same) > 1 and gpu_index in same and os.environ.get("ORB_BENCH_SLICE", "1") != "0":
            order = sorted(cpus)
            per = max(4, len(order) // len(same))
            k = same.index(gpu_index)
            mine = set(order[(k * per) % len(order):(k * per) % len(order) + per]) or cpus
            os.sched_setaffinity(0, mine)
            return "%s (slice %d/%d: %d cores)" % (cpulist, k, len(same), len(mine))
        os.sched_setaffinity(0, cpus)
        return cpulist
    except (OSError, ValueError, subprocess.SubprocessError):
        return None


def summarize_clocks(rows):
    if not rows:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(int(r[0]) for r in rows if r[0].isdigit())
    mx = [int(r[1]) for r in rows if r[1].isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
            "samples": len(rows)}


# ---------------------------------------------------------------------------------------------------------
# CPU restatement (oracle/) on the host cores: cpu_baseline leg and the --impl reference arm
# ---------------------------------------------------------------------------------------------------------
def cpu_arm():
    """('reference', harness) when oracle/_ref/liborb_ref.so is there — the reference's own unmodified ORBextractor.cc compiled over
    the OpenCV stand-in (oracle/Makefile) — else ('port', None): oracle/'s restatement."""
    try:
        from oracle import orb_ref
        if os.path.exists(orb_ref.REF_SO):
            H = orb_ref.Harness(orb_ref.REF_SO)
            H.L.rh_set_monotonic_alloc(0)      # plain malloc for the quadtree nodes, as the reference runs
            return "reference", H
    except Exception as e:                     # noqa: BLE001
        sys.stderr.write("oracle/_ref unavailable (%s): timing the oracle port\n" % e)
    return "port", None


def cpu_extract_throughput(frames, threads, repeat=1, harness=None):
    """frames: uint8 [n, H, W]; every thread owns one CPU extractor (the reference's ORBextractor through the rh_* harness, or
    the oracle port) and takes frames round-robin."""
    if harness is not None:
        exs = [harness.extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH) for _ in range(threads)]
    else:
        from oracle import orb_oracle
        exs = [orb_oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH) for _ in range(threads)]
    counts = [0] * threads

    def work(t):
        for _ in range(repeat):
            for i in range(t, len(frames), threads):
                k, _d = exs[t].extract(frames[i])
                counts[t] += 1
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return sum(counts) / dt, dt, sum(counts)


def cv2_simd_primitive_row(frame, reps=5):
    """BASELINE.md row B5 (sanity, not the headline): OpenCV's own SIMD kernels (the cv2 4.13.0 wheel, one thread) on the primitive
    calls ORBextractor makes for one frame — resize + copyMakeBorder per level, one cv::FAST call per 30-px cell at iniThFAST,
    GaussianBlur per level — driven from Python (the ~800 FAST calls include interpreter overhead).  Shows how far the scalar,
    bit-exact primitives of the CPU arm are from an optimised OpenCV build.  None when cv2 is not importable."""
    try:
        import cv2
    except Exception:      # noqa: BLE001
        return None
    cv2.setNumThreads(1)
    h, w = frame.shape
    dims = level_dims(w, h)
    fast = cv2.FastFeatureDetector_create(INI_TH, True)

    def pyramid():
        lv = [frame]
        for (lw, lh) in dims[1:]:
            lv.append(cv2.resize(lv[-1], (lw, lh), interpolation=cv2.INTER_LINEAR))
        return [cv2.copyMakeBorder(x, 19, 19, 19, 19, cv2.BORDER_REFLECT_101) for x in lv]

    def cells(bordered):
        n = 0
        for b in bordered:
            H_, W_ = b.shape[0] - 38, b.shape[1] - 38
            minb, maxx, maxy = 19 + 16, 19 + W_ - 16, 19 + H_ - 16     # ORBextractor.cc:801-817 (minBorder = 16, maxBorder = dim - 16), bordered coordinates
            ncols, nrows = (maxx - minb) // 30, (maxy - minb) // 30
            wc, hc = -(-(maxx - minb) // ncols), -(-(maxy - minb) // nrows)
            for i in range(nrows):
                y0 = minb + i * hc
                y1 = min(y0 + hc + 6, maxy)
                if y0 >= maxy - 3:
                    continue
                for j in range(ncols):
                    x0 = minb + j * wc
                    x1 = min(x0 + wc + 6, maxx)
                    if x0 >= maxx - 6:
                        continue
                    fast.detect(b[y0:y1, x0:x1])
                    n += 1
        return n

    def blur(bordered):
        for b in bordered:
            cv2.GaussianBlur(b[19:-19, 19:-19].copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)

    def ms(fn, *a):
        fn(*a)
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn(*a)
        return (time.perf_counter() - t0) / reps * 1e3, r
    t_pyr, bordered = ms(pyramid)
    t_fast, ncell = ms(cells, bordered)
    t_blur, _ = ms(blur, bordered)
    return {"pyramid_ms": t_pyr, "fast_cells_ms": t_fast, "fast_calls": ncell, "blur_ms": t_blur, "threads": 1, "cv2": cv2.__version__,
            "note": "BASELINE.md B5 sanity row: cv2 SIMD kernels called from Python for ONE %dx%d frame (no quadtree / orientation / descriptors); "
                    "not the baseline" % (w, h)}


def cpu_hamming_throughput(q, db, threads):
    from oracle import orb_oracle
    parts = np.array_split(np.arange(len(q)), threads)

    def work(t):
        if len(parts[t]):
            orb_oracle.hamming_top2(q[parts[t]], db)
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return len(q) * len(db) / dt, dt


def _median_ms(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def other_configs(device):
    """BASELINE configs 1-3 as the SLAM caller sees them: host buffers in, host results out, one call at a time
    (latency, not batch throughput), each next to the CPU oracle on ONE thread (the reference runs one thread per
    extractor, Frame.cc:79-82).  Results were already checked against the oracle by tests/ and smoke()."""
    from oracle import orb_oracle
    from orb_slam_2_ros_b200 import ORBextractor, ORBmatcher, compute_stereo_matches, synth
    from orb_slam_2_ros_b200.matcher import MODE_TRACK_LAST
    out = {}
    # config 1: one 640x480 frame, 1000 features
    img = synth.synth_frame(0, W, H)
    ex = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=device, max_batch=1)
    oex = orb_oracle.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    ms = _median_ms(lambda: ex(img), 50)
    cms = _median_ms(lambda: oex.extract(img), 5, warm=1)
    out["config1_single_frame"] = {"workload": "one 640x480 frame, nFeatures=1000, host image -> host keypoints + descriptors",
                                   "ms": ms, "frames_per_s": 1e3 / ms, "cpu_oracle_ms_1thread": cms}
    # config 2: 1000 x 1000 brute-force matching + SearchByProjection between two extracted frames
    img_b = synth.shifted_frame(img, 3, -2, 0)
    ka, da = ex(img)
    kb, db = ex(img_b)
    m = ORBmatcher(0.6, True, device=device)
    ms_bf = _median_ms(lambda: m.MatchBruteForce(da, ka["angle"], db, kb["angle"], 50), 30)
    cms_bf = _median_ms(lambda: orb_oracle.match_bruteforce(da, ka["angle"], db, kb["angle"], 50, 0.6, True), 5, warm=1)
    sf = ex.mvScaleFactor
    q_u = (ka["x"] + np.float32(3)).astype(np.float32); q_v = (ka["y"] - np.float32(2)).astype(np.float32)
    q_r = (np.float32(15.0) * sf[ka["octave"]]).astype(np.float32)
    qmin, qmax = ka["octave"] - 1, ka["octave"] + 1
    bounds = (0.0, 0.0, float(W), float(H))
    m9 = ORBmatcher(0.9, True, device=device)

    def sbp():
        return m9.SearchByProjection(MODE_TRACK_LAST, kb, db, bounds, np.zeros(len(kb), np.uint8), q_u, q_v, q_r, qmin, qmax, da,
                                     q_angle=ka["angle"], th_dist=100)
    grid = orb_oracle.Grid(kb, *bounds)

    def sbp_cpu():
        return orb_oracle.search_by_projection(orb_oracle.MODE_TRACK_LAST, grid, db, None, np.zeros(len(kb), np.uint8), q_u, q_v, q_r,
                                               qmin, qmax, da, q_angle=ka["angle"], th_dist=100, nn_ratio=0.9, check_orientation=True)
    ms_sbp = _median_ms(sbp, 30)
    cms_sbp = _median_ms(sbp_cpu, 5, warm=1)
    # SearchForInitialization (mono initialisation, ORBmatcher.cc:406-521): level-0 keypoints, 100-px windows
    prev0 = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)

    def sfi():
        return m9.SearchForInitialization(ka, da, kb, db, bounds, prev0.copy(), 100)
    valid0 = (ka["octave"] == 0).astype(np.uint8)
    zeros0 = np.zeros(len(ka), np.int32)

    def sfi_cpu():
        return orb_oracle.search_by_projection(orb_oracle.MODE_INITIALIZATION, grid, db, None, np.zeros(len(kb), np.uint8),
                                               prev0[:, 0].copy(), prev0[:, 1].copy(), np.full(len(ka), 100, np.float32), zeros0, zeros0,
                                               da, q_angle=ka["angle"], q_valid=valid0, th_dist=50, nn_ratio=0.9, check_orientation=True)
    ms_sfi = _median_ms(sfi, 30)
    cms_sfi = _median_ms(sfi_cpu, 5, warm=1)
    out["config2_matching"] = {"workload": "%d x %d descriptors of two extracted frames" % (len(da), len(db)),
                               "search_for_initialization_ms": ms_sfi, "search_for_initialization_cpu_oracle_ms_1thread": cms_sfi,
                               "search_for_initialization_matches": int(sfi()[0]),
                               "bruteforce_ratio_rothist_ms": ms_bf, "bruteforce_compares_per_s": len(da) * len(db) / (ms_bf * 1e-3),
                               "bruteforce_cpu_oracle_ms_1thread": cms_bf, "bruteforce_matches": int(m.MatchBruteForce(da, ka["angle"], db, kb["angle"], 50)[0]),
                               "search_by_projection_ms": ms_sbp, "search_by_projection_cpu_oracle_ms_1thread": cms_sbp,
                               "search_by_projection_matches": int(sbp()[0])}
    # config 3: KITTI-shape stereo pair, 2000 features: both extractions + ComputeStereoMatches
    left, right, _ = synth.synth_stereo_pair(2, 1241, 376)
    exl = ORBextractor(2000, SCALE, NLEVELS, INI_TH, MIN_TH, device=device, max_batch=1)
    exr = ORBextractor(2000, SCALE, NLEVELS, INI_TH, MIN_TH, device=device, max_batch=1)
    bf, b = 386.1448, 0.53716
    res = {}

    def stereo():
        kl, dl = exl(left)
        kr, dr = exr(right)
        res["n"] = compute_stereo_matches(exl, exr, kl, dl, kr, dr, bf, b)[0]
    oL, oR = orb_oracle.Extractor(2000, SCALE, NLEVELS, INI_TH, MIN_TH), orb_oracle.Extractor(2000, SCALE, NLEVELS, INI_TH, MIN_TH)

    def stereo_cpu():
        kl, dl = oL.extract(left)
        kr, dr = oR.extract(right)
        orb_oracle.stereo_match(oL, oR, kl, dl, kr, dr, bf, b)
    ms_st = _median_ms(stereo, 30)
    cms_st = _median_ms(stereo_cpu, 3, warm=1)

    # the reference extracts the two images on two std::threads (Frame.cc:79-82): the same with a helper thread for the
    # right image (the C call releases the GIL), then ComputeStereoMatches
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(1)

    def stereo_2threads():
        fut = pool.submit(exr, right)
        kl, dl = exl(left)
        kr, dr = fut.result()
        res["n2"] = compute_stereo_matches(exl, exr, kl, dl, kr, dr, bf, b)[0]
    ms_st2 = _median_ms(stereo_2threads, 30)
    pool.shutdown()
    out["config3_stereo"] = {"workload": "1241x376 pair, nFeatures=2000: 2 extractions + ComputeStereoMatches, sequential calls",
                             "ms": ms_st, "pairs_per_s": 1e3 / ms_st, "stereo_matches": int(res["n"]), "cpu_oracle_ms_1thread": cms_st,
                             "ms_two_extractor_threads": ms_st2, "stereo_matches_two_threads": int(res["n2"])}
    # bag-of-words transform (ORBvoc shape k = 10, L = 6, 1.1 M nodes) + SearchByBoW + SearchForTriangulation
    from orb_slam_2_ros_b200 import ORBVocabulary
    P = synth.synth_vocabulary(11, k=10, L=6)
    voc = ORBVocabulary.from_arrays(10, 6, 0, 0, *P, device=device)
    ovoc = orb_oracle.Vocabulary.from_arrays(10, 6, 0, 0, *P)
    ms_bow = _median_ms(lambda: voc.transform(da, 4), 30)
    cms_bow = _median_ms(lambda: ovoc.transform(da, 4), 5, warm=1)
    nbatch = 256
    frames_desc = [da if i % 2 == 0 else db for i in range(nbatch)]
    ms_bowb = _median_ms(lambda: voc.transform_batch(frames_desc, 4), 5, warm=2)
    (_, fva), (_, fvb) = voc.transform_batch([da, db], 4)
    m7 = ORBmatcher(0.7, True, device=device)
    ms_sbb = _median_ms(lambda: m7.SearchByBoW(da, ka["angle"], None, fva, db, kb["angle"], None, fvb), 30)
    cms_sbb = _median_ms(lambda: orb_oracle.search_by_bow(da, ka["angle"], None, fva, db, kb["angle"], None, fvb, 50, False, 0.7, True), 5, warm=1)
    F12 = np.array([[0, 0, -2], [0, 0, -3], [2, 3, 0]], np.float32) * np.float32(0.01)
    sig2 = (sf * sf).astype(np.float32)
    ms_tri = _median_ms(lambda: m.SearchForTriangulation(ka, da, None, None, fva, kb, db, None, None, fvb, F12, 1e6, 1e6, sf, sig2), 30)
    cms_tri = _median_ms(lambda: orb_oracle.search_for_triangulation(ka, da, None, None, fva, kb, db, None, None, fvb, F12, 1e6, 1e6, sf, sig2,
                                                                     False, True), 5, warm=1)
    out["bow_and_mapping_matchers"] = {
        "workload": "synthetic vocabulary k=10 L=6 (%d nodes, ORBvoc shape), %d / %d descriptors of two extracted frames, levelsup 4" % (len(P[0]), len(da), len(db)),
        "bow_transform_ms": ms_bow, "bow_transform_cpu_oracle_ms_1thread": cms_bow,
        "bow_transform_batch%d_ms" % nbatch: ms_bowb, "bow_transform_batch_descriptors_per_s": sum(len(d) for d in frames_desc) / (ms_bowb * 1e-3),
        "search_by_bow_ms": ms_sbb, "search_by_bow_cpu_oracle_ms_1thread": cms_sbb,
        "search_by_bow_matches": int(m7.SearchByBoW(da, ka["angle"], None, fva, db, kb["angle"], None, fvb)[0]),
        "search_for_triangulation_ms": ms_tri, "search_for_triangulation_cpu_oracle_ms_1thread": cms_tri,
        "search_for_triangulation_matches": int(m.SearchForTriangulation(ka, da, None, None, fva, kb, db, None, None, fvb, F12, 1e6, 1e6, sf, sig2)[0])}
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on all host cores.  oracle/_ref/liborb_ref.so is
    the reference's UNMODIFIED ORBextractor.cc compiled from /root/reference over an OpenCV stand-in whose primitives are
    verified against cv2 4.13.0 (kind "reference"); if that library is missing the arm times oracle/'s restatement (kind "port")."""
    if rank != 0:
        return
    from orb_slam_2_ros_b200 import synth
    kind, harness = cpu_arm()
    threads = os.cpu_count() or 1
    sample = max(threads, min(args.ref_frames, 4 * threads))
    frames = synth.synth_batch(0, sample, W, H, unique=min(16, sample))
    for _ in range(args.warmup):
        cpu_extract_throughput(frames[:threads], threads, harness=harness)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        _fps, dt, n = cpu_extract_throughput(frames, threads, harness=harness)
        t_total += dt
        n_total += n
    fps = n_total / t_total
    # Hamming sample: 256 queries x 400k rows per step
    from orb_slam_2_ros_b200 import synth as S
    db = S.synth_descriptors(DB_SEED, 0, 400000)
    q, _, _ = S.synth_queries(DB_SEED, 400000, 256)
    cps, hdt = cpu_hamming_throughput(q, db, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "ORBextractor 640x480 nFeatures=1000 8 levels 1.2 20/7 (BASELINE config 1 shape), "
                               "%d frames per step on %d host threads" % (sample, threads)},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": "%d synthetic frames per step, one %s per host thread; scalar OpenCV-4.13-exact primitives "
                                   "(no SIMD: an OpenCV build with SIMD kernels is roughly 2x faster per frame)"
                                   % (sample, "reference ORBextractor (unmodified ORBextractor.cc, oracle/_ref)" if kind == "reference" else "oracle extractor")},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hamming": {"metric": "Hamming compares/s", "value": cps, "unit": "compares/s", "cores": threads, "kind": "port",
                    "sample": "256 queries x 400000 rows, reference bit-hack DescriptorDistance, %.2f s" % hdt},
        "gpu_launches": 0,
    }
    emit_json_line(line)


# ---------------------------------------------------------------------------------------------------------
# The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner with printf when
# NCCL_DEBUG=VERSION is set in the environment), so file descriptor 1 is pointed at stderr for the whole run and the JSON
# line goes to the saved original.
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="frames per step per GPU (512 x 640x480 = 157 MB > L2)")
    ap.add_argument("--db-rows", type=int, default=10_000_000, help="total database rows of the Hamming leg")
    ap.add_argument("--ref-frames", type=int, default=64)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--noise", type=int, default=8, help="+-grey-level noise of the synthetic frames (SURVEY.md §8d: 8)")
    ap.add_argument("--e2e-callers", type=int, default=2, help="2: also time two extractor instances on two host threads")
    ap.add_argument("--frame-size", default="", help="WxH of the synthetic frames (default 640x480 = the metric's shape; 752x480 = BASELINE config 4)")
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pairs", action="store_true", help="skip the config 2 / 3 pair batches and the config 4 leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    protect_stdout()
    if args.frame_size:
        global W, H
        W, H = (int(x) for x in args.frame_size.lower().split("x"))

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 and os.environ.get("ORB_BENCH_NUMA", "1") != "0" else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner must not precede the JSON line on stdout
        dist.init_process_group("nccl", device_id=dev)
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    import __graft_entry__ as entry
    if not os.path.exists(os.path.join(ROOT, "orb_slam_2_ros_b200", "lib", "liborb_b200.so")):
        if rank == 0:
            entry.build()
        if world > 1:
            dist.barrier()
    from orb_slam_2_ros_b200 import ORBextractor, synth
    from orb_slam_2_ros_b200._lib import KP_DTYPE, TOP2_DTYPE

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    B = args.batch
    # ---- synthetic frames: every rank its own contiguous block of seeds (config 4's sharding) ----
    frames_np = synth.synth_batch(1000 * rank, B, W, H, unique=16, noise=args.noise)
    h_frames = torch.from_numpy(frames_np).pin_memory()
    d_frames = h_frames.to(dev, non_blocking=False)
    ex = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=B)
    cap = ex.max_keypoints
    d_kps = torch.zeros((B, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(dev)        # a real (non-null) stream shared by torch's events and the library
    torch.cuda.set_stream(stream)
    ex.set_stream(stream.cuda_stream)

    def step_device():
        ex.extract_batch_device(d_frames.data_ptr(), B, W, H, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())

    for _ in range(args.warmup):
        step_device()
    barrier()
    ex.profile_enable(True)
    ex.profile_read(reset=True)
    launches0 = ex.launch_count()
    stop, clk = threading.Event(), {}
    sampler = threading.Thread(target=clocks_sampler, args=(stop, clk, local_rank), daemon=True)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    stage_ms, prof_calls, prof_frames = ex.profile_read(reset=True)
    ex.profile_enable(False)
    launches = ex.launch_count() - launches0
    n_kp = int(d_n.sum().item())
    value = world * B * args.steps / (ms_dev * 1e-3)

    # ---- e2e: host frames -> host keypoints + descriptors through the public call ----
    t_kps_h = torch.zeros((B, cap, KP_DTYPE.itemsize), dtype=torch.uint8).pin_memory()   # pinned host outputs
    t_desc_h = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
    kps_h, desc_h = t_kps_h.numpy(), t_desc_h.numpy()
    n_h = np.zeros(B, np.int32)
    from orb_slam_2_ros_b200._lib import check, lib, ptr

    def step_e2e():
        check(lib().orb_extract_batch(ex._h, h_frames.data_ptr(), B, W, H, W, W * H, ptr(kps_h), ptr(desc_h), cap, ptr(n_h)))

    for _ in range(min(args.warmup, 3)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * args.steps / t_e2e

    # ---- the same with TWO extractor instances called from two host threads (the reference's own threading model for
    #      stereo: two ORBextractor instances on two std::threads, Frame.cc:79-82): the pipeline fill / drain of one call
    #      (first chunk's H2D before any kernel, last chunk's kernels + D2H after the last copy) overlaps the other's ----
    e2e2_value = None
    if args.e2e_callers >= 2:
        ex2 = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=B)
        t_kps_h2 = torch.zeros((B, cap, KP_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        t_desc_h2 = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
        kps_h2, desc_h2 = t_kps_h2.numpy(), t_desc_h2.numpy()
        n_h2 = np.zeros(B, np.int32)
        callers = [(ex, kps_h, desc_h, n_h), (ex2, kps_h2, desc_h2, n_h2)]

        def run_caller(idx, nsteps):
            e, k, d, n = callers[idx]
            for _ in range(nsteps):   # ctypes releases the GIL for the duration of the call
                check(lib().orb_extract_batch(e._h, h_frames.data_ptr(), B, W, H, W, W * H, ptr(k), ptr(d), cap, ptr(n)))

        def run_both(total_steps):
            per = [total_steps - total_steps // 2, total_steps // 2]
            th = [threading.Thread(target=run_caller, args=(i, per[i])) for i in range(2)]
            for t in th:
                t.start()
            for t in th:
                t.join()

        run_both(2)
        barrier()
        t0 = time.perf_counter()
        run_both(args.steps)
        torch.cuda.synchronize()
        t_e2e2 = max_over_ranks(time.perf_counter() - t0)
        e2e2_value = world * B * args.steps / t_e2e2
        assert np.array_equal(n_h, n_h2) and np.array_equal(desc_h[0, :int(n_h[0])], desc_h2[0, :int(n_h2[0])])
        del ex2
    h2d = B * W * H
    d2h = int(n_h.sum()) * (KP_DTYPE.itemsize + 32) + 4 * B
    stop.set()

    # ---- what the box's host<->device path gives a plain copy loop: the step's own traffic (B frames in, the keypoints + descriptors
    #      out) as bare cudaMemcpyAsync calls on pinned buffers, 8 chunks each way on two streams, ON ALL RANKS AT THE SAME TIME.
    #      e2e is reported as a fraction of the frame rate this allows (its ceiling on this box at this N). ----
    def pcie_probe(reps=6):
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        nout = max(d2h, 1 << 20)
        ho = torch.empty(nout, dtype=torch.uint8).pin_memory()
        do = torch.empty(nout, dtype=torch.uint8, device=dev)
        hin = h_frames.view(-1)
        din = d_frames.view(-1)
        nin = hin.numel()

        def run(both):
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                with torch.cuda.stream(s_in):
                    c = nin // 8
                    for k in range(8):
                        din[k * c:(k + 1) * c].copy_(hin[k * c:(k + 1) * c], non_blocking=True)
                if both:
                    with torch.cuda.stream(s_out):
                        c = nout // 8
                        for k in range(8):
                            ho[k * c:(k + 1) * c].copy_(do[k * c:(k + 1) * c], non_blocking=True)
            torch.cuda.synchronize()
            return max_over_ranks(time.perf_counter() - t0) / reps
        run(True)
        t_in, t_both = run(False), run(True)
        return {"h2d_gbs_per_gpu": nin / t_in / 1e9, "h2d_gbs_per_gpu_with_d2h": nin / t_both / 1e9, "h2d_gbs_aggregate_with_d2h": world * nin / t_both / 1e9,
                "frames_per_s_ceiling": world * B / t_both, "bytes_in_per_step": nin, "bytes_out_per_step": nout,
                "how": "bare cudaMemcpyAsync on pinned buffers, 8 chunks each way, all %d rank(s) simultaneously, slowest rank" % world}
    probe = pcie_probe()

    # ---- Hamming leg: 2000 queries x db_rows rows sharded contiguously over the ranks ----
    ham = None
    if not args.no_hamming:
        rows_total = args.db_rows
        from orb_slam_2_ros_b200.sharding import shard_range
        r0, r1 = shard_range(rows_total, rank, world)
        # the sharded database is a LIBRARY object: per-shard search, ncclAllGather of the 2000 x 24-byte results and the device merge
        # run inside liborb_b200.so on the shard's stream (orb_db_query_top2_sharded); only the 128-byte NCCL id travels through torch
        from orb_slam_2_ros_b200.matcher import ShardedDescriptorDB, shard_unique_id
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if world > 1:
            if rank == 0:
                uid.copy_(torch.from_numpy(shard_unique_id()))
            dist.broadcast(uid, 0)
        db = ShardedDescriptorDB(max(r1 - r0, 1), r0, rank, world, uid.cpu().numpy() if world > 1 else None, device=local_rank)
        chunk = 1 << 20
        for s in range(r0, r1, chunk):
            db.add(synth.synth_descriptors(DB_SEED, s, min(chunk, r1 - s)))
        q_np, planted, _flips = synth.synth_queries(DB_SEED, rows_total, NQ)
        d_q = torch.from_numpy(q_np).to(dev)
        d_merged = torch.zeros((NQ, TOP2_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        db.set_stream(stream.cuda_stream)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def step_hamming():
            db.query_top2_device(d_q.data_ptr(), NQ, d_merged.data_ptr())   # exact global top-2 on every rank

        for _ in range(args.warmup):
            step_hamming()
        barrier()
        db.profile_enable(True)
        db.profile_read(reset=True)
        hl0 = db.launch_count()
        tot = 0.0
        hsteps = max(3, min(args.steps, 10))
        for _ in range(hsteps):
            flush.zero_()                                       # L2 flush between timed iterations
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step_hamming()
            b.record(stream)
            torch.cuda.synchronize()
            tot += max_over_ranks(a.elapsed_time(b))
        s_ms, m_ms, calls = db.profile_read(reset=True)
        db.profile_enable(False)
        hl = db.launch_count() - hl0
        # result check: the global top-1 of the planted queries must be the planted row (all four fields of all 2000 queries are
        # compared with the oracle at full size by tests/test_gpu_config5.py)
        merged = d_merged.cpu().numpy().view(TOP2_DTYPE).reshape(NQ)
        ok = int((merged["best_idx"][planted >= 0] == planted[planted >= 0]).sum())
        import ctypes
        issue = ctypes.c_double(0)
        check(lib().orb_bench_issue_rate(local_rank, 0, 2000, ctypes.byref(issue)))
        popc_rate = issue.value * 1e9                           # measured POPC issue rate of this GPU (register-only)
        check(lib().orb_bench_issue_rate(local_rank, 1, 2000, ctypes.byref(issue)))
        cmp8_rate = issue.value * 1e9                           # register-only compare with 8 POPC + top-2 update
        check(lib().orb_bench_issue_rate(local_rank, 2, 2000, ctypes.byref(issue)))
        cmp_rate = issue.value * 1e9                            # register-only compare as the kernel does it (carry-save, 4 POPC)
        popc8_peak = popc_rate / 8.0                            # 8 x popc.b32 per 256-bit compare (SURVEY.md §8d's roofline)
        popc4_peak = popc_rate / 4.0                            # the kernel's carry-save popcount needs 4 POPC per compare: its POPC-pipe bound
        kern_cps = NQ * (r1 - r0) / (s_ms / max(calls, 1) * 1e-3) if s_ms > 0 else 0.0
        ham = {
            "metric": "Hamming compares/s", "value": NQ * rows_total * hsteps / (tot * 1e-3), "unit": "compares/s",
            "scaling": "strong", "steps": hsteps, "ms_per_step": tot / hsteps,
            "config": {"workload": "%d queries x %d-row descriptor DB sharded over %d GPU(s): orb_db_query_top2_sharded = per-shard top-2 + "
                                   "ncclAllGather + device merge inside liborb_b200.so" % (NQ, rows_total, world), "l2": "flushed between timed iterations"},
            "planted_top1_found": "%d/%d" % (ok, int((planted >= 0).sum())),
            "roofline": {"bound": "popc pipe (4 POPC per 256-bit compare after the carry-save tree)", "achieved": kern_cps / 1e9,
                         "peak": popc4_peak / 1e9, "unit": "Gcompare/s", "frac": kern_cps / popc4_peak if popc4_peak else None,
                         "peak_source": "POPC issue rate of this GPU measured register-only in this run (orb_bench_issue_rate kind 0) / 4",
                         "popc_per_s": popc_rate, "popc8_roofline_gcompare_s": popc8_peak / 1e9,
                         "frac_vs_popc8_roofline": kern_cps / popc8_peak if popc8_peak else None,
                         "binding_pipe": "ALU pipe (8 LOP3 XOR + 8 LOP3 carry-save adders per compare; the key is packed by IMADs on the FMA pipe "
                                         "and the top-2 min/max runs once per group of 4 rows) next to the XU/POPC pipe; ncu on this kernel: ALU "
                                         "pipe 80 %, XU/POPC pipe 76 %, issue slots 66 % (profiles/r2/ncu_hamming_pipes.txt)",
                         "own_instruction_mix_ceiling_gcompare_s": cmp_rate / 1e9, "register_only_popc8_compare_per_s": cmp8_rate,
                         "kernel": "hamming_top2_kernel", "avg_launch_ms": s_ms / max(calls, 1)},
            "gpu_launches": int(hl),   # search + split-merge kernels (+ the cross-shard merge kernel when sharded)
        }

    # ---- BASELINE configs 2 and 3 as BATCHES OF PAIRS, device-resident, sharded over the ranks like frames:
    #      config 2 = two extractions + 1000 x 1000 brute-force match (ratio test, rotation histogram) + SearchByProjection of A's
    #      keypoints in B per pair; config 3 = two 1241x376 / 2000-feature extractions + Frame::ComputeStereoMatches per pair ----
    def pairs_legs():
        from orb_slam_2_ros_b200._lib import SearchBatch
        from orb_slam_2_ros_b200.matcher import MODE_TRACK_LAST, match_bruteforce_batch_device, search_by_projection_batch_device
        from orb_slam_2_ros_b200.stereo import ComputeStereoMatchesBatchDevice
        out = {}
        steps = max(3, min(args.steps, 5))

        def timed(fn):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(steps):
                fn()
            b.record(stream)
            barrier()
            return max_over_ranks(a.elapsed_time(b)) / steps

        def dev_outputs(P, cap_):
            return (torch.zeros((P, cap_, KP_DTYPE.itemsize), dtype=torch.uint8, device=dev), torch.zeros((P, cap_, 32), dtype=torch.uint8, device=dev),
                    torch.zeros(P, dtype=torch.int32, device=dev))

        # ---------------- config 2 ----------------
        P = 128
        A = synth.synth_batch(5000 + 1000 * rank, P, 640, 480, unique=16, noise=args.noise)
        Bf = np.stack([synth.shifted_frame(A[i], 3, -2, 7000 + i) for i in range(16)])
        Bf = Bf[np.arange(P) % 16]
        dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(np.ascontiguousarray(Bf)).to(dev)
        exa = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=P)
        exb = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=P)
        exa.set_stream(stream.cuda_stream); exb.set_stream(stream.cuda_stream)
        c2 = exa.max_keypoints
        ka, da, na = dev_outputs(P, c2)
        kb, db_, nb = dev_outputs(P, c2)

        def extract2():
            exa.extract_batch_device(dA.data_ptr(), P, 640, 480, 640, 640 * 480, ka.data_ptr(), da.data_ptr(), c2, na.data_ptr())
            exb.extract_batch_device(dB.data_ptr(), P, 640, 480, 640, 640 * 480, kb.data_ptr(), db_.data_ptr(), c2, nb.data_ptr())
        extract2()
        torch.cuda.synchronize()
        # the caller's side of SearchByProjection (Tracking projects the last frame's points): A's keypoints shifted by the known motion,
        # window 15 * scale[octave], octave band +-1 — computed once, the frames are the same every step
        kf = ka.view(torch.float32).view(P, c2, 7)
        octv = ka.view(torch.int32).view(P, c2, 7)[..., 5].clamp(0, NLEVELS - 1)
        sc = torch.from_numpy(exa.GetScaleFactors()).to(dev)
        q = dict(u=(kf[..., 0] + 3.0).contiguous(), v=(kf[..., 1] - 2.0).contiguous(), r=(15.0 * sc[octv.long()]).contiguous(),
                 lo=(octv - 1).to(torch.int32).contiguous(), hi=(octv + 1).to(torch.int32).contiguous(), ang=kf[..., 3].contiguous())
        taken = torch.zeros((P, c2), dtype=torch.uint8, device=dev)
        moq = torch.zeros((P, c2), dtype=torch.int32, device=dev); tq = torch.zeros_like(moq)
        nm_s = torch.zeros(P, dtype=torch.int32, device=dev); nm_b = torch.zeros_like(nm_s); m12 = torch.zeros_like(moq)
        sb = SearchBatch(kb.data_ptr(), db_.data_ptr(), None, nb.data_ptr(), c2, taken.data_ptr(), na.data_ptr(), c2, q["u"].data_ptr(), q["v"].data_ptr(),
                         q["r"].data_ptr(), q["lo"].data_ptr(), q["hi"].data_ptr(), da.data_ptr(), None, None, q["ang"].data_ptr(), None, None,
                         moq.data_ptr(), tq.data_ptr(), nm_s.data_ptr())

        # the two searches of a pair are independent of each other (both only read the extractor's outputs), so the caller runs them on two
        # streams: the brute-force resolve step (one CTA per pair, latency-bound) then overlaps the window search's candidate kernel
        stream_b = torch.cuda.Stream(device=dev)

        def match2():
            stream_b.wait_stream(stream)
            match_bruteforce_batch_device(P, ka.data_ptr(), da.data_ptr(), na.data_ptr(), c2, kb.data_ptr(), db_.data_ptr(), nb.data_ptr(), c2,
                                          m12.data_ptr(), nm_b.data_ptr(), 50, 0.6, True, device=local_rank, stream=stream.cuda_stream)
            with torch.cuda.stream(stream_b):
                taken.zero_()
                search_by_projection_batch_device(MODE_TRACK_LAST, P, sb, (0.0, 0.0, 640.0, 480.0), 100, 0.9, True, device=local_rank,
                                                  stream=stream_b.cuda_stream)
            stream.wait_stream(stream_b)

        def step2():
            extract2()
            match2()
        ms_all, ms_match = timed(step2), timed(match2)
        torch.cuda.synchronize()
        out["config2"] = {"metric": "frame pairs/s: 2 x ORBextractor 640x480 + 1000x1000 brute-force match (ratio 0.6, rot. hist.) + SearchByProjection (th 15)",
                          "value": world * P / (ms_all * 1e-3), "unit": "pairs/s", "pairs_per_step_per_gpu": P, "ms_per_step": ms_all,
                          "matching_only_ms_per_step": ms_match, "matching_only_pairs_per_s": world * P / (ms_match * 1e-3), "scaling": "weak",
                          "bruteforce_matches_per_pair": float(nm_b.float().mean().item()), "projection_matches_per_pair": float(nm_s.float().mean().item()),
                          "entry_points": "orb_extract_batch_device x2, orb_match_bruteforce_batch_device || orb_search_by_projection_batch_device on two streams (device-resident)"}
        del exa, exb, ka, da, kb, db_, dA, dB
        # ---------------- config 3 ----------------
        P3, w3, h3, nf3 = 64, 1241, 376, 2000
        prs = [synth.synth_stereo_pair(9000 + 100 * rank + i, w3, h3)[:2] for i in range(8)]
        L = np.stack([prs[i % 8][0] for i in range(P3)]); R = np.stack([prs[i % 8][1] for i in range(P3)])
        dL, dR = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
        exl = ORBextractor(nf3, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=P3)
        exr = ORBextractor(nf3, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=P3)
        exl.set_stream(stream.cuda_stream); exr.set_stream(stream.cuda_stream)
        c3 = exl.max_keypoints
        kl, dl, nl = dev_outputs(P3, c3)
        kr, dr, nr = dev_outputs(P3, c3)
        ur = torch.zeros((P3, c3), dtype=torch.float32, device=dev); dep = torch.zeros_like(ur)
        nm3 = torch.zeros(P3, dtype=torch.int32, device=dev)
        bf3 = float(np.float32(386.1448)); b3 = float(np.float32(386.1448) / np.float32(718.856))

        def stereo_only():
            ComputeStereoMatchesBatchDevice(exl, exr, P3, kl.data_ptr(), dl.data_ptr(), nl.data_ptr(), kr.data_ptr(), dr.data_ptr(), nr.data_ptr(), c3, bf3, b3,
                                            ur.data_ptr(), dep.data_ptr(), nm3.data_ptr())

        def step3():
            exl.extract_batch_device(dL.data_ptr(), P3, w3, h3, w3, w3 * h3, kl.data_ptr(), dl.data_ptr(), c3, nl.data_ptr())
            exr.extract_batch_device(dR.data_ptr(), P3, w3, h3, w3, w3 * h3, kr.data_ptr(), dr.data_ptr(), c3, nr.data_ptr())
            stereo_only()
        ms3, ms3m = timed(step3), timed(stereo_only)
        torch.cuda.synchronize()
        out["config3"] = {"metric": "stereo pairs/s: 2 x ORBextractor 1241x376 nFeatures=2000 + Frame::ComputeStereoMatches", "value": world * P3 / (ms3 * 1e-3),
                          "unit": "pairs/s", "pairs_per_step_per_gpu": P3, "ms_per_step": ms3, "stereo_matching_only_ms_per_step": ms3m,
                          "stereo_matching_only_pairs_per_s": world * P3 / (ms3m * 1e-3), "scaling": "weak",
                          "stereo_matches_per_pair": float(nm3.float().mean().item()),
                          "entry_points": "orb_extract_batch_device x2, orb_stereo_match_batch_device (device-resident)"}
        return out

    # ---- BASELINE config 4 as specified: 4096 frames of 752x480 STRONG-sharded over the ranks (contiguous blocks of 4096 / N) ----
    def config4_leg():
        total, w4, h4 = 4096, 752, 480
        per = total // world
        b4 = min(512, per)
        nb4 = per // b4
        fr = synth.synth_batch(20000 + 1000 * rank, b4, w4, h4, unique=16, noise=args.noise)
        hf = torch.from_numpy(fr).pin_memory()
        df = hf.to(dev)
        e4 = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=b4)
        e4.set_stream(stream.cuda_stream)
        c4 = e4.max_keypoints
        k4 = torch.zeros((b4, c4, KP_DTYPE.itemsize), dtype=torch.uint8, device=dev); d4 = torch.zeros((b4, c4, 32), dtype=torch.uint8, device=dev)
        n4 = torch.zeros(b4, dtype=torch.int32, device=dev)

        def dev_pass():
            for _ in range(nb4):
                e4.extract_batch_device(df.data_ptr(), b4, w4, h4, w4, w4 * h4, k4.data_ptr(), d4.data_ptr(), c4, n4.data_ptr())
        for _ in range(2):
            dev_pass()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        dev_pass()
        b.record(stream)
        barrier()
        ms = max_over_ranks(a.elapsed_time(b))
        kh = torch.zeros((b4, c4, KP_DTYPE.itemsize), dtype=torch.uint8).pin_memory(); dh = torch.zeros((b4, c4, 32), dtype=torch.uint8).pin_memory()
        nh = np.zeros(b4, np.int32)

        def host_pass():
            for _ in range(nb4):
                check(lib().orb_extract_batch(e4._h, hf.data_ptr(), b4, w4, h4, w4, w4 * h4, ptr(kh.numpy()), ptr(dh.numpy()), c4, ptr(nh)))
        host_pass()
        barrier()
        t0 = time.perf_counter()
        host_pass()
        torch.cuda.synchronize()
        te = max_over_ranks(time.perf_counter() - t0)
        return {"metric": "ORB extract frames/s, 4096 frames of 752x480 (EuRoC shape) sharded over the GPUs", "value": total / (ms * 1e-3), "unit": "frames/s",
                "scaling": "strong", "frames_total": total, "frames_per_gpu": per, "batch": b4, "ms_total": ms,
                "e2e": {"value": total / te, "unit": "frames/s", "h2d_bytes": per * w4 * h4, "callers": 1,
                        "note": "host pinned frames -> host keypoints + descriptors, one caller per rank"}}

    pairs = pairs_legs() if not args.no_pairs else None
    config4 = config4_leg() if not args.no_pairs else None

    def content_leg():
        """The same device-resident step on frames with fewer corner sources (40 rectangles + 20 triangles, +-2 noise instead of
        260 + 120, +-8): FAST's time follows the number of candidate pixels (the exact scorer, NMS and emission run per candidate),
        the other stages do not care.  Raw NMS corners of frame 0 (all levels) are reported for both contents."""
        step_device()
        barrier()
        raw_default = sum(len(ex.debug_raw_corners(l)) for l in range(NLEVELS))
        sparse = synth.synth_batch(500000 + 1000 * rank, B, W, H, unique=16, noise=2, n_rect=40, n_tri=20)
        d_frames.copy_(torch.from_numpy(sparse))
        for _ in range(3):
            step_device()
        barrier()
        ex.profile_enable(True)
        ex.profile_read(reset=True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(5):
            step_device()
        s1.record(stream)
        barrier()
        ms = max_over_ranks(s0.elapsed_time(s1)) / 5
        st, calls, _ = ex.profile_read(reset=True)
        ex.profile_enable(False)
        raw_sparse = sum(len(ex.debug_raw_corners(l)) for l in range(NLEVELS))
        d_frames.copy_(h_frames)
        return {"workload": "same step, frames with 40 rectangles + 20 triangles and +-2 noise (fewer corner sources)", "ms_per_step": ms,
                "frames_per_s": world * B / (ms * 1e-3), "stage_ms": {k: v / max(calls, 1) for k, v in st.items()},
                "raw_nms_corners_frame0": raw_sparse, "raw_nms_corners_frame0_headline_content": raw_default,
                "keypoints_per_frame": float(d_n.float().mean().item())}

    content = content_leg() if not args.no_pairs else None

    sampler.join(timeout=2.0)
    if clk.get("proc"):
        clk["proc"].kill()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback")
        ab = algorithmic_bytes(W, H)
        traffic = {}
        try:   # DRAM bytes per frame and stage from the committed `ncu --set full` capture of one bench step
            import glob
            tf = sorted(glob.glob(os.path.join(ROOT, "profiles", "**", "*_traffic.json"), recursive=True))[-1]
            traffic = {k: v["dram_bytes_per_frame"] * B for k, v in json.load(open(tf))["stages"].items()}
            traffic["_source"] = os.path.relpath(tf, ROOT)
        except (IndexError, OSError, KeyError, ValueError):
            pass
        pipes = {}
        try:   # ALU / FMA / LSU / issue utilisation per stage from the committed ncu capture (tools/ncu_pipes_json.py): names the binding pipe
            import glob
            pf = sorted(glob.glob(os.path.join(ROOT, "profiles", "**", "*_pipes.json"), recursive=True))[-1]
            pj = json.load(open(pf))
            pipes = pj["stages"]
            pipes["_source"] = os.path.relpath(pf, ROOT)
        except (IndexError, OSError, KeyError, ValueError):
            pass
        per_call = {k: v / max(prof_calls, 1) for k, v in stage_ms.items()}
        stages = {}
        for k in ("pyramid", "fast_cells", "blur"):
            gbs = ab[k] * B / (per_call[k] * 1e-3) / 1e9 if per_call[k] > 0 else 0.0
            stages[k] = {"ms": per_call[k], "alg_bytes": ab[k] * B, "gbs": gbs, "frac": gbs / hbm_peak, "traffic": traffic.get(k)}
        for k in ("quadtree", "orient_describe"):
            stages[k] = {"ms": per_call[k], "traffic": traffic.get(k)}
        dom = max(("pyramid", "fast_cells", "blur", "quadtree", "orient_describe"), key=lambda k: per_call[k])
        total_alg = sum(ab.values()) * B
        sum_ms = sum(per_call.values())
        if dom in ab:
            roof = {"bound": "hbm", "achieved": stages[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": stages[dom]["frac"],
                    "traffic": traffic.get(dom), "kernel": dom, "peak_source": peak_src, "avg_launch_ms": per_call[dom]}
        else:
            # the dominant stage moves no image-sized data: report the whole step against the unfused-stage byte model
            gbs = total_alg / (sum_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                    "kernel": dom + " (latency-bound stage; achieved = all stages' algorithmic bytes / step time)",
                    "peak_source": peak_src, "avg_launch_ms": per_call[dom]}
        for k in stages:
            if k in pipes:
                stages[k]["pipes"] = pipes[k]
        if dom in pipes:
            roof["binding_pipe"] = {"stage": dom, "utilisation": pipes[dom], "source": pipes.get("_source"),
                                    "note": "integer stencil: the stage is bounded by instruction issue on the named pipe, not by HBM (DESIGN.md §4)"}
        roof["traffic_source"] = traffic.get("_source")
        roof["stages"] = stages
        roof["step_alg_bytes"] = total_alg
        roof["step_gbs"] = total_alg / (ms_dev / args.steps * 1e-3) / 1e9
        roof["step_frac"] = roof["step_gbs"] / hbm_peak                    # whole step against the unfused-stage byte model
        # SURVEY.md §8(d)'s second denominator, the FUSED lower bound: read the image once, write the public bordered pyramid
        # once, write the outputs (~60 KB of keypoints + descriptors per frame)
        dims_ = level_dims(W, H)
        fused = W * H + sum((a + 38) * (b + 38) for a, b in dims_) + n_kp // B * 60
        roof["step_fused_bound_bytes"] = fused * B
        roof["step_frac_fused_bound"] = fused * B / (ms_dev / args.steps * 1e-3) / 1e9 / hbm_peak

        cpu = None
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = frames_np[:max(threads, 16)]
            kind, harness = cpu_arm()
            fps1, dt1, n1 = cpu_extract_throughput(sample[:threads], threads, harness=harness)      # calibration pass (also warm-up)
            repeat = max(1, int(args.cpu_seconds * fps1 / len(sample)))
            fps, dt, n = cpu_extract_throughput(sample, threads, repeat=repeat, harness=harness)
            cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                   "sample": "%d extractions of %d distinct synthetic frames in %.1f s, one %s per host thread; scalar OpenCV-4.13-exact "
                             "primitives (no SIMD: an OpenCV build with SIMD kernels is roughly 2x faster per frame)"
                             % (n, len(sample), dt, "reference ORBextractor (unmodified ORBextractor.cc, oracle/_ref)" if kind == "reference" else "oracle extractor"),
                   "per_frame_ms_one_thread": 1e3 * threads / fps,
                   "cv2_simd_primitives": cv2_simd_primitive_row(frames_np[0])}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "ORBextractor %dx%d nFeatures=1000 8 levels 1.2 20/7 (BASELINE config 1), batch of %d frames "
                                   "per GPU per step, synthetic frames with +-%d grey-level noise" % (W, H, B, args.noise), "frames_per_step_per_gpu": B, "keypoints_per_step_per_gpu": n_kp,
                       "l2": "inputs larger than L2 (%d MB of frames + %d MB pyramid arena per step)" % (B * W * H >> 20, (B * 1158012) >> 20),
                       "parallelism": "frames sharded over %d GPU(s), no collective" % world,
                       "host_affinity": ("rank 0 bound to the cpus next to its GPU: %s" % numa_cpus) if numa_cpus else "unbound"},
            "e2e": {"value": e2e2_value if e2e2_value else e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "callers": 2 if e2e2_value else 1, "one_caller_value": e2e_value,
                    "copy_probe": probe, "fraction_of_copy_ceiling": (e2e2_value if e2e2_value else e2e_value) / probe["frames_per_s_ceiling"],
                    "note": "every step = one blocking orb_extract_batch of the whole batch (pinned host frames in, pinned host keypoints + "
                            "descriptors out, H2D and D2H inside).  value: the steps are issued by two ORBextractor instances on two host "
                            "threads (the reference's threading model, Frame.cc:79-82; the CPU arm runs one extractor per host thread), so "
                            "one call's pipeline drain overlaps the other's copies; one_caller_value: a single caller, calls back to back"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "hamming": ham,
            "pairs": pairs,
            "config4": config4,
            "content_sensitivity": content,
            "other_configs": other_configs(local_rank) if (world == 1 and not args.no_cpu) else None,
            "clocks": summarize_clocks(clk.get("rows")),
        }
        emit_json_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
