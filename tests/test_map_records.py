"""cv::Mat / cv::KeyPoint record payloads of the reference's saved map (BoostArchiver.h:46-91) —
host codecs of the C ABI against the oracle's struct restatement (CPU), and the descriptor-matrix loader of the device
shard (GPU)."""
import numpy as np
import pytest

from oracle import orb_oracle
from orb_slam_2_ros_b200 import map_records as mr
from orb_slam_2_ros_b200._lib import KP_DTYPE, OrbError


def _kps(n, seed=0):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    k["x"], k["y"] = rng.uniform(0, 640, n), rng.uniform(0, 480, n)
    k["size"] = 31.0 * 1.2 ** rng.integers(0, 8, n)
    k["angle"], k["response"] = rng.uniform(0, 360, n), rng.integers(7, 200, n)
    k["octave"], k["class_id"] = rng.integers(0, 8, n), -1
    return k


@pytest.mark.parametrize("rows,cols,dtype,cvtype", [(1000, 32, np.uint8, 0), (0, 32, np.uint8, 0), (1, 32, np.uint8, 0),
                                                      (4, 4, np.float32, 5), (3, 3, np.float64, 6)])
def test_mat_record_matches_oracle_and_round_trips(rows, cols, dtype, cvtype):
    rng = np.random.default_rng(rows + cols)
    mat = (rng.integers(0, 256, (rows, cols)).astype(dtype) if dtype == np.uint8 else rng.standard_normal((rows, cols)).astype(dtype))
    enc = mr.encode_mat_record(mat)
    assert enc == orb_oracle.mat_record_encode(mat, cvtype)            # byte-exact against the restatement
    assert enc[:8] == np.array([cols, rows], "<i4").tobytes()           # cols BEFORE rows (BoostArchiver.h:68-69)
    back, used = mr.decode_mat_record(enc + b"tail")
    assert used == len(enc) and back.dtype == mat.dtype and np.array_equal(back.reshape(rows, cols), mat)
    r, c, es, et, data, n = orb_oracle.mat_record_decode(enc)
    assert (r, c, es, et, n) == (rows, cols, mat.dtype.itemsize, cvtype, len(enc)) and np.array_equal(data, mat.view(np.uint8).ravel())


def test_mat_record_non_contiguous_is_cloned_and_bad_input_rejected():
    big = np.arange(64 * 64, dtype=np.uint8).reshape(64, 64)
    view = big[::2, :32]                                                # not continuous: the reference clones first
    assert mr.encode_mat_record(view) == orb_oracle.mat_record_encode(np.ascontiguousarray(view), 0)
    enc = mr.encode_mat_record(big[:10, :32])
    for cut in (0, 5, 23, 24 + 31, len(enc) - 1):
        if cut == 0:
            continue
        with pytest.raises(OrbError):
            mr.decode_mat_record(enc[:cut])


def test_keypoint_records_keep_the_reference_size_bug():
    k = _kps(257)
    enc = mr.encode_keypoint_records(k)
    assert len(enc) == 28 * len(k) and enc == orb_oracle.keypoint_records_encode(k)
    # response is written twice, size never (BoostArchiver.h:49-57)
    rec = np.frombuffer(enc, "<f4").reshape(-1, 7)
    assert np.array_equal(rec[:, 3], k["response"]) and np.array_equal(rec[:, 4], k["response"])
    back = mr.decode_keypoint_records(enc, len(k))
    ob = orb_oracle.keypoint_records_decode(enc, len(k))
    for f in KP_DTYPE.names:
        assert np.array_equal(back[f], ob[f]), f
    assert np.all(back["size"] == 0) and np.any(k["size"] != 0)         # a loaded map has KeyPoint.size = 0
    for f in ("x", "y", "angle", "response", "octave", "class_id"):
        assert np.array_equal(back[f], k[f]), f
    assert mr.encode_keypoint_records(k[:0]) == b"" and len(mr.decode_keypoint_records(b"", 0)) == 0
    with pytest.raises(ValueError):
        mr.decode_keypoint_records(enc[:-1], len(k))


@pytest.mark.gpu
def test_db_loads_descriptor_records_from_a_map_stream():
    from orb_slam_2_ros_b200 import DescriptorDB, synth
    rows = [700, 0, 1, 2049]                                           # four keyframes' mDescriptors, one of them empty
    mats = [synth.synth_descriptors(7, 10_000 * i, n) for i, n in enumerate(rows)]
    stream = b"".join(b"\x01\x02\x03" * i + mr.encode_mat_record(m) for i, m in enumerate(mats))   # records between other archive bytes
    db_a, db_b = DescriptorDB(sum(rows), index_base=5), DescriptorDB(sum(rows), index_base=5)
    off = 0
    for i, m in enumerate(mats):
        off += 3 * i
        added, used = mr.db_add_mat_record(db_a, stream, off)
        assert added == len(m)
        off += used
        db_b.add(m)
    assert off == len(stream) and len(db_a) == len(db_b) == sum(rows)
    q = np.concatenate([mats[0][:50], mats[3][-50:], synth.synth_descriptors(9, 0, 100)])
    ta, tb = db_a.query_top2(q), db_b.query_top2(q)
    otop = orb_oracle.hamming_top2(q, np.concatenate(mats))
    for f in ("best_dist", "second_dist", "best_idx"):
        assert np.array_equal(ta[f], tb[f]), f
    assert np.array_equal(ta["best_idx"], otop["best_idx"] + 5) and np.array_equal(ta["best_dist"], otop["best_dist"])
    with pytest.raises(OrbError):                                       # a 4x4 float matrix is not a descriptor matrix
        mr.db_add_mat_record(db_a, mr.encode_mat_record(np.eye(4, dtype=np.float32)))
    with pytest.raises(OrbError):                                       # capacity
        mr.db_add_mat_record(db_a, mr.encode_mat_record(mats[2]))


def test_mat_record_decode_rejects_overflowing_and_inconsistent_headers():
    """Map files are untrusted input: a header whose cols * rows * elemSize wraps 64 bits, or whose elemSize
    does not match its type, must be rejected instead of describing more data than the buffer holds."""
    import ctypes as C
    import struct
    from orb_slam_2_ros_b200._lib import lib
    L = lib()
    L.orb_mat_record_decode.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 6
    def decode(buf):
        b = (C.c_uint8 * len(buf)).from_buffer_copy(buf)
        return L.orb_mat_record_decode(b, len(buf), None, None, None, None, None, None)
    ok = struct.pack("<iiQQ", 32, 2, 1, 0) + bytes(64)
    assert decode(ok) == 0
    assert decode(struct.pack("<iiQQ", 2**31 - 1, 2**31 - 1, 64, 6 | (7 << 3)) + bytes(64)) != 0      # product wraps 2^64
    assert decode(struct.pack("<iiQQ", 2**31 - 1, 2**31 - 1, 4, 5) + bytes(64)) != 0                 # 2^64-ish payload, 64 bytes there
    assert decode(struct.pack("<iiQQ", 32, 2, 2, 0) + bytes(128)) != 0                               # elemSize 2 for CV_8UC1
    assert decode(struct.pack("<iiQQ", 32, 3, 1, 0) + bytes(64)) != 0                                # truncated payload
