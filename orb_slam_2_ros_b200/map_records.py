"""Host mirror of the map-file record payloads (reference BoostArchiver.h:46-91): the cv::Mat and
cv::KeyPoint records a KeyFrame writes into System::SaveMap's boost binary archive (KeyFrame.cc:858-864), and the loader
that appends a decoded descriptor matrix to a device-resident DescriptorDB shard.  All work happens in liborb_b200.so."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib, ptr

CV_8U = 0   # cv::Mat::type() of mDescriptors (CV_8UC1)
KP_RECORD_BYTES = 28


def encode_mat_record(mat, elem_type=None):
    """bytes of `ar & mat` (BoostArchiver.h:61-76): cols, rows, elemSize, type, data."""
    mat = np.ascontiguousarray(mat)   # the reference clones non-continuous matrices first (:65-66)
    if mat.ndim != 2:
        raise ValueError("2-D matrix expected")
    rows, cols = mat.shape
    es = mat.dtype.itemsize
    if elem_type is None:
        elem_type = {np.dtype(np.uint8): 0, np.dtype(np.int8): 1, np.dtype(np.uint16): 2, np.dtype(np.int16): 3,
                     np.dtype(np.int32): 4, np.dtype(np.float32): 5, np.dtype(np.float64): 6}[mat.dtype]
    need = C.c_size_t()
    check(lib().orb_mat_record_bytes(rows, cols, es, C.byref(need)))
    out = np.zeros(need.value, np.uint8)
    wr = C.c_size_t()
    check(lib().orb_mat_record_encode(ptr(mat) if mat.size else None, rows, cols, es, elem_type, ptr(out), out.size, C.byref(wr)))
    return out[:wr.value].tobytes()


def decode_mat_record(buf, offset=0):
    """-> (matrix as a numpy array, bytes consumed); `load` of BoostArchiver.h:78-91."""
    raw = np.frombuffer(buf, np.uint8)[offset:]
    rows, cols = C.c_int32(), C.c_int32()
    es, et, used = C.c_size_t(), C.c_size_t(), C.c_size_t()
    data = C.c_void_p()
    check(lib().orb_mat_record_decode(ptr(raw) if raw.size else None, raw.size, C.byref(rows), C.byref(cols), C.byref(es),
                                      C.byref(et), C.byref(data), C.byref(used)))
    dt = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}.get(et.value & 7, np.uint8)
    start = used.value - rows.value * cols.value * es.value
    mat = raw[start:used.value].view(dt if np.dtype(dt).itemsize == es.value else np.uint8)
    return mat.reshape(rows.value, cols.value).copy(), used.value


def encode_keypoint_records(kps):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.zeros(len(kps) * KP_RECORD_BYTES, np.uint8)
    check(lib().orb_keypoint_records_encode(ptr(kps) if len(kps) else None, len(kps), ptr(out) if len(kps) else None))
    return out.tobytes()


def decode_keypoint_records(buf, n, offset=0):
    raw = np.frombuffer(buf, np.uint8)[offset:offset + n * KP_RECORD_BYTES]
    if raw.size < n * KP_RECORD_BYTES:
        raise ValueError("truncated keypoint records")
    kps = np.zeros(n, KP_DTYPE)
    check(lib().orb_keypoint_records_decode(ptr(raw) if n else None, n, ptr(kps) if n else None))
    return kps


def db_add_mat_record(db, buf, offset=0):
    """Append the N x 32 descriptor matrix stored at buf[offset:] to the shard; -> (rows added, bytes consumed)."""
    raw = np.frombuffer(buf, np.uint8)[offset:]
    used, rows = C.c_size_t(), C.c_int64()
    check(lib().orb_db_add_mat_record(db._h, ptr(raw) if raw.size else None, raw.size, C.byref(used), C.byref(rows)))
    return rows.value, used.value
