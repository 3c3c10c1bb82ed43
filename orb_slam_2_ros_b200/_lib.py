"""ctypes binding of liborb_b200.so (include/orb_b200.h).  The product has NO CPU fallback: if the shared
library is missing this module raises, and every compute call returns ORB_ERR_NO_DEVICE without a GPU."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liborb_b200.so")

ORB_OK, ORB_ERR_INVALID, ORB_ERR_CUDA, ORB_ERR_CAPACITY, ORB_ERR_NO_DEVICE, ORB_ERR_TOO_SMALL = 0, -1, -2, -3, -4, -5

PIX_GRAY8, PIX_BGR8, PIX_RGB8, PIX_BGRA8, PIX_RGBA8 = 0, 1, 2, 3, 4

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
TOP2_DTYPE = np.dtype([("best_dist", "<i4"), ("second_dist", "<i4"), ("best_idx", "<i8"), ("second_idx", "<i8")])
assert KP_DTYPE.itemsize == 28 and TOP2_DTYPE.itemsize == 24


class SearchParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("th_dist", C.c_int32), ("nn_ratio", C.c_float), ("check_orientation", C.c_int32),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float)]


class SearchBatch(C.Structure):
    """orb_search_batch (include/orb_b200.h): device pointers of a batch of (target frame, query set) pairs"""
    _fields_ = [("d_kps_un", C.c_void_p), ("d_desc", C.c_void_p), ("d_u_right", C.c_void_p), ("d_n", C.c_void_p), ("cap_n", C.c_int32),
                ("d_taken", C.c_void_p), ("d_nq", C.c_void_p), ("cap_q", C.c_int32),
                ("d_q_u", C.c_void_p), ("d_q_v", C.c_void_p), ("d_q_radius", C.c_void_p), ("d_q_min_level", C.c_void_p), ("d_q_max_level", C.c_void_p),
                ("d_q_desc", C.c_void_p), ("d_q_ur", C.c_void_p), ("d_q_er_max", C.c_void_p), ("d_q_angle", C.c_void_p), ("d_q_valid", C.c_void_p),
                ("d_q_obs", C.c_void_p), ("d_match_of_query", C.c_void_p), ("d_target_query", C.c_void_p), ("d_nmatches", C.c_void_p)]


class OrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("liborb_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol include/orb_b200.h declares (tests check that the .so exports all of them)
EXPORTS = [
    "orb_last_error", "orb_device_count", "orb_version", "orb_create", "orb_destroy", "orb_get_tables",
    "orb_max_keypoints", "orb_extract", "orb_extract_batch", "orb_extract_batch_device", "orb_extract_batch_pix", "orb_extract_batch_device_pix", "orb_set_stream", "orb_sync",
    "orb_launch_count", "orb_profile_enable", "orb_profile_read", "orb_level_dims", "orb_pyramid_level", "orb_pyramid_levels", "orb_debug_blurred", "orb_debug_raw_corners",
    "orb_debug_tie_counts", "orb_hamming_top2", "orb_hamming_top2_csr", "orb_db_create", "orb_db_destroy", "orb_db_add", "orb_db_add_device",
    "orb_db_size", "orb_db_set_stream", "orb_db_query_top2", "orb_db_query_top2_device", "orb_db_launch_count", "orb_db_profile_enable", "orb_db_profile_read",
    "orb_top2_merge", "orb_top2_merge_device", "orb_search_by_projection", "orb_match_bruteforce", "orb_stereo_match", "orb_stereo_match_batch_device", "orb_search_by_projection_batch_device", "orb_match_bruteforce_batch_device",
    "orb_search_by_bow", "orb_search_for_triangulation", "orb_search_by_sim3", "orb_distinctive_descriptors", "orb_fuse_search", "orb_voc_create", "orb_voc_load_text", "orb_voc_destroy", "orb_voc_info", "orb_bow_transform_features",
    "orb_bow_transform_features_device", "orb_bow_transform", "orb_bow_transform_device",
    "orb_mat_record_bytes", "orb_mat_record_encode", "orb_mat_record_decode", "orb_keypoint_records_encode",
    "orb_keypoint_records_decode", "orb_db_add_mat_record",
    "orb_shard_unique_id", "orb_db_create_sharded", "orb_db_query_top2_sharded", "orb_db_query_top2_sharded_host", "orb_bench_issue_rate", "orb_debug_sincos_range",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
    L.orb_last_error.restype = C.c_char_p
    L.orb_create.argtypes = [C.POINTER(vp), i32, f32, i32, i32, i32, i32, i32]
    L.orb_destroy.argtypes = [vp]
    L.orb_destroy.restype = None
    L.orb_get_tables.argtypes = [vp] * 6
    L.orb_max_keypoints.argtypes = [vp]
    L.orb_extract.argtypes = [vp, vp, i32, i32, sz, vp, vp, i32, C.POINTER(i32)]
    L.orb_extract_batch.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp]
    L.orb_extract_batch_device.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp]
    L.orb_extract_batch_pix.argtypes = [vp, vp, i32, i32, i32, i32, sz, sz, vp, vp, i32, vp]
    L.orb_extract_batch_device_pix.argtypes = [vp, vp, i32, i32, i32, i32, sz, sz, vp, vp, i32, vp]
    L.orb_set_stream.argtypes = [vp, vp]
    L.orb_sync.argtypes = [vp]
    L.orb_launch_count.argtypes = [vp]
    L.orb_launch_count.restype = i64
    L.orb_profile_enable.argtypes = [vp, i32]
    L.orb_profile_read.argtypes = [vp, vp, C.POINTER(i64), C.POINTER(i64), i32]
    L.orb_level_dims.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.orb_pyramid_level.argtypes = [vp, i32, i32, vp, sz]
    L.orb_pyramid_levels.argtypes = [vp, i32, vp, sz, vp, vp, vp]
    L.orb_debug_blurred.argtypes = [vp, i32, i32, vp, sz]
    L.orb_debug_raw_corners.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32)]
    L.orb_debug_tie_counts.argtypes = [vp, i32, vp]
    L.orb_hamming_top2.argtypes = [i32, vp, i32, vp, i64, vp]
    L.orb_hamming_top2_csr.argtypes = [i32, vp, i32, vp, i64, vp, vp, vp]
    L.orb_db_create.argtypes = [C.POINTER(vp), i32, i64, i64]
    L.orb_db_destroy.argtypes = [vp]
    L.orb_db_destroy.restype = None
    L.orb_db_add.argtypes = [vp, vp, i64]
    L.orb_db_add_device.argtypes = [vp, vp, i64]
    L.orb_db_size.argtypes = [vp]
    L.orb_db_size.restype = i64
    L.orb_db_set_stream.argtypes = [vp, vp]
    L.orb_db_query_top2.argtypes = [vp, vp, i32, vp]
    L.orb_db_query_top2_device.argtypes = [vp, vp, i32, vp]
    L.orb_db_launch_count.argtypes = [vp]
    L.orb_db_launch_count.restype = i64
    L.orb_db_profile_enable.argtypes = [vp, i32]
    L.orb_db_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64), i32]
    L.orb_top2_merge.argtypes = [vp, i32, i32, vp]
    L.orb_top2_merge_device.argtypes = [i32, vp, i32, i32, vp, vp]
    L.orb_search_by_projection.argtypes = [i32, C.POINTER(SearchParams), vp, vp, vp, i32, vp, i32] + [vp] * 13 + [C.POINTER(i32)]
    L.orb_match_bruteforce.argtypes = [i32, vp, vp, i32, vp, vp, i32, i32, f32, i32, vp, C.POINTER(i32)]
    L.orb_stereo_match.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, f32, f32, vp, vp, C.POINTER(i32)]
    L.orb_search_by_projection_batch_device.argtypes = [i32, C.POINTER(SearchParams), i32, C.POINTER(SearchBatch), vp]
    L.orb_match_bruteforce_batch_device.argtypes = [i32, i32, vp, vp, vp, i32, vp, vp, vp, i32, i32, f32, i32, vp, vp, vp]
    L.orb_stereo_match_batch_device.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, f32, f32, vp, vp, vp]
    pi32 = C.POINTER(i32)
    L.orb_search_for_triangulation.argtypes = [i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, f32, f32, vp, vp, i32, i32, i32, vp, pi32]
    L.orb_search_by_sim3.argtypes = [i32, vp, vp, i32, vp, vp, vp, i32, vp] + [vp] * 12 + [i32, vp, pi32]
    L.orb_distinctive_descriptors.argtypes = [i32, vp, vp, i32, vp, vp]
    L.orb_fuse_search.argtypes = [i32, vp, vp, vp, i32, vp, vp, i32, i32] + [vp] * 9
    L.orb_search_by_bow.argtypes = [i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, f32, i32, vp, vp, pi32]
    L.orb_voc_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    L.orb_voc_load_text.argtypes = [C.POINTER(vp), i32, C.c_char_p]
    L.orb_voc_destroy.argtypes = [vp]
    L.orb_voc_destroy.restype = None
    L.orb_voc_info.argtypes = [vp] + [pi32] * 6
    L.orb_bow_transform_features.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.orb_bow_transform_features_device.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    L.orb_bow_transform.argtypes = [vp, vp, vp, i32, i32] + [vp] * 7
    L.orb_bow_transform_device.argtypes = [vp, vp, vp, i32, i32, i32, i32] + [vp] * 9
    L.orb_shard_unique_id.argtypes = [vp]
    L.orb_db_create_sharded.argtypes = [C.POINTER(vp), i32, i64, i64, i32, i32, vp]
    L.orb_db_query_top2_sharded.argtypes = [vp, vp, i32, vp]
    L.orb_db_query_top2_sharded_host.argtypes = [vp, vp, i32, vp]
    L.orb_debug_sincos_range.argtypes = [i32, C.c_uint32, C.c_longlong, vp, vp]
    L.orb_bench_issue_rate.argtypes = [i32, i32, i32, C.POINTER(C.c_double)]
    psz = C.POINTER(C.c_size_t)
    L.orb_mat_record_bytes.argtypes = [i32, i32, C.c_size_t, psz]
    L.orb_mat_record_encode.argtypes = [vp, i32, i32, C.c_size_t, C.c_size_t, vp, C.c_size_t, psz]
    L.orb_mat_record_decode.argtypes = [vp, C.c_size_t, pi32, pi32, psz, psz, C.POINTER(vp), psz]
    L.orb_keypoint_records_encode.argtypes = [vp, i32, vp]
    L.orb_keypoint_records_decode.argtypes = [vp, i32, vp]
    L.orb_db_add_mat_record.argtypes = [vp, vp, C.c_size_t, psz, C.POINTER(C.c_int64)]
    _lib = L
    return L


def check(rc):
    if rc != ORB_OK:
        raise OrbError(rc, lib().orb_last_error().decode("utf-8", "replace"))


def ptr(a):
    """host numpy array -> void* (None passes NULL)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
