"""Frame::ComputeStereoMatches (reference orb_slam2/src/Frame.cc:502-676) over the C ABI."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib, ptr


def compute_stereo_matches(extractor_left, extractor_right, kps_l, desc_l, kps_r, desc_r, bf, b):
    """The two extractors must have just processed the left / right image (their pyramids are read, like
    mpORBextractorLeft/Right->mvImagePyramid).  Returns (nmatches, mvuRight[N], mvDepth[N])."""
    kps_l = np.ascontiguousarray(kps_l, KP_DTYPE); kps_r = np.ascontiguousarray(kps_r, KP_DTYPE)
    desc_l = np.ascontiguousarray(desc_l, np.uint8); desc_r = np.ascontiguousarray(desc_r, np.uint8)
    n = len(kps_l)
    ur = np.full(n, -1, np.float32); depth = np.full(n, -1, np.float32); nm = C.c_int32(0)
    check(lib().orb_stereo_match(extractor_left._h, extractor_right._h, ptr(kps_l), ptr(desc_l), n, ptr(kps_r), ptr(desc_r),
                                 len(kps_r), bf, b, ptr(ur), ptr(depth), C.byref(nm)))
    return nm.value, ur, depth


def ComputeStereoMatchesBatchDevice(extractor_left, extractor_right, npairs, d_kps_l, d_desc_l, d_n_l, d_kps_r, d_desc_r, d_n_r, cap, bf, b,
                                    d_u_right, d_depth, d_nmatches):
    """Frame::ComputeStereoMatches for npairs pairs at once, everything device-resident (raw device pointers: the outputs of
    the two extractors' extract_batch_device calls); asynchronous on the left extractor's stream."""
    import ctypes as C
    vp = C.c_void_p
    check(lib().orb_stereo_match_batch_device(extractor_left._h, extractor_right._h, npairs, vp(d_kps_l), vp(d_desc_l), vp(d_n_l), vp(d_kps_r),
                                              vp(d_desc_r), vp(d_n_r), cap, bf, b, vp(d_u_right), vp(d_depth), vp(d_nmatches)))
