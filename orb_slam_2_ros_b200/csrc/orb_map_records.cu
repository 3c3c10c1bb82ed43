// orb_map_records.cu — SURVEY.md §8f N4: the record payloads of the reference's saved map that feed the GPU descriptor
// database (reference orb_slam2/include/BoostArchiver.h:46-91; written by KeyFrame::serialize, KeyFrame.cc:858-864, into
// a boost binary archive opened with no_header, System.cc:627: primitives are raw native-endian bytes).
// Host-side codecs + one loader that appends a decoded N x 32 descriptor matrix to a device shard.  Only these two
// payloads are interpreted; the archive's class-id / object-tracking preambles and the Map pointer graph are the caller's.
#include <cstring>

#include "orb_internal.cuh"

namespace {
constexpr size_t MAT_HEADER = 4 + 4 + 8 + 8;   // int cols, int rows, size_t elem_size, size_t elem_type (BoostArchiver.h:68-71)
constexpr size_t KP_RECORD = 28;               // 7 x 4 bytes (BoostArchiver.h:49-57)
}

extern "C" {

int orb_mat_record_bytes(int rows, int cols, size_t elem_size, size_t* bytes) {
    if (!bytes || rows < 0 || cols < 0) return ORB_ERR_INVALID;
    *bytes = MAT_HEADER + (size_t)rows * (size_t)cols * elem_size;   // data_size = cols * rows * elem_size (BoostArchiver.h:73)
    return ORB_OK;
}

int orb_mat_record_encode(const uint8_t* data, int rows, int cols, size_t elem_size, size_t elem_type, uint8_t* out, size_t cap,
                          size_t* written) {
    size_t need;
    if (orb_mat_record_bytes(rows, cols, elem_size, &need) != ORB_OK || !out || (!data && need > MAT_HEADER)) return ORB_ERR_INVALID;
    if (cap < need) { orb_set_error("orb_mat_record_encode: %zu bytes needed, %zu given", need, cap); return ORB_ERR_CAPACITY; }
    const int32_t c = cols, r = rows;
    const uint64_t es = elem_size, et = elem_type;
    memcpy(out, &c, 4); memcpy(out + 4, &r, 4); memcpy(out + 8, &es, 8); memcpy(out + 16, &et, 8);   // cols BEFORE rows (:68-69)
    if (need > MAT_HEADER) memcpy(out + MAT_HEADER, data, need - MAT_HEADER);
    if (written) *written = need;
    return ORB_OK;
}

int orb_mat_record_decode(const uint8_t* buf, size_t len, int* rows, int* cols, size_t* elem_size, size_t* elem_type,
                          const uint8_t** data, size_t* consumed) {
    if (!buf) return ORB_ERR_INVALID;
    if (len < MAT_HEADER) { orb_set_error("orb_mat_record_decode: truncated header (%zu bytes)", len); return ORB_ERR_INVALID; }
    int32_t c, r;
    uint64_t es, et;
    memcpy(&c, buf, 4); memcpy(&r, buf + 4, 4); memcpy(&es, buf + 8, 8); memcpy(&et, buf + 16, 8);   // load order (:82-85)
    if (c < 0 || r < 0 || es > 64) { orb_set_error("orb_mat_record_decode: implausible header %d x %d x %llu", r, c, (unsigned long long)es); return ORB_ERR_INVALID; }
    // map files are untrusted input: c, r < 2^31 and es <= 64 could still wrap a 64-bit product, so bound the factors against the
    // bytes that are actually there before multiplying
    const size_t avail = len - MAT_HEADER;
    if (es == 0 ? (false) : (c != 0 && r != 0 && ((size_t)c > avail / (size_t)es || (size_t)r > avail / (size_t)es / (size_t)c))) {
        orb_set_error("orb_mat_record_decode: truncated payload");
        return ORB_ERR_INVALID;
    }
    {   // elemSize must be the size CV_ELEM_SIZE gives for the stored type (depth in the low 3 bits, channels - 1 above)
        static const unsigned depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        const uint64_t want = (uint64_t)depth_bytes[et & 7] * (((et >> 3) & 511) + 1);
        if (et > 0xFFF || want != es) { orb_set_error("orb_mat_record_decode: elemSize %llu does not match type %llu", (unsigned long long)es, (unsigned long long)et); return ORB_ERR_INVALID; }
    }
    const size_t payload = (size_t)c * (size_t)r * (size_t)es;
    if (avail < payload) { orb_set_error("orb_mat_record_decode: truncated payload"); return ORB_ERR_INVALID; }
    if (rows) *rows = r;
    if (cols) *cols = c;
    if (elem_size) *elem_size = (size_t)es;
    if (elem_type) *elem_type = (size_t)et;
    if (data) *data = buf + MAT_HEADER;
    if (consumed) *consumed = MAT_HEADER + payload;
    return ORB_OK;
}

int orb_keypoint_records_encode(const orb_kp* kps, int n, uint8_t* out) {
    if (n < 0 || (n && (!kps || !out))) return ORB_ERR_INVALID;
    for (int i = 0; i < n; ++i) {
        uint8_t* o = out + (size_t)i * KP_RECORD;
        // ar & angle & class_id & octave & response & response & pt.x & pt.y   (size is never written)
        memcpy(o, &kps[i].angle, 4); memcpy(o + 4, &kps[i].class_id, 4); memcpy(o + 8, &kps[i].octave, 4);
        memcpy(o + 12, &kps[i].response, 4); memcpy(o + 16, &kps[i].response, 4);
        memcpy(o + 20, &kps[i].x, 4); memcpy(o + 24, &kps[i].y, 4);
    }
    return ORB_OK;
}

int orb_keypoint_records_decode(const uint8_t* buf, int n, orb_kp* kps) {
    if (n < 0 || (n && (!kps || !buf))) return ORB_ERR_INVALID;
    for (int i = 0; i < n; ++i) {
        const uint8_t* o = buf + (size_t)i * KP_RECORD;
        memcpy(&kps[i].angle, o, 4); memcpy(&kps[i].class_id, o + 4, 4); memcpy(&kps[i].octave, o + 8, 4);
        memcpy(&kps[i].response, o + 12, 4);
        memcpy(&kps[i].response, o + 16, 4);   // the second read lands in response again (BoostArchiver.h:53-54)
        memcpy(&kps[i].x, o + 20, 4); memcpy(&kps[i].y, o + 24, 4);
        kps[i].size = 0.f;                     // cv::KeyPoint() default: the archive carries no size
    }
    return ORB_OK;
}

int orb_db_add_mat_record(orb_db* db, const uint8_t* buf, size_t len, size_t* consumed, int64_t* rows_added) {
    int rows, cols;
    size_t es, et;
    const uint8_t* data;
    size_t used;
    const int rc = orb_mat_record_decode(buf, len, &rows, &cols, &es, &et, &data, &used);
    if (rc != ORB_OK) return rc;
    if (rows > 0 && (cols != 32 || es != 1)) {   // mDescriptors is N x 32 CV_8UC1 (ORBextractor.cc:1112)
        orb_set_error("orb_db_add_mat_record: %d x %d matrix with %zu-byte elements is not a descriptor matrix", rows, cols, es);
        return ORB_ERR_INVALID;
    }
    const int rc2 = rows > 0 ? orb_db_add(db, data, rows) : (db ? ORB_OK : ORB_ERR_INVALID);
    if (rc2 != ORB_OK) return rc2;
    if (consumed) *consumed = used;
    if (rows_added) *rows_added = rows;
    return ORB_OK;
}

}  // extern "C"
